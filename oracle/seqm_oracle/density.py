"""Density matrix from the Fock matrix: eigensolver route and SP2 purification.

Restates: seqm/seqm_functions/pack.py:8-96 (packed layout [4*nHeavy heavy AOs][nHydro H s AOs]),
          seqm/seqm_functions/diag.py:110-241 (sym_eig_trunc; the Gershgorin-shifted padding of
          diag.py:168-204 only exists to batch unequal sizes through one eigh call and has no effect on
          P or on the physical eigenpairs, so molecules are solved at their own size here),
          seqm/seqm_functions/SP2.py:9-85.
"""
import numpy as np


def packed_index(nheavy, nhydro):
    """Indices (into the padded 4*molsize basis) of the real orbitals in packed order (pack.py:8-16)."""
    return np.concatenate([np.arange(4 * nheavy), 4 * nheavy + 4 * np.arange(nhydro)])


def eig_packed(F, nheavy, nhydro):
    idx = packed_index(nheavy, nhydro)
    e, v = np.linalg.eigh(F[np.ix_(idx, idx)], UPLO="U")
    return idx, e, v


def density_from_fock(F, nHeavy, nHydro, nocc, mols=None, want_eig=False):
    """P = 2 C_occ C_occ^T per molecule, returned in the padded layout (diag.py:206-232, pack.py:85-96)."""
    nmol, N, _ = F.shape
    D = np.zeros_like(F)
    E = np.zeros((nmol, N))
    V = None
    if want_eig:
        nmax = int(np.max(4 * nHeavy + nHydro))
        V = np.zeros((nmol, nmax, nmax))
    for m in range(nmol):
        if mols is not None and not mols[m]:
            continue
        idx, e, v = eig_packed(F[m], int(nHeavy[m]), int(nHydro[m]))
        c = v[:, : int(nocc[m])]
        D[m][np.ix_(idx, idx)] = 2.0 * (c @ c.T)
        E[m, : e.shape[0]] = e
        if want_eig:
            n = e.shape[0]
            V[m, :n, :n] = v
            for k in range(n, V.shape[1]):
                V[m, k, k] = 1.0
    return D, E, V


def sp2_packed(a, nocc, eps):
    """SP2 on one packed symmetric matrix; returns (2*X, number of X^2 products).  SP2.py:9-85."""
    eps = min(max(eps, 1.0e-7), 1.0e-3)
    n = a.shape[0]
    aii = np.diag(a)
    ri = np.sum(np.abs(a), axis=1) - np.abs(aii)
    h1 = np.min(aii - ri)
    hN = np.max(aii + ri)
    x = (np.eye(n) * hN - a) / (hN - h1)
    errm0 = abs(np.trace(x) - nocc)
    errm1 = errm0
    k = 0
    while True:
        x2 = x @ x
        tr2 = np.trace(x2)
        if abs(tr2 - nocc) < abs(2.0 * np.trace(x) - tr2 - nocc):
            x = x2
        else:
            x = 2.0 * x - x2
        errm1, errm0 = errm0, abs(np.trace(x) - nocc)
        k += 1
        if errm0 < eps and errm1 < eps:
            break
        if k > 10000:
            raise RuntimeError("SP2 did not converge")
    return 2.0 * x, k


def sp2_density(F, nHeavy, nHydro, nocc, eps, mols=None, reference_padding=True):
    """SP2 density per molecule.

    reference_padding=True reproduces the reference exactly: pack() zero-pads every packed matrix to the
    largest orbital count among the molecules handed to it (pack.py:76-77; the SCF loop hands it only the
    not-yet-converged ones, scf_loop.py:66-78), so smaller molecules carry extra zero eigenvalues through
    SP2 and their density differs at the O(eps) level from the unpadded purification.
    reference_padding=False purifies each molecule at its own size (what pyseqm_b200 does; identical
    for uniform batches such as MD replicas or a single molecule).
    """
    nmol = F.shape[0]
    D = np.zeros_like(F)
    act = [m for m in range(nmol) if mols is None or mols[m]]
    nmax = max(int(4 * nHeavy[m] + nHydro[m]) for m in act) if act else 0
    for m in act:
        idx = packed_index(int(nHeavy[m]), int(nHydro[m]))
        n = idx.shape[0]
        a = F[m][np.ix_(idx, idx)]
        if reference_padding and n < nmax:
            ap = np.zeros((nmax, nmax))
            ap[:n, :n] = a
            a = ap
        d, _ = sp2_packed(a, float(nocc[m]), eps)
        D[m][np.ix_(idx, idx)] = d[:n, :n]
    return D
