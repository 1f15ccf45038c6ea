"""SCF drivers: constant mixing [0, alpha], adaptive mixing [1], Pulay DIIS [2].

Restates seqm/seqm_functions/scf_loop.py: get_error 106-147, scf_forward0 164-347,
compute_fac/adaptive_mix 350-420, scf_forward1 424-635, scf_forward2 639-1132 (nDirect1 = nAdapt = 0,
nFock = 10, batch-global DIIS reset), MAX_ITER = 1000 (line 29).
All batch-coupled control flow of the reference (the global DIIS reset, the `torch.all(done)` exit of
the adaptive-mix renormalisation) is kept, because it decides iteration counts.
"""
import numpy as np

from .density import density_from_fock, sp2_density
from .energy import elec_energy
from .hamiltonian import build_fock

MAX_ITER = 1000
DM_ERR_FACTOR = 2.0  # scf_loop.py:50
DM_ELEM_FACTOR = 15.0  # scf_loop.py:51
DIIS_FACTOR = 50.0  # scf_loop.py:52


class _State:
    pass


def _get_error(S, Pold, Pm, nc, Eel_new, eps, diis_error=None):
    """scf_loop.py:106-147; mutates S.err, S.dm_err, S.dm_elem."""
    S.err[nc] = Eel_new[nc] - S.Eel[nc]
    bad = np.abs(S.err) > eps
    if diis_error is not None:
        bad = bad | (diis_error > DIIS_FACTOR * eps)
    dm_mask = nc & ~bad
    if np.any(dm_mask):
        dP = Pm[dm_mask] - Pold[dm_mask]
        S.dm_err[dm_mask] = np.sqrt(np.sum(dP * dP, axis=(1, 2))) / S.msize[dm_mask]
        S.dm_elem[dm_mask] = np.max(np.abs(dP), axis=(1, 2))
    return bad | (S.dm_err > eps * DM_ERR_FACTOR) | (S.dm_elem > eps * DM_ELEM_FACTOR)


def _adaptive_mix(k, P_prev, P_cur, old2_diag):
    """scf_loop.py:361-420 on the active sub-batch."""
    is_third = k % 3 == 0
    DAMP = 0.05 if k > 4 else 1.0e10
    d_prev = np.diagonal(P_prev, axis1=1, axis2=2).copy()
    d_cur = np.diagonal(P_cur, axis1=1, axis2=2).copy()
    nb = P_cur.shape[0]
    if is_third:
        diff1 = d_cur - d_prev
        diff2 = d_cur - 2.0 * d_prev + old2_diag
        num = np.sum(diff1**2, axis=1)
        den = np.sum(diff2**2, axis=1)
        valid = (den > 0) & (num < 100.0 * den)
        FAC = np.zeros(nb)
        FAC[valid] = np.sqrt(num[valid] / den[valid])
        Pmix = (1.0 + FAC)[:, None, None] * P_cur - FAC[:, None, None] * P_prev
    else:
        FAC = np.zeros(nb)
        Pmix = P_cur.copy()
    delta = d_cur - d_prev
    cap = np.abs(delta) > DAMP
    di = np.where(cap, d_prev + np.sign(delta) * DAMP, d_cur + FAC[:, None] * delta)
    di = np.clip(di, 0.0, 2.0)
    SUM0 = np.sum(d_cur, axis=1)
    for _ in range(20):
        SUM2 = np.sum(di, axis=1)
        large = SUM2 > 1.0e-3
        SUM3 = np.zeros_like(SUM2)
        SUM3[large] = SUM0[large] / SUM2[large]
        done = (~large) | (np.abs(SUM3 - 1.0) <= 1.0e-5)
        if np.all(done):
            break
        scaled = np.maximum(di * SUM3[:, None], 0.0)
        full = scaled > 2.0
        di = np.where(full, 2.0, scaled)
        SUM0 = SUM0 - np.sum(full, axis=1) * 2.0
    ar = np.arange(Pmix.shape[1])
    Pmix[:, ar, ar] = di
    return Pmix, d_prev


def _diis_coeff(EVEC, cF):
    """Pseudo-inverse solve of the Pulay system, lower triangle only (scf_loop.py:1011-1035)."""
    L, Q = np.linalg.eigh(EVEC, UPLO="L")  # batched over active molecules
    absv = np.abs(L)
    with np.errstate(divide="ignore", invalid="ignore"):
        cond = np.max(absv, axis=-1) / np.min(absv, axis=-1)
        inv = np.where(absv > 1.0e-13, 1.0 / L, 0.0)
    coeff = -np.einsum("bki,bi,bi->bk", Q[:, :cF, :], inv, Q[:, -1, :])
    return coeff, cond


def run_scf(P, par, H, w, D0, eps, converger=(2,), sp2=(False,), verbose=False, fock_fn=None, density_fn=None, msize=None):
    """Returns (D, notconverged, n_iter).  D0 is not modified.
    fock_fn(Pm) / density_fn(F, mask) / msize: the PM6 d-orbital path (pm6d.py) plugs its 9-slot Fock build, packed
    eigensolver and matrix_size_sqrt = 9 nSH + 4 nHeavy + 4 nHydro (scf_loop.py:245) into the same loop."""
    build = fock_fn if fock_fn is not None else (lambda Pm_: build_fock(P, par, H, w, Pm_))
    nmol = P.nmol
    N = H.shape[1]
    S = _State()
    S.err = np.ones(nmol)
    S.dm_err = np.ones(nmol)
    S.dm_elem = np.ones(nmol)
    S.msize = (4 * P.nHeavy + 4 * P.nHydro).astype(np.float64) if msize is None else np.asarray(msize, dtype=np.float64)  # scf_loop.py:728
    Pm = D0.copy()
    Pold = np.zeros_like(Pm)
    Pnew = np.zeros_like(Pm)
    nc = np.ones(nmol, dtype=bool)

    def make_pnew(F, mask):
        if density_fn is not None:
            return density_fn(F, mask)
        if sp2[0]:
            return sp2_density(F, P.nHeavy, P.nHydro, P.nocc, sp2[1], mask)
        return density_from_fock(F, P.nHeavy, P.nHydro, P.nocc, mask)[0]

    F = build(Pm)
    S.Eel = elec_energy(Pm, F, H)
    Eel_new = np.zeros(nmol)
    kind = converger[0]
    n_iter = 0

    if kind in (0, 1):
        alpha = converger[1] if kind == 0 else None
        old2 = np.zeros((nmol, N))
        ks = range(MAX_ITER + 1) if kind == 0 else range(1, MAX_ITER + 1)
        for k in ks:
            Pnew[nc] = make_pnew(F, nc)[nc]
            Pold[nc] = Pm[nc]
            if kind == 0:
                Pm[nc] = alpha * Pm[nc] + (1.0 - alpha) * Pnew[nc]
            else:
                Pmix, dprev = _adaptive_mix(k, Pm[nc], Pnew[nc], old2[nc])
                Pm[nc] = Pmix
                old2[nc] = dprev
            F = build(Pm)
            Eel_new[nc] = elec_energy(Pm[nc], F[nc], H[nc])
            nc_new = _get_error(S, Pold, Pm, nc, Eel_new, eps)
            nc = nc_new
            S.Eel[nc] = Eel_new[nc]
            n_iter = k
            if not np.any(nc):
                break
        return Pm, nc, n_iter

    if kind != 2:
        raise ValueError("scf_converger must be [0, alpha], [1] or [2]")
    nFock = 10
    iu = np.triu_indices(N)
    FPPF = np.zeros((nmol, nFock, iu[0].shape[0]))
    FOCK = np.zeros((nmol, nFock, N, N))

    def fresh_emat():
        return np.tile(np.tril(np.eye(nFock + 1) - 1.0)[None], (nmol, 1, 1))

    EMAT = fresh_emat()
    counter, cF = -1, 0
    diis_error = np.full(nmol, np.finfo(np.float64).max)
    k = 0
    for k in range(0, MAX_ITER + 1):
        if not np.any(nc):
            break
        cF = cF + 1 if cF < nFock else nFock
        counter = (counter + 1) % nFock
        act = np.nonzero(nc)[0]
        FOCK[act, counter] = F[act]
        C = F[act] @ Pm[act] - Pm[act] @ F[act]
        Cp = C[:, iu[0], iu[1]]
        FPPF[act, counter] = Cp
        diis_error[act] = np.max(np.abs(Cp), axis=1)
        EMAT[act, counter, :cF] = np.einsum("at,ajt->aj", Cp, FPPF[act, :cF])
        reset = False
        if cF >= 2:
            EVEC = EMAT[act, : cF + 1, : cF + 1].copy()
            denom = np.maximum(EVEC[:, counter, counter], 1.0e-15)
            EVEC[:, :cF, :cF] /= denom[:, None, None]
            coeff, cond = _diis_coeff(EVEC, cF)
            reset = bool(np.any(cond > 1.0e7))
            F[act] = np.matmul(coeff[:, None, :], FOCK[act, :cF].reshape(act.shape[0], cF, N * N)).reshape(-1, N, N)
        Pnew[nc] = make_pnew(F, nc)[nc]
        Pold[nc] = Pm[nc]
        if cF < 2:
            Pm[nc] = 0.5 * Pm[nc] + 0.5 * Pnew[nc]
        else:
            Pm[nc] = Pnew[nc]
        F = build(Pm)
        Eel_new[nc] = elec_energy(Pm[nc], F[nc], H[nc])
        nc = _get_error(S, Pold, Pm, nc, Eel_new, eps, diis_error)
        S.Eel[nc] = Eel_new[nc]
        if reset:
            counter, cF = -1, 0
            FPPF[:] = 0.0
            FOCK[:] = 0.0
            EMAT = fresh_emat()
    else:
        k = MAX_ITER + 1
    return Pm, nc, k
