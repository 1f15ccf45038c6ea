"""KSA-XL-BOMD (oracle): finite-temperature density, canonical density-matrix perturbation theory and the rank-m Krylov
approximation of the kernel that drives the field density.  TEST INFRASTRUCTURE (see oracle/README): nothing in
pyseqm_b200/ imports this.

Restates: seqm/seqm_functions/fermi_q.py:8-72 (Fermi_Q: Newton iteration for the chemical potential with a BATCH-GLOBAL stop
          test, D0 = 2 Q f Q^t, entropy),
          seqm/seqm_functions/canon_dm_prt.py:6-39 (Canon_DM_PRT, Alg. 2 of JCTC 16, 3628 (2020): recursive Fermi-operator
          expansion of the first-order response in the eigenbasis, chemical-potential correction),
          seqm/dynamics/xlbomd.py:201-341 (Krylov branch of EnergyXL.forward; CANON_DM_PRT_ITER = 10, xlbomd.py:55),
          seqm/seqm_functions/G_XL_LR.py:7 (G = the Fock build without the one-electron part),
          seqm/MolecularDynamics.py:1608-1619 (KSA_XL_BOMD._propagate_P) and 1576-1582 (dP2dt2 = 0 at t = 0).
Pinned by tests/test_oracle_golden.py to tests/golden/ksa_operators.npz and md_ksa_*.npz (tools/make_golden_ksa.py).
"""
import numpy as np

from .density import eig_packed, packed_index
from .energy import elec_energy_xl, isolated_atom_energy, molecule_sums, pair_nuclear_energy
from .gradient import hf_gradient
from .hamiltonian import build_fock, build_hcore
from .integrals import atom_multipoles, rho0_eff
from .parser import parse
from .tables import Tables, method_parameters

KB = 8.61739e-5  # eV/K (xlbomd.py:207)
CANON_DM_PRT_ITER = 10


def fermi_q(F, T_el, nocc, nHeavy, nHydro):
    """-> D0 (padded layout), S, list of (idx, e, Q) per molecule, occupations f (nmol, nmax), mu (nmol,)"""
    nmol = F.shape[0]
    norb = 4 * nHeavy + nHydro
    nmax = int(norb.max())
    beta = 1.0 / (KB * T_el)
    eig = [eig_packed(F[m], int(nHeavy[m]), int(nHydro[m])) for m in range(nmol)]
    e = np.zeros((nmol, nmax))
    for m in range(nmol):
        e[m, : norb[m]] = eig[m][1]
    mask = (np.arange(nmax)[None, :] < norb[:, None]).astype(np.float64)
    ar = np.arange(nmol)
    mu = 0.5 * (e[ar, nocc - 1] + e[ar, nocc])
    f = None
    for _ in range(64):
        f = mask / (1.0 + np.exp(beta * (e - mu[:, None])))
        occ = f.sum(axis=1)
        docc = np.maximum((beta * f * (1.0 - f)).sum(axis=1), 1e-30)
        if np.all(np.abs(nocc - occ) <= 1e-9):  # every molecule keeps iterating until ALL have converged
            break
        mu = mu + (nocc - occ) / docc
    D0 = np.zeros_like(F)
    for m in range(nmol):
        idx, _, Q = eig[m]
        D0[m][np.ix_(idx, idx)] = 2.0 * (Q * f[m, : norb[m]][None, :]) @ Q.T
    ok = (f > 1e-14) & ((1.0 - f) > 1e-14)
    p = np.where(ok, f, 0.5)
    S = np.sum(np.where(ok, -KB * (p * np.log(p) + (1.0 - p) * np.log(1.0 - p)), 0.0), axis=1)
    return D0, S, eig, f, mu


def canon_dm_prt(FO1, T_el, eig, mu, m_iter=CANON_DM_PRT_ITER):
    """First-order response of the finite-temperature density to the perturbation FO1 (padded layout in and out)."""
    beta = 1.0 / (KB * T_el)
    cnst = 2.0 ** (-2 - m_iter) * beta
    P1 = np.zeros_like(FO1)
    for m in range(FO1.shape[0]):
        idx, e, Q = eig[m]
        X = Q.T @ FO1[m][np.ix_(idx, idx)] @ Q
        p0 = (0.5 - cnst * (e - mu[m]))[:, None]
        X = -cnst * X
        for _ in range(m_iter):
            p02 = p0 * p0
            dX = p0 * X + X * p0.T
            iD0 = 1.0 / (2.0 * (p02 - p0) + 1.0)
            p0 = iD0 * p02
            X = iD0 * (dX + 2.0 * (X - dX) * p0.T)
        dpdmu = beta * p0 * (1.0 - p0)
        dmu1 = -np.trace(X) / dpdmu.sum()
        X = X + np.diag(dpdmu[:, 0]) * dmu1
        P1[m][np.ix_(idx, idx)] = Q @ X @ Q.T
    return P1


def scf_ksa(P, par, H, w, D0, eps, xl, max_iter=1000):
    """scf_forward3 (scf_loop.py:1135-1381): field <- field - sum_k alpha_k V_k per iteration, stop on |dEelec| <= eps per
    molecule; CANON_DM_PRT_ITER = 8 in this file of the reference (scf_loop.py:47).  Converged molecules keep their field (and
    therefore their Fock matrix and Fermi data), which is what the reference's refresh of the unconverged subset amounts to.
    -> field density, notconverged, iterations"""
    from .energy import elec_energy

    T_el = xl["T_el"]
    field = np.array(D0, dtype=np.float64)
    nmol = field.shape[0]
    F = build_fock(P, par, H, w, field)
    Eelec = np.zeros(nmol)
    err = np.ones(nmol)
    notconv = np.ones(nmol, dtype=bool)
    n_iter = 0
    while notconv.any() and n_iter < max_iter:
        n_iter += 1
        D, S, eig, f, mu = fermi_q(F, T_el, P.nocc, P.nHeavy, P.nHydro)
        d2, _ = krylov_kernel(P, par, w, D, field, eig, mu, T_el, xl["max_rank"], xl["err_threshold"], m_iter=8)
        field[notconv] += d2[notconv]  # d2 = -sum_k alpha_k V_k
        F = build_fock(P, par, H, w, field)
        Enew = elec_energy(field, F, H)
        err[notconv] = np.abs(Enew - Eelec)[notconv]
        Eelec[notconv] = Enew[notconv]
        notconv = err > eps
    return field, notconv, n_iter


def krylov_kernel(P, par, w, D, field, eig, mu, T_el, max_rank, err_threshold, m_iter=CANON_DM_PRT_ITER):
    """Rank-m approximation of the kernel acting on the residual D - field (xlbomd.py:238-341) -> dP2dt2, Error."""
    fro = lambda A: np.sqrt(np.sum(A * A, axis=(1, 2)))  # noqa: E731
    dDS = D - field
    nrm = fro(dDS)
    V, W = [], []
    dW = dDS
    Error = np.full(D.shape[0], 10.0)
    H0 = np.zeros_like(D)
    alpha = None
    while len(V) < max_rank and Error.max() > err_threshold:
        v = dW.copy()
        for vj in V:
            v = v - np.sum(v.transpose(0, 2, 1) * vj, axis=(1, 2))[:, None, None] * vj
        v = v / fro(v)[:, None, None]
        V.append(v)
        FO1 = build_fock(P, par, H0, w, v)  # G(dD): the Fock build without Hcore (G_XL_LR.py:7)
        PO1 = canon_dm_prt(FO1, T_el, eig, mu, m_iter)
        W.append(PO1 - v)
        dW = W[-1]
        r = len(W)
        O = np.array([[np.sum(W[a].transpose(0, 2, 1) * W[b], axis=(1, 2)) for b in range(r)] for a in range(r)]).transpose(2, 0, 1)
        rhs = np.array([np.sum(W[a].transpose(0, 2, 1) * dDS, axis=(1, 2)) for a in range(r)]).T
        alpha = np.linalg.solve(O, rhs[:, :, None])[:, :, 0]
        ident = sum(W[a] * alpha[:, a][:, None, None] for a in range(r))
        Error = fro(ident - dDS) / nrm
    dP2dt2 = -sum(V[a] * alpha[:, a][:, None, None] for a in range(len(V)))
    return dP2dt2, Error


def ksa_forward(species, coordinates, seqm_parameters, field, xl):
    """Electronic_Structure.forward(dm_prop="XL-BOMD", xl_bomd_params with max_rank): dict(force, dm, Etot, Electronic_entropy,
    dP2dt2, Krylov_Error, Fermi_occ, e_gap, ...)."""
    T = Tables.get()
    method = seqm_parameters["method"]
    P = parse(species, coordinates)
    par = method_parameters(method, P.Z)
    mp = atom_multipoles(P.Z, par)
    hc = build_hcore(P, par, mp)
    H, w = hc["H"], hc["w"]
    F = build_fock(P, par, H, w, field)
    T_el = xl["T_el"]
    D, S, eig, f, mu = fermi_q(F, T_el, P.nocc, P.nHeavy, P.nHydro)
    dP2dt2, Error = krylov_kernel(P, par, w, D, field, eig, mu, T_el, xl["max_rank"], xl["err_threshold"])
    e_gap = np.array([eig[m][1][P.nocc[m]] - eig[m][1][P.nocc[m] - 1] for m in range(P.nmol)])
    Eelec = elec_energy_xl(D, field, F, H)
    EnucAB = pair_nuclear_energy(method, P.ni, P.nj, P.idxi, P.idxj, P.rij, w[:, 0, 0], par, rho0=rho0_eff(par, mp))
    Enuc = molecule_sums(EnucAB, P.pair_molid, P.nmol)
    Etot = Eelec + Enuc
    Eiso = molecule_sums(isolated_atom_energy(P.Z, par), P.atom_molid, P.nmol)
    Hf = Etot - Eiso + molecule_sums(T.eheat[P.Z], P.atom_molid, P.nmol)
    force = -hf_gradient(P, par, method, D, mp, field=field)
    return dict(force=force, dm=D, Etot=Etot, Hf=Hf, Eelec=Eelec, Enuc=Enuc, Eiso=Eiso, e_gap=e_gap,
                Electronic_entropy=-2.0 * T_el * S, dP2dt2=dP2dt2, Krylov_Error=Error, Fermi_occ=f)  # fmt: skip


def run_ksa_md(species, coordinates, velocities, seqm_parameters, timestep, steps, xl):
    """KSA_XL_BOMD driven step by step from the converged SCF at t = 0 (dP2dt2 = 0 in the first propagation)."""
    from .api import single_point
    from .md import ACC_SCALE, kinetic_energy, xl_coefficients

    T = Tables.get()
    species = np.asarray(species)
    x = np.array(coordinates, dtype=np.float64)
    v = np.array(velocities, dtype=np.float64)
    mass = T.mass[species][:, :, None]
    minv = np.where(species[:, :, None] > 0, 1.0 / np.where(mass > 0, mass, 1.0), 0.0)
    r = single_point(species, x, seqm_parameters)
    k = xl["k"]
    m = k + 1
    kappa, coeff = xl_coefficients(k)
    Pf = r["dm"].copy()
    Pt = np.stack([Pf.copy() for _ in range(m)])
    d2 = np.zeros_like(Pf)
    acc = r["force"] * minv * ACC_SCALE
    out = dict(Etot=[], Ek=[], Electronic_entropy=[], Krylov_Error=[])
    for i in range(steps):
        v += 0.5 * acc * timestep
        x += v * timestep
        cindx = i % m
        Pf = kappa * (d2 + Pf) + np.tensordot(coeff[cindx : cindx + m], Pt, axes=(0, 0))
        Pt[m - 1 - cindx] = Pf
        r = ksa_forward(species, x, seqm_parameters, Pf, xl)
        d2 = r["dP2dt2"]
        acc = r["force"] * minv * ACC_SCALE
        v += 0.5 * acc * timestep
        out["Etot"].append(r["Etot"].copy())
        out["Ek"].append(kinetic_energy(mass, v))
        out["Electronic_entropy"].append(r["Electronic_entropy"].copy())
        out["Krylov_Error"].append(r["Krylov_Error"].copy())
    res = {k_: np.stack(v_) for k_, v_ in out.items()}
    res.update(coordinates=x, velocities=v, force=r["force"], dm=r["dm"], dP2dt2=d2, Fermi_occ=r["Fermi_occ"])
    return res
