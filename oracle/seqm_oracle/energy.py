"""Energies.  Restates seqm/seqm_functions/energy.py:8-216."""
import numpy as np

from .tables import Tables


def elec_energy(D, F, H):
    """Eelec = 1/2 sum D (H + F), H symmetric (energy.py:26-53)."""
    return 0.5 * np.sum(D * (H + F), axis=(1, 2))


def elec_energy_xl(Dm, Pm, F, H):
    """XL-BOMD shadow energy sum D F - 1/2 (F - H) P   (energy.py:76-88)."""
    return np.sum(Dm * F - 0.5 * (F - H) * Pm, axis=(1, 2))


def pair_nuclear_energy(method, ni, nj, idxi, idxj, rij, gam, par, rho0=None):
    """Core-core repulsion per pair for MNDO / AM1 / PM3 (energy.py:91-139) and PM6_SP (energy.py:140-171;
    rho0 = per-atom rho_0 with rho_core substituted where non-zero, two_elec_two_center_int.py:273-281)."""
    T = Tables.get()
    rija = rij * T.a0
    if method == "PM6_SP":
        ng = par["_ngauss"]
        K = np.stack([par[f"Gaussian{g}_K"] for g in range(1, ng + 1)], axis=1)
        L = np.stack([par[f"Gaussian{g}_L"] for g in range(1, ng + 1)], axis=1)
        M = np.stack([par[f"Gaussian{g}_M"] for g in range(1, ng + 1)], axis=1)
        zz = T.tore[ni] * T.tore[nj]
        t4 = zz / rija
        t5 = np.sum(K[idxi] * np.exp(-L[idxi] * (rija[:, None] - M[idxi]) ** 2), axis=1)
        t6 = np.sum(K[idxj] * np.exp(-L[idxj] * (rija[:, None] - M[idxj]) ** 2), axis=1)
        alp, chi = par["_alp"][ni, nj], par["_chi"][ni, nj]
        unpol = 1.0e-8 * ((T.atomic_num[ni] ** (1.0 / 3.0) + T.atomic_num[nj] ** (1.0 / 3.0)) / rija) ** 12
        g0 = zz * T.ev / np.sqrt(rij * rij + (rho0[idxi] + rho0[idxj]) ** 2)
        XH = ((ni == 6) | (ni == 7) | (ni == 8)) & (nj == 1)
        scale = np.where(XH, 1.0 + 2.0 * chi * np.exp(-alp * rija**2), 1.0 + 2.0 * chi * np.exp(-alp * (rija + 0.0003 * rija**6)))
        E = unpol + g0 * scale
        E = E + np.where((ni == 6) & (nj == 6), g0 * 9.28 * np.exp(-5.98 * rija), 0.0)
        E = E - np.where((ni == 14) & (nj == 8), g0 * 0.0007 * np.exp(-((rij - 2.9) ** 2)), 0.0)
        return E + t4 * (t5 + t6)
    alpha = par["alpha"]
    t1 = T.tore[ni] * T.tore[nj] * gam
    XH = ((ni == 7) | (ni == 8)) & (nj == 1)
    tmp = np.exp(-alpha[idxi] * rija)
    t2 = np.where(XH, tmp * rija, tmp)
    t3 = np.exp(-alpha[idxj] * rija)
    E = t1 * (1.0 + t2 + t3)
    if method == "MNDO":
        return E
    if method in ("AM1", "PM3"):
        ng = par["_ngauss"]
        K = np.stack([par[f"Gaussian{g}_K"] for g in range(1, ng + 1)], axis=1)
        L = np.stack([par[f"Gaussian{g}_L"] for g in range(1, ng + 1)], axis=1)
        M = np.stack([par[f"Gaussian{g}_M"] for g in range(1, ng + 1)], axis=1)
        t4 = T.tore[ni] * T.tore[nj] / rija
        t5 = np.sum(K[idxi] * np.exp(-L[idxi] * (rija[:, None] - M[idxi]) ** 2), axis=1)
        t6 = np.sum(K[idxj] * np.exp(-L[idxj] * (rija[:, None] - M[idxj]) ** 2), axis=1)
        return E + t4 * (t5 + t6)
    raise ValueError("Supported Method: MNDO, AM1, PM3")


def isolated_atom_energy(Z, par):
    """energy.py:8-23."""
    T = Tables.get()
    return (
        par["U_ss"] * T.ussc[Z]
        + par["U_pp"] * T.uppc[Z]
        + par["g_ss"] * T.gssc[Z]
        + par["g_pp"] * T.gppc[Z]
        + par["g_sp"] * T.gspc[Z]
        + par["g_p2"] * T.gp2c[Z]
        + par["h_sp"] * T.hspc[Z]
    )


def molecule_sums(vals, molid, nmol):
    out = np.zeros(nmol)
    np.add.at(out, molid, vals)
    return out
