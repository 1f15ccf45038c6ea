"""Hellmann-Feynman nuclear gradient of the SCF energy at fixed density.

Restates what seqm/seqm_functions/anal_grad.py:16-225 computes (pair gradient dE_pair/dR_i contracted
with the density, scattered +/- onto the two atoms).  The reference differentiates the two-electron
integrals analytically (anal_grad.py:561-635, 718-1605) and the overlaps by central finite differences
with delta = 1e-5 Angstrom (anal_grad.py:641-715).  This oracle takes the central difference of the WHOLE
pair energy (same delta); it is pinned to the reference's analytic and autograd forces at <= 1e-6 eV/A
by tests/test_oracle_golden.py.
"""
import numpy as np

from .energy import pair_nuclear_energy
from .integrals import PACK, PACK_COL, PACK_ROW, WEIGHT, atom_multipoles, overlap_sp, rho0_eff, two_center_integrals_geom
from .tables import Tables

DELTA = 1.0e-5  # Angstrom, anal_grad.py:13


def _pair_energy(P, par, mp, method, xij, rij, PAi, PBj, Dab, pki, pkj, xl=None):
    """Energy terms that depend on the geometry of each pair, at fixed density blocks.
    xl = (field blocks) switches to the XL-BOMD shadow energy sum D o F(P) - 1/2 (F(P)-h) o P:
    one-electron terms with D, two-electron terms with (D - P/2) x P  (energy.py:76-88)."""
    T = Tables.get()
    w, e1b, e2a, _, _ = two_center_integrals_geom(P.ni, P.nj, P.idxi, P.idxj, xij, rij, mp, T)
    zeta = np.stack([par["zeta_s"], par["zeta_p"]], axis=1)
    di = overlap_sp(P.ni, P.nj, xij, rij, zeta[P.idxi], zeta[P.idxj])
    bA = np.stack([par["beta_s"]] + [par["beta_p"]] * 3, axis=1)
    bsum = bA[P.idxi][:, :, None] + bA[P.idxj][:, None, :]  # 2 * (beta_i+beta_j)/2: both triangles
    E = np.sum(Dab * di * bsum, axis=(1, 2))
    sym = np.triu(np.ones((4, 4)), 1) + np.triu(np.ones((4, 4)))  # 1 on diagonal, 2 above
    E += np.sum(PAi * e1b * sym, axis=(1, 2)) + np.sum(PBj * e2a * sym, axis=(1, 2))
    w4 = w[:, PACK[:, :, None, None], PACK[None, None, :, :]]  # (p, mu, nu, lam, sig)
    wx = w4.transpose(0, 1, 3, 2, 4).reshape(-1, 16, 16)
    d16 = Dab.reshape(-1, 16)
    if xl is None:
        E += np.matmul(pki[:, None, :], np.matmul(w, pkj[:, :, None]))[:, 0, 0]
        E += -0.5 * np.matmul(d16[:, None, :], np.matmul(wx, d16[:, :, None]))[:, 0, 0]
    else:
        fki, fkj, Fab = xl  # packed weighted diagonal blocks and off-diagonal block of the field P
        xi, xj = pki - 0.5 * fki, pkj - 0.5 * fkj
        E += np.matmul(xi[:, None, :], np.matmul(w, fkj[:, :, None]))[:, 0, 0]
        E += np.matmul(fki[:, None, :], np.matmul(w, xj[:, :, None]))[:, 0, 0]
        x16 = d16 - 0.5 * Fab.reshape(-1, 16)
        E += -np.matmul(x16[:, None, :], np.matmul(wx, Fab.reshape(-1, 16)[:, :, None]))[:, 0, 0]
    E += pair_nuclear_energy(method, P.ni, P.nj, P.idxi, P.idxj, rij, w[:, 0, 0], par, rho0=rho0_eff(par, mp))
    return E


def hf_gradient(P, par, method, Dm, mp=None, field=None):
    """grad (nmol, molsize, 3) in eV/Angstrom; force = -grad.  (anal_grad.py:16-92, 95-225)
    field: optional XL-BOMD field density P (then Dm is the density D solved from F(P); xlbomd.py:536-551)."""
    T = Tables.get()
    nmol, molsize = P.nmol, P.molsize
    if mp is None:
        mp = atom_multipoles(P.Z, par)
    Db = Dm.reshape(nmol, molsize, 4, molsize, 4).transpose(0, 1, 3, 2, 4)
    PA = Db[P.atom_molid, P.atom_pos, P.atom_pos]
    mi, ai, aj = P.pair_molid, P.atom_pos[P.idxi], P.atom_pos[P.idxj]
    Dab = Db[mi, ai, aj]
    pk = PA[:, PACK_ROW, PACK_COL] * WEIGHT
    args = (PA[P.idxi], PA[P.idxj], Dab, pk[P.idxi], pk[P.idxj])
    xl = None
    if field is not None:
        Fb = field.reshape(nmol, molsize, 4, molsize, 4).transpose(0, 1, 3, 2, 4)
        FA = Fb[P.atom_molid, P.atom_pos, P.atom_pos]
        fk = FA[:, PACK_ROW, PACK_COL] * WEIGHT
        xl = (fk[P.idxi], fk[P.idxj], Fb[mi, ai, aj])
    Xij = P.xij * (P.rij * T.a0)[:, None]  # R_j - R_i in Angstrom
    g = np.zeros((P.rij.shape[0], 3))
    for c in range(3):
        Es = []
        for s in (+1.0, -1.0):
            # displace atom i by s*DELTA  =>  Xij -= s*DELTA   (anal_grad.py:683-691)
            X = Xij.copy()
            X[:, c] -= s * DELTA
            d = np.sqrt(np.sum(X * X, axis=1))
            Es.append(_pair_energy(P, par, mp, method, X / d[:, None], d / T.a0, *args, xl=xl))
        g[:, c] = (Es[0] - Es[1]) / (2.0 * DELTA)
    nat = P.Z.shape[0]
    ga = np.zeros((nat, 3))
    np.add.at(ga, P.idxi, g)
    np.add.at(ga, P.idxj, -g)
    grad = np.zeros((nmol * molsize, 3))
    grad[P.real_atoms] = ga
    return grad.reshape(nmol, molsize, 3)
