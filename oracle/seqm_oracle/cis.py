"""CIS sigma-vector building block (oracle): contraction of AO transition densities with the two-electron integrals.
TEST INFRASTRUCTURE: nothing in pyseqm_b200/ imports this.

Restates seqm/seqm_functions/rcis_batch.py:296-403 (makeA_pi_batched) and 406-506 (makeA_pi_symm_batch): the symmetric part of
the density goes through the two-electron part of the Fock build (Coulomb + exchange + one-centre terms, identical to
fock.py), the antisymmetric part only feels exchange: off-diagonal blocks -1/2 sum P_anti (mu nu | la sg), antisymmetric, and
the one-centre terms (s,p): (hsp - gsp)/2, (p,p'): gpp/4 - 3 gp2/4.
Pinned by tests/test_md.py to tests/golden/cis_sigma_methanal.npz (tools/make_golden_ksa.py).
"""
import numpy as np

from .hamiltonian import build_fock

_IND = np.array([[0, 1, 3, 6], [1, 2, 4, 7], [3, 4, 5, 8], [6, 7, 8, 9]])


def sigma_ao(P, par, w, X, all_symmetric=False):
    """X, result: dense padded (nmol, 4 molsize, 4 molsize)."""
    Xs = 0.5 * (X + X.transpose(0, 2, 1))
    F = build_fock(P, par, np.zeros_like(X), w, Xs)
    if all_symmetric:
        return F
    Xa = 0.5 * (X - X.transpose(0, 2, 1))
    for p in range(P.idxi.shape[0]):
        i, j = int(P.idxi[p]), int(P.idxj[p])  # global atom indices
        m = int(P.pair_molid[p])
        li, lj = int(P.atom_pos[i]), int(P.atom_pos[j])
        blk = Xa[m, 4 * li : 4 * li + 4, 4 * lj : 4 * lj + 4]
        K = np.zeros((4, 4))
        for a in range(4):
            for b in range(4):
                K[a, b] = -0.5 * np.sum(blk * w[p][np.ix_(_IND[a], _IND[b])])
        F[m, 4 * li : 4 * li + 4, 4 * lj : 4 * lj + 4] += K
        F[m, 4 * lj : 4 * lj + 4, 4 * li : 4 * li + 4] -= K.T
    for a in range(P.Z.shape[0]):
        m = int(P.atom_molid[a])
        la = int(P.atom_pos[a])
        blk = Xa[m, 4 * la : 4 * la + 4, 4 * la : 4 * la + 4]
        one = np.zeros((4, 4))
        for i in range(1, 4):
            one[0, i] = blk[0, i] * (0.5 * par["h_sp"][a] - 0.5 * par["g_sp"][a])
        for i, j in ((1, 2), (1, 3), (2, 3)):
            one[i, j] = blk[i, j] * (0.25 * par["g_pp"][a] - 0.75 * par["g_p2"][a])
        F[m, 4 * la : 4 * la + 4, 4 * la : 4 * la + 4] += one - one.T
    return F
