"""Core Hamiltonian and Fock matrix in the reference's padded dense layout
(nmol, 4*molsize, 4*molsize); orbital index = 4*atom_position + (s, px, py, pz).

Restates: seqm/seqm_functions/hcore.py:9-179, seqm/seqm_functions/fock.py:132-347,
          seqm/seqm_functions/scf_loop.py:2066-2085 (initial guess).
Unlike the reference, Hcore is returned already symmetric (the reference keeps the upper triangle and
symmetrises inside elec_energy, energy.py:37-38); `hcore_upper` reproduces the reference's M.
"""
import numpy as np

from .integrals import PACK, PACK_COL, PACK_ROW, WEIGHT, overlap_sp, two_center_integrals
from .tables import Tables


def _blocks_view(X, nmol, molsize):
    """(nmol, 4m, 4m) -> (nmol, m, m, 4, 4) view of atom blocks."""
    return X.reshape(nmol, molsize, 4, molsize, 4).transpose(0, 1, 3, 2, 4)


class Segments:
    """Sorted-segment sums replacing the reference's index_add_ scatters."""

    def __init__(self, idx, n):
        self.order = np.argsort(idx, kind="stable")
        s = idx[self.order]
        self.starts = np.nonzero(np.concatenate([[True], s[1:] != s[:-1]]))[0] if s.size else np.zeros(0, int)
        self.targets = s[self.starts] if s.size else np.zeros(0, int)
        self.n = n

    def add(self, out, vals):
        if vals.shape[0] == 0:
            return
        v = vals[self.order]
        out[self.targets] += np.add.reduceat(v, self.starts, axis=0)


def build_hcore(P, par, mp=None):
    """Returns dict(H (symmetric, dense), w, e1b, e2a, di, rho0i, rho0j).  hcore.py:9-179."""
    nmol, molsize = P.nmol, P.molsize
    nat = P.Z.shape[0]
    w, e1b, e2a, rho0i, rho0j = two_center_integrals(P, par, mp)
    zeta = np.stack([par["zeta_s"], par["zeta_p"]], axis=1)
    di = overlap_sp(P.ni, P.nj, P.xij, P.rij, zeta[P.idxi], zeta[P.idxj])
    # diagonal blocks: U_ss/U_pp + sum_B core attraction (upper triangle)  hcore.py:131-150
    D = np.zeros((nat, 4, 4))
    D[:, 0, 0] = par["U_ss"]
    for k in (1, 2, 3):
        D[:, k, k] = par["U_pp"]
    P.seg_i = getattr(P, "seg_i", None) or Segments(P.idxi, nat)
    P.seg_j = getattr(P, "seg_j", None) or Segments(P.idxj, nat)
    P.seg_i.add(D, e1b)
    P.seg_j.add(D, e2a)
    D = D + np.triu(D, 1).transpose(0, 2, 1)
    # off-diagonal blocks: di * (beta_mu^A + beta_nu^B)/2     hcore.py:155-173
    bA = np.stack([par["beta_s"]] + [par["beta_p"]] * 3, axis=1)
    bsum = 0.5 * (bA[P.idxi][:, :, None] + bA[P.idxj][:, None, :])
    Hab = di * bsum
    H = np.zeros((nmol, 4 * molsize, 4 * molsize))
    Hb = _blocks_view(H, nmol, molsize)
    Hb[P.atom_molid, P.atom_pos, P.atom_pos] = D
    mi, ai, aj = P.pair_molid, P.atom_pos[P.idxi], P.atom_pos[P.idxj]
    Hb[mi, ai, aj] = Hab
    Hb[mi, aj, ai] = Hab.transpose(0, 2, 1)
    return dict(H=H, w=w, e1b=e1b, e2a=e2a, di=di, rho0i=rho0i, rho0j=rho0j)


def hcore_upper(H, P):
    """The reference's block tensor M (nmol*molsize^2,4,4): upper blocks / upper triangles only."""
    nmol, molsize = P.nmol, P.molsize
    U = np.triu(H)
    return _blocks_view(U, nmol, molsize).reshape(nmol * molsize * molsize, 4, 4).copy()


def initial_density(P):
    """Diagonal guess: tore/4 on heavy s,p ; 1 on H s   (scf_loop.py:2066-2081)."""
    T = Tables.get()
    nmol, molsize = P.nmol, P.molsize
    D = np.zeros((nmol, 4 * molsize, 4 * molsize))
    Db = _blocks_view(D, nmol, molsize)
    heavy = P.Z > 1
    val = T.tore[P.Z] / 4.0
    for k in range(4):
        Db[P.atom_molid[heavy], P.atom_pos[heavy], P.atom_pos[heavy], k, k] = val[heavy]
    hyd = P.Z == 1
    Db[P.atom_molid[hyd], P.atom_pos[hyd], P.atom_pos[hyd], 0, 0] = 1.0
    return D


def exchange_view(P, w):
    """(npairs,16,16) gather of w with rows (mu,lam) and columns (nu,sig); cached per integral tensor."""
    c = getattr(P, "_wx_cache", None)
    if c is None or c[0] is not w:
        w4 = w[:, PACK[:, :, None, None], PACK[None, None, :, :]]  # (p, mu, nu, lam, sig)
        c = (w, np.ascontiguousarray(w4.transpose(0, 1, 3, 2, 4)).reshape(-1, 16, 16))
        P._wx_cache = c
    return c[1]


def build_fock(P, par, H, w, Dm, mols=None):
    """F = H + G(D) for the dense symmetric density Dm (fock.py:132-347).

    `mols`: optional boolean mask of molecules to (re)build; others are returned as zeros.
    """
    nmol, molsize = P.nmol, P.molsize
    nat = P.Z.shape[0]
    Db = _blocks_view(Dm, nmol, molsize)
    PA = Db[P.atom_molid, P.atom_pos, P.atom_pos]  # (nat,4,4) diagonal blocks
    gss, gpp, gsp, gp2, hsp = par["g_ss"], par["g_pp"], par["g_sp"], par["g_p2"], par["h_sp"]
    # one-centre terms (fock.py:187-231)
    Pss = PA[:, 0, 0]
    Ppt = PA[:, 1, 1] + PA[:, 2, 2] + PA[:, 3, 3]
    G = np.zeros((nat, 4, 4))
    G[:, 0, 0] = 0.5 * Pss * gss + Ppt * (gsp - 0.5 * hsp)
    for k in (1, 2, 3):
        Pk = PA[:, k, k]
        G[:, k, k] = Pss * (gsp - 0.5 * hsp) + 0.5 * Pk * gpp + (Ppt - Pk) * (1.25 * gp2 - 0.25 * gpp)
        G[:, 0, k] = PA[:, 0, k] * (1.5 * hsp - 0.5 * gsp)
    for a, b in ((1, 2), (1, 3), (2, 3)):
        G[:, a, b] = PA[:, a, b] * (0.75 * gpp - 1.25 * gp2)
    # two-centre Coulomb (fock.py:278-294)
    pk = PA[:, PACK_ROW, PACK_COL] * WEIGHT  # (nat,10)
    JA = np.matmul(w, pk[P.idxj][:, :, None])[:, :, 0]  # onto atom i
    JB = np.matmul(pk[P.idxi][:, None, :], w)[:, 0, :]  # onto atom j
    J = np.zeros((nat, 10))
    P.seg_i.add(J, JA)
    P.seg_j.add(J, JB)
    G[:, PACK_ROW, PACK_COL] += J
    G = G + np.triu(G, 1).transpose(0, 2, 1)
    # two-centre exchange (fock.py:297-345): K[mu,lam] = -1/2 sum_{nu,sig} D_AB[nu,sig] w[pack(mu,nu), pack(lam,sig)]
    mi, ai, aj = P.pair_molid, P.atom_pos[P.idxi], P.atom_pos[P.idxj]
    Dab = Db[mi, ai, aj]
    K = -0.5 * np.matmul(exchange_view(P, w), Dab.reshape(-1, 16, 1)).reshape(-1, 4, 4)
    F = H.copy()
    Fb = _blocks_view(F, nmol, molsize)
    Fb[P.atom_molid, P.atom_pos, P.atom_pos] += G
    Fb[mi, ai, aj] += K
    Fb[mi, aj, ai] += K.transpose(0, 2, 1)
    return F
