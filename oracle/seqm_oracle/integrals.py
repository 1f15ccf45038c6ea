"""Two-centre two-electron integrals (MNDO multipole model) and Slater overlaps.

Restates:
  seqm/seqm_functions/cal_par.py:11-28 (dd_qq), 112-169 (rho1 secant), 198-257 (rho2 secant)
  seqm/seqm_functions/two_elec_two_center_int.py:98-283 (per-atom prologue + gather),
      1384-1574 (w_withquaternion: local -> molecular frame, e1b/e2a), 1576-1631 (quaternion rotation)
  seqm/seqm_functions/two_elec_two_center_int_local_frame.py:18-293 (the 22 local-frame integrals)
  seqm/seqm_functions/diat_overlap_PM6_SP.py:6-444 (+ SET 451, aintgs 464, bintgs 522)

The local-frame integrals are written here in the Dewar-Thiel point-charge form (every charge
distribution = monopole/dipole/quadrupole point-charge configuration, interaction =
sum c_p c_q ev / sqrt(d^2 + (rho_a+rho_b)^2)); the reference's 22 hand-expanded formulas are the same
sums with identical terms merged.  The rotation is the generic pair-product transform
w = T^T L T, algebraically identical to the reference's 100 expanded elements.
"""
import math

import numpy as np

from .tables import Tables

# packed pair index used everywhere: (0:ss 1:xs 2:xx 3:ys 4:yx 5:yy 6:zs 7:zx 8:zy 9:zz)
PACK = np.array([[0, 1, 3, 6], [1, 2, 4, 7], [3, 4, 5, 8], [6, 7, 8, 9]])
PACK_ROW = np.array([0, 0, 1, 0, 1, 2, 0, 1, 2, 3])  # the smaller orbital index of packed entry
PACK_COL = np.array([0, 1, 1, 2, 2, 2, 3, 3, 3, 3])  # the larger orbital index
WEIGHT = np.array([1.0, 2.0, 1.0, 2.0, 2.0, 1.0, 2.0, 2.0, 2.0, 1.0])


def atom_multipoles(Z, par):
    """dd, qq, rho0, rho1, rho2 per atom (two_elec_two_center_int.py:116-247; cal_par.py)."""
    T = Tables.get()
    ev = T.ev
    nat = Z.shape[0]
    qn = T.qn[Z]
    gss, gpp, gp2, hsp = par["g_ss"], par["g_pp"], par["g_p2"], par["h_sp"]
    zs, zp = par["zeta_s"], par["zeta_p"]
    hpp = np.maximum(0.5 * (gpp - gp2), 0.1)  # clamp_min(0.1), two_elec_two_center_int.py:122
    isX = Z > 2
    dd = np.zeros(nat)
    qq = np.zeros(nat)
    rho1 = np.zeros(nat)
    rho2 = np.zeros(nat)
    with np.errstate(divide="ignore", invalid="ignore"):
        rho0 = 0.5 * ev / gss
    if np.any(isX):
        q, s, p_ = qn[isX], zs[isX], zp[isX]
        dd[isX] = (2.0 * q + 1.0) * (4.0 * s * p_) ** (q + 0.5) / (s + p_) ** (2.0 * q + 2.0) / np.sqrt(3.0)
        qq[isX] = np.sqrt((4.0 * q**2 + 6.0 * q + 2.0) / 20.0) / p_
        # rho1: 5 secant steps on hsp = d/2 - 1/(2 sqrt(4 D^2 + 1/d^2))   (cal_par.py:133-150)
        D1 = dd[isX]
        h = hsp[isX] / ev
        d1 = (np.abs(h) / D1**2) ** (1.0 / 3.0)
        d1 = np.where(h < 0.0, -d1, d1)
        d2 = d1 + 0.04
        f1 = lambda d: 0.5 * d - 0.5 / np.sqrt(4.0 * D1**2 + 1.0 / d**2)  # noqa: E731
        for _ in range(5):
            h1, h2 = f1(d1), f1(d2)
            with np.errstate(divide="ignore", invalid="ignore"):
                d3 = np.where(np.abs(h2 - h1) > 1.0e-16, d1 + (d2 - d1) * (h - h1) / (h2 - h1), d2)
            d1, d2 = d2, d3
        rho1[isX] = 0.5 / d2
        # rho2: same on hpp = q/4 - 1/(2 sqrt(4 D^2+1/q^2)) + 1/(4 sqrt(8 D^2 + 1/q^2))  (cal_par.py:219-242)
        D2 = qq[isX]
        h = hpp[isX] / ev
        q1 = (np.abs(h) / 3.0 / D2**4) ** 0.2
        q1 = np.where(h < 0.0, -q1, q1)
        q2 = q1 + 0.04
        f2 = lambda q_: (  # noqa: E731
            0.25 * q_ - 0.5 / np.sqrt(4.0 * D2**2 + 1.0 / q_**2) + 0.25 / np.sqrt(8.0 * D2**2 + 1.0 / q_**2)
        )
        for _ in range(5):
            h1, h2 = f2(q1), f2(q2)
            with np.errstate(divide="ignore", invalid="ignore"):
                q3 = np.where(np.abs(h2 - h1) > 1.0e-16, q1 + (q2 - q1) * (h - h1) / (h2 - h1), q2)
            q1, q2 = q2, q3
        rho2[isX] = 0.5 / q2
    return dd, qq, rho0, rho1, rho2


def rho0_eff(par, mp):
    """rho_0 per atom with rho_core substituted where it is non-zero (two_elec_two_center_int.py:273-281)."""
    rho0 = mp[2]
    rc = par.get("rho_core")
    if rc is None:
        return rho0
    return np.where(rc != 0.0, rc, rho0)


# --- point-charge multipole configurations -------------------------------------------------------
# each entry: list of (coefficient, x, y, z) in units of the charge separation D (z = bond axis)
def _cfg(kind, D1, D2):
    z0 = 0.0
    if kind == "q":
        return [(1.0, z0, z0, z0)]
    if kind == "mz":
        return [(0.5, z0, z0, D1), (-0.5, z0, z0, -D1)]
    if kind == "mx":
        return [(0.5, D1, z0, z0), (-0.5, -D1, z0, z0)]
    if kind == "Qzz":
        return [(0.25, z0, z0, 2.0 * D2), (-0.5, z0, z0, z0), (0.25, z0, z0, -2.0 * D2)]
    if kind == "Qxx":
        return [(0.25, 2.0 * D2, z0, z0), (-0.5, z0, z0, z0), (0.25, -2.0 * D2, z0, z0)]
    if kind == "Qyy":
        return [(0.25, z0, 2.0 * D2, z0), (-0.5, z0, z0, z0), (0.25, z0, -2.0 * D2, z0)]
    if kind == "Qxz":
        return [(0.25, D2, z0, D2), (-0.25, D2, z0, -D2), (-0.25, -D2, z0, D2), (0.25, -D2, z0, -D2)]
    raise KeyError(kind)


_ORDER = {"q": 0, "mz": 1, "mx": 1, "Qzz": 2, "Qxx": 2, "Qyy": 2, "Qxz": 2}


def _mm(r, ka, kb, A, B, ev):
    """[multipole ka on A | multipole kb on B]; A sits at +r along the local axis seen from B."""
    rho = A["rho"][_ORDER[ka]] + B["rho"][_ORDER[kb]]
    add = rho * rho
    tot = 0.0
    for ca, xa, ya, za in _cfg(ka, A["D1"], A["D2"]):
        for cb, xb, yb, zb in _cfg(kb, B["D1"], B["D2"]):
            dz = za - zb + r
            dx = xa - xb
            dy = ya - yb
            tot = tot + (ca * cb * ev) / np.sqrt(dz * dz + dx * dx + dy * dy + add)
    return tot


def local_frame_integrals(r, A, B, kind):
    """Local-frame integrals for one pair class (two_elec_two_center_int_local_frame.py:77-274).

    kind 'HH' -> (n,1): (ss|ss); 'XH' -> (n,4): ri[0..3]; 'XX' -> (n,22) in the reference's order.
    A/B: dict(D1=dd, D2=qq, rho=(rho0,rho1,rho2)) gathered per pair.
    """
    ev = Tables.get().ev
    qq_ = _mm(r, "q", "q", A, B, ev)
    if kind == "HH":
        return np.stack([qq_], axis=1)
    mzq = _mm(r, "mz", "q", A, B, ev)
    Qzzq = _mm(r, "Qzz", "q", A, B, ev)
    Qxxq = _mm(r, "Qxx", "q", A, B, ev)
    if kind == "XH":
        return np.stack([qq_, mzq, qq_ + Qzzq, qq_ + Qxxq], axis=1)
    qmz = _mm(r, "q", "mz", A, B, ev)
    qQzz = _mm(r, "q", "Qzz", A, B, ev)
    qQxx = _mm(r, "q", "Qxx", A, B, ev)
    QxxQxx = _mm(r, "Qxx", "Qxx", A, B, ev)
    QxxQyy = _mm(r, "Qxx", "Qyy", A, B, ev)
    ri = [
        qq_,  # 0  (ss|ss)
        mzq,  # 1  (so|ss)
        qq_ + Qzzq,  # 2  (oo|ss)
        qq_ + Qxxq,  # 3  (pp|ss)
        qmz,  # 4  (ss|os)
        _mm(r, "mz", "mz", A, B, ev),  # 5  (so|so)
        _mm(r, "mx", "mx", A, B, ev),  # 6  (sp|sp)
        qmz + _mm(r, "Qzz", "mz", A, B, ev),  # 7  (oo|so)
        qmz + _mm(r, "Qxx", "mz", A, B, ev),  # 8  (pp|so)
        _mm(r, "Qxz", "mx", A, B, ev),  # 9  (po|sp)
        qq_ + qQzz,  # 10 (ss|oo)
        qq_ + qQxx,  # 11 (ss|pp)
        mzq + _mm(r, "mz", "Qzz", A, B, ev),  # 12 (so|oo)
        mzq + _mm(r, "mz", "Qxx", A, B, ev),  # 13 (so|pp)
        _mm(r, "mx", "Qxz", A, B, ev),  # 14 (sp|op)
        qq_ + qQzz + Qzzq + _mm(r, "Qzz", "Qzz", A, B, ev),  # 15 (oo|oo)
        qq_ + qQzz + Qxxq + _mm(r, "Qxx", "Qzz", A, B, ev),  # 16 (pp|oo)
        qq_ + qQxx + Qzzq + _mm(r, "Qzz", "Qxx", A, B, ev),  # 17 (oo|pp)
        qq_ + qQxx + Qxxq + QxxQxx,  # 18 (pp|pp)
        _mm(r, "Qxz", "Qxz", A, B, ev),  # 19 (po|po)
        qq_ + qQxx + Qxxq + QxxQyy,  # 20 (pp|p*p*)
        0.5 * (QxxQxx - QxxQyy),  # 21 (p*p|p*p)
    ]
    return np.stack(ri, axis=1)


# local packed index: 0:ss 1:os 2:oo 3:ps 4:po 5:pp 6:p*s 7:p*o 8:p*p 9:p*p*   (o = sigma)
_L_MAP = [
    (0, 0, 0), (1, 0, 1), (2, 0, 2), (5, 0, 3), (9, 0, 3),
    (0, 1, 4), (1, 1, 5), (3, 3, 6), (6, 6, 6), (2, 1, 7), (5, 1, 8), (9, 1, 8), (4, 3, 9), (7, 6, 9),
    (0, 2, 10), (0, 5, 11), (0, 9, 11), (1, 2, 12), (1, 5, 13), (1, 9, 13), (3, 4, 14), (6, 7, 14),
    (2, 2, 15), (5, 2, 16), (9, 2, 16), (2, 5, 17), (2, 9, 17), (5, 5, 18), (9, 9, 18),
    (4, 4, 19), (7, 7, 19), (5, 9, 20), (9, 5, 20), (8, 8, 21),
]  # fmt: skip


def rotation_rows(v):
    """Rows of the rotation taking unit vector v onto the x axis (two_elec_two_center_int.py:1576-1628)."""
    n = v.shape[0]
    q = np.zeros((n, 4))
    q[:, 1] = v[:, 2]
    q[:, 2] = -v[:, 1]
    q[:, 3] = 1.0 + v[:, 0]
    anti = np.abs(q[:, 3]) < 1.0e-7
    q[anti] = np.array([0.0, 0.0, 1.0, 0.0])
    q = q / np.sqrt(np.sum(q * q, axis=1, keepdims=True))
    qy, qz, qw = q[:, 1], q[:, 2], q[:, 3]
    rot = np.empty((n, 3, 3))
    rot[:, 0, 0] = 1 - 2 * (qy * qy + qz * qz)
    rot[:, 0, 1] = -2 * (qz * qw)
    rot[:, 0, 2] = 2 * (qy * qw)
    rot[:, 1, 0] = 2 * (qz * qw)
    rot[:, 1, 1] = 1 - 2 * (qz * qz)
    rot[:, 1, 2] = 2 * (qy * qz)
    rot[:, 2, 0] = -2 * (qy * qw)
    rot[:, 2, 1] = 2 * (qy * qz)
    rot[:, 2, 2] = 1 - 2 * (qy * qy)
    return rot


def _pair_product_transform(rot):
    """T[local packed][molecular packed] for products of (s, p) orbitals."""
    n = rot.shape[0]
    R = np.zeros((n, 4, 4))  # R[a][k]: local orbital a in terms of molecular orbital k
    R[:, 0, 0] = 1.0
    R[:, 1:, 1:] = rot
    T = np.zeros((n, 10, 10))
    for KL in range(10):
        b, a = PACK_ROW[KL], PACK_COL[KL]  # a >= b local orbitals
        for kl in range(10):
            l, k = PACK_ROW[kl], PACK_COL[kl]
            t = R[:, a, k] * R[:, b, l]
            if a != b:
                t = t + R[:, b, k] * R[:, a, l]
            T[:, KL, kl] = t
    return T


def two_center_integrals(P, par, mp=None):
    """w (npairs,10,10), e1b, e2a (npairs,4,4 upper triangles) -- hcore.py:97-122 ->
    two_elec_two_center_int.py:98-283 / 1384-1574.  Also returns rho0 gathered on pairs."""
    T = Tables.get()
    if mp is None:
        mp = atom_multipoles(P.Z, par)
    dd, qq, rho0, rho1, rho2 = mp
    return two_center_integrals_geom(P.ni, P.nj, P.idxi, P.idxj, P.xij, P.rij, mp, T)


def two_center_integrals_geom(ni, nj, idxi, idxj, xij, rij, mp, T=None):
    T = T or Tables.get()
    dd, qq, rho0, rho1, rho2 = mp
    npairs = rij.shape[0]
    HH = (ni == 1) & (nj == 1)
    XH = (ni > 1) & (nj == 1)
    XX = (ni > 1) & (nj > 1)
    w = np.zeros((npairs, 10, 10))

    def side(idx, m):
        a = idx[m]
        return dict(D1=dd[a], D2=qq[a], rho=(rho0[a], rho1[a], rho2[a]))

    rot = rotation_rows(-xij)
    if np.any(HH):
        w[HH, 0, 0] = local_frame_integrals(rij[HH], side(idxi, HH), side(idxj, HH), "HH")[:, 0]
    for m, kind in ((XH, "XH"), (XX, "XX")):
        if not np.any(m):
            continue
        ri = local_frame_integrals(rij[m], side(idxi, m), side(idxj, m), kind)
        L = np.zeros((ri.shape[0], 10, 10))
        for a, b, k in _L_MAP:
            if k < ri.shape[1] and (kind == "XX" or b == 0):
                L[:, a, b] = ri[:, k]
        Tm = _pair_product_transform(rot[m])
        wm = np.matmul(Tm.transpose(0, 2, 1), np.matmul(L, Tm))
        if kind == "XH":
            wm[:, :, 1:] = 0.0
        w[m] = wm
    e1b = np.zeros((npairs, 4, 4))
    e2a = np.zeros((npairs, 4, 4))
    e1b[:, PACK_ROW, PACK_COL] = -T.tore[nj][:, None] * w[:, :, 0]
    e2a[:, PACK_ROW, PACK_COL] = -T.tore[ni][:, None] * w[:, 0, :]
    return w, e1b, e2a, rho0[idxi], rho0[idxj]


# --- Slater overlaps ------------------------------------------------------------------------------
def _aintgs(x, kmax):
    """A_k(x) = int_1^inf t^k exp(-x t) dt, upward recurrence (diat_overlap_PM6_SP.py:464-519)."""
    a = [np.exp(-x) / x]
    for k in range(1, kmax + 1):
        a.append(a[0] + k * a[k - 1] / x)
    return a


def _bintgs(x, kmax, x_regime=None):
    """B_k(x) = int_-1^1 t^k exp(-x t) dt with the reference's three regimes
    (|x|>0.5 recurrence, 1e-6<|x|<=0.5 four-term series, else x=0 limit) -- diat_overlap_PM6_SP.py:522-670.
    x_regime: the argument that selects the regime (finite-difference gradients freeze the choice at the undisplaced
    geometry: the truncated series and the recurrence differ by ~1e-7 at |x| = 0.5, which a stencil must not straddle)."""
    absx = np.abs(x if x_regime is None else x_regime)
    big = absx > 0.5
    mid = (absx <= 0.5) & (absx > 1.0e-6)
    b = []
    xs = np.where(big, x, 1.0)
    tx = np.exp(xs) / xs
    tmx = -np.exp(-xs) / xs
    xm = np.where(mid, x, 0.0)
    for k in range(kmax + 1):
        lim = 2.0 / (k + 1.0) if k % 2 == 0 else 0.0
        if k == 0:
            rec = tx + tmx
        else:
            rec = (tx if k % 2 == 0 else -tx) + tmx + k * b_rec_prev / xs
        b_rec_prev = rec
        if k % 2 == 0:
            ser = (
                2.0 / (k + 1.0)
                + xm**2 / ((k + 3.0) * 1.0)
                + xm**4 / ((k + 5.0) * 12.0)
                + xm**6 / ((k + 7.0) * 360.0)
            )
        else:
            ser = -2.0 / (k + 2.0) * xm - xm**3 / ((k + 4.0) * 3.0) - xm**5 / ((k + 6.0) * 60.0)
        b.append(np.where(big, rec, np.where(mid, ser, lim)))
    return b


def _poly_mul(p, q):
    out = np.zeros((p.shape[0] + q.shape[0] - 1, p.shape[1] + q.shape[1] - 1))
    for i in range(p.shape[0]):
        for j in range(p.shape[1]):
            if p[i, j] != 0:
                out[i : i + q.shape[0], j : j + q.shape[1]] += p[i, j] * q
    return out


def _poly_pow(p, n):
    out = np.ones((1, 1))
    for _ in range(n):
        out = _poly_mul(out, p)
    return out


_XI_P_ETA = np.array([[0.0, 1.0], [1.0, 0.0]])  # xi + eta   (index [power of xi][power of eta])
_XI_M_ETA = np.array([[0.0, -1.0], [1.0, 0.0]])  # xi - eta
_ONE_P = np.array([[1.0, 0.0], [0.0, 1.0]])  # 1 + xi eta
_M_ONE_P = np.array([[-1.0, 0.0], [0.0, 1.0]])  # xi eta - 1
_XI2M1 = np.array([[-1.0], [0.0], [1.0]])  # xi^2 - 1
_1META2 = np.array([[1.0, 0.0, -1.0]])  # 1 - eta^2


def overlap_poly(na, nb, kind):
    """Integer polynomial in (xi, eta) and angular constant of the prolate-spheroidal overlap integrand."""
    if kind == "ss":
        return 0.5, _poly_mul(_poly_pow(_XI_P_ETA, na), _poly_pow(_XI_M_ETA, nb))
    if kind == "os":  # p-sigma on A, s on B
        return math.sqrt(3.0) / 2.0, _poly_mul(
            _poly_mul(_poly_pow(_XI_P_ETA, na - 1), _ONE_P), _poly_pow(_XI_M_ETA, nb)
        )
    if kind == "so":
        return math.sqrt(3.0) / 2.0, _poly_mul(
            _poly_mul(_poly_pow(_XI_P_ETA, na), _M_ONE_P), _poly_pow(_XI_M_ETA, nb - 1)
        )
    if kind == "oo":
        return 1.5, _poly_mul(
            _poly_mul(_poly_pow(_XI_P_ETA, na - 1), _poly_pow(_XI_M_ETA, nb - 1)), _poly_mul(_ONE_P, _M_ONE_P)
        )
    if kind == "pp":
        return 0.75, _poly_mul(
            _poly_mul(_poly_pow(_XI_P_ETA, na - 1), _poly_pow(_XI_M_ETA, nb - 1)), _poly_mul(_XI2M1, _1META2)
        )
    raise KeyError(kind)


def _sto_overlap(na, nb, za, zb, r, kind):
    c, poly = overlap_poly(na, nb, kind)
    alpha = 0.5 * r * (za + zb)
    beta = 0.5 * r * (za - zb)
    kmax = na + nb
    A = _aintgs(alpha, kmax)
    B = _bintgs(beta, kmax)
    tot = 0.0
    for k in range(poly.shape[0]):
        for l in range(poly.shape[1]):
            if poly[k, l] != 0.0:
                tot = tot + poly[k, l] * A[k] * B[l]
    norm = (
        (2.0 * za) ** (na + 0.5)
        * (2.0 * zb) ** (nb + 0.5)
        / math.sqrt(math.factorial(2 * na) * math.factorial(2 * nb))
        * (0.5 * r) ** (na + nb + 1)
    )
    return c * norm * tot


def overlap_sp(ni, nj, xij, rij, zeta_a, zeta_b):
    """di (npairs,4,4) = <mu on i | nu on j> in the molecular frame (diat_overlap_PM6_SP.py:6-444);
    zero beyond the 40 bohr cutoff (hcore.py:81-92).  zeta_* (npairs,2) = (zeta_s, zeta_p)."""
    T = Tables.get()
    npairs = rij.shape[0]
    di = np.zeros((npairs, 4, 4))
    qa_all, qb_all = T.qn_int[ni], T.qn_int[nj]
    within = rij <= T.overlap_cutoff
    for na in np.unique(qa_all):
        for nb in np.unique(qb_all):
            m = (qa_all == na) & (qb_all == nb) & within
            if not np.any(m):
                continue
            if not (1 <= nb <= na <= 3):
                raise ValueError("\nError from diat.py, overlap matrix\nSome elements are not supported yet")
            r = rij[m]
            e = xij[m]
            zsa, zpa, zsb, zpb = zeta_a[m, 0], zeta_a[m, 1], zeta_b[m, 0], zeta_b[m, 1]
            blk = np.zeros((r.shape[0], 4, 4))
            blk[:, 0, 0] = _sto_overlap(na, nb, zsa, zsb, r, "ss")
            if na > 1:
                blk[:, 1:, 0] = _sto_overlap(na, nb, zpa, zsb, r, "os")[:, None] * e
            if nb > 1:
                blk[:, 0, 1:] = _sto_overlap(na, nb, zsa, zpb, r, "so")[:, None] * e
                soo = _sto_overlap(na, nb, zpa, zpb, r, "oo")
                spp = _sto_overlap(na, nb, zpa, zpb, r, "pp")
                ee = e[:, :, None] * e[:, None, :]
                blk[:, 1:, 1:] = (soo - spp)[:, None, None] * ee + spp[:, None, None] * np.eye(3)[None]
            di[m] = blk
    return di
