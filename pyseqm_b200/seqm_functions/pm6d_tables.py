"""Host-side tables of the PM6 d-orbital path: everything that depends on the element only.

The reference rebuilds these per atom on every forward -- Slater-Condon radial integrals, the golden-section search for
the additive terms, the 243-entry one-centre list (two_elec_two_center_int.py:16-97, 163-247; cal_par.py:283-393;
build_two_elec_one_center_int_D.py:15-202).  Here they are evaluated once per element and per process and live on
the device as rows of the per-element parameter table and as four small constant tables the kernels index:

  d_rows(...)                 rows SEQM_P_UDD .. SEQM_P_RHO2D of the element table
  multipole_coefficients()    c[45][7][5]: expansion of the 45 local orbital products in point-charge multipoles
  overlap_polynomials()       [4][4][14][9][9]: prolate-spheroidal polynomials of the 14 local s/p/d overlaps
  one_center_integrals(...)   [Z][45*45]: one-centre (kl|mn) that contain a d orbital

Plain numpy / math on the host: parameter preparation, not part of the per-step path.
"""
import math

import numpy as np

EV = 27.21  # constants.py:4
S3, S5, S15 = math.sqrt(3.0), math.sqrt(5.0), math.sqrt(15.0)
D_ROWS = ("U_dd", "zeta_d", "beta_d", "qnd", "dp", "ds", "ddq", "rho3", "rho4", "rho5", "rho6", "rho2d")
TRI = [(a, b) for a in range(9) for b in range(a + 1)]
PAIR = np.zeros((9, 9), dtype=np.int64)
for _k, (_a, _b) in enumerate(TRI):
    PAIR[_a, _b] = PAIR[_b, _a] = _k
TRI_A = np.array([t[0] for t in TRI])
TRI_B = np.array([t[1] for t in TRI])


def d_shell(z):
    """nSuperHeavy set of the reference (basics.py:258-269)."""
    return (12 < z < 18) or (20 < z < 30) or (32 < z < 36) or (38 < z < 48) or (50 < z < 54) or (70 < z < 80) or z == 57


def _transition(z):
    return (20 < z < 30) or (38 < z < 48) or (70 < z < 80) or z == 57


def _binom(a, b):
    return math.factorial(a) / (math.factorial(b) * math.factorial(a - b))


def radial_integral(k, na, ea, nb, eb, nc, ec, nd, ed):
    """Slater-Condon parameter R^k(ab, cd) over Slater functions, eV (MOPAC rsc; two_elec_two_center_int.py:1309-1364)."""
    nab, ncd, eab, ecd = na + nb, nc + nd, ea + eb, ec + ed
    e, n = eab + ecd, nab + ncd
    ae = math.log(e)
    c = math.exp(math.log(math.factorial(n - 1)) + na * math.log(ea) + nb * math.log(eb) + nc * math.log(ec)
                 + nd * math.log(ed) + 0.5 * (math.log(ea) + math.log(eb) + math.log(ec) + math.log(ed))
                 + math.log(2) * (n + 2)
                 - 0.5 * (math.log(math.factorial(2 * na)) + math.log(math.factorial(2 * nb))
                          + math.log(math.factorial(2 * nc)) + math.log(math.factorial(2 * nd))) - ae * n) * EV  # fmt: skip
    s0, s1, s2 = 1 / e, 0, 0
    m = ncd - k
    for i in range(1, m + 1):
        s0 = s0 * e / ecd
        s1 = s1 + s0 * (_binom(ncd - k - 1, i - 1) - _binom(ncd + k, i - 1)) / _binom(n - 1, i - 1)
    m2 = ncd + k + 1
    for i in range(m + 1, m2 + 1):
        s0 = s0 * e / ecd
        s2 = s2 + s0 * _binom(m2 - 1, i - 1) / _binom(n - 1, i - 1)
    s3 = math.exp(ae * n - math.log(ecd) * m2 - math.log(eab) * (nab - k)) / _binom(n - 1, m2 - 1)
    return c * (s1 - s2 + s3)


def radial_moment(z1, z2, n1, n2, L):
    """<r^L> between two Slater functions (cal_par.py:377-393)."""
    if z1 == 0 or z2 == 0:
        return 0.0
    a = math.factorial(n1 + n2 + L) / math.sqrt(math.factorial(2 * n1) * math.factorial(2 * n2))
    return (a * (2 * z1 / (z1 + z2)) ** n1 * math.sqrt(2 * z1 / (z1 + z2)) * (2 * z2 / (z1 + z2)) ** n2
            * math.sqrt(2 * z2 / (z1 + z2)) / (z1 + z2) ** L)  # fmt: skip


def additive_term(L, D, FG):
    """rho with [unit multipole L, separation D] self-interaction = FG, by the reference's golden-section search on
    [0.1, 5] (cal_par.py:283-359): its bracket end, not the exact root, is what the reference's integrals contain."""
    if L == 0:
        return 0.5 * EV / FG
    if FG == 0.0:
        return 0.0
    dsq = D * D
    a1, a2 = 0.1, 5.0
    f1 = f2 = 0.0
    for _ in range(100):
        delta = a2 - a1
        if delta < 1.0e-8:
            break
        y1, y2 = a1 + delta * 0.382, a1 + delta * 0.618
        if L == 1:
            f1 = (EV * 0.25 * (1.0 / y1 - 1.0 / math.sqrt(y1**2 + dsq)) - FG) ** 2
            f2 = (EV * 0.25 * (1.0 / y2 - 1.0 / math.sqrt(y2**2 + dsq)) - FG) ** 2
        else:
            f1 = (EV / 8.0 * (1.0 / y1 - 2.0 / math.sqrt(y1**2 + dsq * 0.5) + 1.0 / math.sqrt(y1**2 + dsq)) - FG) ** 2
            f2 = (EV / 8.0 * (1.0 / y2 - 2.0 / math.sqrt(y2**2 + dsq * 0.5) + 1.0 / math.sqrt(y2**2 + dsq)) - FG) ** 2
        if f1 < f2:
            a2 = y2
        else:
            a1 = y1
    return a2 if f1 >= f2 else a1


def d_rows(z, qn, qnd, row):
    """The 12 d-shell rows of element z; `row` maps parameter-file column names to values."""
    out = dict.fromkeys(D_ROWS, 0.0)
    if z > 2 and row.get("zeta_p", 0.0) > 0.0:
        qq = math.sqrt((4.0 * qn * qn + 6.0 * qn + 2.0) / 20.0) / row["zeta_p"]
        out["rho2d"] = additive_term(2, qq * math.sqrt(2.0), 0.5 * (row["g_pp"] - row["g_p2"]))
    if not d_shell(z) or row.get("zeta_d", 0.0) == 0.0:
        return out
    zs, zp, zd = row["s_orb_exp_tail"], row["p_orb_exp_tail"], row["d_orb_exp_tail"]
    qd = qn - 1 if _transition(z) else qn
    R = radial_integral
    dp_add = (4.0 / 15.0) * R(1, qn, zp, qd, zd, qn, zp, qd, zd)
    if _transition(z) and row["G2SD"] > 1.0e-9:
        ds_add = 0.2 * row["G2SD"]
    else:
        ds_add = 0.2 * R(2, qn, zs, qd, zd, qn, zs, qd, zd)
    dd_add = (4.0 / 49.0) * R(2, qd, zd, qd, zd, qd, zd, qd, zd)
    dd0 = R(0, qd, zd, qd, zd, qd, zd, qd, zd)
    dd4 = R(4, qd, zd, qd, zd, qd, zd, qd, zd)
    dp3 = (27.0 / 245.0) * R(3, qn, zp, qd, zd, qn, zp, qd, zd)
    r_pd = radial_moment(row["zeta_p"], row["zeta_d"], qn, qd, 1)
    r2_sd = radial_moment(row["zeta_s"], row["zeta_d"], qn, qd, 2)
    r2_dd = radial_moment(row["zeta_d"], row["zeta_d"], qd, qd, 2)
    out.update(U_dd=row["U_dd"], zeta_d=row["zeta_d"], beta_d=row["beta_d"], qnd=float(qnd))
    out["dp"] = r_pd / math.sqrt(5)
    out["ds"] = math.sqrt(r2_sd * math.sqrt(1.0 / 15.0)) * math.sqrt(2.0)
    out["rho5"] = additive_term(2, out["ds"], ds_add)
    fg = dd0 + dd_add + 4 / 49 * dd4
    fg1 = dd0 + 0.5 * dd_add - 24 / 441 * dd4
    fg2 = dd0 - dd_add + 6 / 441 * dd4
    out["rho3"] = additive_term(0, 1.0, 0.2 * (fg + 2.0 * fg1 + 2.0 * fg2))
    out["rho4"] = additive_term(1, r_pd / math.sqrt(5.0), (dp_add + dp3) - 1.8 * (3 / 49 * 245 / 27 * dp3))
    out["ddq"] = math.sqrt(2.0 * (r2_dd / 7.0))
    out["rho6"] = additive_term(2, out["ddq"], (3 / 4 * dd_add + 20 / 441 * dd4) - (20.0 / 35.0) * (35 / 441 * dd4))
    return out


# ---- angular algebra -------------------------------------------------------------------------------------------------
def _sphere():
    t, wt = np.polynomial.legendre.leggauss(12)
    nph = 24
    ph = (np.arange(nph) + 0.5) * (2 * math.pi / nph)
    ct, pp = np.meshgrid(t, ph, indexing="ij")
    st = np.sqrt(1 - ct * ct)
    return (st * np.cos(pp)).ravel(), (st * np.sin(pp)).ravel(), ct.ravel(), (np.repeat(wt, nph) / nph) / 2.0


def _local_functions(x, y, z):  # s, p(z, x, y), d(z2, xz, yz, x2-y2, xy); <f f> = 1 under dOmega / 4 pi
    return np.stack([np.ones_like(x), S3 * z, S3 * x, S3 * y, 0.5 * S5 * (3 * z * z - 1.0), S15 * x * z, S15 * y * z,
                     0.5 * S15 * (x * x - y * y), S15 * x * y])  # fmt: skip


def _molecular_functions(x, y, z):  # s, px, py, pz, d(x2-y2, xz, z2, yz, xy)
    return np.stack([np.ones_like(x), S3 * x, S3 * y, S3 * z, 0.5 * S15 * (x * x - y * y), S15 * x * z,
                     0.5 * S5 * (3 * z * z - 1.0), S15 * y * z, S15 * x * y])  # fmt: skip


def multipole_coefficients():
    """(c, c_yx), each (45, 7, 5): c[kl][source][m] with sources (ss/pp monopole, sp dipole, pp quadrupole, sd quadrupole,
    pd dipole, dd monopole, dd quadrupole) and m slots (0, 1c, 1s, 2c, 2s).  c = <f_k f_l C_lm> / (g_source kappa_lm):
    g converts <r^l> of the product type into the charge separation (1/sqrt3, 1/5, 1/sqrt15, 1/sqrt5, 1/7) and kappa is the
    multipole moment of the unit point-charge configuration per D^l (1; 3/2 for (2,0); sqrt3 for (2,1), (2,2)).
    Reference conventions reproduced: cosine-type coefficients carry 6 decimals, sine-type ones full precision; in a
    (d, sp-heavy) pair the d-sigma d-delta quadrupole terms have the opposite sign (c_yx)."""
    x, y, z, wq = _sphere()
    f = _local_functions(x, y, z)
    C = np.stack([np.ones_like(x), z, x, y, 0.5 * (3 * z * z - 1.0), S3 * x * z, S3 * y * z, 0.5 * S3 * (x * x - y * y), S3 * x * y])
    l_of = [0, 1, 1, 1, 2, 2, 2, 2, 2]
    src = {(0, 0): {0: 0}, (0, 1): {1: 1}, (1, 1): {0: 0, 2: 2}, (0, 2): {2: 3}, (1, 2): {1: 4}, (2, 2): {0: 5, 2: 6}}
    g = {1: 1.0 / S3, 2: 0.2, 3: 1.0 / S15, 4: 1.0 / S5, 6: 1.0 / 7.0}
    kappa = [1.0, 1.0, 1.0, 1.0, 1.5, S3, S3, S3, S3]
    mslot = [0, 0, 1, 2, 0, 1, 2, 3, 4]
    c = np.zeros((45, 7, 5))
    for kl, (a, b) in enumerate(TRI):
        ty = tuple(sorted((l_of[a], l_of[b])))
        for lm in range(9):
            l = 0 if lm == 0 else (1 if lm < 4 else 2)
            if l not in src[ty]:
                continue
            s = src[ty][l]
            A = float(np.sum(wq * f[a] * f[b] * C[lm]))
            if abs(A) < 1e-12:
                continue
            val = A if l == 0 else A / (g[s] * kappa[lm])
            c[kl, s, mslot[lm]] = val if mslot[lm] in (2, 4) else round(val, 6)
    cyx = c.copy()
    for kl in (PAIR[7, 4], PAIR[8, 4]):
        cyx[kl, 6, 3:5] *= -1.0
    return c, cyx


def _pmul(p, q):
    out = np.zeros((p.shape[0] + q.shape[0] - 1, p.shape[1] + q.shape[1] - 1))
    for i in range(p.shape[0]):
        for j in range(p.shape[1]):
            if p[i, j] != 0:
                out[i : i + q.shape[0], j : j + q.shape[1]] += p[i, j] * q
    return out


def _ppow(p, n):
    out = np.ones((1, 1))
    for _ in range(n):
        out = _pmul(out, p)
    return out


OVERLAP_KINDS = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (1, 1, 1), (2, 0, 0), (0, 2, 0), (2, 1, 0), (1, 2, 0), (2, 1, 1),
                 (1, 2, 1), (2, 2, 0), (2, 2, 1), (2, 2, 2)]  # (l_a, l_b, m); spd_kind() in spd_kernels.cuh  # fmt: skip


def overlap_polynomials():
    """(4, 4, 14, 9, 9): poly[na-1][nb-1][kind][k][l] multiplies A_k(alpha) B_l(beta) in the overlap of the Slater function
    (na, la, m) at the origin with (nb, lb, m) at +R on the local z axis; the angular constant is folded in.
    Prolate spheroidal coordinates in units of R/2: r_a = xi + eta, z_a = 1 + xi eta, r_b = xi - eta, z_b = xi eta - 1,
    rho^2 = (xi^2 - 1)(1 - eta^2), volume element (xi^2 - eta^2) = r_a r_b."""
    xpe = np.array([[0.0, 1.0], [1.0, 0.0]])
    xme = np.array([[0.0, -1.0], [1.0, 0.0]])
    za = np.array([[1.0, 0.0], [0.0, 1.0]])
    zb = np.array([[-1.0, 0.0], [0.0, 1.0]])
    rho2 = _pmul(np.array([[-1.0], [0.0], [1.0]]), np.array([[1.0, 0.0, -1.0]]))

    def side(n, l, m, zp, rp):
        if n < l + 1:
            return None
        base = _ppow(rp, n - l)
        if l == 0:
            return 1.0, base
        if l == 1:
            return S3, (_pmul(base, zp) if m == 0 else base)
        if m == 0:
            q, r2 = 3.0 * _pmul(zp, zp), _pmul(rp, rp)
            tot = np.zeros((max(q.shape[0], r2.shape[0]), max(q.shape[1], r2.shape[1])))
            tot[: q.shape[0], : q.shape[1]] += q
            tot[: r2.shape[0], : r2.shape[1]] -= r2
            return 0.5 * S5, _pmul(base, tot)
        return (S15, _pmul(base, zp)) if m == 1 else (0.5 * S15, base)

    out = np.zeros((4, 4, 14, 9, 9))
    for na in range(1, 5):
        for nb in range(1, 5):
            for kind, (la, lb, m) in enumerate(OVERLAP_KINDS):
                A, B = side(na, la, m, za, xpe), side(nb, lb, m, zb, xme)
                if A is None or B is None:
                    continue
                poly = _pmul(A[1], B[1])
                for _ in range(m):
                    poly = _pmul(poly, rho2)
                out[na - 1, nb - 1, kind, : poly.shape[0], : poly.shape[1]] = (0.5 if m == 0 else 0.25) * A[0] * B[0] * poly
    return out


_ANGULAR = None


def _angular_factors():
    """Ang[k][kl, mn] = <f_k f_l P_k(cos gamma_12) f_m f_n>: the angular part of (kl|mn) = sum_k Ang[k] R^k."""
    global _ANGULAR
    if _ANGULAR is None:
        x, y, z, wq = _sphere()
        f = _molecular_functions(x, y, z)
        prod = f[TRI_A] * f[TRI_B] * wq[None, :]
        cosg = np.clip(x[:, None] * x[None, :] + y[:, None] * y[None, :] + z[:, None] * z[None, :], -1.0, 1.0)
        _ANGULAR = []
        for k in range(5):
            A = prod @ np.polynomial.legendre.legval(cosg, [0] * k + [1]) @ prod.T
            A[np.abs(A) < 1e-13] = 0.0
            _ANGULAR.append(A)
    return _ANGULAR


def one_center_integrals(qn, qnd, zs, zp, zd, f0sd, g2sd):
    """(45, 45): one-centre (kl|mn) of one d-shell element, zero for quadruples without a d orbital (those are the g_ss ...
    h_sp parameters).  Slater-Condon expansion with the internal exponents; F0(ss,dd) and G2(sd,sd) are replaced by the
    F0SD / G2SD parameters when set (build_two_elec_one_center_int_D.py:56-60)."""
    n, ex = {0: qn, 1: qn, 2: qnd}, {0: zs, 1: zp, 2: zd}
    Ang = _angular_factors()
    lo = [0, 1, 1, 1, 2, 2, 2, 2, 2]
    out = np.zeros((45, 45))
    cache = {}
    for kl in range(45):
        a, b = lo[TRI_A[kl]], lo[TRI_B[kl]]
        for mn in range(45):
            c, d = lo[TRI_A[mn]], lo[TRI_B[mn]]
            if max(a, b, c, d) < 2:
                continue
            tot = 0.0
            for k in range(5):
                ang = Ang[k][kl, mn]
                if ang == 0.0:
                    continue
                ab, cd = tuple(sorted((a, b))), tuple(sorted((c, d)))
                key = (k,) + min(ab, cd) + max(ab, cd)
                if key not in cache:
                    (p, q), (r, s) = min(ab, cd), max(ab, cd)
                    val = radial_integral(k, n[p], ex[p], n[q], ex[q], n[r], ex[r], n[s], ex[s])
                    if key == (0, 0, 0, 2, 2) and abs(f0sd) > 1.0e-9:
                        val = f0sd
                    if key == (2, 0, 2, 0, 2) and abs(g2sd) > 1.0e-9:
                        val = g2sd
                    cache[key] = val
                tot += ang * cache[key]
            out[kl, mn] = tot
    return out
