"""PM6 with d orbitals: CPU (numpy, fp64) restatement of the reference's spd path.

TEST INFRASTRUCTURE ONLY (see the package docstring).  Restates, in one generic formulation:
  * per-element d-shell multipole parameters -- two_elec_two_center_int.py:16-97 (`_pm6_d_param_from_key`), 116-247;
    `GetSlaterCondonParameter` :1309-1364; `AIJL`, `POIJ` cal_par.py:283-393
  * local-frame two-centre integrals over 45 x 45 orbital-pair products as interactions of point-charge multipoles
    (Thiel & Voityuk, TCA 81, 391 (1992)) -- two_elec_two_center_int_local_frame_d_orbitals.py:23-4164
  * rotation to the molecular frame -- RotationMatrixD.py:5-310, two_elec_two_center_int.py:800-1306
  * spd Slater overlaps -- diat_overlapD.py:4-5370
  * one-centre two-electron integrals with d orbitals -- build_two_elec_one_center_int_D.py:15-202, fock.py:29-83,237-253
  * Hcore / Fock with 9 x 9 atom blocks -- hcore.py:61-179, fock.py:132-347
  * packed layout 9 nSH + 4 nHeavy + nHydro -- packd.py:8-218, diag_d.py:18-150

Conventions (probe-verified against the reference, tools/make_golden_pm6d.py):
  molecular orbital order per atom: s, px, py, pz, d(x2-y2), d(xz), d(z2), d(yz), d(xy)   (MOPAC order)
  local (diatomic) frame: z axis from atom j to atom i, atom i at the origin, atom j at z = -r;
  local orbital order: s, p-sigma, p-pi(x), p-pi(y), d-sigma(z2), d-pi(xz), d-pi(yz), d-delta(x2-y2), d-delta(xy)
  w[p, kl, mn] = (kl on atom i | mn on atom j) with kl, mn packed lower-triangle indices (45 each); the reference stores
  the transpose for method="PM6" (hcore.py:143-146, fock.py:277).
"""
import math

import numpy as np

from .tables import Tables

EV = 27.21


# --- element classes ---------------------------------------------------------------------------------------------
def d_shell(Z):
    """The reference's nSuperHeavy set (basics.py:258-269): elements treated with 9 orbitals by method='PM6'."""
    Z = np.asarray(Z)
    return (((Z > 12) & (Z < 18)) | ((Z > 20) & (Z < 30)) | ((Z > 32) & (Z < 36)) | ((Z > 38) & (Z < 48))
            | ((Z > 50) & (Z < 54)) | ((Z > 70) & (Z < 80)) | (Z == 57))  # fmt: skip


def _transition(Z):  # "category A" of two_elec_two_center_int.py:163: d shell has principal quantum number n - 1
    return ((Z > 20) and (Z < 30)) or ((Z > 38) and (Z < 48)) or ((Z > 70) and (Z < 80)) or Z == 57


def norb_of(Z):
    Z = np.asarray(Z)
    return np.where(d_shell(Z), 9, np.where(Z > 1, 4, np.where(Z == 1, 1, 0)))


# --- radial integrals and the additive terms (parameter preparation, per element) ------------------------------------
def _binom(a, b):
    return math.factorial(a) / (math.factorial(b) * math.factorial(a - b))


def slater_condon(K, NA, EA, NB, EB, NC, EC, ND, ED):
    """Radial part R^K(ab, cd) of a one-centre two-electron integral over Slater functions, in eV
    (two_elec_two_center_int.py:1309-1364; MOPAC's rsc)."""
    NA, NB, NC, ND = int(NA), int(NB), int(NC), int(ND)
    AEA, AEB, AEC, AED = math.log(EA), math.log(EB), math.log(EC), math.log(ED)
    NAB, NCD = NA + NB, NC + ND
    ECD, EAB = EC + ED, EA + EB
    E = ECD + EAB
    N = NAB + NCD
    AE, A2, ACD, AAB = math.log(E), math.log(2), math.log(ECD), math.log(EAB)
    C = math.exp(math.log(math.factorial(N - 1)) + NA * AEA + NB * AEB + NC * AEC + ND * AED
                 + 0.5 * (AEA + AEB + AEC + AED) + A2 * (N + 2)
                 - 0.5 * (math.log(math.factorial(2 * NA)) + math.log(math.factorial(2 * NB))
                          + math.log(math.factorial(2 * NC)) + math.log(math.factorial(2 * ND))) - AE * N)  # fmt: skip
    C = C * EV
    S0, S1, S2 = 1 / E, 0, 0
    M = NCD - K
    for I in range(1, M + 1):
        S0 = S0 * E / ECD
        S1 = S1 + S0 * (_binom(NCD - K - 1, I - 1) - _binom(NCD + K + 1 - 1, I - 1)) / _binom(N - 1, I - 1)
    M2 = NCD + K + 1
    for I in range(M + 1, M2 + 1):
        S0 = S0 * E / ECD
        S2 = S2 + S0 * _binom(M2 - 1, I - 1) / _binom(N - 1, I - 1)
    S3 = math.exp(AE * N - ACD * M2 - AAB * (NAB - K)) / _binom(N - 1, M2 - 1)
    return C * (S1 - S2 + S3)


def aijl(Z1, Z2, N1, N2, L):
    """<r^L> between two Slater functions (cal_par.py:377-393)."""
    N1, N2 = int(N1), int(N2)
    if Z1 == 0 or Z2 == 0:
        return 0.0
    a = math.factorial(N1 + N2 + L) / math.sqrt(math.factorial(2 * N1) * math.factorial(2 * N2))
    return (a * (2 * Z1 / (Z1 + Z2)) ** N1 * math.sqrt(2 * Z1 / (Z1 + Z2)) * (2 * Z2 / (Z1 + Z2)) ** N2
            * math.sqrt(2 * Z2 / (Z1 + Z2)) / (Z1 + Z2) ** L)  # fmt: skip


def poij(L, D, FG):
    """Additive term rho that makes the one-centre limit of the point-charge multipole interaction equal FG: golden-
    section search on [0.1, 5] with the reference's exact arithmetic (cal_par.py:283-359) -- the bracket end it
    returns (and hence the last digits of rho) is part of the reference's numbers."""
    if L == 0:
        return 0.5 * EV / FG
    if FG == 0.0:
        return 0.0
    dsq = D * D
    EV4, EV8 = EV * 0.25, EV / 8.0
    A1, A2, G1, G2 = 0.1, 5.0, 0.382, 0.618
    F1 = F2 = 0.0
    for _ in range(100):
        DELTA = A2 - A1
        if DELTA < 1.0e-8:
            break
        Y1 = A1 + DELTA * G1
        Y2 = A1 + DELTA * G2
        if L == 1:
            F1 = (EV4 * (1.0 / Y1 - 1.0 / math.sqrt(Y1**2 + dsq)) - FG) ** 2
            F2 = (EV4 * (1.0 / Y2 - 1.0 / math.sqrt(Y2**2 + dsq)) - FG) ** 2
        else:
            F1 = (EV8 * (1.0 / Y1 - 2.0 / math.sqrt(Y1**2 + dsq * 0.5) + 1.0 / math.sqrt(Y1**2 + dsq)) - FG) ** 2
            F2 = (EV8 * (1.0 / Y2 - 2.0 / math.sqrt(Y2**2 + dsq * 0.5) + 1.0 / math.sqrt(Y2**2 + dsq)) - FG) ** 2
        if F1 < F2:
            A2 = Y2
        else:
            A1 = Y1
    return A2 if F1 >= F2 else A1


def d_element_multipoles(Z, qn, zetas, zetap, zetad, zs, zp, zd, g2sd):
    """Charge separations (dp, ds, dd) and additive terms (rho3..rho6) of one d-shell element
    (two_elec_two_center_int.py:31-97 and 210-243).  zs/zp/zd: the internal ('tail') exponents."""
    qd = qn - 1 if _transition(Z) else qn
    sc = slater_condon
    dp_add = (4.0 / 15.0) * sc(1, qn, zp, qd, zd, qn, zp, qd, zd)
    if _transition(Z) and g2sd > 1.0e-9:
        ds_add = 0.2 * g2sd
    else:
        ds_add = 0.2 * sc(2, qn, zs, qd, zd, qn, zs, qd, zd)
    dd_add = (4.0 / 49.0) * sc(2, qd, zd, qd, zd, qd, zd, qd, zd)
    dd0_add = sc(0, qd, zd, qd, zd, qd, zd, qd, zd)
    dd4 = sc(4, qd, zd, qd, zd, qd, zd, qd, zd)
    dp3 = (27.0 / 245.0) * sc(3, qn, zp, qd, zd, qn, zp, qd, zd)
    aij52 = aijl(zetap, zetad, qn, qd, 1)
    aij43 = aijl(zetas, zetad, qn, qd, 2)
    aij63 = aijl(zetad, zetad, qd, qd, 2)
    out = {}
    out["dp"] = aij52 / math.sqrt(5)
    D = math.sqrt(aij43 * math.sqrt(1.0 / 15.0)) * math.sqrt(2.0)
    out["ds"] = D
    out["rho5"] = poij(2, D, ds_add)
    FG = dd0_add + dd_add + 4 / 49 * dd4
    FG1 = dd0_add + 0.5 * dd_add - 24 / 441 * dd4
    FG2 = dd0_add - dd_add + 6 / 441 * dd4
    out["rho3"] = poij(0, 1.0, 0.2 * (FG + 2.0 * FG1 + 2.0 * FG2))
    D = aij52 / math.sqrt(5.0)
    FG = dp_add + dp3
    FG1 = 3 / 49 * 245 / 27 * dp3
    out["rho4"] = poij(1, D, FG - 1.8 * FG1)
    D = math.sqrt(2.0 * (aij63 / 7.0))
    out["ddq"] = D
    FG = 3 / 4 * dd_add + 20 / 441 * dd4
    FG1 = 35 / 441 * dd4
    out["rho6"] = poij(2, D, FG - (20.0 / 35.0) * FG1)
    return out


_ELEMENT_CACHE = {}


def atom_multipoles_spd(Z, par, mp):
    """Per-atom arrays of every multipole parameter the spd pair code needs.  `mp` = (dd, qq, rho0, rho1, rho2) of the
    sp path (integrals.atom_multipoles).  rho2d = POIJ(2, qq sqrt2, (gpp - gp2)/2) replaces rho2 wherever a d orbital
    takes part (two_elec_two_center_int.py:247, 260)."""
    T = Tables.get()
    dd, qq, rho0, rho1, rho2 = mp
    n = Z.shape[0]
    out = {k: np.zeros(n) for k in ("dp", "ds", "ddq", "rho3", "rho4", "rho5", "rho6", "rho2d")}
    isd = d_shell(Z)
    for a in range(n):
        z = int(Z[a])
        if z > 2:
            key = ("r2d", float(qq[a]), float(par["g_pp"][a]), float(par["g_p2"][a]))
            if key not in _ELEMENT_CACHE:
                _ELEMENT_CACHE[key] = poij(2, float(qq[a]) * math.sqrt(2), 0.5 * (float(par["g_pp"][a]) - float(par["g_p2"][a])))
            out["rho2d"][a] = _ELEMENT_CACHE[key]
        if not isd[a]:
            continue
        if par["zeta_d"][a] == 0.0:
            raise NotImplementedError(f"PM6: element Z={z} is in the reference's d-shell set but has no d parameters")
        if par["rho_core"][a] != 0.0:
            raise NotImplementedError(f"PM6: element Z={z} uses rho_core (not covered)")
        key = (z, float(par["zeta_s"][a]), float(par["zeta_p"][a]), float(par["zeta_d"][a]), float(par["s_orb_exp_tail"][a]),
               float(par["p_orb_exp_tail"][a]), float(par["d_orb_exp_tail"][a]), float(par["G2SD"][a]))  # fmt: skip
        if key not in _ELEMENT_CACHE:
            _ELEMENT_CACHE[key] = d_element_multipoles(z, int(T.qn_int[z]), *key[1:])
        for k, v in _ELEMENT_CACHE[key].items():
            out[k][a] = v
    out.update(dd=dd, qq=qq, rho0=rho0, rho1=rho1, rho2=rho2)
    return out


# --- angular algebra: real harmonics, pair products, multipole coefficients -----------------------------------------------
L_OF = np.array([0, 1, 1, 1, 2, 2, 2, 2, 2])
TRI = [(a, b) for a in range(9) for b in range(a + 1)]  # packed lower triangle: index = a (a + 1) / 2 + b
PAIR = np.zeros((9, 9), dtype=np.int64)
for _k, (_a, _b) in enumerate(TRI):
    PAIR[_a, _b] = PAIR[_b, _a] = _k
TRI_A = np.array([t[0] for t in TRI])
TRI_B = np.array([t[1] for t in TRI])
WEIGHT45 = np.where(TRI_A == TRI_B, 1.0, 2.0)  # fock.py:18-27
S3, S5, S15 = math.sqrt(3.0), math.sqrt(5.0), math.sqrt(15.0)


def _local_orbitals(x, y, z):
    """Angular parts (normalised to <f f> = 1 under dOmega / 4 pi) in LOCAL order: s, p(z,x,y), d(z2, xz, yz, x2-y2, xy)."""
    return np.stack([np.ones_like(x), S3 * z, S3 * x, S3 * y, 0.5 * S5 * (3 * z * z - 1.0), S15 * x * z, S15 * y * z,
                     0.5 * S15 * (x * x - y * y), S15 * x * y])  # fmt: skip


def _racah(x, y, z):
    """C_lm (Racah-normalised real harmonics): index 0 -> (0,0); 1..3 -> (1,0),(1,1c),(1,1s); 4..8 -> (2,0),(2,1c),(2,1s),(2,2c),(2,2s)."""
    return np.stack([np.ones_like(x), z, x, y, 0.5 * (3 * z * z - 1.0), S3 * x * z, S3 * y * z, 0.5 * S3 * (x * x - y * y), S3 * x * y])


# multipole sources of an atom: (name, l, charge separation key, additive term key)
SOURCES = [("ss0", 0), ("sp1", 1), ("pp2", 2), ("sd2", 2), ("pd1", 1), ("dd0", 0), ("dd2", 2)]
SRC_OF_TYPE = {(0, 0): {0: 0}, (0, 1): {1: 1}, (1, 1): {0: 0, 2: 2}, (0, 2): {2: 3}, (1, 2): {1: 4}, (2, 2): {0: 5, 2: 6}}
# <r^l> -> D^l conversion of each product type and moment of the unit point-charge configuration per D^l
_G = {1: 1.0 / S3, 2: 1.0 / 5.0, 3: 1.0 / S15, 4: 1.0 / S5, 6: 1.0 / 7.0}
_KAPPA = {0: 1.0, 1: 1.0, 2: 1.0, 3: 1.0, 4: 1.5, 5: S3, 6: S3, 7: S3, 8: S3}
M_INDEX = {0: 0, 1: 0, 2: 1, 3: 2, 4: 0, 5: 1, 6: 2, 7: 3, 8: 4}  # C_lm index -> m slot (0, 1c, 1s, 2c, 2s)
_COEF = None


def multipole_coefficients():
    """c[kl, source, m]: expansion of the 45 local orbital products in point-charge multipoles (l <= 2, Thiel-Voityuk):
    c = <f_k f_l C_lm> / (g_type kappa_lm).  The reference carries the cosine-type (m = 0, 1c, 2c) coefficients with 6
    decimals (0.666667, 1.154701, 0.577350 ...) and the sine-type ones (1s, 2s) in full precision; both are reproduced."""
    global _COEF
    if _COEF is not None:
        return _COEF
    t, wt = np.polynomial.legendre.leggauss(12)
    ph = (np.arange(24) + 0.5) * (2 * math.pi / 24)
    ct, pp = np.meshgrid(t, ph, indexing="ij")
    st = np.sqrt(1 - ct * ct)
    x, y, z = (st * np.cos(pp)).ravel(), (st * np.sin(pp)).ravel(), ct.ravel()
    wq = (np.repeat(wt, 24) / 24.0) / 2.0  # integrates dOmega / 4 pi
    f, C = _local_orbitals(x, y, z), _racah(x, y, z)
    c = np.zeros((45, len(SOURCES), 5))
    for kl, (a, b) in enumerate(TRI):
        ty = tuple(sorted((L_OF[a], L_OF[b])))
        for lm in range(9):
            l = 0 if lm == 0 else (1 if lm < 4 else 2)
            if l not in SRC_OF_TYPE[ty]:
                continue
            s = SRC_OF_TYPE[ty][l]
            A = float(np.sum(wq * f[a] * f[b] * C[lm]))
            if abs(A) < 1e-12:
                continue
            val = A if l == 0 else A / (_G[s] * _KAPPA[lm])
            c[kl, s, M_INDEX[lm]] = val if M_INDEX[lm] in (2, 4) else round(val, 6)
    _COEF = c
    return c


def multipole_coefficients_yx():
    """Coefficients of the d atom in a (d element, sp-only heavy element) pair: in that branch the reference carries the
    d-sigma x d-delta quadrupole terms with the opposite sign (two_elec_two_center_int_local_frame_d_orbitals.py:1281,
    1386, 1389 against 3369, 3911 of the d-d branch).  Part of the reference's numbers, reproduced."""
    c = multipole_coefficients().copy()
    for kl in (PAIR[7, 4], PAIR[8, 4]):
        c[kl, 6, 3:5] *= -1.0
    return c


def _configuration(l, m, D):
    """Point charges (q, x, y, z) of the unit multipole (l, m) with separation parameter D (npairs,): D is the dipole
    half-length for l = 1 and the UNSCALED quadrupole length for l = 2."""
    z0 = np.zeros_like(D)
    if l == 0:
        return [(1.0, z0, z0, z0)]
    if l == 1:
        ax = {0: 2, 1: 0, 2: 1}[m]
        out = []
        for s in (1.0, -1.0):
            p = [z0, z0, z0]
            p[ax] = s * D
            out.append((0.5 * s, *p))
        return out
    r2 = math.sqrt(2.0) * D
    if m == 0:  # Q~zx + 1/2 Q~xy: +1/4 at z = +-sqrt2 D, -1/8 at x = +-sqrt2 D and at y = +-sqrt2 D
        return [(0.25, z0, z0, r2), (0.25, z0, z0, -r2), (-0.125, r2, z0, z0), (-0.125, -r2, z0, z0),
                (-0.125, z0, r2, z0), (-0.125, z0, -r2, z0)]  # fmt: skip
    if m in (1, 2):  # (2,1): +-1/4 at (+-D, +-D) in the xz (yz) plane
        out = []
        for sa in (1.0, -1.0):
            for sb in (1.0, -1.0):
                p = [z0, z0, sb * D]
                p[0 if m == 1 else 1] = sa * D
                out.append((0.25 * sa * sb, *p))
        return out
    if m == 3:  # (2,2c): +1/4 at x = +-sqrt2 D, -1/4 at y = +-sqrt2 D
        return [(0.25, r2, z0, z0), (0.25, -r2, z0, z0), (-0.25, z0, r2, z0), (-0.25, z0, -r2, z0)]
    out = []  # (2,2s): the same square turned by 45 degrees
    for sa in (1.0, -1.0):
        for sb in (1.0, -1.0):
            out.append((0.25 * sa * sb, sa * D, sb * D, z0))
    return out


def _source_params(mpd, idx, use_rho2d):
    """(D, rho) per source for the atoms `idx`; quadrupole lengths are returned unscaled (ds, ddq are stored with their sqrt2)."""
    r2 = math.sqrt(2.0)
    return [
        (None, mpd["rho0"][idx]), (mpd["dd"][idx], mpd["rho1"][idx]),
        (mpd["qq"][idx], (mpd["rho2d"] if use_rho2d else mpd["rho2"])[idx]),
        (mpd["ds"][idx] / r2, mpd["rho5"][idx]), (mpd["dp"][idx], mpd["rho4"][idx]), (None, mpd["rho3"][idx]),
        (mpd["ddq"][idx] / r2, mpd["rho6"][idx]),
    ]  # fmt: skip


def _interaction_table(r, pa, pb):
    """V[pair, s, t, m] = interaction (eV) of multipole (l_s, m) on atom i (origin) with (l_t, m) on atom j (z = -r)."""
    n = r.shape[0]
    V = np.zeros((n, len(SOURCES), len(SOURCES), 5))
    for s, (_, ls) in enumerate(SOURCES):
        for t, (_, lt) in enumerate(SOURCES):
            add = (pa[s][1] + pb[t][1]) ** 2
            for m in range(5):
                am = 0 if m == 0 else (1 if m < 3 else 2)
                if am > ls or am > lt:
                    continue
                Da = pa[s][0] if pa[s][0] is not None else np.zeros(n)
                Db = pb[t][0] if pb[t][0] is not None else np.zeros(n)
                tot = np.zeros(n)
                for qa, xa, ya, za in _configuration(ls, m, Da):
                    for qb, xb, yb, zb in _configuration(lt, m, Db):
                        tot = tot + qa * qb * EV / np.sqrt((xa - xb) ** 2 + (ya - yb) ** 2 + (za - zb + r) ** 2 + add)
                V[:, s, t, m] = tot
    return V


def local_integrals_spd(r, mpd, idxi, idxj, norb_i, norb_j):
    """L (npairs, 45, 45): local-frame (kl on i | mn on j) for every product pair in which a d orbital takes part, with
    rho2d as the p-p quadrupole term.  The sp x sp block is NOT this expansion: the reference keeps the original MNDO
    point-charge formulas of its sp code for it (RotationMatrixD.py:308), see two_center_integrals_spd."""
    c = multipole_coefficients()
    VD = _interaction_table(r, _source_params(mpd, idxi, True), _source_params(mpd, idxj, True))
    L = np.einsum("ksm,pstm,ltm->pkl", c, VD, c, optimize=True)
    yx = (np.asarray(norb_i) == 9) & (np.asarray(norb_j) == 4)
    if np.any(yx):
        L[yx] = np.einsum("ksm,pstm,ltm->pkl", multipole_coefficients_yx(), VD[yx], c, optimize=True)
    # one product pair carries the 6-decimal constant where its neighbours carry the full one:
    # (d-sigma p-pi(y) | p-pi(y) s) = -0.577350 [pd dipole | sp dipole], ...local_frame_d_orbitals.py:1260, 3276
    L[:, PAIR[4, 3], PAIR[3, 0]] += (round(-1.0 / S3, 6) + 1.0 / S3) * VD[:, 4, 1, 2]
    L[:, :10, :10] = 0.0
    valid = np.array([[max(a, b) < n for (a, b) in TRI] for n in (0, 1, 4, 9)])  # by norb class
    cls = {0: 0, 1: 1, 4: 2, 9: 3}
    vi = valid[[cls[int(n)] for n in norb_i]]
    vj = valid[[cls[int(n)] for n in norb_j]]
    return L * vi[:, :, None] * vj[:, None, :]


# --- rotation to the molecular frame ------------------------------------------------------------------------------------
# quadratic forms of the five d functions, molecular order (x2-y2, xz, z2, yz, xy) and local order (z2, xz, yz, x2-y2, xy)
def _dform(name):
    M = np.zeros((3, 3))
    if name == "z2":
        M[0, 0] = M[1, 1] = -1.0 / S3
        M[2, 2] = 2.0 / S3
    elif name == "x2-y2":
        M[0, 0], M[1, 1] = 1.0, -1.0
    else:
        i, j = {"xz": (0, 2), "yz": (1, 2), "xy": (0, 1)}[name]
        M[i, j] = M[j, i] = 1.0
    return M  # tr(M M') = 2 delta


_DMOL = np.stack([_dform(n) for n in ("x2-y2", "xz", "z2", "yz", "xy")])
_DLOC = np.stack([_dform(n) for n in ("z2", "xz", "yz", "x2-y2", "xy")])


def local_axes(ez):
    """(ex, ey) completing the local z axis, the reference's choice (RotationMatrixD.py:11-45; MOPAC rotmat): with
    ez = (ca sb, sa sb, cb): ex = (ca cb, sa cb, -sb), ey = (-sa, ca, 0); for ez along +-z: ex = (1, 0, 0), ey = (0, +-1, 0).
    The choice matters at the 1e-8 level only, through the reference's unequal x-/y-type constants (see above)."""
    xy = np.sqrt(ez[:, 0] ** 2 + ez[:, 1] ** 2)
    ok = xy >= 1.0e-10
    sgn = np.sign(ez[:, 2])
    xs = np.where(ok, xy, 1.0)
    ca = np.where(ok, ez[:, 0] / xs, sgn)
    sa = np.where(ok, ez[:, 1] / xs, 0.0)
    cb = np.where(ok, ez[:, 2], sgn)
    sb = np.where(ok, xy, 0.0)
    ex = np.stack([ca * cb, sa * cb, -sb], axis=1)
    ey = np.stack([-sa, ca, np.zeros_like(ca)], axis=1)
    return ex, ey


def orbital_rotation(ez):
    """R (n, 9, 9): molecular orbital a = sum_b R[a, b] local orbital b."""
    n = ez.shape[0]
    ex, ey = local_axes(ez)
    E = np.stack([ex, ey, ez], axis=1)  # rows: local axes in molecular components
    R = np.zeros((n, 9, 9))
    R[:, 0, 0] = 1.0
    R[:, 1:4, 1], R[:, 1:4, 2], R[:, 1:4, 3] = ez, ex, ey  # local p order: sigma (z), pi (x), pi (y)
    # local d_b(r) = (E r)^T Mloc_b (E r) = r^T (E^T Mloc_b E) r ; expand in the molecular forms
    Q = np.einsum("nia,bij,njc->nbac", E, _DLOC, E)
    R[:, 4:, 4:] = 0.5 * np.einsum("nbac,dac->ndb", Q, _DMOL)
    return R


def pair_transform(R):
    """T (n, 45, 45): molecular product kl = sum_mn T[kl, mn] local product mn."""
    a, b = TRI_A, TRI_B
    T = R[:, a][:, :, a] * R[:, b][:, :, b] + R[:, a][:, :, b] * R[:, b][:, :, a]
    T[:, :, a == b] *= 0.5
    return T


def two_center_integrals_spd(ni, nj, idxi, idxj, xij, rij, mpd):
    """w (npairs,45,45) molecular frame, plus the core-attraction blocks e_i = -tore_j (kl_i | ss_j), e_j = -tore_i (ss_i | mn_j)
    as (npairs, 9, 9) upper triangles (two_elec_two_center_int.py:800-1306)."""
    from .integrals import two_center_integrals_geom

    T = Tables.get()
    L = local_integrals_spd(rij, mpd, idxi, idxj, norb_of(ni), norb_of(nj))
    Tp = pair_transform(orbital_rotation(-xij))
    w = np.matmul(Tp, np.matmul(L, Tp.transpose(0, 2, 1)))
    # sp x sp block: the sp path (MNDO formulas, rho2 from the secant iteration), already in the molecular frame
    mp = (mpd["dd"], mpd["qq"], mpd["rho0"], mpd["rho1"], mpd["rho2"])
    w[:, :10, :10] = two_center_integrals_geom(ni, nj, idxi, idxj, xij, rij, mp, T)[0]
    n = rij.shape[0]
    e_i = np.zeros((n, 9, 9))
    e_j = np.zeros((n, 9, 9))
    e_i[:, TRI_B, TRI_A] = -T.tore[nj][:, None] * w[:, :, 0]
    e_j[:, TRI_B, TRI_A] = -T.tore[ni][:, None] * w[:, 0, :]
    return w, e_i, e_j


# --- Slater overlaps with d functions --------------------------------------------------------------------------------------
def _qn_tables():
    import json
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "..", "..", "pyseqm_b200", "data", "element_tables.json")) as f:
        d = json.load(f)
    return np.asarray(d["qn_int"], dtype=np.int64), np.asarray(d["qnD_int"], dtype=np.int64)


def _sigma_pi_delta_poly(na, la, nb, lb, m):
    """Polynomial in (xi, eta) of the prolate-spheroidal overlap integrand of (na, la, m) on A (origin) with (nb, lb, m)
    on B (at +R on the local z axis), lengths in units of R/2, and the angular constant in front of it."""
    from .integrals import _1META2, _M_ONE_P, _ONE_P, _XI2M1, _XI_M_ETA, _XI_P_ETA, _poly_mul, _poly_pow

    rho2 = _poly_mul(_XI2M1, _1META2)  # (xi^2 - 1)(1 - eta^2) = (x^2 + y^2) / (R/2)^2

    def radial_angular(n, l, zpoly, rpoly):
        # r^(n-1) (f_lm r^l / r^l): (r)^(n - l) once the volume element (xi^2 - eta^2) = r_A r_B is shared out
        base = _poly_pow(rpoly, n - l)
        if l == 0:
            return 1.0, base
        if l == 1:
            return S3, (_poly_mul(base, zpoly) if m == 0 else base)
        if m == 0:  # (sqrt5 / 2) (3 z^2 - r^2)
            q = 3.0 * _poly_mul(zpoly, zpoly)
            r2 = _poly_mul(rpoly, rpoly)
            k0, k1 = max(q.shape[0], r2.shape[0]), max(q.shape[1], r2.shape[1])
            tot = np.zeros((k0, k1))
            tot[: q.shape[0], : q.shape[1]] += q
            tot[: r2.shape[0], : r2.shape[1]] -= r2
            return 0.5 * S5, _poly_mul(base, tot)
        if m == 1:
            return S15, _poly_mul(base, zpoly)
        return 0.5 * S15, base

    ka, pa = radial_angular(na, la, _ONE_P, _XI_P_ETA)
    kb, pb = radial_angular(nb, lb, _M_ONE_P, _XI_M_ETA)
    poly = _poly_mul(pa, pb)
    for _ in range(m):  # the rho factors of both functions: rho^(2m)
        poly = _poly_mul(poly, rho2)
    phi = 0.5 if m == 0 else 0.25  # (2 pi or pi) / 4 pi
    return phi * ka * kb, poly


def _sto_overlap_general(na, la, za, nb, lb, zb, m, r, r_regime=None):
    from .integrals import _aintgs, _bintgs

    c, poly = _sigma_pi_delta_poly(na, la, nb, lb, m)
    alpha = 0.5 * r * (za + zb)
    beta = 0.5 * r * (za - zb)
    kmax = max(poly.shape) - 1
    A = _aintgs(alpha, kmax)
    B = _bintgs(beta, kmax, None if r_regime is None else 0.5 * r_regime * (za - zb))
    tot = 0.0
    for k in range(poly.shape[0]):
        for l in range(poly.shape[1]):
            if poly[k, l] != 0.0:
                tot = tot + poly[k, l] * A[k] * B[l]
    norm = ((2.0 * za) ** (na + 0.5) * (2.0 * zb) ** (nb + 0.5) / math.sqrt(math.factorial(2 * na) * math.factorial(2 * nb))
            * (0.5 * r) ** (na + nb + 1))  # fmt: skip
    return c * norm * tot


def overlap_spd(ni, nj, xij, rij, zeta_a, zeta_b, rij_regime=None):
    """di (npairs, 9, 9) = <mu on i | nu on j>, molecular frame (diat_overlapD.py:4-5148); zeta_* (npairs, 3) = s, p, d.
    rij_regime: distances that select the B-integral regime (see integrals._bintgs), default rij."""
    T = Tables.get()
    qn, qnd = _qn_tables()
    npairs = rij.shape[0]
    di = np.zeros((npairs, 9, 9))
    within = rij <= T.overlap_cutoff
    nob_i, nob_j = norb_of(ni), norb_of(nj)
    # local frame: z from i to j; local orbital order s, p(z, x, y), d(z2, xz, yz, x2-y2, xy)
    loc_m = [0, 0, 1, 1, 0, 1, 1, 2, 2]
    partner = {2: 2, 3: 3, 5: 5, 6: 6, 7: 7, 8: 8}
    same_m_pairs = [(a, b) for a in range(9) for b in range(9)
                    if loc_m[a] == loc_m[b] and ((loc_m[a] == 0) or (a in (2, 5, 7)) == (b in (2, 5, 7)))]  # fmt: skip
    keys = np.stack([ni, nj], axis=1)
    R = orbital_rotation(xij)
    for zi, zj in {tuple(k) for k in keys.tolist()}:
        msk = (ni == zi) & (nj == zj) & within
        if not np.any(msk):
            continue
        r = rij[msk]
        na = [int(qn[zi])] * 4 + [int(qnd[zi])] * 5
        nb = [int(qn[zj])] * 4 + [int(qnd[zj])] * 5
        noa, nobj = int(norb_of(zi)), int(norb_of(zj))
        Sl = np.zeros((r.shape[0], 9, 9))
        cache = {}
        for a, b in same_m_pairs:
            if a >= noa or b >= nobj:
                continue
            la, lb, m = int(L_OF[a]), int(L_OF[b]), loc_m[a]
            key = (la, lb, m)
            if key not in cache:
                za = zeta_a[msk, la]
                zb = zeta_b[msk, lb]
                cache[key] = _sto_overlap_general(na[a], la, za, nb[b], lb, zb, m, r,
                                                  None if rij_regime is None else rij_regime[msk])
            Sl[:, a, b] = cache[key]
        Rm = R[msk]
        blk = np.matmul(Rm, np.matmul(Sl, Rm.transpose(0, 2, 1)))
        if noa == 9 and nobj == 9:
            # the reference's (d_yz, d_xy) element carries its delta-bar term with the wrong sign
            # (diat_overlapD.py:5104-5116: "+ ca sb cb (2 ca^2 - 1)" where the rotation gives "-"); reproduced
            e = xij[msk]
            xy = np.sqrt(e[:, 0] ** 2 + e[:, 1] ** 2)
            ok = xy >= 1.0e-10
            ca = np.where(ok, e[:, 0] / np.where(ok, xy, 1.0), np.sign(e[:, 2]))
            cb = np.where(ok, e[:, 2], np.sign(e[:, 2]))
            sb = np.where(ok, xy, 0.0)
            fix = 2.0 * Sl[:, 7, 7] * ca * sb * cb * (2.0 * ca * ca - 1.0)
            blk[:, 7, 8] += fix
            blk[:, 8, 7] += fix
        di[msk] = blk
    return di


# --- one-centre two-electron integrals with d orbitals -------------------------------------------------------------------
# molecular orbital order s, px, py, pz, d(x2-y2), d(xz), d(z2), d(yz), d(xy)
def _molecular_orbitals(x, y, z):
    return np.stack([np.ones_like(x), S3 * x, S3 * y, S3 * z, 0.5 * S15 * (x * x - y * y), S15 * x * z,
                     0.5 * S5 * (3 * z * z - 1.0), S15 * y * z, S15 * x * y])  # fmt: skip


_ANG = None


def _angular_factors():
    """Ang[k][mu nu, lam sig] = <f_mu f_nu P_k(cos gamma_12) f_lam f_sig> over both unit spheres (dOmega / 4 pi each): the
    angular part of (mu nu | lam sig) = sum_k Ang[k] R^k  (Slater-Condon expansion; Gauss-Legendre x uniform-phi product
    quadrature, exact for the degree <= 12 polynomials that occur)."""
    global _ANG
    if _ANG is not None:
        return _ANG
    t, wt = np.polynomial.legendre.leggauss(10)
    nph = 20
    ph = (np.arange(nph) + 0.5) * (2 * math.pi / nph)
    ct, pp = np.meshgrid(t, ph, indexing="ij")
    st = np.sqrt(1 - ct * ct)
    x, y, z = (st * np.cos(pp)).ravel(), (st * np.sin(pp)).ravel(), ct.ravel()
    wq = (np.repeat(wt, nph) / nph) / 2.0
    f = _molecular_orbitals(x, y, z)
    prod = f[TRI_A] * f[TRI_B] * wq[None, :]  # (45, N)
    cosg = np.clip(x[:, None] * x[None, :] + y[:, None] * y[None, :] + z[:, None] * z[None, :], -1.0, 1.0)
    out = []
    for k in range(5):
        Pk = np.polynomial.legendre.legval(cosg, [0] * k + [1])
        A = prod @ Pk @ prod.T
        A[np.abs(A) < 1e-13] = 0.0
        out.append(A)
    _ANG = out
    return out


def one_center_integrals_d(Z, zs, zp, zd, f0sd, g2sd):
    """I (45, 45) = (kl | mn) on one atom for every quadruple that contains a d orbital (zero for pure sp quadruples,
    which the g_ss ... h_sp parameters cover), from Slater-Condon radial integrals over the internal exponents
    (build_two_elec_one_center_int_D.py:15-202: R016 -> F0SD and R244 -> G2SD when those parameters are set)."""
    qn, qnd = _qn_tables()
    n = {0: int(qn[Z]), 1: int(qn[Z]), 2: int(qnd[Z])}
    ex = {0: zs, 1: zp, 2: zd}
    Ang = _angular_factors()
    lo = np.array([0, 1, 1, 1, 2, 2, 2, 2, 2])
    I = np.zeros((45, 45))
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
                key = (k,) + (min(ab, cd) + max(ab, cd))
                if key not in cache:
                    (p, q), (r, s) = min(ab, cd), max(ab, cd)
                    val = slater_condon(k, n[p], ex[p], n[q], ex[q], n[r], ex[r], n[s], ex[s])
                    if key == (0, 0, 0, 2, 2) and abs(f0sd) > 1.0e-9:
                        val = f0sd
                    if key == (2, 0, 2, 0, 2) and abs(g2sd) > 1.0e-9:
                        val = g2sd
                    cache[key] = val
                tot += ang * cache[key]
            I[kl, mn] = tot
    return I


_ONE_CENTER_CACHE = {}


def one_center_fock_d(Z, par, PA):
    """F_A += sum_{lam sig} P_A[lam sig] ((mu nu | lam sig) - 1/2 (mu lam | nu sig)) over the d-containing integrals
    (fock.py:237-253 with W of calc_integral).  PA (nat, 9, 9) symmetric diagonal density blocks; returns (nat, 9, 9)."""
    out = np.zeros_like(PA)
    for a in np.nonzero(d_shell(Z))[0]:
        key = (int(Z[a]), float(par["s_orb_exp_tail"][a]), float(par["p_orb_exp_tail"][a]), float(par["d_orb_exp_tail"][a]),
               float(par["F0SD"][a]), float(par["G2SD"][a]))  # fmt: skip
        if key not in _ONE_CENTER_CACHE:
            I = one_center_integrals_d(*key)
            _ONE_CENTER_CACHE[key] = I[PAIR[:, :, None, None], PAIR[None, None, :, :]]  # (mu, nu, lam, sig)
        I4 = _ONE_CENTER_CACHE[key]
        out[a] = np.einsum("mnls,ls->mn", I4, PA[a]) - 0.5 * np.einsum("mlns,ls->mn", I4, PA[a])
    return out


# --- Hcore and Fock in the 9-slot dense layout (nmol, 9 molsize, 9 molsize) ------------------------------------------------
def _blocks9(X, nmol, molsize):
    return X.reshape(nmol, molsize, 9, molsize, 9).transpose(0, 1, 3, 2, 4)


def check_d_first(P):
    """packd/unpackd (packd.py:195-218) take the first nSuperHeavy atoms of a molecule as the d-shell atoms: with rows
    sorted by descending Z that fails when an sp-only element is heavier than a d-shell element of the same molecule."""
    d = d_shell(P.species)
    heavy_sp = (P.species > 1) & ~d
    first_sp = np.where(heavy_sp.any(axis=1), heavy_sp.argmax(axis=1), P.molsize)
    last_d = np.where(d.any(axis=1), P.molsize - 1 - d[:, ::-1].argmax(axis=1), -1)
    if np.any(last_d > first_sp):
        raise ValueError("PM6: d-shell elements must precede the sp-only elements of a molecule in the Z-sorted order")


def build_hcore_spd(P, par, mpd):
    """Dense symmetric Hcore (hcore.py:61-179, PM6 branch) + the integrals it is made of."""
    from .hamiltonian import Segments

    nmol, molsize = P.nmol, P.molsize
    nat = P.Z.shape[0]
    w, e_i, e_j = two_center_integrals_spd(P.ni, P.nj, P.idxi, P.idxj, P.xij, P.rij, mpd)
    zeta = np.stack([par["zeta_s"], par["zeta_p"], par["zeta_d"]], axis=1)
    di = overlap_spd(P.ni, P.nj, P.xij, P.rij, zeta[P.idxi], zeta[P.idxj])
    nob = norb_of(P.Z)
    D = np.zeros((nat, 9, 9))
    D[:, 0, 0] = par["U_ss"]
    for k in (1, 2, 3):
        D[:, k, k] = par["U_pp"]
    for k in range(4, 9):
        D[:, k, k] = par["U_dd"]
    P.seg_i = getattr(P, "seg_i", None) or Segments(P.idxi, nat)
    P.seg_j = getattr(P, "seg_j", None) or Segments(P.idxj, nat)
    P.seg_i.add(D, e_i)
    P.seg_j.add(D, e_j)
    D = D + np.triu(D, 1).transpose(0, 2, 1)
    live = np.arange(9)[None, :] < nob[:, None]
    D = D * live[:, :, None] * live[:, None, :]
    bA = np.stack([par["beta_s"]] + [par["beta_p"]] * 3 + [par["beta_d"]] * 5, axis=1)
    Hab = di * 0.5 * (bA[P.idxi][:, :, None] + bA[P.idxj][:, None, :])
    Hab = Hab * live[P.idxi][:, :, None] * live[P.idxj][:, None, :]
    H = np.zeros((nmol, 9 * molsize, 9 * molsize))
    Hb = _blocks9(H, nmol, molsize)
    Hb[P.atom_molid, P.atom_pos, P.atom_pos] = D
    mi, ai, aj = P.pair_molid, P.atom_pos[P.idxi], P.atom_pos[P.idxj]
    Hb[mi, ai, aj] = Hab
    Hb[mi, aj, ai] = Hab.transpose(0, 2, 1)
    return dict(H=H, w=w, di=di)


def initial_density_spd(P):
    """tore/4 on s and p of heavy atoms (d shells start empty), 1 on hydrogen s (scf_loop.py:2066-2081)."""
    T = Tables.get()
    nmol, molsize = P.nmol, P.molsize
    D = np.zeros((nmol, 9 * molsize, 9 * molsize))
    Db = _blocks9(D, nmol, molsize)
    heavy = P.Z > 1
    val = T.tore[P.Z] / 4.0
    for k in range(4):
        Db[P.atom_molid[heavy], P.atom_pos[heavy], P.atom_pos[heavy], k, k] = val[heavy]
    hyd = P.Z == 1
    Db[P.atom_molid[hyd], P.atom_pos[hyd], P.atom_pos[hyd], 0, 0] = 1.0
    return D


def build_fock_spd(P, par, H, w, Dm):
    """F = H + G(D), 9 x 9 blocks (fock.py:132-347 with themethod == 'PM6')."""
    nmol, molsize = P.nmol, P.molsize
    nat = P.Z.shape[0]
    Db = _blocks9(Dm, nmol, molsize)
    PA = Db[P.atom_molid, P.atom_pos, P.atom_pos]
    gss, gpp, gsp, gp2, hsp = par["g_ss"], par["g_pp"], par["g_sp"], par["g_p2"], par["h_sp"]
    Pss = PA[:, 0, 0]
    Ppt = PA[:, 1, 1] + PA[:, 2, 2] + PA[:, 3, 3]
    G = np.zeros((nat, 9, 9))
    G[:, 0, 0] = 0.5 * Pss * gss + Ppt * (gsp - 0.5 * hsp)
    for k in (1, 2, 3):
        Pk = PA[:, k, k]
        G[:, k, k] = Pss * (gsp - 0.5 * hsp) + 0.5 * Pk * gpp + (Ppt - Pk) * (1.25 * gp2 - 0.25 * gpp)
        G[:, 0, k] = PA[:, 0, k] * (1.5 * hsp - 0.5 * gsp)
    for a, b in ((1, 2), (1, 3), (2, 3)):
        G[:, a, b] = PA[:, a, b] * (0.75 * gpp - 1.25 * gp2)
    G = G + np.triu(G, 1).transpose(0, 2, 1)
    G += one_center_fock_d(P.Z, par, PA)
    pk = PA[:, TRI_A, TRI_B] * WEIGHT45
    JA = np.matmul(w, pk[P.idxj][:, :, None])[:, :, 0]
    JB = np.matmul(pk[P.idxi][:, None, :], w)[:, 0, :]
    J = np.zeros((nat, 45))
    P.seg_i.add(J, JA)
    P.seg_j.add(J, JB)
    Jm = np.zeros((nat, 9, 9))
    Jm[:, TRI_A, TRI_B] = J
    Jm = Jm + np.tril(Jm, -1).transpose(0, 2, 1)
    G += Jm
    mi, ai, aj = P.pair_molid, P.atom_pos[P.idxi], P.atom_pos[P.idxj]
    Dab = Db[mi, ai, aj]
    w4 = w[:, PAIR[:, :, None, None], PAIR[None, None, :, :]]  # (p, mu, nu, lam, sig)
    K = -0.5 * np.einsum("pmnls,pns->pml", w4, Dab)
    nob = norb_of(P.Z)
    live = np.arange(9)[None, :] < nob[:, None]
    G = G * live[:, :, None] * live[:, None, :]
    K = K * live[P.idxi][:, :, None] * live[P.idxj][:, None, :]
    F = H.copy()
    Fb = _blocks9(F, nmol, molsize)
    Fb[P.atom_molid, P.atom_pos, P.atom_pos] += G
    Fb[mi, ai, aj] += K
    Fb[mi, aj, ai] += K.transpose(0, 2, 1)
    return F


# --- density, energies, gradient, driver --------------------------------------------------------------------------------
def packed_index_spd(species_row):
    """Indices (into the 9-slot padded basis) of the real orbitals of one molecule in atom order = the reference's
    packed order [9 per d atom][4 per sp heavy atom][1 per hydrogen] (packd.py:8-85) given d-first sorted rows."""
    nob = norb_of(species_row)
    return np.concatenate([9 * a + np.arange(n) for a, n in enumerate(nob) if n > 0])


def density_from_fock_spd(F, species, nocc, mols=None, want_eig=False):
    """P = 2 C_occ C_occ^T per molecule in the padded layout (diag_d.py:18-150)."""
    nmol, N, _ = F.shape
    D = np.zeros_like(F)
    E = np.zeros((nmol, N))
    V = []
    for m in range(nmol):
        if mols is not None and not mols[m]:
            V.append(None)
            continue
        idx = packed_index_spd(species[m])
        e, v = np.linalg.eigh(F[m][np.ix_(idx, idx)], UPLO="U")
        c = v[:, : int(nocc[m])]
        D[m][np.ix_(idx, idx)] = 2.0 * (c @ c.T)
        E[m, : e.shape[0]] = e
        V.append(v)
    return (D, E, V) if want_eig else (D, E)


def _pair_energy_spd(P, par, mpd, xij, rij, PAi, PBj, Dab, rij_regime=None):
    from .energy import pair_nuclear_energy
    from .integrals import rho0_eff

    w, e_i, e_j = two_center_integrals_spd(P.ni, P.nj, P.idxi, P.idxj, xij, rij, mpd)
    zeta = np.stack([par["zeta_s"], par["zeta_p"], par["zeta_d"]], axis=1)
    di = overlap_spd(P.ni, P.nj, xij, rij, zeta[P.idxi], zeta[P.idxj], rij_regime)
    bA = np.stack([par["beta_s"]] + [par["beta_p"]] * 3 + [par["beta_d"]] * 5, axis=1)
    bsum = bA[P.idxi][:, :, None] + bA[P.idxj][:, None, :]
    E = np.sum(Dab * di * bsum, axis=(1, 2))
    pki = PAi[:, TRI_A, TRI_B] * WEIGHT45
    pkj = PBj[:, TRI_A, TRI_B] * WEIGHT45
    E += np.sum(pki * e_i[:, TRI_B, TRI_A], axis=1) + np.sum(pkj * e_j[:, TRI_B, TRI_A], axis=1)
    E += np.matmul(pki[:, None, :], np.matmul(w, pkj[:, :, None]))[:, 0, 0]
    w4 = w[:, PAIR[:, :, None, None], PAIR[None, None, :, :]]
    E += -0.5 * np.einsum("pml,pmnls,pns->p", Dab, w4, Dab, optimize=True)
    mp = (mpd["dd"], mpd["qq"], mpd["rho0"], mpd["rho1"], mpd["rho2"])
    E += pair_nuclear_energy("PM6_SP", P.ni, P.nj, P.idxi, P.idxj, rij, w[:, 0, 0], par, rho0=rho0_eff(par, mp))
    return E


def hf_gradient_spd(P, par, mpd, Dm, delta=1.0e-4):
    """dE/dR at fixed density by central differences of the pair energy (the reference differentiates the same
    expression with autograd, basics.py:1315-1336; there is no analytic PM6 gradient, anal_grad.py:50-51).  delta = 1e-4 A:
    with 45 x 45 integral blocks the pair energies are large sums and 1e-5 leaves 1e-5 eV/A of round-off (PCl3)."""
    T = Tables.get()
    nmol, molsize = P.nmol, P.molsize
    Db = _blocks9(Dm, nmol, molsize)
    PA = Db[P.atom_molid, P.atom_pos, P.atom_pos]
    mi, ai, aj = P.pair_molid, P.atom_pos[P.idxi], P.atom_pos[P.idxj]
    args = (PA[P.idxi], PA[P.idxj], Db[mi, ai, aj])
    Xij = P.xij * (P.rij * T.a0)[:, None]
    g = np.zeros((P.rij.shape[0], 3))
    for c in range(3):
        Es = {}
        for s in (+1.0, -1.0, +2.0, -2.0):  # five-point stencil: compressed diatomics carry 50-80 eV/A, large 3rd derivative
            X = Xij.copy()
            X[:, c] -= s * delta
            d = np.sqrt(np.sum(X * X, axis=1))
            Es[s] = _pair_energy_spd(P, par, mpd, X / d[:, None], d / T.a0, *args, rij_regime=P.rij)
        g[:, c] = (8.0 * (Es[1.0] - Es[-1.0]) - (Es[2.0] - Es[-2.0])) / (12.0 * delta)
    nat = P.Z.shape[0]
    ga = np.zeros((nat, 3))
    np.add.at(ga, P.idxi, g)
    np.add.at(ga, P.idxj, -g)
    grad = np.zeros((nmol * molsize, 3))
    grad[P.real_atoms] = ga
    return grad.reshape(nmol, molsize, 3)


def single_point_pm6d(species, coordinates, seqm_parameters, P0=None, do_force=True, charges=0):
    """method='PM6' with d-shell elements: the result contract of SURVEY 8(a15) in the 9-slot layout."""
    from .energy import elec_energy, isolated_atom_energy, molecule_sums, pair_nuclear_energy
    from .integrals import atom_multipoles, rho0_eff
    from .parser import parse
    from .scf import run_scf
    from .tables import method_parameters

    T = Tables.get()
    eps = float(seqm_parameters["scf_eps"])
    conv = seqm_parameters.get("scf_converger", [2])
    if seqm_parameters.get("sp2", [False])[0]:
        raise NotImplementedError("oracle: PM6-d with SP2 is not covered")
    P = parse(species, coordinates, charges=charges, outer_cutoff=seqm_parameters.get("pair_outer_cutoff", 1.0e10))
    check_d_first(P)
    par = method_parameters("PM6", P.Z)
    mp = atom_multipoles(P.Z, par)
    mpd = atom_multipoles_spd(P.Z, par, mp)
    hc = build_hcore_spd(P, par, mpd)
    H, w = hc["H"], hc["w"]
    nSH = np.sum(d_shell(P.species), axis=1)
    nHeavy = np.sum((P.species > 1) & ~d_shell(P.species), axis=1)
    D0 = initial_density_spd(P) if P0 is None else np.asarray(P0, dtype=np.float64)
    D, notconv, n_iter = run_scf(
        P, par, H, w, D0, eps, conv, (False,),
        fock_fn=lambda Pm: build_fock_spd(P, par, H, w, Pm),
        density_fn=lambda F, mask: density_from_fock_spd(F, P.species, P.nocc, mask)[0],
        msize=9 * nSH + 4 * nHeavy + 4 * P.nHydro,
    )  # fmt: skip
    F = build_fock_spd(P, par, H, w, D)
    _, e_mo, V = density_from_fock_spd(F, P.species, P.nocc, want_eig=True)
    Eelec = elec_energy(D, F, H)
    EnucAB = pair_nuclear_energy("PM6_SP", P.ni, P.nj, P.idxi, P.idxj, P.rij, w[:, 0, 0], par, rho0=rho0_eff(par, mp))
    Enuc = molecule_sums(EnucAB, P.pair_molid, P.nmol)
    Etot = Eelec + Enuc
    Eiso = molecule_sums(isolated_atom_energy(P.Z, par), P.atom_molid, P.nmol)
    Hf = Etot - Eiso
    if seqm_parameters.get("Hf_flag", True):
        Hf = Hf + molecule_sums(T.eheat[P.Z], P.atom_molid, P.nmol)
    ar = np.arange(P.nmol)
    e_gap = e_mo[ar, P.nocc] - e_mo[ar, P.nocc - 1]
    q = T.tore[P.species] - np.diagonal(D, axis1=1, axis2=2).reshape(P.nmol, P.molsize, 9).sum(axis=2)
    out = dict(Etot=Etot, Hf=Hf, Eelec=Eelec, Enuc=Enuc, Eiso=Eiso, e_mo=e_mo, e_gap=e_gap, dm=D, F=F, H=H, w=w, q=q,
               notconverged=notconv, n_scf_iter=n_iter, molecular_orbitals=V, parsed=P)  # fmt: skip
    if do_force:
        out["force"] = -hf_gradient_spd(P, par, mpd, D)
    return out
