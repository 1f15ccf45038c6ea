// spd_kernels.cuh -- PM6 with d orbitals (method SEQM_PM6_D).
//
//   spd_pair_kernel        one CTA per pair with a d atom ("Y pair"): the (np_i x np_j) block of two-centre integrals
//                          (np = 45 / 10 / 1 orbital products of a d / sp heavy / hydrogen atom) and the beta-scaled
//                          9 x 9 overlap block
//   spd_hcore_kernel       one CTA per molecule: packed Hcore with 9 x 9 / 4 x 4 / 1 x 1 atom blocks
//   spd_fock_kernel        one CTA per molecule: F = H + one-centre (sp parameters + Slater-Condon d integrals) + J - K/2
//   spd_pair_gradient_kernel  one CTA per Y pair: five-point stencil of the d-dependent pair energy, evaluated in the local frame
//
// Replaces (reference file:line, lanl/PYSEQM v2.0.0):
//   two_elec_two_center_int_local_frame_d_orbitals.py:23-4164   local-frame integrals with d orbitals
//   RotationMatrixD.py:5-310, two_elec_two_center_int.py:800-1306   rotation, assembly of w(45,45), e1b / e2a
//   diat_overlapD.py:4-5148   Slater overlaps s, p, d (n <= 4 here)
//   hcore.py:61-179, fock.py:132-347 (`_d_contrib_one_center`, PM6 branch of `_two_center`)
//
// Formulation (not a transcription of the reference's ~9.5 k unrolled lines): every product of two local orbitals is a
// sum of point-charge multipoles (monopole, dipole, quadrupole; Thiel & Voityuk, TCA 81, 391) with tabulated
// coefficients c[kl][source][m] (b.mp_coef, derived by quadrature on the host), so the local integral is
//     (kl | mn) = sum_{s,t,m} c_i[kl][s][m] V[s][t][|m|] c_j[mn][t][m]
// with V the interaction of two unit multipoles of the two atoms (<= 36 inverse square roots each).  The molecular
// frame follows from the 45 x 45 pair-product transform T of the 9 x 9 orbital rotation: w = T L T^t.  The sp x sp
// sub-block is NOT part of that expansion: as in the reference it comes from the MNDO formulas of the sp path
// (pair_core.cuh), already rotated.  Where the reference's numbers deviate from the clean scheme (6-decimal constants
// on the cosine-type terms, a sign on the d-sigma d-delta term of (d, sp) pairs, one overlap element) the deviation
// is reproduced and marked "reference quirk".
#pragma once
#include "pair_helpers.cuh"

#ifndef SEQM_HOSTEMU
#define SPD_NOINLINE __device__ __noinline__
#else
#define SPD_NOINLINE static
#endif
#define SPD_NSRC 7
#define SPD_THREADS 128
#define SPD_SQRT3 1.7320508075688772
#define SPD_SQRT2 1.4142135623730951

// multipole sources: 0 ss/pp monopole (rho0), 1 sp dipole, 2 pp quadrupole, 3 sd quadrupole, 4 pd dipole,
// 5 dd monopole, 6 dd quadrupole
SEQM_HD int spd_src_l(int s) { return (s == 0 || s == 5) ? 0 : ((s == 1 || s == 4) ? 1 : 2); }
// packed lower-triangle product index -> (a >= b)
SEQM_HD void spd_unpack(int kl, int* a, int* b) {
  int x = 0;
  while ((x + 1) * (x + 2) / 2 <= kl) ++x;
  *a = x;
  *b = kl - x * (x + 1) / 2;
}

struct SpdAtom {  // what the pair code needs of one atom
  double D[SPD_NSRC], rho[SPD_NSRC];  // charge separation (quadrupoles: unscaled) and additive term per source
  double zeta[3], beta[3];
  int n[3];  // principal quantum numbers of s, p, d
  int norb, nprod;
};
SEQM_HD void spd_load_atom(const seqm_batch_t& b, int a, bool has_d, SpdAtom& A) {
  const int Z = b.atom_Z[a];
  A.norb = has_d ? 9 : (Z > 1 ? 4 : 1);
  A.nprod = has_d ? 45 : (Z > 1 ? 10 : 1);
  A.D[0] = 0.0; A.rho[0] = par(b, SEQM_P_RHO0, a);
  A.D[1] = par(b, SEQM_P_DD, a); A.rho[1] = par(b, SEQM_P_RHO1, a);
  A.D[2] = par(b, SEQM_P_QQ, a); A.rho[2] = par(b, SEQM_P_RHO2D, a);
  A.D[3] = par(b, SEQM_P_DS, a) * (1.0 / SPD_SQRT2); A.rho[3] = par(b, SEQM_P_RHO5, a);
  A.D[4] = par(b, SEQM_P_DP, a); A.rho[4] = par(b, SEQM_P_RHO4, a);
  A.D[5] = 0.0; A.rho[5] = par(b, SEQM_P_RHO3, a);
  A.D[6] = par(b, SEQM_P_DDQ, a) * (1.0 / SPD_SQRT2); A.rho[6] = par(b, SEQM_P_RHO6, a);
  A.zeta[0] = par(b, SEQM_P_ZS, a); A.zeta[1] = par(b, SEQM_P_ZP, a); A.zeta[2] = par(b, SEQM_P_ZD, a);
  A.beta[0] = par(b, SEQM_P_BS, a); A.beta[1] = par(b, SEQM_P_BP, a); A.beta[2] = par(b, SEQM_P_BD, a);
  A.n[0] = A.n[1] = (int)par(b, SEQM_P_QN, a);
  A.n[2] = (int)par(b, SEQM_P_QND, a);
}

// point charges (q, x, y, z) of the unit multipole (l, |m|) with separation D: 1 (monopole), 2 (dipole), 6 ((2,0)) or
// 4 ((2,1), (2,2)) of them.  N is the compile-time count, so the arrays live in registers.
SEQM_HD int spd_ncharge(int l, int am) { return l == 0 ? 1 : (l == 1 ? 2 : (am == 0 ? 6 : 4)); }
template <int N>
SEQM_D void spd_configuration(int am, double D, double* q, double* x, double* y, double* z) {
#pragma unroll
  for (int k = 0; k < N; ++k) q[k] = x[k] = y[k] = z[k] = 0.0;
  if (N == 1) {
    q[0] = 1.0;
  } else if (N == 2) {
    q[0] = 0.5; q[1] = -0.5;
    if (am == 0) { z[0] = D; z[1] = -D; } else { x[0] = D; x[1] = -D; }
  } else if (N == 6) {  // Q~zx + 1/2 Q~xy: +1/4 at z = +-sqrt2 D, -1/8 at x = +-sqrt2 D and y = +-sqrt2 D
    const double r2 = SPD_SQRT2 * D;
    q[0] = q[1] = 0.25; z[0] = r2; z[1] = -r2;
    q[2] = q[3] = -0.125; x[2] = r2; x[3] = -r2;
    q[4] = q[5] = -0.125; y[4] = r2; y[5] = -r2;
  } else if (am == 1) {  // (2,1): +-1/4 at (+-D, 0, +-D)
    q[0] = 0.25; x[0] = D; z[0] = D;
    q[1] = -0.25; x[1] = D; z[1] = -D;
    q[2] = -0.25; x[2] = -D; z[2] = D;
    q[3] = 0.25; x[3] = -D; z[3] = -D;
  } else {  // (2,2): +1/4 at x = +-sqrt2 D, -1/4 at y = +-sqrt2 D
    const double r2 = SPD_SQRT2 * D;
    q[0] = q[1] = 0.25; x[0] = r2; x[1] = -r2;
    q[2] = q[3] = -0.25; y[2] = r2; y[3] = -r2;
  }
}
template <int NA, int NB>
SEQM_D double spd_interaction_t(int am, double Da, double Db, double add, double r) {
  double qa[NA], xa[NA], ya[NA], za[NA], qb[NB], xb[NB], yb[NB], zb[NB];
  spd_configuration<NA>(am, Da, qa, xa, ya, za);
  spd_configuration<NB>(am, Db, qb, xb, yb, zb);
  double tot = 0.0;
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const double dx = xa[i] - xb[j], dy = ya[i] - yb[j], dz = za[i] - zb[j] + r;
      tot += qa[i] * qb[j] * seqm_rsqrt(dx * dx + dy * dy + dz * dz + add);
    }
  return tot;
}
// interaction (eV) of multipole (ls, am) of atom i at the origin with (lt, am) of atom j at z = -r
SPD_NOINLINE double spd_interaction(int ls, int lt, int am, double Da, double Db, double add, double r) {
  const int na = spd_ncharge(ls, am), nb = spd_ncharge(lt, am);
  double t;
#define SPD_CASE(A_, B_) if (na == A_ && nb == B_) t = spd_interaction_t<A_, B_>(am, Da, Db, add, r); else
  SPD_CASE(1, 1) SPD_CASE(1, 2) SPD_CASE(1, 4) SPD_CASE(1, 6) SPD_CASE(2, 1) SPD_CASE(2, 2) SPD_CASE(2, 4) SPD_CASE(2, 6)
  SPD_CASE(4, 1) SPD_CASE(4, 2) SPD_CASE(4, 4) SPD_CASE(4, 6) SPD_CASE(6, 1) SPD_CASE(6, 2) SPD_CASE(6, 4) SPD_CASE(6, 6)
  t = 0.0;
#undef SPD_CASE
  return SEQM_EV * t;
}

// local axes of the reference (RotationMatrixD.py:11-45, MOPAC rotmat) for the unit vector w (local z):
// w = (ca sb, sa sb, cb) -> u = (ca cb, sa cb, -sb), v = (-sa, ca, 0); w along +-z: u = (1, 0, 0), v = (0, +-1, 0)
SEQM_HD void spd_local_axes(const double w[3], double u[3], double v[3], double* ca_o, double* sb_o, double* cb_o) {
  const double xy = sqrt(w[0] * w[0] + w[1] * w[1]);
  double ca, sa, cb, sb;
  if (xy >= 1.0e-10) {
    ca = w[0] / xy; sa = w[1] / xy; cb = w[2]; sb = xy;
  } else {
    const double sg = (w[2] > 0.0) ? 1.0 : ((w[2] < 0.0) ? -1.0 : 0.0);
    ca = sg; sa = 0.0; cb = sg; sb = 0.0;
  }
  u[0] = ca * cb; u[1] = sa * cb; u[2] = -sb;
  v[0] = -sa; v[1] = ca; v[2] = 0.0;
  if (ca_o) { *ca_o = ca; *sb_o = sb; *cb_o = cb; }
}
// R[a][b] (9 x 9, row-major in R81): molecular orbital a = sum_b R[a][b] local orbital b.
// molecular order s, px, py, pz, d(x2-y2), d(xz), d(z2), d(yz), d(xy); local order s, p(z, x, y), d(z2, xz, yz, x2-y2, xy).
// The d block expands the quadratic form Q of each local d function (in molecular components) in the molecular forms.
SEQM_HD double spd_dform(int bq, int i, int j, const double u[3], const double v[3], const double w[3]) {
  if (bq == 0) return (2.0 * w[i] * w[j] - u[i] * u[j] - v[i] * v[j]) * (1.0 / SPD_SQRT3);
  if (bq == 1) return u[i] * w[j] + w[i] * u[j];
  if (bq == 2) return v[i] * w[j] + w[i] * v[j];
  if (bq == 3) return u[i] * u[j] - v[i] * v[j];
  return u[i] * v[j] + v[i] * u[j];
}
SEQM_HD double spd_rotation_element(int a, int bq, const double u[3], const double v[3], const double w[3]) {
  if (a == 0 || bq == 0) return (a == bq) ? 1.0 : 0.0;
  if (a < 4) return (bq == 1) ? w[a - 1] : (bq == 2 ? u[a - 1] : (bq == 3 ? v[a - 1] : 0.0));
  if (bq < 4) return 0.0;
  const int d = bq - 4;
  if (a == 4) return 0.5 * (spd_dform(d, 0, 0, u, v, w) - spd_dform(d, 1, 1, u, v, w));
  if (a == 5) return spd_dform(d, 0, 2, u, v, w);
  if (a == 6) return 0.5 * SPD_SQRT3 * spd_dform(d, 2, 2, u, v, w);
  if (a == 7) return spd_dform(d, 1, 2, u, v, w);
  return spd_dform(d, 0, 1, u, v, w);
}
// ---- Slater overlaps with d functions -------------------------------------------------------------------------------
// 14 local overlaps (la, lb, m): the polynomial of the prolate-spheroidal integrand times its angular constant comes from
// b.ovl_poly[na-1][nb-1][kind][k][l] (host-built, pm6d_tables.py), auxiliary integrals A_k, B_l as in the sp path.
SEQM_HD void spd_kind(int kind, int* la, int* lb, int* m) {
  const int LA[14] = {0, 1, 0, 1, 1, 2, 0, 2, 1, 2, 1, 2, 2, 2};
  const int LB[14] = {0, 0, 1, 1, 1, 0, 2, 1, 2, 1, 2, 2, 2, 2};
  const int MM[14] = {0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 2};
  *la = LA[kind]; *lb = LB[kind]; *m = MM[kind];
}
// B_k(x) with the regime (recurrence | four-term series | x = 0 limit, diat_overlapD.py:5222-5370) selected by xr: the
// stencil of the gradient kernel freezes the choice at the undisplaced geometry, because the truncated series and the
// recurrence differ by ~1e-7 at |x| = 0.5 and a finite difference must not straddle that step (the reference's
// autograd differentiates inside the branch the geometry is in).
SEQM_HD void spd_aux_B(double x, double xr, int kmax, double* B) {
  const double ax = fabs(xr);
  if (ax > 0.5) {
    const double tx = exp(x) / x, tmx = -(exp(-x) / x);
    B[0] = tx + tmx;
    for (int k = 1; k <= kmax; ++k) B[k] = ((k & 1) ? (tmx - tx) : (tx + tmx)) + (double)k * B[k - 1] / x;
  } else if (ax > 1.0e-6) {
    const double x2 = x * x;
    for (int k = 0; k <= kmax; ++k) {
      if ((k & 1) == 0)
        B[k] = 2.0 / (k + 1.0) + x2 / (k + 3.0) + x2 * x2 / ((k + 5.0) * 12.0) + x2 * x2 * x2 / ((k + 7.0) * 360.0);
      else
        B[k] = (-2.0 / (k + 2.0)) * x - x2 * x / ((k + 4.0) * 3.0) - x2 * x2 * x / ((k + 6.0) * 60.0);
    }
  } else {
    for (int k = 0; k <= kmax; ++k) B[k] = (k & 1) ? 0.0 : 2.0 / (k + 1.0);
  }
}
SPD_NOINLINE double spd_local_overlap(const seqm_batch_t& b, const SpdAtom& A, const SpdAtom& B, int kind, double r, double r_regime) {
  int la, lb, m;
  spd_kind(kind, &la, &lb, &m);
  const int na = A.n[la], nb = B.n[lb];
  if (na < 1 || na > 4 || nb < 1 || nb > 4) return 0.0;
  const double za = A.zeta[la], zb = B.zeta[lb];
  double Ak[10], Bk[10];
  const int kmax = 8;
  aux_A((0.5 * (za + zb)) * r, kmax, Ak);
  spd_aux_B((0.5 * (za - zb)) * r, (0.5 * (za - zb)) * r_regime, kmax, Bk);
  const double* poly = b.ovl_poly + ((long long)((na - 1) * 4 + (nb - 1)) * 14 + kind) * 81;
  double tot = 0.0;
  for (int k = 0; k <= kmax; ++k)
    for (int l = 0; l <= kmax; ++l) {
      const double c = poly[k * 9 + l];
      if (c != 0.0) tot += c * (Ak[k] * Bk[l]);
    }
  double fa = 1.0, fb = 1.0;
  for (int k = 2; k <= 2 * na; ++k) fa *= k;
  for (int k = 2; k <= 2 * nb; ++k) fb *= k;
  const double pre = pow(2.0 * za, na + 0.5) * pow(2.0 * zb, nb + 0.5) / sqrt(fa * fb);
  return pre * ipow(0.5 * r, na + nb + 1) * tot;
}

// Sparse form of the multipole coefficient tables in shared memory: per orbital product at most 4 (here: 3) non-zero
// (source, m) entries.  cv[side][45][4] values, cc[side][45][4] codes s * 5 + m5 (-1 = unused); side 0 = atom i (the
// "yx" table when the partner is an sp-only heavy atom: reference quirk), side 1 = atom j.  All threads call it.
SEQM_D void spd_stage_coefficients(const seqm_batch_t& b, bool yx, double* cv, int* cc) {
  const double* tab_i = yx ? b.mp_coef_yx : b.mp_coef;
  for (int t = threadIdx.x; t < 90; t += blockDim.x) {
    const int side = t / 45, kl = t % 45;
    const double* tab = side ? b.mp_coef : tab_i;
    int cnt = 0;
    for (int q = 0; q < SPD_NSRC * 5 && cnt < 4; ++q) {
      const double cval = tab[kl * SPD_NSRC * 5 + q];
      if (cval != 0.0) {
        cv[side * 180 + kl * 4 + cnt] = cval;
        cc[side * 180 + kl * 4 + cnt] = q;
        ++cnt;
      }
    }
    for (; cnt < 4; ++cnt) cc[side * 180 + kl * 4 + cnt] = -1;
  }
}
// local integral (kl | mn) from the staged coefficients and the unit multipole interactions V[s][t][|m|]
SEQM_D double spd_local_integral(int kl, int mn, const double* cv, const int* cc, const double* V) {
  double L = 0.0;
  for (int x = 0; x < 4; ++x) {
    const int ca_code = cc[kl * 4 + x];
    if (ca_code < 0) break;
    const int s = ca_code / 5, m5 = ca_code % 5;
    const int am = (m5 == 0) ? 0 : (m5 < 3 ? 1 : 2);
    double in = 0.0;
    for (int y = 0; y < 4; ++y) {
      const int cb_code = cc[180 + mn * 4 + y];
      if (cb_code < 0) break;
      if (cb_code % 5 == m5) in += V[(s * SPD_NSRC + cb_code / 5) * 3 + am] * cv[180 + mn * 4 + y];
    }
    L += cv[kl * 4 + x] * in;
  }
  // reference quirk: (d-sigma p-pi(y) | p-pi(y) s) carries -0.577350 where its neighbours carry -1/sqrt3
  if (kl == 13 && mn == 6) L += (-0.577350 + 1.0 / SPD_SQRT3) * V[(4 * SPD_NSRC + 1) * 3 + 1];
  return L;
}

// ---- the block of one Y pair in shared memory ---------------------------------------------------------------------------
// layout of the CTA's dynamic shared memory (doubles)
#define SPD_OFF_T 0       /* 45 x 45 pair-product transform */
#define SPD_OFF_L 2025    /* local integrals, then the rotated block w (np_i x np_j, row stride np_j) */
#define SPD_OFF_TMP 4050  /* half-rotated block */
#define SPD_OFF_R 6075    /* 9 x 9 orbital rotation */
#define SPD_OFF_V 6156    /* 7 x 7 x 3 unit multipole interactions */
#define SPD_OFF_S 6303    /* 9 x 9 local overlaps, then beta-scaled molecular-frame block */
#define SPD_OFF_CV 6384   /* staged sparse coefficients: values [2][45][4] */
#define SPD_OFF_CC 6744   /* and codes (ints) */
#define SPD_SMEM_DOUBLES 6924

// Fills sm[SPD_OFF_L ..] with w (np_i x np_j) and sm[SPD_OFF_S ..] with hab (9 x 9, beta-scaled overlaps) for atoms
// i (a d atom) and j at positions Ri, Rj (Angstrom).  wsp: the pair's 10 x 10 block of the sp path at this geometry.
// All threads of the CTA call it; it ends with a barrier.
SPD_NOINLINE void spd_pair_block(const seqm_batch_t& b, int i, int j, bool dj, const double* Ri, const double* Rj,
                           const double* wsp, double* sm) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  __shared__ SpdAtom sAB[2];  // the two atoms' parameters: one copy per CTA instead of a stack copy per thread
  if (tid == 0) spd_load_atom(b, i, true, sAB[0]);
  if (tid == 1 || nthr == 1) spd_load_atom(b, j, dj, sAB[1]);
  SEQM_SYNC();
  const SpdAtom& A = sAB[0];
  const SpdAtom& B = sAB[1];
  const int npi = A.nprod, npj = B.nprod;
  double e[3] = {Rj[0] - Ri[0], Rj[1] - Ri[1], Rj[2] - Ri[2]};
  const double dist = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  for (int c = 0; c < 3; ++c) e[c] /= dist;
  const double r = dist * (1.0 / SEQM_A0);
  double* T = sm + SPD_OFF_T;
  double* L = sm + SPD_OFF_L;
  double* tmp = sm + SPD_OFF_TMP;
  double* R = sm + SPD_OFF_R;
  double* V = sm + SPD_OFF_V;
  double* S = sm + SPD_OFF_S;
  if (pair_cut(b, r)) {
    for (int t = tid; t < npi * npj; t += nthr) L[t] = 0.0;
    for (int t = tid; t < 81; t += nthr) S[t] = 0.0;
    SEQM_SYNC();
    return;
  }
  // local z axis from j to i (MOPAC convention, two_elec_two_center_int.py:287-290)
  const double w3[3] = {-e[0], -e[1], -e[2]};
  double u3[3], v3[3];
  spd_local_axes(w3, u3, v3, nullptr, nullptr, nullptr);
  for (int t = tid; t < 81; t += nthr) R[t] = spd_rotation_element(t / 9, t % 9, u3, v3, w3);
  spd_stage_coefficients(b, !dj && B.norb == 4, sm + SPD_OFF_CV, reinterpret_cast<int*>(sm + SPD_OFF_CC));
  // unit multipole interactions
  for (int t = tid; t < SPD_NSRC * SPD_NSRC * 3; t += nthr) {
    const int s = t / (SPD_NSRC * 3), tt = (t / 3) % SPD_NSRC, am = t % 3;
    const int ls = spd_src_l(s), lt = spd_src_l(tt);
    double val = 0.0;
    const bool have = (tt < 3 || dj) && (tt == 0 || B.norb > 1) && (s >= 3 || tt >= 3);  // sources in play
    if (am <= ls && am <= lt && have) {
      const double add = (A.rho[s] + B.rho[tt]) * (A.rho[s] + B.rho[tt]);
      val = spd_interaction(ls, lt, am, A.D[s], B.D[tt], add, r);
    }
    V[t] = val;
  }
  // local overlaps (frame of the integrals: j sits at -z, hence the parity factors)
  for (int t = tid; t < 81; t += nthr) S[t] = 0.0;
  SEQM_SYNC();
  if (r <= SEQM_OVERLAP_CUTOFF) {
    for (int kind = tid; kind < 14; kind += nthr) {
      int la, lb, m;
      spd_kind(kind, &la, &lb, &m);
      const int oa = (la == 0) ? 0 : (la == 1 ? 1 : 4), ob = (lb == 0) ? 0 : (lb == 1 ? 1 : 4);
      if (oa >= A.norb || ob >= B.norb) continue;
      const double sv = spd_local_overlap(b, A, B, kind, r, r);
      // local orbital of (l, m): p: sigma 1, pi 2,3 ; d: sigma 4, pi 5,6, delta 7,8.  Reflection z -> -z: (-1)^(l+m)
      const double sg = (((la + m) & 1) ? -1.0 : 1.0) * (((lb + m) & 1) ? -1.0 : 1.0);
      const int ia = (m == 0) ? oa : (la == 1 ? 2 : (m == 1 ? 5 : 7));
      const int ib = (m == 0) ? ob : (lb == 1 ? 2 : (m == 1 ? 5 : 7));
      S[ia * 9 + ib] = sg * sv;
      if (m > 0) S[(ia + 1) * 9 + ib + 1] = sg * sv;
    }
  }
  // pair-product transform
  for (int t = tid; t < 2025; t += nthr) {
    const int kl = t / 45, mn = t % 45;
    int a, c, bb, d;
    spd_unpack(kl, &a, &bb);
    spd_unpack(mn, &c, &d);
    T[t] = (c == d) ? R[a * 9 + c] * R[bb * 9 + c] : R[a * 9 + c] * R[bb * 9 + d] + R[a * 9 + d] * R[bb * 9 + c];
  }
  SEQM_SYNC();
  // local integrals
  const double* cv = sm + SPD_OFF_CV;
  const int* cc = reinterpret_cast<const int*>(sm + SPD_OFF_CC);
  for (int t = tid; t < npi * npj; t += nthr) {
    const int kl = t / npj, mn = t % npj;
    L[t] = (kl >= 10 || mn >= 10) ? spd_local_integral(kl, mn, cv, cc, V) : 0.0;
  }
  SEQM_SYNC();
  // tmp[k'l'][mn] = sum_{m'n'} L[k'l'][m'n'] T[mn][m'n']
  for (int t = tid; t < npi * npj; t += nthr) {
    const int kl = t / npj, mn = t % npj;
    double acc = 0.0;
    for (int q = 0; q < npj; ++q) acc += L[kl * npj + q] * T[mn * 45 + q];
    tmp[t] = acc;
  }
  SEQM_SYNC();
  // w[kl][mn] = sum_{k'l'} T[kl][k'l'] tmp[k'l'][mn]; the sp x sp sub-block comes from the sp path
  for (int t = tid; t < npi * npj; t += nthr) {
    const int kl = t / npj, mn = t % npj;
    double acc;
    if (kl < 10 && mn < 10) {
      acc = wsp[kl * 10 + mn];
    } else {
      acc = 0.0;
      for (int q = 0; q < npi; ++q) acc += T[kl * 45 + q] * tmp[q * npj + mn];
    }
    L[t] = acc;
  }
  // overlaps to the molecular frame: S_mol = R S_loc R^t (tmp is free again after the barrier)
  SEQM_SYNC();
  for (int t = tid; t < 81; t += nthr) {
    const int a = t / 9, bq = t % 9;
    double acc = 0.0;
    for (int q = 0; q < 9; ++q) acc += S[a * 9 + q] * R[bq * 9 + q];
    tmp[t] = acc;
  }
  SEQM_SYNC();
  double fix = 0.0;
  if (dj) {  // reference quirk: the delta-bar term of the (d_yz, d_xy) element has the wrong sign, diat_overlapD.py:5104-5116
    double ca, sb, cb, uu[3], vv[3];
    spd_local_axes(e, uu, vv, &ca, &sb, &cb);
    fix = 2.0 * S[7 * 9 + 7] * ca * sb * cb * (2.0 * ca * ca - 1.0);
  }
  SEQM_SYNC();
  for (int t = tid; t < 81; t += nthr) {
    const int a = t / 9, bq = t % 9;
    double acc = 0.0;
    for (int q = 0; q < 9; ++q) acc += R[a * 9 + q] * tmp[q * 9 + bq];
    if ((a == 7 && bq == 8) || (a == 8 && bq == 7)) acc += fix;
    const double ba = A.beta[a == 0 ? 0 : (a < 4 ? 1 : 2)], bb = B.beta[bq == 0 ? 0 : (bq < 4 ? 1 : 2)];
    S[t] = (a < A.norb && bq < B.norb) ? acc * 0.5 * (ba + bb) : 0.0;
  }
  SEQM_SYNC();
}

SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS(SPD_THREADS) spd_pair_kernel(seqm_batch_t b, const double* __restrict__ xyz,
                                                                 const double* __restrict__ w10, double* __restrict__ wd,
                                                                 double* __restrict__ hab_d) {
  SEQM_DYN_SMEM(double, sm);
  const int slot = blockIdx.x;
  const int p = b.ypairs[slot];
  const int i = b.pair_i[p], j = b.pair_j[p];
  const int mol = b.atom_mol[i];
  const bool dj = (j - b.mol_atom0[mol]) < b.mol_nsh[mol];
  spd_pair_block(b, i, j, dj, xyz + 3 * (long long)i, xyz + 3 * (long long)j, w10 + (long long)p * 100, sm);
  const long long o0 = b.pair_wd0[p], cnt = b.pair_wd0[p + 1] - o0;
  for (int t = threadIdx.x; t < cnt; t += blockDim.x) wd[o0 + t] = sm[SPD_OFF_L + t];
  for (int t = threadIdx.x; t < 81; t += blockDim.x) hab_d[(long long)slot * 81 + t] = sm[SPD_OFF_S + t];
}

// (kl on atom a | mn on atom o) of one molecule, whichever array holds the pair
struct SpdPairRef {
  const double* base;
  int sk, sm;  // strides of kl and mn
};
SEQM_HD SpdPairRef spd_pair_ref(const seqm_batch_t& b, const MolView& v, const double* w10, int a, int o) {
  const bool first = a < o;
  const int lo = first ? a : o, hi = first ? o : a;
  const int p = v.p0 + pair_local(v, lo, hi);
  SpdPairRef r;
  if (lo < v.nsh) {  // Y pair: ragged block (np_lo x np_hi)
    const int nhi = prod_cnt(v, hi);
    r.base = b.wd + b.pair_wd0[p];
    r.sk = first ? nhi : 1;
    r.sm = first ? 1 : nhi;
  } else {
    r.base = w10 + (long long)p * 100;
    r.sk = first ? 10 : 1;
    r.sm = first ? 1 : 10;
  }
  return r;
}

// One CTA per molecule: packed, fully symmetric Hcore (hcore.py:124-173 with 9 x 9 blocks).
SEQM_GLOBAL void spd_hcore_kernel(seqm_batch_t b, const double* __restrict__ w10, const double* __restrict__ hab,
                                  double* __restrict__ H) {
  const MolView v = mol_view(b, b.mol_order[blockIdx.x]);
  double* Hm = H + v.mat0;
  const int n = v.n;
  for (int t = threadIdx.x; t < v.npair * 81; t += blockDim.x) {
    const int pl = t / 81, mu = (t / 9) % 9, nu = t % 9;
    const int p = v.p0 + pl;
    const int i = b.pair_i[p] - v.a0, j = b.pair_j[p] - v.a0;
    if (mu >= orb_cnt(v, i) || nu >= orb_cnt(v, j)) continue;
    double h;
    if (i < v.nsh) h = b.hab_d[(long long)b.ypair_slot[p] * 81 + mu * 9 + nu];
    else h = hab[(long long)p * 16 + mu * 4 + nu];
    const int r = orb_off(v, i) + mu, c = orb_off(v, j) + nu;
    Hm[r * n + c] = h;
    Hm[c * n + r] = h;
  }
  for (int t = threadIdx.x; t < v.na * 45; t += blockDim.x) {
    const int a = t / 45, kl = t % 45;
    if (kl >= prod_cnt(v, a)) continue;
    int mu, nu;
    spd_unpack(kl, &mu, &nu);
    const int ga = v.a0 + a;
    double acc = 0.0;
    if (mu == nu) acc = (mu == 0) ? par(b, SEQM_P_USS, ga) : (mu < 4 ? par(b, SEQM_P_UPP, ga) : par(b, SEQM_P_UDD, ga));
    for (int o = 0; o < v.na; ++o) {
      if (o == a) continue;
      const SpdPairRef pr = spd_pair_ref(b, v, w10, a, o);
      acc -= par(b, SEQM_P_TORE, v.a0 + o) * pr.base[kl * pr.sk];
    }
    const int oa = orb_off(v, a);
    Hm[(oa + mu) * n + oa + nu] = acc;
    Hm[(oa + nu) * n + oa + mu] = acc;
  }
}

// One CTA per molecule: F = Hcore + G(P).  shared: sP[n*n] | pk[na*45] (weighted packed diagonal blocks)
SEQM_GLOBAL void spd_fock_kernel(seqm_batch_t b, const double* __restrict__ P, const double* __restrict__ H,
                                 const double* __restrict__ w10, double* __restrict__ F, const int32_t* __restrict__ active) {
  const int m = b.mol_order[blockIdx.x];
  if (active && !active[m]) return;
  const MolView v = mol_view(b, m);
  const int n = v.n;
  SEQM_DYN_SMEM(double, sP);
  double* pk = sP + n * n;
  const double* Pm = P + v.mat0;
  const double* Hm = H + v.mat0;
  double* Fm = F + v.mat0;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) sP[t] = Pm[t];
  SEQM_SYNC();
  for (int t = threadIdx.x; t < v.na * 45; t += blockDim.x) {
    const int a = t / 45, kl = t % 45;
    double x = 0.0;
    if (kl < prod_cnt(v, a)) {
      int mu, nu;
      spd_unpack(kl, &mu, &nu);
      const int oa = orb_off(v, a);
      x = sP[(oa + mu) * n + oa + nu] * (mu == nu ? 1.0 : 2.0);
    }
    pk[t] = x;
  }
  SEQM_SYNC();
  // exchange blocks: F_AB[mu,la] = H_AB[mu,la] - 1/2 sum_{nu in A, sg in B} P_AB[nu,sg] (mu nu | la sg).
  // Work items: the sp x sp corner (mu, la < 4) of every pair, then the remaining elements of the pairs with a d atom --
  // those are the first nY pairs of the molecule (pairs are ordered by their first atom and d atoms come first) --
  // instead of 81 slots for every pair (in an organic molecule nine tenths of those would be idle).
  const int nY = v.nsh * (v.na - 1) - v.nsh * (v.nsh - 1) / 2;
  const int n_sp = v.npair * 16;
  for (int t = threadIdx.x; t < n_sp + nY * 81; t += blockDim.x) {
    int pl, mu, la;
    if (t < n_sp) {
      pl = t >> 4;
      mu = (t >> 2) & 3;
      la = t & 3;
    } else {
      const int u = t - n_sp;
      pl = u / 81;
      mu = (u / 9) % 9;
      la = u % 9;
      if (mu < 4 && la < 4) continue;
    }
    const int p = v.p0 + pl;
    const int i = b.pair_i[p] - v.a0, j = b.pair_j[p] - v.a0;
    const int ni = orb_cnt(v, i), nj = orb_cnt(v, j);
    if (mu >= ni || la >= nj) continue;
    const int oi = orb_off(v, i), oj = orb_off(v, j);
    const SpdPairRef pr = spd_pair_ref(b, v, w10, i, j);
    double k = 0.0;
    for (int nu = 0; nu < ni; ++nu)
      for (int sg = 0; sg < nj; ++sg) k += sP[(oi + nu) * n + oj + sg] * pr.base[pack2(mu, nu) * pr.sk + pack2(la, sg) * pr.sm];
    const int r = oi + mu, c = oj + la;
    const double f = Hm[r * n + c] - 0.5 * k;
    Fm[r * n + c] = f;
    Fm[c * n + r] = f;
  }
  // diagonal blocks
  for (int t = threadIdx.x; t < v.na * 45; t += blockDim.x) {
    const int a = t / 45, kl = t % 45;
    if (kl >= prod_cnt(v, a)) continue;
    int mu, nu;
    spd_unpack(kl, &mu, &nu);
    const int oa = orb_off(v, a), ga = v.a0 + a;
    double g = 0.0;
    if (mu < 4) {  // sp one-centre terms (fock.py:187-231)
      const double gss = par(b, SEQM_P_GSS, ga), gsp = par(b, SEQM_P_GSP, ga), gpp = par(b, SEQM_P_GPP, ga);
      const double gp2 = par(b, SEQM_P_GP2, ga), hsp = par(b, SEQM_P_HSP, ga);
      const double Pss = sP[oa * n + oa];
      double Ppt = 0.0;
      if (a < v.nheavy) Ppt = sP[(oa + 1) * n + oa + 1] + sP[(oa + 2) * n + oa + 2] + sP[(oa + 3) * n + oa + 3];
      if (mu == 0)
        g = 0.5 * Pss * gss + Ppt * (gsp - 0.5 * hsp);
      else if (nu == 0)
        g = sP[oa * n + oa + mu] * (1.5 * hsp - 0.5 * gsp);
      else if (mu == nu) {
        const double Pk = sP[(oa + mu) * n + oa + mu];
        g = Pss * (gsp - 0.5 * hsp) + 0.5 * Pk * gpp + (Ppt - Pk) * (1.25 * gp2 - 0.25 * gpp);
      } else
        g = sP[(oa + nu) * n + oa + mu] * (0.75 * gpp - 1.25 * gp2);
    }
    if (a < v.nsh) {  // one-centre integrals containing a d orbital (fock.py:237-253): J - K/2 with the element's table
      const int Z = b.atom_Z[ga];
      const double* I = b.onecenter_d + (long long)(Z < b.oc_dim ? Z : 0) * 2025;
      double jj = 0.0;
      for (int mn = 0; mn < 45; ++mn) jj += I[kl * 45 + mn] * pk[a * 45 + mn];
      double kk = 0.0;
      for (int la = 0; la < 9; ++la)
        for (int sg = 0; sg < 9; ++sg) kk += I[pack2(mu, la) * 45 + pack2(nu, sg)] * sP[(oa + la) * n + oa + sg];
      g += jj - 0.5 * kk;
    }
    for (int o = 0; o < v.na; ++o) {  // Coulomb from every other atom
      if (o == a) continue;
      const SpdPairRef pr = spd_pair_ref(b, v, w10, a, o);
      const int no = prod_cnt(v, o);
      double j = 0.0;
      for (int mn = 0; mn < no; ++mn) j += pk[o * 45 + mn] * pr.base[kl * pr.sk + mn * pr.sm];
      g += j;
    }
    const double f = Hm[(oa + mu) * n + oa + nu] + g;
    Fm[(oa + mu) * n + oa + nu] = f;
    Fm[(oa + nu) * n + oa + mu] = f;
  }
}

// ---- gradient of the Y pairs -----------------------------------------------------------------------------------------------
// dE_pair/dR_i of the part of a Y pair's energy that the sp gradient kernel does not cover: the resonance term over
// the 9 x 9 overlap block and every two-electron / core-attraction term that contains a d orbital.  The reference has
// no analytic PM6 gradient either (anal_grad.py:50-51: autograd through the same expression); here a five-point
// stencil (delta = 1e-4 Angstrom, error O(delta^4)) runs through an evaluation that needs no 45 x 45 transform at all:
// the energy is frame invariant, so the three density blocks are rotated INTO the local frame (9 x 9 x 9 products)
// and contracted with the local integrals, which are assembled on the fly from the unit multipole interactions.
// shared layout (doubles): Pi 81 | Pj 81 | Pij 81 | Ai 81 | Aj 81 | Aij 81 | tmp 243 | R 81 | V 147 | S 81 | pk 90 | red 40
#define SPG_PI 0
#define SPG_PJ 81
#define SPG_PIJ 162
#define SPG_AI 243
#define SPG_AJ 324
#define SPG_AIJ 405
#define SPG_TMP 486
#define SPG_R 729
#define SPG_V 810
#define SPG_S 957
#define SPG_PK 1038
#define SPG_RED 1128
#define SPG_CV 1176    /* sparse multipole coefficients: values [2][45][4] (atom i table, atom j table) */
#define SPG_CC 1536    /* their codes s * 5 + m5 (-1 = unused), ints stored in [2][45][4] int slots = 180 doubles */
#define SPG_AB 1716    /* auxiliary integrals A_k, B_k (k <= 8) of the 9 (l_a, l_b) zeta combinations: [9][18] */
#define SPG_PART 1878  /* partial sums of the overlap polynomials [14][9] */
#define SPG_PRE 2004   /* geometry-independent overlap prefactors [14] */
#define SPG_SMEM_DOUBLES 2020

SPD_NOINLINE double spd_pair_energy_local(const seqm_batch_t& b, int i, int j, bool dj, const SpdAtom& A, const SpdAtom& B,
                                          const double* Ri, const double* Rj, double r_regime, double* sm) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  double e[3] = {Rj[0] - Ri[0], Rj[1] - Ri[1], Rj[2] - Ri[2]};
  const double dist = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  for (int c = 0; c < 3; ++c) e[c] /= dist;
  const double r = dist * (1.0 / SEQM_A0);
  if (pair_cut(b, r)) return 0.0;
  double* R = sm + SPG_R;
  double* V = sm + SPG_V;
  double* S = sm + SPG_S;
  double* pk = sm + SPG_PK;
  double* tmp = sm + SPG_TMP;
  const int npi = A.nprod, npj = B.nprod;
  const double w3[3] = {-e[0], -e[1], -e[2]};
  double u3[3], v3[3];
  spd_local_axes(w3, u3, v3, nullptr, nullptr, nullptr);
  for (int t = tid; t < 81; t += nthr) R[t] = spd_rotation_element(t / 9, t % 9, u3, v3, w3);
  for (int t = tid; t < SPD_NSRC * SPD_NSRC * 3; t += nthr) {
    const int s = t / (SPD_NSRC * 3), tt = (t / 3) % SPD_NSRC, am = t % 3;
    const int ls = spd_src_l(s), lt = spd_src_l(tt);
    double val = 0.0;
    const bool have = (tt < 3 || dj) && (tt == 0 || B.norb > 1) && (s >= 3 || tt >= 3);  // sp x sp sources: not needed here
    if (am <= ls && am <= lt && have) {
      const double add = (A.rho[s] + B.rho[tt]) * (A.rho[s] + B.rho[tt]);
      val = spd_interaction(ls, lt, am, A.D[s], B.D[tt], add, r);
    }
    V[t] = val;
  }
  for (int t = tid; t < 81; t += nthr) S[t] = 0.0;
  // auxiliary integrals of the nine (l_a, l_b) exponent combinations, one thread each
  double* AB = sm + SPG_AB;
  for (int c = tid; c < 9; c += nthr) {
    const int la = c / 3, lb = c % 3;
    const double za = A.zeta[la], zb = B.zeta[lb];
    if (za > 0.0 && zb > 0.0) {
      aux_A((0.5 * (za + zb)) * r, 8, AB + c * 18);
      spd_aux_B((0.5 * (za - zb)) * r, (0.5 * (za - zb)) * r_regime, 8, AB + c * 18 + 9);
    } else {
      for (int k = 0; k < 18; ++k) AB[c * 18 + k] = 0.0;
    }
  }
  SEQM_SYNC();
  if (r <= SEQM_OVERLAP_CUTOFF) {
    // 14 kinds x 9 rows of the (xi, eta) polynomial: one (kind, k) row per thread, then 14 threads add their 9 rows
    double* part = sm + SPG_PART;
    for (int t = tid; t < 126; t += nthr) {
      const int kind = t / 9, k = t % 9;
      int la, lb, m;
      spd_kind(kind, &la, &lb, &m);
      const int na = A.n[la], nb = B.n[lb];
      double acc = 0.0;
      if (na >= 1 && na <= 4 && nb >= 1 && nb <= 4) {
        const double* poly = b.ovl_poly + ((long long)((na - 1) * 4 + (nb - 1)) * 14 + kind) * 81 + k * 9;
        const double* Bk = AB + (la * 3 + lb) * 18 + 9;
        for (int l = 0; l < 9; ++l) acc += poly[l] * Bk[l];
        acc *= AB[(la * 3 + lb) * 18 + k];
      }
      part[t] = acc;
    }
    SEQM_SYNC();
    for (int kind = tid; kind < 14; kind += nthr) {
      int la, lb, m;
      spd_kind(kind, &la, &lb, &m);
      const int oa = (la == 0) ? 0 : (la == 1 ? 1 : 4), ob = (lb == 0) ? 0 : (lb == 1 ? 1 : 4);
      if (oa >= A.norb || ob >= B.norb) continue;
      double tot = 0.0;
      for (int k = 0; k < 9; ++k) tot += part[kind * 9 + k];
      const double sv = sm[SPG_PRE + kind] * ipow(0.5 * r, A.n[la] + B.n[lb] + 1) * tot;
      const double sg = (((la + m) & 1) ? -1.0 : 1.0) * (((lb + m) & 1) ? -1.0 : 1.0);  // atom j sits at -z
      const int ia = (m == 0) ? oa : (la == 1 ? 2 : (m == 1 ? 5 : 7));
      const int ib = (m == 0) ? ob : (lb == 1 ? 2 : (m == 1 ? 5 : 7));
      const double bb = 0.5 * (A.beta[la] + B.beta[lb]);
      S[ia * 9 + ib] = sg * sv * bb;
      if (m > 0) S[(ia + 1) * 9 + ib + 1] = sg * sv * bb;
    }
  }
  // densities into the local frame: X_loc = R^t X R for the three blocks (tmp = X R, then R^t tmp)
  for (int t = tid; t < 243; t += nthr) {
    const int blk = t / 81, a = (t % 81) / 9, q = t % 9;
    const double* X = sm + SPG_PI + 81 * blk;
    double acc = 0.0;
    for (int k = 0; k < 9; ++k) acc += X[a * 9 + k] * R[k * 9 + q];
    tmp[t] = acc;
  }
  SEQM_SYNC();
  for (int t = tid; t < 243; t += nthr) {
    const int blk = t / 81, a = (t % 81) / 9, q = t % 9;
    double acc = 0.0;
    for (int k = 0; k < 9; ++k) acc += R[k * 9 + a] * tmp[blk * 81 + k * 9 + q];
    sm[SPG_AI + t] = acc;
  }
  SEQM_SYNC();
  const double* Ai = sm + SPG_AI;
  const double* Aj = sm + SPG_AJ;
  const double* Aij = sm + SPG_AIJ;
  for (int t = tid; t < 90; t += nthr) {
    const int side = t / 45, kl = t % 45;
    int mu, nu;
    spd_unpack(kl, &mu, &nu);
    pk[t] = (side ? Aj : Ai)[mu * 9 + nu] * (mu == nu ? 1.0 : 2.0);
  }
  SEQM_SYNC();
  double en = 0.0;
  // resonance term: 2 sum D_ij o (beta S)
  for (int t = tid; t < 81; t += nthr) en += 2.0 * Aij[t] * S[t];
  // two-electron and core-attraction terms with a d orbital; the multipole coefficients come from the sparse lists
  // (<= 4 non-zero (source, m) entries per product) that the kernel staged in shared memory
  const double* cv = sm + SPG_CV;
  const int* cc = reinterpret_cast<const int*>(sm + SPG_CC);
  const double ti = par(b, SEQM_P_TORE, i), tj = par(b, SEQM_P_TORE, j);
  for (int t = tid; t < npi * npj; t += nthr) {
    const int kl = t / npj, mn = t % npj;
    if (kl < 10 && mn < 10) continue;
    const double L = spd_local_integral(kl, mn, cv, cc, V);
    double c = pk[kl] * pk[45 + mn];
    if (mn == 0) c -= pk[kl] * tj;
    if (kl == 0) c -= pk[45 + mn] * ti;
    int mu, nu, la, sg;
    spd_unpack(kl, &mu, &nu);
    spd_unpack(mn, &la, &sg);
    double x = Aij[mu * 9 + la] * Aij[nu * 9 + sg];
    if (mu != nu) x += Aij[nu * 9 + la] * Aij[mu * 9 + sg];
    if (la != sg) {
      x += Aij[mu * 9 + sg] * Aij[nu * 9 + la];
      if (mu != nu) x += Aij[nu * 9 + sg] * Aij[mu * 9 + la];
    }
    en += (c - 0.5 * x) * L;
  }
  if (dj && tid == 0 && r <= SEQM_OVERLAP_CUTOFF) {  // reference quirk of the (d_yz, d_xy) overlap element (molecular frame)
    double ca, sb, cb, uu[3], vv[3];
    spd_local_axes(e, uu, vv, &ca, &sb, &cb);
    const double bdd = 0.5 * (A.beta[2] + B.beta[2]);
    const double s333 = (bdd != 0.0) ? S[7 * 9 + 7] / bdd : 0.0;
    const double fix = 2.0 * s333 * ca * sb * cb * (2.0 * ca * ca - 1.0);
    en += 2.0 * bdd * fix * (sm[SPG_PIJ + 7 * 9 + 8] + sm[SPG_PIJ + 8 * 9 + 7]);
  }
  return block_sum(en, sm + SPG_RED);
}

#ifndef SPD_GRAD_MINB
#define SPD_GRAD_MINB 8  // 64 registers, 8 CTAs per SM: measured 13 % faster than the unconstrained 96 (barrier-latency bound)
#endif
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS2(SPD_THREADS, SPD_GRAD_MINB) spd_pair_gradient_kernel(seqm_batch_t b, const double* __restrict__ xyz,
                                                                          const double* __restrict__ P, double* __restrict__ gp) {
  SEQM_DYN_SMEM(double, sm);
  const int slot = blockIdx.x;
  const int p = b.ypairs[slot];
  const int i = b.pair_i[p], j = b.pair_j[p];
  const MolView v = mol_view(b, b.atom_mol[i]);
  const int li = i - v.a0, lj = j - v.a0;
  const bool dj = lj < v.nsh;
  const int noj = orb_cnt(v, lj);
  const int oi = orb_off(v, li), oj = orb_off(v, lj), n = v.n;
  const double* Pm = P + v.mat0;
  for (int t = threadIdx.x; t < 243; t += blockDim.x) {  // molecular-frame density blocks, zero beyond the atoms' orbitals
    const int blk = t / 81, a = (t % 81) / 9, q = t % 9;
    const int ro = (blk == 1) ? oj : oi, co = (blk == 0) ? oi : oj;
    const int nr = (blk == 1) ? noj : 9, nc = (blk == 0) ? 9 : noj;
    sm[SPG_PI + t] = (a < nr && q < nc) ? Pm[(ro + a) * n + co + q] : 0.0;
  }
  __shared__ SpdAtom sAB[2];
  if (threadIdx.x == 0) spd_load_atom(b, i, true, sAB[0]);
  if (threadIdx.x == 1 || blockDim.x == 1) spd_load_atom(b, j, dj, sAB[1]);
  SEQM_SYNC();
  const SpdAtom& A = sAB[0];
  const SpdAtom& B = sAB[1];
  {  // geometry-independent staging: sparse multipole coefficients of both atoms, overlap prefactors
    spd_stage_coefficients(b, !dj && noj == 4, sm + SPG_CV, reinterpret_cast<int*>(sm + SPG_CC));
    for (int kind = threadIdx.x; kind < 14; kind += blockDim.x) {
      int la, lb, m;
      spd_kind(kind, &la, &lb, &m);
      const int na = A.n[la], nb = B.n[lb];
      double pre = 0.0;
      if (na >= 1 && na <= 4 && nb >= 1 && nb <= 4 && A.zeta[la] > 0.0 && B.zeta[lb] > 0.0) {
        double fa = 1.0, fb = 1.0;
        for (int k = 2; k <= 2 * na; ++k) fa *= k;
        for (int k = 2; k <= 2 * nb; ++k) fb *= k;
        pre = pow(2.0 * A.zeta[la], na + 0.5) * pow(2.0 * B.zeta[lb], nb + 0.5) / sqrt(fa * fb);
      }
      sm[SPG_PRE + kind] = pre;
    }
  }
  SEQM_SYNC();
  const double delta = 1.0e-4;
  const double Rj[3] = {xyz[3 * (long long)j], xyz[3 * (long long)j + 1], xyz[3 * (long long)j + 2]};
  const double Ri0[3] = {xyz[3 * (long long)i], xyz[3 * (long long)i + 1], xyz[3 * (long long)i + 2]};
  // undisplaced distance (bohr): selects the B-integral regime of every stencil point
  const double r0 = sqrt((Rj[0] - Ri0[0]) * (Rj[0] - Ri0[0]) + (Rj[1] - Ri0[1]) * (Rj[1] - Ri0[1]) +
                         (Rj[2] - Ri0[2]) * (Rj[2] - Ri0[2])) * (1.0 / SEQM_A0);
  double g[3];
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    double E[4];
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const double s = (k == 0) ? 1.0 : (k == 1 ? -1.0 : (k == 2 ? 2.0 : -2.0));
      double Ri[3] = {Ri0[0], Ri0[1], Ri0[2]};
      Ri[c] += s * delta;
      E[k] = spd_pair_energy_local(b, i, j, dj, A, B, Ri, Rj, r0, sm);
      SEQM_SYNC();
    }
    g[c] = (8.0 * (E[0] - E[1]) - (E[2] - E[3])) / (12.0 * delta);
  }
  if (threadIdx.x == 0) {  // the sp gradient kernel has already written the sp x sp part of this pair
    gp[3 * (long long)p] += g[0];
    gp[3 * (long long)p + 1] += g[1];
    gp[3 * (long long)p + 2] += g[2];
  }
}
