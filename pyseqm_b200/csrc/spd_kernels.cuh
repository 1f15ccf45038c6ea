// spd_kernels.cuh -- PM6 with d orbitals (method SEQM_PM6_D).
//
//   spd_pair_kernel        one CTA per pair with a d atom ("Y pair"): the (np_i x np_j) block of two-centre integrals
//                          (np = 45 / 10 / 1 orbital products of a d / sp heavy / hydrogen atom) and the beta-scaled
//                          9 x 9 overlap block
//   spd_hcore_kernel       one CTA per molecule: packed Hcore with 9 x 9 / 4 x 4 / 1 x 1 atom blocks
//   spd_fock_kernel        one CTA per molecule: F = H + one-centre (sp parameters + Slater-Condon d integrals) + J - K/2
//   spd_pair_gradient_kernel  one CTA per Y pair: dE_pair/dR_i by a five-point stencil through the same pair code
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

// point charges (q, x, y, z) of the unit multipole (l, |m|) with separation D; returns their number (<= 6)
SEQM_HD int spd_configuration(int l, int am, double D, double q[6], double x[6], double y[6], double z[6]) {
  for (int k = 0; k < 6; ++k) q[k] = x[k] = y[k] = z[k] = 0.0;
  if (l == 0) { q[0] = 1.0; return 1; }
  if (l == 1) {
    q[0] = 0.5; q[1] = -0.5;
    if (am == 0) { z[0] = D; z[1] = -D; } else { x[0] = D; x[1] = -D; }
    return 2;
  }
  const double r2 = SPD_SQRT2 * D;
  if (am == 0) {  // Q~zx + 1/2 Q~xy: +1/4 at z = +-sqrt2 D, -1/8 at x = +-sqrt2 D and y = +-sqrt2 D
    q[0] = q[1] = 0.25; z[0] = r2; z[1] = -r2;
    q[2] = q[3] = -0.125; x[2] = r2; x[3] = -r2;
    q[4] = q[5] = -0.125; y[4] = r2; y[5] = -r2;
    return 6;
  }
  if (am == 1) {  // +-1/4 at (+-D, 0, +-D)
    int k = 0;
    for (int sa = 1; sa >= -1; sa -= 2)
      for (int sb = 1; sb >= -1; sb -= 2) { q[k] = 0.25 * sa * sb; x[k] = sa * D; z[k] = sb * D; ++k; }
    return 4;
  }
  q[0] = q[1] = 0.25; x[0] = r2; x[1] = -r2;  // (2,2): +1/4 at x = +-sqrt2 D, -1/4 at y = +-sqrt2 D
  q[2] = q[3] = -0.25; y[2] = r2; y[3] = -r2;
  return 4;
}
// interaction (eV) of multipole (ls, am) of atom i at the origin with (lt, am) of atom j at z = -r
SPD_NOINLINE double spd_interaction(int ls, int lt, int am, double Da, double Db, double add, double r) {
  double qa[6], xa[6], ya[6], za[6], qb[6], xb[6], yb[6], zb[6];
  const int na = spd_configuration(ls, am, Da, qa, xa, ya, za);
  const int nb = spd_configuration(lt, am, Db, qb, xb, yb, zb);
  double tot = 0.0;
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) {
      const double dx = xa[i] - xb[j], dy = ya[i] - yb[j], dz = za[i] - zb[j] + r;
      tot += qa[i] * qb[j] / sqrt(dx * dx + dy * dy + dz * dz + add);
    }
  return SEQM_EV * tot;
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
// molecular order s, px, py, pz, d(x2-y2), d(xz), d(z2), d(yz), d(xy); local order s, p(z, x, y), d(z2, xz, yz, x2-y2, xy)
SPD_NOINLINE void spd_orbital_rotation(const double u[3], const double v[3], const double w[3], double* R81) {
  for (int k = 0; k < 81; ++k) R81[k] = 0.0;
  R81[0] = 1.0;
  for (int c = 0; c < 3; ++c) {
    R81[(1 + c) * 9 + 1] = w[c];
    R81[(1 + c) * 9 + 2] = u[c];
    R81[(1 + c) * 9 + 3] = v[c];
  }
  const double is3 = 1.0 / SPD_SQRT3;
  for (int bq = 0; bq < 5; ++bq) {  // quadratic form Q of the local d function bq in molecular components
    double Q[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double t;
        if (bq == 0) t = (2.0 * w[i] * w[j] - u[i] * u[j] - v[i] * v[j]) * is3;
        else if (bq == 1) t = u[i] * w[j] + w[i] * u[j];
        else if (bq == 2) t = v[i] * w[j] + w[i] * v[j];
        else if (bq == 3) t = u[i] * u[j] - v[i] * v[j];
        else t = u[i] * v[j] + v[i] * u[j];
        Q[i][j] = t;
      }
    R81[4 * 9 + 4 + bq] = 0.5 * (Q[0][0] - Q[1][1]);
    R81[5 * 9 + 4 + bq] = Q[0][2];
    R81[6 * 9 + 4 + bq] = 0.5 * SPD_SQRT3 * Q[2][2];
    R81[7 * 9 + 4 + bq] = Q[1][2];
    R81[8 * 9 + 4 + bq] = Q[0][1];
  }
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

// ---- the block of one Y pair in shared memory ---------------------------------------------------------------------------
// layout of the CTA's dynamic shared memory (doubles)
#define SPD_OFF_T 0       /* 45 x 45 pair-product transform */
#define SPD_OFF_L 2025    /* local integrals, then the rotated block w (np_i x np_j, row stride np_j) */
#define SPD_OFF_TMP 4050  /* half-rotated block */
#define SPD_OFF_R 6075    /* 9 x 9 orbital rotation */
#define SPD_OFF_V 6156    /* 7 x 7 x 3 unit multipole interactions */
#define SPD_OFF_S 6303    /* 9 x 9 local overlaps, then beta-scaled molecular-frame block */
#define SPD_OFF_SP 6384   /* 10 x 10 sp x sp block of the sp path (gradient kernel recomputes it per geometry) */
#define SPD_OFF_X 6484    /* scratch: densities of the gradient kernel (81 + 45 + 45) + reductions */
#define SPD_SMEM_DOUBLES 6720

// Fills sm[SPD_OFF_L ..] with w (np_i x np_j) and sm[SPD_OFF_S ..] with hab (9 x 9, beta-scaled overlaps) for atoms
// i (a d atom) and j at positions Ri, Rj (Angstrom).  wsp: the pair's 10 x 10 block of the sp path at this geometry.
// All threads of the CTA call it; it ends with a barrier.
SPD_NOINLINE void spd_pair_block(const seqm_batch_t& b, int i, int j, bool dj, const double* Ri, const double* Rj,
                           const double* wsp, double* sm, double r_regime) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  SpdAtom A, B;
  spd_load_atom(b, i, true, A);
  spd_load_atom(b, j, dj, B);
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
  if (tid == 0) spd_orbital_rotation(u3, v3, w3, R);
  // unit multipole interactions
  for (int t = tid; t < SPD_NSRC * SPD_NSRC * 3; t += nthr) {
    const int s = t / (SPD_NSRC * 3), tt = (t / 3) % SPD_NSRC, am = t % 3;
    const int ls = spd_src_l(s), lt = spd_src_l(tt);
    double val = 0.0;
    const bool have = (tt < 3 || dj) && (tt == 0 || B.norb > 1);  // sources the partner atom carries
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
      const double sv = spd_local_overlap(b, A, B, kind, r, r_regime > 0.0 ? r_regime : r);
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
  const double* ci = (!dj && B.norb == 4) ? b.mp_coef_yx : b.mp_coef;  // reference quirk: sign of d-sigma d-delta in (d, sp) pairs
  const double* cj = b.mp_coef;
  for (int t = tid; t < npi * npj; t += nthr) {
    const int kl = t / npj, mn = t % npj;
    double acc = 0.0;
    if (kl >= 10 || mn >= 10) {
      for (int s = 0; s < SPD_NSRC; ++s)
        for (int m5 = 0; m5 < 5; ++m5) {
          const double ca = ci[(kl * SPD_NSRC + s) * 5 + m5];
          if (ca == 0.0) continue;
          const int am = (m5 == 0) ? 0 : (m5 < 3 ? 1 : 2);
          double in = 0.0;
          for (int tt = 0; tt < SPD_NSRC; ++tt) {
            const double cb = cj[(mn * SPD_NSRC + tt) * 5 + m5];
            if (cb != 0.0) in += V[(s * SPD_NSRC + tt) * 3 + am] * cb;
          }
          acc += ca * in;
        }
      // reference quirk: (d-sigma p-pi(y) | p-pi(y) s) carries -0.577350 where its neighbours carry -1/sqrt3
      if (kl == 13 && mn == 6) acc += (-0.577350 + 1.0 / SPD_SQRT3) * V[(4 * SPD_NSRC + 1) * 3 + 1];
    }
    L[t] = acc;
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
  spd_pair_block(b, i, j, dj, xyz + 3 * (long long)i, xyz + 3 * (long long)j, w10 + (long long)p * 100, sm, -1.0);
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
  // exchange blocks: F_AB[mu,la] = H_AB[mu,la] - 1/2 sum_{nu in A, sg in B} P_AB[nu,sg] (mu nu | la sg)
  for (int t = threadIdx.x; t < v.npair * 81; t += blockDim.x) {
    const int pl = t / 81, mu = (t / 9) % 9, la = t % 9;
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

// energy of one Y pair at fixed density from the block in shared memory (all threads; result valid in all threads)
SPD_NOINLINE double spd_pair_energy(const seqm_batch_t& b, int i, int j, int npi, int npj, int noi, int noj, double r_bohr,
                              const double* sm, double* red) {
  const double* W = sm + SPD_OFF_L;
  const double* S = sm + SPD_OFF_S;
  const double* Dij = sm + SPD_OFF_X;        // 9 x 9 off-diagonal density block
  const double* pki = sm + SPD_OFF_X + 81;   // weighted packed diagonal blocks
  const double* pkj = sm + SPD_OFF_X + 126;
  double e = 0.0;
  const double ti = par(b, SEQM_P_TORE, i), tj = par(b, SEQM_P_TORE, j);
  for (int t = threadIdx.x; t < 81; t += blockDim.x) e += 2.0 * Dij[t] * S[t];
  for (int t = threadIdx.x; t < npi * npj; t += blockDim.x) {
    const int kl = t / npj, mn = t % npj;
    const double wv = W[t];
    double c = pki[kl] * pkj[mn];
    if (mn == 0) c -= pki[kl] * tj;
    if (kl == 0) c -= pkj[mn] * ti;
    int mu, nu, la, sg;
    spd_unpack(kl, &mu, &nu);
    spd_unpack(mn, &la, &sg);
    // exchange: -1/2 sum over the ordered orbital quadruples that map onto (kl, mn)
    double x = Dij[mu * 9 + la] * Dij[nu * 9 + sg];
    if (mu != nu) x += Dij[nu * 9 + la] * Dij[mu * 9 + sg];
    if (la != sg) {
      x += Dij[mu * 9 + sg] * Dij[nu * 9 + la];
      if (mu != nu) x += Dij[nu * 9 + sg] * Dij[mu * 9 + la];
    }
    c -= 0.5 * x;
    e += c * wv;
  }
  e = block_sum(e, red);
  double alp, chi;
  pair_pw(b, i, j, alp, chi);
  (void)noi; (void)noj;
  return e + core_core(b.method, b.atom_Z[i], b.atom_Z[j], load_core(b, i), load_core(b, j), r_bohr, W[0], alp, chi);
}

// dE_pair/dR_i of the Y pairs by the five-point stencil (delta = 1e-4 Angstrom, error O(delta^4)) through spd_pair_block:
// the reference has no analytic PM6 gradient either (anal_grad.py:50-51; it differentiates the same expression by autograd).
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS(SPD_THREADS) spd_pair_gradient_kernel(seqm_batch_t b, const double* __restrict__ xyz,
                                                                          const double* __restrict__ P, double* __restrict__ gp) {
  SEQM_DYN_SMEM(double, sm);
  __shared__ double red[33];
  const int slot = blockIdx.x;
  const int p = b.ypairs[slot];
  const int i = b.pair_i[p], j = b.pair_j[p];
  const int mol = b.atom_mol[i];
  const MolView v = mol_view(b, mol);
  const int li = i - v.a0, lj = j - v.a0;
  const bool dj = lj < v.nsh;
  const int noi = 9, noj = orb_cnt(v, lj), npi = 45, npj = prod_cnt(v, lj);
  const int oi = orb_off(v, li), oj = orb_off(v, lj), n = v.n;
  const double* Pm = P + v.mat0;
  double* X = sm + SPD_OFF_X;
  for (int t = threadIdx.x; t < 81; t += blockDim.x) {
    const int mu = t / 9, la = t % 9;
    X[t] = (mu < noi && la < noj) ? Pm[(oi + mu) * n + oj + la] : 0.0;
  }
  for (int t = threadIdx.x; t < 90; t += blockDim.x) {
    const int side = t / 45, kl = t % 45;
    int mu, nu;
    spd_unpack(kl, &mu, &nu);
    const int o = side ? oj : oi, no = side ? noj : noi;
    X[81 + t] = (mu < no) ? Pm[(o + mu) * n + o + nu] * (mu == nu ? 1.0 : 2.0) : 0.0;
  }
  SEQM_SYNC();
  const double delta = 1.0e-4;
  double Rj[3] = {xyz[3 * (long long)j], xyz[3 * (long long)j + 1], xyz[3 * (long long)j + 2]};
  double r0;  // undisplaced distance (bohr): selects the B-integral regime of every stencil point
  {
    const double dx = Rj[0] - xyz[3 * (long long)i], dy = Rj[1] - xyz[3 * (long long)i + 1], dz = Rj[2] - xyz[3 * (long long)i + 2];
    r0 = sqrt(dx * dx + dy * dy + dz * dz) * (1.0 / SEQM_A0);
  }
  double g[3];
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    double E[4];
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const double s = (k == 0) ? 1.0 : (k == 1 ? -1.0 : (k == 2 ? 2.0 : -2.0));
      double Ri[3] = {xyz[3 * (long long)i], xyz[3 * (long long)i + 1], xyz[3 * (long long)i + 2]};
      Ri[c] += s * delta;
      double* wsp = sm + SPD_OFF_SP;
      if (threadIdx.x == 0) {  // sp x sp block of the sp path at this geometry
        PairGeom<double> pg;
        const double dx = Rj[0] - Ri[0], dy = Rj[1] - Ri[1], dz = Rj[2] - Ri[2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        pg.e[0] = dx / d; pg.e[1] = dy / d; pg.e[2] = dz / d;
        pg.r = d * (1.0 / SEQM_A0);
        double wl[10][10];
        for (int a = 0; a < 10; ++a)
          for (int q = 0; q < 10; ++q) wl[a][q] = 0.0;
        pair_w(b, i, j, pg, wl, noj > 1 ? 22 : 4);
        for (int a = 0; a < 10; ++a)
          for (int q = 0; q < 10; ++q) wsp[a * 10 + q] = (noj > 1 || q == 0) ? wl[a][q] : 0.0;
      }
      SEQM_SYNC();
      spd_pair_block(b, i, j, dj, Ri, Rj, wsp, sm, r0);
      const double dx = Rj[0] - Ri[0], dy = Rj[1] - Ri[1], dz = Rj[2] - Ri[2];
      const double rb = sqrt(dx * dx + dy * dy + dz * dz) * (1.0 / SEQM_A0);
      E[k] = pair_cut(b, rb) ? 0.0 : spd_pair_energy(b, i, j, npi, npj, noi, noj, rb, sm, red);
      SEQM_SYNC();
    }
    g[c] = (8.0 * (E[0] - E[1]) - (E[2] - E[3])) / (12.0 * delta);
  }
  if (threadIdx.x == 0) {
    gp[3 * (long long)p] = g[0];
    gp[3 * (long long)p + 1] = g[1];
    gp[3 * (long long)p + 2] = g[2];
  }
}
