// pair_core.cuh -- per-atom-pair physics, templated on the scalar (double | Dual3).
//
// Replaces (reference file:line, lanl/PYSEQM v2.0.0):
//   per-atom multipole prologue      seqm/seqm_functions/two_elec_two_center_int.py:116-247, cal_par.py:11-28,112-169,198-257
//   22 local-frame integrals         seqm/seqm_functions/two_elec_two_center_int_local_frame.py:77-274
//   local -> molecular rotation      seqm/seqm_functions/two_elec_two_center_int.py:1384-1574, quaternion 1576-1631
//   Slater overlaps (n = 1..3, sp)   seqm/seqm_functions/diat_overlap_PM6_SP.py:6-444, aintgs 464, bintgs 522
//   core-core repulsion              seqm/seqm_functions/energy.py:91-139
//
// Formulation (not a transcription): every charge distribution is a Dewar-Thiel point-charge multipole
// (q, mu_z, mu_x, Q_zz, Q_xx, Q_xz); the 22 integrals are sums of the 69 distinct
// ev*c/sqrt(z^2 + x^2 + (rho_a+rho_b)^2) terms derived from those configurations.  The molecular-frame
// tensor is w = T^t L T with T the (1 + 3x3 + 6x6) pair-product transform of the rotation rows.
#pragma once
#include "dual.cuh"

#define SEQM_EV 27.21      // constants.py:4
#define SEQM_A0 0.529167   // constants.py:9
#define SEQM_OVERLAP_CUTOFF 40.0  // bohr, constants.py:23

struct AtomMultipole {  // per atom, atomic units
  double dd, qq, rho0, rho1, rho2;
};

// ---- per-atom prologue --------------------------------------------------------------------------
// dd, qq from zeta_s, zeta_p and the principal quantum number; rho1/rho2 by exactly five secant steps.
SEQM_HD AtomMultipole atom_multipole(int Z, double qn, double zs, double zp, double gss, double gpp, double gp2, double hsp) {
  AtomMultipole m;
  m.dd = m.qq = m.rho1 = m.rho2 = 0.0;
  m.rho0 = 0.5 * SEQM_EV / gss;
  if (Z > 2) {
    double hpp = 0.5 * (gpp - gp2);
    if (hpp < 0.1) hpp = 0.1;
    m.dd = (2.0 * qn + 1.0) * pow(4.0 * zs * zp, qn + 0.5) / pow(zs + zp, 2.0 * qn + 2.0) / sqrt(3.0);
    m.qq = sqrt((4.0 * qn * qn + 6.0 * qn + 2.0) / 20.0) / zp;
    {
      const double D = m.dd, h = hsp / SEQM_EV;
      double d1 = pow(fabs(h) / (D * D), 1.0 / 3.0);
      if (h < 0.0) d1 = -d1;
      double d2 = d1 + 0.04;
      for (int it = 0; it < 5; ++it) {
        double h1 = 0.5 * d1 - 0.5 / sqrt(4.0 * D * D + 1.0 / (d1 * d1));
        double h2 = 0.5 * d2 - 0.5 / sqrt(4.0 * D * D + 1.0 / (d2 * d2));
        double d3 = (fabs(h2 - h1) > 1.0e-16) ? d1 + (d2 - d1) * (h - h1) / (h2 - h1) : d2;
        d1 = d2;
        d2 = d3;
      }
      m.rho1 = 0.5 / d2;
    }
    {
      const double D = m.qq, h = hpp / SEQM_EV;
      double q1 = pow(fabs(h) / 3.0 / (D * D * D * D), 0.2);
      if (h < 0.0) q1 = -q1;
      double q2 = q1 + 0.04;
      for (int it = 0; it < 5; ++it) {
        double h1 = 0.25 * q1 - 0.5 / sqrt(4.0 * D * D + 1.0 / (q1 * q1)) + 0.25 / sqrt(8.0 * D * D + 1.0 / (q1 * q1));
        double h2 = 0.25 * q2 - 0.5 / sqrt(4.0 * D * D + 1.0 / (q2 * q2)) + 0.25 / sqrt(8.0 * D * D + 1.0 / (q2 * q2));
        double q3 = (fabs(h2 - h1) > 1.0e-16) ? q1 + (q2 - q1) * (h - h1) / (h2 - h1) : q2;
        q1 = q2;
        q2 = q3;
      }
      m.rho2 = 0.5 / q2;
    }
  }
  return m;
}

// ---- local-frame integrals ----------------------------------------------------------------------
// f(z, x2, a) = 1/sqrt(z^2 + x2 + a): one point-charge interaction, z axial and x2 squared transverse
// separation, a = (rho_a + rho_b)^2.
template <class T>
SEQM_HD T pc(const T& z, double x2, double a) { return inv_sqrt(z * z + (x2 + a)); }

// ri[0..21] in the order  (ss|ss) (so|ss) (oo|ss) (pp|ss) (ss|os) (so|so) (sp|sp) (oo|so) (pp|so) (po|sp)
// (ss|oo) (ss|pp) (so|oo) (so|pp) (sp|op) (oo|oo) (pp|oo) (oo|pp) (pp|pp) (po|po) (pp|p*p*) (p*p|p*p)
// (o = p-sigma, p/p* = the two p-pi).  Atom A sits at +r on the local axis seen from B.
// nint = 1 (H-H), 4 (X-H: only ri[0..3]) or 22 (X-X).
template <class T>
SEQM_HD void local_integrals(const T& r, const AtomMultipole& A, const AtomMultipole& B, int nint, T* ri) {
  const double ev = SEQM_EV;
  const double a00 = (A.rho0 + B.rho0) * (A.rho0 + B.rho0);
  const T qq = ev * pc(r, 0.0, a00);
  ri[0] = qq;
  if (nint == 1) return;
  const double Da = A.dd, Qa2 = 2.0 * A.qq, Qa = A.qq;
  const double a10 = (A.rho1 + B.rho0) * (A.rho1 + B.rho0);
  const double a20 = (A.rho2 + B.rho0) * (A.rho2 + B.rho0);
  const T mzq = (0.5 * ev) * (pc(r + Da, 0.0, a10) - pc(r - Da, 0.0, a10));
  const T f20 = pc(r, 0.0, a20);
  const T Qzzq = (0.25 * ev) * (pc(r + Qa2, 0.0, a20) + pc(r - Qa2, 0.0, a20)) - (0.5 * ev) * f20;
  const T Qxxq = (0.5 * ev) * (pc(r, Qa2 * Qa2, a20) - f20);
  ri[1] = mzq;
  ri[2] = qq + Qzzq;
  ri[3] = qq + Qxxq;
  if (nint == 4) return;
  const double Db = B.dd, Qb2 = 2.0 * B.qq, Qb = B.qq;
  const double a01 = (A.rho0 + B.rho1) * (A.rho0 + B.rho1);
  const double a02 = (A.rho0 + B.rho2) * (A.rho0 + B.rho2);
  const double a11 = (A.rho1 + B.rho1) * (A.rho1 + B.rho1);
  const double a21 = (A.rho2 + B.rho1) * (A.rho2 + B.rho1);
  const double a12 = (A.rho1 + B.rho2) * (A.rho1 + B.rho2);
  const double a22 = (A.rho2 + B.rho2) * (A.rho2 + B.rho2);
  // monopole / dipole / quadrupole on B seen by the monopole on A
  const T qmz = (0.5 * ev) * (pc(r - Db, 0.0, a01) - pc(r + Db, 0.0, a01));
  const T f02 = pc(r, 0.0, a02);
  const T qQzz = (0.25 * ev) * (pc(r - Qb2, 0.0, a02) + pc(r + Qb2, 0.0, a02)) - (0.5 * ev) * f02;
  const T qQxx = (0.5 * ev) * (pc(r, Qb2 * Qb2, a02) - f02);
  ri[4] = qmz;
  ri[10] = qq + qQzz;
  ri[11] = qq + qQxx;
  // dipole-dipole
  ri[5] = (0.25 * ev) * (pc(r + (Da - Db), 0.0, a11) + pc(r - (Da - Db), 0.0, a11) - pc(r + (Da + Db), 0.0, a11) -
                         pc(r - (Da + Db), 0.0, a11));
  ri[6] = (0.5 * ev) * (pc(r, (Da - Db) * (Da - Db), a11) - pc(r, (Da + Db) * (Da + Db), a11));
  // quadrupole(A)-dipole(B)
  {
    const T m = pc(r - Db, 0.0, a21), p = pc(r + Db, 0.0, a21);
    const T Qzzmz = (0.125 * ev) * (pc(r + (Qa2 - Db), 0.0, a21) - pc(r + (Qa2 + Db), 0.0, a21) +
                                    pc(r - (Qa2 + Db), 0.0, a21) - pc(r - (Qa2 - Db), 0.0, a21)) -
                    (0.25 * ev) * (m - p);
    const T Qxxmz = (0.25 * ev) * (pc(r - Db, Qa2 * Qa2, a21) - pc(r + Db, Qa2 * Qa2, a21)) - (0.25 * ev) * (m - p);
    ri[7] = qmz + Qzzmz;
    ri[8] = qmz + Qxxmz;
    const double xm = (Qa - Db) * (Qa - Db), xp = (Qa + Db) * (Qa + Db);
    ri[9] = (0.25 * ev) * (pc(r + Qa, xm, a21) - pc(r - Qa, xm, a21) - pc(r + Qa, xp, a21) + pc(r - Qa, xp, a21));
  }
  // dipole(A)-quadrupole(B)
  {
    const T p = pc(r + Da, 0.0, a12), m = pc(r - Da, 0.0, a12);
    const T mzQzz = (0.125 * ev) * (pc(r + (Da - Qb2), 0.0, a12) + pc(r + (Da + Qb2), 0.0, a12) -
                                    pc(r - (Da + Qb2), 0.0, a12) - pc(r - (Da - Qb2), 0.0, a12)) -
                    (0.25 * ev) * (p - m);
    const T mzQxx = (0.25 * ev) * (pc(r + Da, Qb2 * Qb2, a12) - pc(r - Da, Qb2 * Qb2, a12)) - (0.25 * ev) * (p - m);
    ri[12] = mzq + mzQzz;
    ri[13] = mzq + mzQxx;
    const double xm = (Da - Qb) * (Da - Qb), xp = (Da + Qb) * (Da + Qb);
    ri[14] = (0.25 * ev) * (pc(r - Qb, xm, a12) - pc(r + Qb, xm, a12) - pc(r - Qb, xp, a12) + pc(r + Qb, xp, a12));
  }
  // quadrupole-quadrupole
  {
    const T f0 = pc(r, 0.0, a22);
    const T fa = pc(r + Qa2, 0.0, a22) + pc(r - Qa2, 0.0, a22);
    const T fb = pc(r + Qb2, 0.0, a22) + pc(r - Qb2, 0.0, a22);
    const T fxa = pc(r, Qa2 * Qa2, a22), fxb = pc(r, Qb2 * Qb2, a22);
    const T QzzQzz = (0.0625 * ev) * (pc(r + (Qa2 - Qb2), 0.0, a22) + pc(r + (Qa2 + Qb2), 0.0, a22) +
                                      pc(r - (Qa2 + Qb2), 0.0, a22) + pc(r - (Qa2 - Qb2), 0.0, a22)) -
                     (0.125 * ev) * (fa + fb) + (0.25 * ev) * f0;
    const T QxxQzz = (0.125 * ev) * (pc(r - Qb2, Qa2 * Qa2, a22) + pc(r + Qb2, Qa2 * Qa2, a22)) - (0.25 * ev) * fxa -
                     (0.125 * ev) * fb + (0.25 * ev) * f0;
    const T QzzQxx = (0.125 * ev) * (pc(r + Qa2, Qb2 * Qb2, a22) + pc(r - Qa2, Qb2 * Qb2, a22)) - (0.25 * ev) * fxb -
                     (0.125 * ev) * fa + (0.25 * ev) * f0;
    const T tail = (0.25 * ev) * (f0 - fxa - fxb);
    const T QxxQxx = (0.125 * ev) * (pc(r, (Qa2 - Qb2) * (Qa2 - Qb2), a22) + pc(r, (Qa2 + Qb2) * (Qa2 + Qb2), a22)) + tail;
    const T QxxQyy = (0.25 * ev) * pc(r, Qa2 * Qa2 + Qb2 * Qb2, a22) + tail;
    ri[15] = qq + qQzz + Qzzq + QzzQzz;
    ri[16] = qq + qQzz + Qxxq + QxxQzz;
    ri[17] = qq + qQxx + Qzzq + QzzQxx;
    ri[18] = qq + qQxx + Qxxq + QxxQxx;
    ri[20] = qq + qQxx + Qxxq + QxxQyy;
    ri[21] = 0.5 * (QxxQxx - QxxQyy);
    const double xm = (Qa - Qb) * (Qa - Qb), xp = (Qa + Qb) * (Qa + Qb);
    ri[19] = (0.125 * ev) * (pc(r + (Qa - Qb), xm, a22) - pc(r + (Qa + Qb), xm, a22) - pc(r - (Qa + Qb), xm, a22) +
                             pc(r - (Qa - Qb), xm, a22) - pc(r + (Qa - Qb), xp, a22) + pc(r + (Qa + Qb), xp, a22) +
                             pc(r - (Qa + Qb), xp, a22) - pc(r - (Qa - Qb), xp, a22));
  }
}

// ---- rotation -----------------------------------------------------------------------------------
// Rows of the rotation that takes the unit vector v onto the local x axis, via the quaternion
// (0, v_z, -v_y, 1 + v_x)/N; antipodal case |1+v_x| < 1e-7 -> fixed 180 degree flip (and, as in the
// reference, zero derivative).  rot[a][k]: component k of local axis a.
template <class T>
SEQM_HD void rotation_rows(const T v[3], T rot[3][3]) {
  T qy = v[2], qz = -v[1], qw = 1.0 + v[0];
  if (fabs(val(qw)) < 1.0e-7) {
    qy = T(0.0);
    qz = T(1.0);
    qw = T(0.0);
  }
  const T inv = inv_sqrt(qy * qy + qz * qz + qw * qw);
  qy = qy * inv;
  qz = qz * inv;
  qw = qw * inv;
  rot[0][0] = 1.0 - 2.0 * (qy * qy + qz * qz);
  rot[0][1] = -2.0 * (qz * qw);
  rot[0][2] = 2.0 * (qy * qw);
  rot[1][0] = 2.0 * (qz * qw);
  rot[1][1] = 1.0 - 2.0 * (qz * qz);
  rot[1][2] = 2.0 * (qy * qz);
  rot[2][0] = -2.0 * (qy * qw);
  rot[2][1] = 2.0 * (qy * qz);
  rot[2][2] = 1.0 - 2.0 * (qy * qy);
}

// Packed pair index (molecular and local frames alike), orbital order (s, x|sigma, y|pi, z|pi*):
//   0:(ss) 1:(x s) 2:(x x) 3:(y s) 4:(y x) 5:(y y) 6:(z s) 7:(z x) 8:(z y) 9:(z z)
SEQM_HD int pack2(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

// Local-frame tensor L[KL][MN] in terms of ri: 34 non-zeros.  Entry = {KL, MN, ri index}.
struct LEntry { signed char kl, mn, k; };
#define SEQM_NL 34
SEQM_HD LEntry l_entry(int i) {
  const LEntry t[SEQM_NL] = {
      {0, 0, 0},  {1, 0, 1},  {2, 0, 2},  {5, 0, 3},  {9, 0, 3},  {0, 1, 4},  {1, 1, 5},  {3, 3, 6},  {6, 6, 6},
      {2, 1, 7},  {5, 1, 8},  {9, 1, 8},  {4, 3, 9},  {7, 6, 9},  {0, 2, 10}, {0, 5, 11}, {0, 9, 11}, {1, 2, 12},
      {1, 5, 13}, {1, 9, 13}, {3, 4, 14}, {6, 7, 14}, {2, 2, 15}, {5, 2, 16}, {9, 2, 16}, {2, 5, 17}, {2, 9, 17},
      {5, 5, 18}, {9, 9, 18}, {4, 4, 19}, {7, 7, 19}, {5, 9, 20}, {9, 5, 20}, {8, 8, 21}};
  return t[i];
}

// T[KL][kl]: coefficient of the local pair product KL in the molecular pair product kl.
template <class T>
SEQM_HD void pair_transform(const T rot[3][3], T Tm[10][10]) {
#pragma unroll
  for (int i = 0; i < 10; ++i)
#pragma unroll
    for (int j = 0; j < 10; ++j) Tm[i][j] = T(0.0);
  Tm[0][0] = T(1.0);
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int k = 0; k < 3; ++k) Tm[pack2(a + 1, 0)][pack2(k + 1, 0)] = rot[a][k];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          if (b > a || l > k) continue;
          T t = rot[a][k] * rot[b][l];
          if (a != b) t = t + rot[b][k] * rot[a][l];
          Tm[pack2(a + 1, b + 1)][pack2(k + 1, l + 1)] = t;
        }
}

// Index classes of the packed pair index: 0 -> ss, {1,3,6} -> p s, {2,4,5,7,8,9} -> p p.
SEQM_HD int pack_class(int kl) { return (kl == 0) ? 0 : ((kl == 1 || kl == 3 || kl == 6) ? 1 : 2); }

// w[kl][mn] = sum_{KL,MN} T[KL][kl] L[KL][MN] T[MN][mn]; only same-class entries of T are non-zero.
// ncolA / ncolB: 1 for a hydrogen (only the ss product exists), 10 for a heavy atom.
template <class T>
SEQM_HD void rotate_to_molecular(const T* ri, int nint, const T Tm[10][10], T w[10][10]) {
  const int nB = (nint == 22) ? 10 : 1, nA = (nint == 1) ? 1 : 10;
  if (nint == 1) {  // H-H: (ss|ss) is rotation invariant
    w[0][0] = ri[0];
    return;
  }
  T U[10][10];  // U[KL][mn] = sum_MN L[KL][MN] T[MN][mn]
#pragma unroll
  for (int i = 0; i < 10; ++i)
#pragma unroll
    for (int j = 0; j < 10; ++j)
      if (j < nB) { U[i][j] = T(0.0); w[i][j] = T(0.0); }
#pragma unroll
  for (int e = 0; e < SEQM_NL; ++e) {
    const LEntry le = l_entry(e);
    if (le.k >= nint || le.mn >= nB) continue;
    const int c = pack_class(le.mn);
#pragma unroll
    for (int mn = 0; mn < 10; ++mn)
      if (mn < nB && pack_class(mn) == c) U[le.kl][mn] = U[le.kl][mn] + ri[le.k] * Tm[le.mn][mn];
  }
#pragma unroll
  for (int kl = 0; kl < 10; ++kl) {
    if (kl >= nA) continue;
    const int c = pack_class(kl);
#pragma unroll
    for (int KL = 0; KL < 10; ++KL) {
      if (pack_class(KL) != c) continue;
#pragma unroll
      for (int mn = 0; mn < 10; ++mn)
        if (mn < nB) w[kl][mn] = w[kl][mn] + Tm[KL][kl] * U[KL][mn];
    }
  }
}

// ---- Slater overlaps ----------------------------------------------------------------------------
// A_k(x) = int_1^inf t^k e^{-xt} dt by upward recurrence; B_k(x) = int_-1^1 t^k e^{-xt} dt with the
// reference's three regimes (|x| > 0.5 recurrence, 1e-6 < |x| <= 0.5 four-term series, else x = 0).
#define SEQM_KMAX 6
template <class T>
SEQM_HD void aux_A(const T& x, int kmax, T* A) {
  A[0] = e_xp(-x) / x;
  for (int k = 1; k <= kmax; ++k) A[k] = A[0] + (double)k * A[k - 1] / x;
}
template <class T>
SEQM_HD void aux_B(const T& x, int kmax, T* B) {
  const double ax = fabs(val(x));
  if (ax > 0.5) {
    const T tx = e_xp(x) / x, tmx = -(e_xp(-x) / x);
    B[0] = tx + tmx;
    for (int k = 1; k <= kmax; ++k) B[k] = ((k & 1) ? (tmx - tx) : (tx + tmx)) + (double)k * B[k - 1] / x;
  } else if (ax > 1.0e-6) {
    const T x2 = x * x;
    for (int k = 0; k <= kmax; ++k) {
      if ((k & 1) == 0)
        B[k] = 2.0 / (k + 1.0) + x2 / (k + 3.0) + x2 * x2 / ((k + 5.0) * 12.0) + x2 * x2 * x2 / ((k + 7.0) * 360.0);
      else
        B[k] = (-2.0 / (k + 2.0)) * x - x2 * x / ((k + 4.0) * 3.0) - x2 * x2 * x / ((k + 6.0) * 60.0);
    }
  } else {
    for (int k = 0; k <= kmax; ++k) B[k] = T((k & 1) ? 0.0 : 2.0 / (k + 1.0));
  }
}

// Integer polynomial tables of the prolate-spheroidal integrands (filled at library init, see
// overlap_tables.h): poly[na-1][nb-1][kind][k][l] multiplies A_k B_l.
// kind: 0 (s|s), 1 (p-sigma_A|s_B), 2 (s_A|p-sigma_B), 3 (p-sigma|p-sigma), 4 (p-pi|p-pi);
// both p-sigma lobes point along +e = R_B - R_A.
struct OverlapTables {
  signed char poly[3][3][5][SEQM_KMAX + 1][SEQM_KMAX + 1];
  double norm[3][3];  // 1/sqrt((2na)! (2nb)!)
};

template <class T>
SEQM_HD T sto_overlap(const OverlapTables& tab, int na, int nb, int kind, double za, double zb, const T& r) {
  const int kmax = na + nb;
  T A[SEQM_KMAX + 1], B[SEQM_KMAX + 1];
  aux_A((0.5 * (za + zb)) * r, kmax, A);
  aux_B((0.5 * (za - zb)) * r, kmax, B);
  T tot = T(0.0);
  for (int k = 0; k <= kmax; ++k)
    for (int l = 0; l <= kmax; ++l) {
      const int c = tab.poly[na - 1][nb - 1][kind][k][l];
      if (c != 0) tot = tot + (double)c * (A[k] * B[l]);
    }
  const double ang = (kind == 0) ? 0.5 : ((kind == 1 || kind == 2) ? 0.8660254037844386 : (kind == 3 ? 1.5 : 0.75));
  const double pre = pow(2.0 * za, na + 0.5) * pow(2.0 * zb, nb + 0.5) * tab.norm[na - 1][nb - 1] * ang;
  return pre * ipow(0.5 * r, na + nb + 1) * tot;
}

// S[mu][nu] = <mu on A | nu on B> in the molecular frame; e = unit vector A -> B; r in bohr.
// na/nb principal quantum numbers; heavyA/heavyB: atom carries p orbitals.
template <class T>
SEQM_HD void overlap_block(const OverlapTables& tab, int na, int nb, bool heavyA, bool heavyB, double zsa, double zpa,
                           double zsb, double zpb, const T& r, const T e[3], T S[4][4]) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) S[i][j] = T(0.0);
  if (val(r) > SEQM_OVERLAP_CUTOFF) return;
  if (na < 1 || na > 3 || nb < 1 || nb > 3) return;  // tables cover n <= 3; PM6 pairs with a d atom (n up to 4) are redone by spd_pair_kernel
  S[0][0] = sto_overlap(tab, na, nb, 0, zsa, zsb, r);
  if (heavyA) {
    const T os = sto_overlap(tab, na, nb, 1, zpa, zsb, r);
    for (int k = 0; k < 3; ++k) S[k + 1][0] = os * e[k];
  }
  if (heavyB) {
    const T so = sto_overlap(tab, na, nb, 2, zsa, zpb, r);
    for (int k = 0; k < 3; ++k) S[0][k + 1] = so * e[k];
  }
  if (heavyA && heavyB) {
    const T oo = sto_overlap(tab, na, nb, 3, zpa, zpb, r);
    const T pp = sto_overlap(tab, na, nb, 4, zpa, zpb, r);
    for (int k = 0; k < 3; ++k)
      for (int l = 0; l < 3; ++l) {
        T t = (oo - pp) * (e[k] * e[l]);
        if (k == l) t = t + pp;
        S[k + 1][l + 1] = t;
      }
  }
}

// ---- core-core repulsion ------------------------------------------------------------------------
// method: 0 MNDO, 1 AM1 (4 gaussians), 2 PM3 (2 gaussians), 3 PM6_SP (4 gaussians + pairwise alpha/chi terms).
// r in bohr, gam = (ss|ss).  energy.py:91-174.
struct CorePar {
  double tore, alpha, gK[4], gL[4], gM[4], rho0eff, atnum;
};
template <class T>
SEQM_HD T core_core(int method, int ni, int nj, const CorePar& A, const CorePar& B, const T& r, const T& gam, double alp,
                    double chi) {
  const T ra = r * SEQM_A0;
  T E;
  if (method >= 3) {  // PM6_SP and PM6 with d orbitals share the core-core function (energy.py:140-171)
    // PM6 core-core (energy.py:140-171): Voityuk-type unpolarisable-core term + scaled (ss|ss)-like interaction
    const double za3 = pow(A.atnum, 1.0 / 3.0) + pow(B.atnum, 1.0 / 3.0);
    const T q = za3 / ra;
    const T q2 = q * q, q4 = q2 * q2;
    const T unpol = 1.0e-8 * (q4 * q4 * q4);
    const double rs = A.rho0eff + B.rho0eff;
    const T g0 = (A.tore * B.tore * SEQM_EV) * inv_sqrt(r * r + rs * rs);
    const bool xh = (ni == 6 || ni == 7 || ni == 8) && (nj == 1);
    T scale;
    if (xh) {
      scale = 1.0 + (2.0 * chi) * e_xp(-(alp * (ra * ra)));
    } else {
      const T ra2 = ra * ra;
      scale = 1.0 + (2.0 * chi) * e_xp(-(alp * (ra + 0.0003 * (ra2 * ra2 * ra2))));
    }
    E = unpol + g0 * scale;
    if (ni == 6 && nj == 6) E = E + g0 * (9.28 * e_xp(-(5.98 * ra)));
    if (ni == 14 && nj == 8) {
      const T d = r - 2.9;
      E = E - g0 * (0.0007 * e_xp(-(d * d)));
    }
  } else {
    const bool xh = ((ni == 7) || (ni == 8)) && (nj == 1);
    T t2 = e_xp(-(A.alpha * ra));
    if (xh) t2 = t2 * ra;
    const T t3 = e_xp(-(B.alpha * ra));
    E = (A.tore * B.tore) * gam * (1.0 + t2 + t3);
  }
  if (method != 0) {
    const int ng = (method == 2) ? 2 : 4;
    T g = T(0.0);
    for (int k = 0; k < ng; ++k) {
      const T da = ra - A.gM[k], db = ra - B.gM[k];
      g = g + A.gK[k] * e_xp(-(A.gL[k] * (da * da))) + B.gK[k] * e_xp(-(B.gL[k] * (db * db)));
    }
    E = E + (A.tore * B.tore) / ra * g;
  }
  return E;
}
