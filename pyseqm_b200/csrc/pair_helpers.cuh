// pair_helpers.cuh -- per-pair helper functions shared by the sp pair kernels (pair_kernels.cuh) and the PM6 d-orbital
// kernels (spd_kernels.cuh, compiled in their own translation unit): parameter loads, pair geometry, the sp block
// w (10 x 10) and overlap (4 x 4) of one pair.  Device functions only, no kernels.
#pragma once
#include "common.cuh"

SEQM_HD AtomMultipole load_multipole(const seqm_batch_t& b, int a) {
  AtomMultipole m;
  m.dd = par(b, SEQM_P_DD, a);
  m.qq = par(b, SEQM_P_QQ, a);
  m.rho0 = par(b, SEQM_P_RHO0, a);
  m.rho1 = par(b, SEQM_P_RHO1, a);
  m.rho2 = par(b, SEQM_P_RHO2, a);
  return m;
}
SEQM_HD CorePar load_core(const seqm_batch_t& b, int a) {
  CorePar c;
  c.tore = par(b, SEQM_P_TORE, a);
  c.alpha = par(b, SEQM_P_ALPHA, a);
  for (int k = 0; k < 4; ++k) {
    c.gK[k] = par(b, SEQM_P_K1 + k, a);
    c.gL[k] = par(b, SEQM_P_L1 + k, a);
    c.gM[k] = par(b, SEQM_P_M1 + k, a);
  }
  const double rc = par(b, SEQM_P_RHOCORE, a);
  c.rho0eff = (rc != 0.0) ? rc : par(b, SEQM_P_RHO0, a);  // two_elec_two_center_int.py:273-281
  c.atnum = par(b, SEQM_P_ATNUM, a);
  return c;
}
SEQM_HD void pair_pw(const seqm_batch_t& b, int i, int j, double& alp, double& chi) {
  alp = chi = 0.0;
  if (b.pw_alpha) {
    const long long k = (long long)b.atom_Z[i] * b.pw_dim + b.atom_Z[j];
    alp = b.pw_alpha[k];
    chi = b.pw_chi[k];
  }
}

// Geometry of a pair as scalars of type T.  For T = Dual3 the derivative slots are d/dR_i.
template <class T>
struct PairGeom {
  T r;     // bohr
  T e[3];  // unit vector i -> j
};
SEQM_HD void pair_geom(const double* xyz, int i, int j, PairGeom<double>& g) {
  const double dx = xyz[3 * j] - xyz[3 * i], dy = xyz[3 * j + 1] - xyz[3 * i + 1], dz = xyz[3 * j + 2] - xyz[3 * i + 2];
  const double d = sqrt(dx * dx + dy * dy + dz * dz);
  g.e[0] = dx / d;
  g.e[1] = dy / d;
  g.e[2] = dz / d;
  g.r = d * (1.0 / SEQM_A0);
}
SEQM_HD void pair_geom(const double* xyz, int i, int j, PairGeom<Dual3>& g) {
  // X = R_j - R_i ; dX/dR_i = -1
  const Dual3 dx(xyz[3 * j] - xyz[3 * i], -1.0, 0.0, 0.0);
  const Dual3 dy(xyz[3 * j + 1] - xyz[3 * i + 1], 0.0, -1.0, 0.0);
  const Dual3 dz(xyz[3 * j + 2] - xyz[3 * i + 2], 0.0, 0.0, -1.0);
  const Dual3 d = sq_root(dx * dx + dy * dy + dz * dz);
  g.e[0] = dx / d;
  g.e[1] = dy / d;
  g.e[2] = dz / d;
  g.r = d * (1.0 / SEQM_A0);
}

// Parser's outer cutoff (basics.py:326: a pair is kept when |R_i - R_j|^2 < cutoff^2).  The dense pair list keeps the
// pair; every kernel that evaluates pair physics from the geometry returns zero for it instead.
SEQM_HD bool pair_cut(const seqm_batch_t& b, double r_bohr) {
  return b.pair_outer_cutoff > 0.0 && r_bohr * SEQM_A0 >= b.pair_outer_cutoff;
}

// w of one pair in the molecular frame (only the entries that exist for the pair class are non-zero)
// Only the entries that exist for the pair class are written: [0][0] (H-H), [0..9][0] (X-H), all (X-X).
template <class T>
SEQM_HD void pair_w(const seqm_batch_t& b, int i, int j, const PairGeom<T>& g, T w[10][10], int nint) {
  T ri[22];
  local_integrals(g.r, load_multipole(b, i), load_multipole(b, j), nint, ri);
  if (nint == 1) {
    w[0][0] = ri[0];
    return;
  }
  T v[3] = {-g.e[0], -g.e[1], -g.e[2]};
  T rot[3][3];
  rotation_rows(v, rot);
  T Tm[10][10];
  pair_transform(rot, Tm);
  rotate_to_molecular(ri, nint, Tm, w);
}

#ifdef SEQM_PAIR_TU  // needs the overlap tables, which live in seqm_pair.cu
template <class T>
SEQM_HD void pair_overlap(const seqm_batch_t& b, int i, int j, const PairGeom<T>& g, T S[4][4]) {
  const bool hi = b.atom_Z[i] > 1, hj = b.atom_Z[j] > 1;
  overlap_block(c_ovl, (int)par(b, SEQM_P_QN, i), (int)par(b, SEQM_P_QN, j), hi, hj, par(b, SEQM_P_ZS, i),
                par(b, SEQM_P_ZP, i), par(b, SEQM_P_ZS, j), par(b, SEQM_P_ZP, j), g.r, g.e, S);
}
#endif

