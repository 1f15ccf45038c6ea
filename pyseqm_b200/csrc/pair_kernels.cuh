// pair_kernels.cuh -- the kernels that run the full pair code (local integrals, rotations, Slater overlaps) per atom pair.
// They are by far the slowest part of the library to COMPILE (forward-mode duals through the pair code), so they live in
// their own translation unit, seqm_pair.cu; the primary unit reaches them through the pairtu_* launchers.
//   pair_integrals_kernel    per pair    w (10x10), beta-scaled overlap block
//   nuclear_energy_kernel    per pair    core-core repulsion
//   pair_gradient_kernel     per pair    dE_pair/dR_i by forward-mode duals through the same pair code
//   pair_gradient_forward_kernel         forward-mode-everywhere reference of the same quantity
#pragma once
#include "common.cuh"
#include "pair_helpers.cuh"

// CLS: 0 H-H, 1 X-H, 2 X-X -- the kernel walks the class's pair list, so every warp is divergence-free and the
// block sizes (1 | 10 orbital products, 1 | 4 | 22 local integrals) are compile-time constants.
#ifndef SEQM_PI_MINB
#define SEQM_PI_MINB 0
#endif
#ifndef SEQM_PG_MINB
#define SEQM_PG_MINB 4  // 128 registers: measured 13 % faster than the unconstrained 168-188 (latency bound on its stack arrays)
#endif
template <int CLS>
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS2(128, (CLS == 2) ? SEQM_PI_MINB : 0) pair_integrals_kernel(seqm_batch_t b, const double* __restrict__ xyz, double* __restrict__ w,
                                       double* __restrict__ hab) {
  constexpr int nA = (CLS >= 1) ? 10 : 1, nB = (CLS == 2) ? 10 : 1, nint = (CLS == 2) ? 22 : ((CLS == 1) ? 4 : 1);
  const int q0 = b.pair_cls_off[CLS], q1 = b.pair_cls_off[CLS + 1];
  for (int q = q0 + blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += gridDim.x * blockDim.x) {
    const int p = b.pair_perm[q];
    const int i = b.pair_i[p], j = b.pair_j[p];
    PairGeom<double> g;
    pair_geom(xyz, i, j, g);
    double* wp = w + (long long)p * 100;
    if (pair_cut(b, g.r)) {
      for (int k = 0; k < 100; ++k) wp[k] = 0.0;
      for (int k = 0; k < 16; ++k) hab[(long long)p * 16 + k] = 0.0;
      continue;
    }
    double wl[10][10];
    pair_w(b, i, j, g, wl, nint);
    for (int k = 0; k < 10; ++k)
      for (int l = 0; l < 10; ++l) wp[k * 10 + l] = (k < nA && l < nB) ? wl[k][l] : 0.0;
    double S[4][4];
    pair_overlap(b, i, j, g, S);
    const double bsi = par(b, SEQM_P_BS, i), bpi = par(b, SEQM_P_BP, i);
    const double bsj = par(b, SEQM_P_BS, j), bpj = par(b, SEQM_P_BP, j);
    double* hp = hab + (long long)p * 16;
    for (int mu = 0; mu < 4; ++mu)
      for (int nu = 0; nu < 4; ++nu)
        hp[mu * 4 + nu] = S[mu][nu] * (0.5 * ((mu ? bpi : bsi) + (nu ? bpj : bsj)));
  }
}

SEQM_GLOBAL void nuclear_energy_kernel(seqm_batch_t b, const double* __restrict__ xyz, const double* __restrict__ w,
                                       double* __restrict__ EnucAB) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < b.npairs; p += gridDim.x * blockDim.x) {
    const int i = b.pair_i[p], j = b.pair_j[p];
    PairGeom<double> g;
    pair_geom(xyz, i, j, g);
    if (pair_cut(b, g.r)) {
      EnucAB[p] = 0.0;
      continue;
    }
    double alp, chi;
    pair_pw(b, i, j, alp, chi);
    EnucAB[p] = core_core(b.method, b.atom_Z[i], b.atom_Z[j], load_core(b, i), load_core(b, j), g.r,
                          w[(long long)p * 100], alp, chi);
  }
}
// Hellmann-Feynman pair gradient dE_pair/dR_i at fixed density (what anal_grad.py:16-225 assembles):
//   E_pair = 2 sum P_AB o (beta S) + sum_A P o e1b + sum_B P o e2a + Coulomb + exchange + core-core
// evaluated in REVERSE mode: all density contractions collapse into one coefficient matrix C (E_2e = C : w), the
// adjoints dE/dri (via C rotated into the local frame) and dE/drot (via dE/dT) are formed in plain doubles, and
// only three small pieces are differentiated forward: the 22 local integrals, the 5 Slater overlaps and the
// core-core function in r (Dual1), and the 3x3 quaternion rotation in the bond direction (Dual3).
// The forward-mode-everywhere version of this kernel (Dual3 through w = T^t L T) is kept as
// pair_gradient_forward_kernel: same numbers, 3.5x the local-memory traffic; tests compare the two.
SEQM_HD int cls_of(int kl) { return pack_class(kl); }

// Two densities: D multiplies the one-electron terms and (D - P/2) the two-electron terms built from P, which is
// the XL-BOMD shadow energy  E = sum D o F(P) - 1/2 (F(P) - h) o P + E_nuc  (energy.py:76-88, xlbomd.py:430-447);
// with D == P it is the ordinary SCF energy.  (No __restrict__ on D/P: they may alias.)
template <int CLS>
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS2(128, (CLS >= 1) ? SEQM_PG_MINB : 0) pair_gradient_kernel(seqm_batch_t b, const double* __restrict__ xyz, const double* D, const double* P,
                                      double* __restrict__ gpair) {
  constexpr bool hi = (CLS >= 1), hj = (CLS == 2);
  constexpr int ni = hi ? 4 : 1, nj = hj ? 4 : 1;
  const int q0 = b.pair_cls_off[CLS], q1 = b.pair_cls_off[CLS + 1];
  for (int q = q0 + blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += gridDim.x * blockDim.x) {
    const int p = b.pair_perm[q];
    const int i = b.pair_i[p], j = b.pair_j[p];
    const MolView v = mol_view(b, b.atom_mol[i]);
    const double* Pm = P + v.mat0;
    const double* Dm = D + v.mat0;
    const int n = v.n, oi = orb_off(v, i - v.a0), oj = orb_off(v, j - v.a0);
    // PM6 pair with a d atom: this kernel differentiates its sp x sp part only (two-electron sp block, core attraction
    // of the sp products, core-core); the resonance term (9 x 9 overlaps) and everything with a d orbital is added by
    // spd_pair_gradient_kernel
    const bool ypair = (i - v.a0) < v.nsh;
    PairGeom<double> g;
    pair_geom(xyz, i, j, g);
    if (pair_cut(b, g.r)) {
      gpair[3 * (long long)p] = gpair[3 * (long long)p + 1] = gpair[3 * (long long)p + 2] = 0.0;
      continue;
    }
    const double dist = g.r * SEQM_A0;
    const Dual1 r1(g.r, 1.0);
    double dEdr = 0.0, dEde[3] = {0.0, 0.0, 0.0};

    // ---- resonance integrals: Q_mu,nu = D_mu,nu (beta_mu^A + beta_nu^B)
    if (!ypair && g.r <= SEQM_OVERLAP_CUTOFF) {
      const double bsi = par(b, SEQM_P_BS, i), bpi = par(b, SEQM_P_BP, i);
      const double bsj = par(b, SEQM_P_BS, j), bpj = par(b, SEQM_P_BP, j);
      const int na = (int)par(b, SEQM_P_QN, i), nb = (int)par(b, SEQM_P_QN, j);
      const double zsa = par(b, SEQM_P_ZS, i), zpa = par(b, SEQM_P_ZP, i);
      const double zsb = par(b, SEQM_P_ZS, j), zpb = par(b, SEQM_P_ZP, j);
      const Dual1 ss = sto_overlap(c_ovl, na, nb, 0, zsa, zsb, r1);
      dEdr += Dm[oi * n + oj] * (bsi + bsj) * ss.d;
      if (hi) {
        const Dual1 os = sto_overlap(c_ovl, na, nb, 1, zpa, zsb, r1);
        for (int k = 0; k < 3; ++k) {
          const double q = Dm[(oi + k + 1) * n + oj] * (bpi + bsj);
          dEdr += q * os.d * g.e[k];
          dEde[k] += q * os.v;
        }
      }
      if (hj) {
        const Dual1 so = sto_overlap(c_ovl, na, nb, 2, zsa, zpb, r1);
        for (int k = 0; k < 3; ++k) {
          const double q = Dm[oi * n + oj + k + 1] * (bsi + bpj);
          dEdr += q * so.d * g.e[k];
          dEde[k] += q * so.v;
        }
      }
      if (hi && hj) {
        const Dual1 oo = sto_overlap(c_ovl, na, nb, 3, zpa, zpb, r1);
        const Dual1 pp = sto_overlap(c_ovl, na, nb, 4, zpa, zpb, r1);
        const double bb = bpi + bpj;
        double tr = 0.0, qee = 0.0;
        for (int k = 0; k < 3; ++k) {
          tr += Dm[(oi + k + 1) * n + oj + k + 1];
          double row = 0.0;
          for (int l = 0; l < 3; ++l) {
            const double qs = Dm[(oi + k + 1) * n + oj + l + 1] + Dm[(oi + l + 1) * n + oj + k + 1];
            row += qs * g.e[l];
            qee += Dm[(oi + k + 1) * n + oj + l + 1] * g.e[k] * g.e[l];
          }
          dEde[k] += bb * (oo.v - pp.v) * row;
        }
        dEdr += bb * ((oo.d - pp.d) * qee + pp.d * tr);
      }
    }

    // ---- two-electron + core-attraction terms: E_2e = sum C[kl][mn] w[kl][mn]
    constexpr int nA = hi ? 10 : 1, nB = hj ? 10 : 1;
    constexpr int nint = (hi && hj) ? 22 : (hi ? 4 : 1);
    double Cm[10][10];
    {
      double pa[10], pb[10], da[10], db[10];  // weighted packed diagonal blocks of P and D
      #pragma unroll
      for (int kl = 0; kl < 10; ++kl) {
        int mu = 0;
        while ((mu + 1) * (mu + 2) / 2 <= kl) ++mu;
        const int nu = kl - mu * (mu + 1) / 2;
        const double wt = (mu == nu) ? 1.0 : 2.0;
        pa[kl] = (kl < nA) ? wt * Pm[(oi + mu) * n + oi + nu] : 0.0;
        pb[kl] = (kl < nB) ? wt * Pm[(oj + mu) * n + oj + nu] : 0.0;
        da[kl] = (kl < nA) ? wt * Dm[(oi + mu) * n + oi + nu] : 0.0;
        db[kl] = (kl < nB) ? wt * Dm[(oj + mu) * n + oj + nu] : 0.0;
      }
      const double ti = par(b, SEQM_P_TORE, i), tj = par(b, SEQM_P_TORE, j);
      #pragma unroll
      for (int kl = 0; kl < nA; ++kl)
        #pragma unroll
        for (int mn = 0; mn < nB; ++mn)
          Cm[kl][mn] = (da[kl] - 0.5 * pa[kl]) * pb[mn] + pa[kl] * (db[mn] - 0.5 * pb[mn]);
      #pragma unroll
      for (int kl = 0; kl < nA; ++kl) Cm[kl][0] -= tj * da[kl];
      #pragma unroll
      for (int mn = 0; mn < nB; ++mn) Cm[0][mn] -= ti * db[mn];
      #pragma unroll
      for (int mu = 0; mu < ni; ++mu)
        #pragma unroll
        for (int nu = 0; nu < ni; ++nu)
          #pragma unroll
          for (int la = 0; la < nj; ++la)
            #pragma unroll
            for (int sg = 0; sg < nj; ++sg)
              Cm[pack2(mu, nu)][pack2(la, sg)] -=
                  (Dm[(oi + mu) * n + oj + la] - 0.5 * Pm[(oi + mu) * n + oj + la]) * Pm[(oi + nu) * n + oj + sg];
    }
    Dual1 ri[22];
    local_integrals(r1, load_multipole(b, i), load_multipole(b, j), nint, ri);
    if (nint == 1) {
      dEdr += Cm[0][0] * ri[0].d;
    } else {
      double vdir[3] = {-g.e[0], -g.e[1], -g.e[2]};
      double rot[3][3], Tm[10][10];
      rotation_rows(vdir, rot);
      pair_transform(rot, Tm);
      // U2 = T C (rows: local pair index, cols: molecular mn) ; U1 = T C^t
      double U1[10][10], U2[10][10], GT[10][10];
      #pragma unroll
      for (int a = 0; a < 10; ++a)
        #pragma unroll
        for (int c = 0; c < 10; ++c) { U1[a][c] = 0.0; U2[a][c] = 0.0; GT[a][c] = 0.0; }
      #pragma unroll
      for (int KL = 0; KL < nA; ++KL)
        #pragma unroll
        for (int kl = 0; kl < nA; ++kl) {
          if (cls_of(KL) != cls_of(kl)) continue;
          const double t = Tm[KL][kl];
          #pragma unroll
          for (int c = 0; c < nB; ++c) U2[KL][c] += t * Cm[kl][c];
        }
      #pragma unroll
      for (int MN = 0; MN < nB; ++MN)
        #pragma unroll
        for (int mn = 0; mn < nB; ++mn) {
          if (cls_of(MN) != cls_of(mn)) continue;
          const double t = Tm[MN][mn];
          #pragma unroll
          for (int c = 0; c < nA; ++c) U1[MN][c] += t * Cm[c][mn];
        }
      #pragma unroll
      for (int e = 0; e < SEQM_NL; ++e) {
        const LEntry le = l_entry(e);
        if (le.k >= nint || le.mn >= nB) continue;
        const int cK = cls_of(le.kl), cM = cls_of(le.mn);
        double cl = 0.0;  // (T C T^t)[KL][MN]
        #pragma unroll
        for (int c = 0; c < nB; ++c)
          if (cls_of(c) == cM) cl += U2[le.kl][c] * Tm[le.mn][c];
        dEdr += cl * ri[le.k].d;
        const double lv = ri[le.k].v;
        #pragma unroll
        for (int c = 0; c < nA; ++c)
          if (cls_of(c) == cK) GT[le.kl][c] += lv * U1[le.mn][c];
        #pragma unroll
        for (int c = 0; c < nB; ++c)
          if (cls_of(c) == cM) GT[le.mn][c] += lv * U2[le.kl][c];
      }
      // dE/drot from dE/dT
      double Gr[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      #pragma unroll
      for (int a = 0; a < 3; ++a)
        #pragma unroll
        for (int k = 0; k < 3; ++k) Gr[a][k] += GT[pack2(a + 1, 0)][pack2(k + 1, 0)];
      #pragma unroll
      for (int a = 0; a < 3; ++a)
        #pragma unroll
        for (int c = 0; c <= a; ++c)
          #pragma unroll
          for (int k = 0; k < 3; ++k)
            #pragma unroll
            for (int l = 0; l <= k; ++l) {
              const double gt = GT[pack2(a + 1, c + 1)][pack2(k + 1, l + 1)];
              Gr[a][k] += gt * rot[c][l];
              Gr[c][l] += gt * rot[a][k];
              if (a != c) {
                Gr[c][k] += gt * rot[a][l];
                Gr[a][l] += gt * rot[c][k];
              }
            }
      // d rot / d v by forward mode on the 3x3 rotation only ; e = -v
      Dual3 vd[3] = {Dual3(vdir[0], 1.0, 0.0, 0.0), Dual3(vdir[1], 0.0, 1.0, 0.0), Dual3(vdir[2], 0.0, 0.0, 1.0)};
      Dual3 rd[3][3];
      rotation_rows(vd, rd);
      #pragma unroll
      for (int a = 0; a < 3; ++a)
        #pragma unroll
        for (int k = 0; k < 3; ++k) {
          dEde[0] -= Gr[a][k] * rd[a][k].d0;
          dEde[1] -= Gr[a][k] * rd[a][k].d1;
          dEde[2] -= Gr[a][k] * rd[a][k].d2;
        }
    }
    // ---- core-core repulsion (depends on r only)
    {
      double alp, chi;
      pair_pw(b, i, j, alp, chi);
      dEdr += core_core(b.method, b.atom_Z[i], b.atom_Z[j], load_core(b, i), load_core(b, j), r1, ri[0], alp, chi).d;
    }
    // ---- chain rule: X = R_j - R_i, r = |X|/a0, e = X/|X| ; gradient with respect to R_i is -dE/dX
    const double ede = dEde[0] * g.e[0] + dEde[1] * g.e[1] + dEde[2] * g.e[2];
    for (int c = 0; c < 3; ++c)
      gpair[3 * (long long)p + c] = -(dEdr * g.e[c] * (1.0 / SEQM_A0) + (dEde[c] - ede * g.e[c]) / dist);
  }
}

// Forward-mode reference implementation of the same quantity (Dual3 through the whole pair code).
SEQM_GLOBAL void pair_gradient_forward_kernel(seqm_batch_t b, const double* __restrict__ xyz, const double* __restrict__ P,
                                              double* __restrict__ gpair) {

  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < b.npairs; p += gridDim.x * blockDim.x) {
    const int i = b.pair_i[p], j = b.pair_j[p];
    const MolView v = mol_view(b, b.atom_mol[i]);
    const double* Pm = P + v.mat0;
    const int n = v.n, oi = orb_off(v, i - v.a0), oj = orb_off(v, j - v.a0);
    const int ni = orb_cnt(v, i - v.a0), nj = orb_cnt(v, j - v.a0);
    PairGeom<Dual3> g;
    pair_geom(xyz, i, j, g);
    if (pair_cut(b, g.r.v)) {
      gpair[3 * (long long)p] = gpair[3 * (long long)p + 1] = gpair[3 * (long long)p + 2] = 0.0;
      continue;
    }
    Dual3 E(0.0);
    {  // resonance term
      Dual3 S[4][4];
      pair_overlap(b, i, j, g, S);
      const double bsi = par(b, SEQM_P_BS, i), bpi = par(b, SEQM_P_BP, i);
      const double bsj = par(b, SEQM_P_BS, j), bpj = par(b, SEQM_P_BP, j);
      for (int mu = 0; mu < ni; ++mu)
        for (int nu = 0; nu < nj; ++nu)
          E += (Pm[(oi + mu) * n + oj + nu] * ((mu ? bpi : bsi) + (nu ? bpj : bsj))) * S[mu][nu];
    }
    Dual3 w[10][10];
    pair_w(b, i, j, g, w, (ni == 4 && nj == 4) ? 22 : (ni == 4 ? 4 : 1));
    const int nA = (ni == 4) ? 10 : 1, nB = (nj == 4) ? 10 : 1;
    double pa[10], pb[10];  // weighted packed diagonal-block densities
    for (int kl = 0; kl < 10; ++kl) {
      int mu = 0;
      while ((mu + 1) * (mu + 2) / 2 <= kl) ++mu;
      const int nu = kl - mu * (mu + 1) / 2;
      const double wt = (mu == nu) ? 1.0 : 2.0;
      pa[kl] = (kl < nA) ? wt * Pm[(oi + mu) * n + oi + nu] : 0.0;
      pb[kl] = (kl < nB) ? wt * Pm[(oj + mu) * n + oj + nu] : 0.0;
    }
    const double ti = par(b, SEQM_P_TORE, i), tj = par(b, SEQM_P_TORE, j);
    for (int kl = 0; kl < nA; ++kl) {
      E += (-tj * pa[kl]) * w[kl][0];  // electrons on i, core of j
      for (int mn = 0; mn < nB; ++mn) E += (pa[kl] * pb[mn]) * w[kl][mn];
    }
    for (int mn = 0; mn < nB; ++mn) E += (-ti * pb[mn]) * w[0][mn];
    for (int mu = 0; mu < ni; ++mu)
      for (int la = 0; la < nj; ++la)
        for (int nu = 0; nu < ni; ++nu)
          for (int sg = 0; sg < nj; ++sg)
            E += (-0.5 * Pm[(oi + mu) * n + oj + la] * Pm[(oi + nu) * n + oj + sg]) * w[pack2(mu, nu)][pack2(la, sg)];
    {
      double alp, chi;
      pair_pw(b, i, j, alp, chi);
      E += core_core(b.method, b.atom_Z[i], b.atom_Z[j], load_core(b, i), load_core(b, j), g.r, w[0][0], alp, chi);
    }
    gpair[3 * (long long)p] = E.d0;
    gpair[3 * (long long)p + 1] = E.d1;
    gpair[3 * (long long)p + 2] = E.d2;
  }
}
