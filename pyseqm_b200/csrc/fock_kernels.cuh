// fock_kernels.cuh -- one CTA per molecule, density staged in shared memory.
//   fock_kernel         F = Hcore + one-centre + two-centre J/K        (fock.py:132-347)
//   elec_energy_kernel  Eelec = 1/2 sum P o (H + F)                     (energy.py:26-53)
// Diagonal blocks are accumulated atom-centrically (each (atom, packed kl) work item sums over all
// partner atoms), so no atomics and a fixed summation order; each off-diagonal block belongs to one pair.
#pragma once
#include "common.cuh"

SEQM_GLOBAL void fock_kernel(seqm_batch_t b, const double* __restrict__ P, const double* __restrict__ H,
                             const double* __restrict__ w, double* __restrict__ F, const int32_t* __restrict__ active) {
  const int m = b.mol_order[blockIdx.x];
  if (active && !active[m]) return;
  const MolView v = mol_view(b, m);
  const int n = v.n;
  SEQM_DYN_SMEM(double, sP);
  const double* Pm = P + v.mat0;
  const double* Hm = H + v.mat0;
  double* Fm = F + v.mat0;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) sP[t] = Pm[t];
  SEQM_SYNC();
  // exchange: F_AB[mu,la] = H_AB[mu,la] - 1/2 sum_{nu in A, sg in B} P_AB[nu,sg] w[pack(mu,nu)][pack(la,sg)]
  for (int t = threadIdx.x; t < v.npair * 16; t += blockDim.x) {
    const int pl = t >> 4, mu = (t >> 2) & 3, la = t & 3;
    const int p = v.p0 + pl;
    const int i = b.pair_i[p] - v.a0, j = b.pair_j[p] - v.a0;
    const int ni = orb_cnt(v, i), nj = orb_cnt(v, j);
    if (mu >= ni || la >= nj) continue;
    const int oi = orb_off(v, i), oj = orb_off(v, j);
    const double* wp = w + (long long)p * 100;
    double k = 0.0;
    for (int nu = 0; nu < ni; ++nu)
      for (int sg = 0; sg < nj; ++sg) k += sP[(oi + nu) * n + oj + sg] * wp[pack2(mu, nu) * 10 + pack2(la, sg)];
    const int r = oi + mu, c = oj + la;
    const double f = Hm[r * n + c] - 0.5 * k;
    Fm[r * n + c] = f;
    Fm[c * n + r] = f;
  }
  // diagonal blocks
  for (int t = threadIdx.x; t < v.na * 10; t += blockDim.x) {
    const int a = t / 10, kl = t % 10;
    if (a >= v.nheavy && kl > 0) continue;
    int mu = 0;
    while ((mu + 1) * (mu + 2) / 2 <= kl) ++mu;
    const int nu = kl - mu * (mu + 1) / 2;  // mu >= nu
    const int oa = orb_off(v, a), ga = v.a0 + a;
    const double gss = par(b, SEQM_P_GSS, ga), gsp = par(b, SEQM_P_GSP, ga), gpp = par(b, SEQM_P_GPP, ga);
    const double gp2 = par(b, SEQM_P_GP2, ga), hsp = par(b, SEQM_P_HSP, ga);
    const double Pss = sP[oa * n + oa];
    double Ppt = 0.0;
    if (a < v.nheavy) Ppt = sP[(oa + 1) * n + oa + 1] + sP[(oa + 2) * n + oa + 2] + sP[(oa + 3) * n + oa + 3];
    double g;
    if (mu == 0)
      g = 0.5 * Pss * gss + Ppt * (gsp - 0.5 * hsp);
    else if (nu == 0)
      g = sP[oa * n + oa + mu] * (1.5 * hsp - 0.5 * gsp);
    else if (mu == nu) {
      const double Pk = sP[(oa + mu) * n + oa + mu];
      g = Pss * (gsp - 0.5 * hsp) + 0.5 * Pk * gpp + (Ppt - Pk) * (1.25 * gp2 - 0.25 * gpp);
    } else
      g = sP[(oa + nu) * n + oa + mu] * (0.75 * gpp - 1.25 * gp2);
    // Coulomb from every other atom o:  sum_mn wt_mn P_oo[mn] (kl on a | mn on o)
    for (int o = 0; o < v.na; ++o) {
      if (o == a) continue;
      const int oo = orb_off(v, o), no = orb_cnt(v, o);
      const bool first = a < o;
      const double* wp = w + (long long)(v.p0 + (first ? pair_local(v, a, o) : pair_local(v, o, a))) * 100;
      const int sk = first ? 10 : 1, sm = first ? 1 : 10;  // strides of (kl, mn) in this pair's w
      double j = sP[oo * n + oo] * wp[kl * sk];
      if (no == 4) {
        for (int x = 1; x < 4; ++x) {
          j += 2.0 * sP[oo * n + oo + x] * wp[kl * sk + pack2(x, 0) * sm];
          for (int y = 1; y <= x; ++y)
            j += (x == y ? 1.0 : 2.0) * sP[(oo + y) * n + oo + x] * wp[kl * sk + pack2(x, y) * sm];
        }
      }
      g += j;
    }
    const double f = Hm[(oa + mu) * n + oa + nu] + g;
    Fm[(oa + mu) * n + oa + nu] = f;
    Fm[(oa + nu) * n + oa + mu] = f;
  }
}

SEQM_GLOBAL void elec_energy_kernel(seqm_batch_t b, const double* __restrict__ P, const double* __restrict__ H,
                                    const double* __restrict__ F, double* __restrict__ E,
                                    const int32_t* __restrict__ active) {
  __shared__ double red[33];
  const int m = b.mol_order[blockIdx.x];
  if (active && !active[m]) return;
  const MolView v = mol_view(b, m);
  const int nn = v.n * v.n;
  double s = 0.0;
  for (int t = threadIdx.x; t < nn; t += blockDim.x) s += P[v.mat0 + t] * (H[v.mat0 + t] + F[v.mat0 + t]);
  s = block_sum(s, red);
  if (threadIdx.x == 0) E[m] = 0.5 * s;
}
