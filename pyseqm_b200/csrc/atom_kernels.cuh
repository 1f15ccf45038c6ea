// atom_kernels.cuh -- light per-atom / per-molecule kernels around the pair code (primary translation unit):
//   atom_multipoles_kernel   per atom      dd, qq, rho0..2
//   hcore_kernel             per molecule  packed Hcore from U, w[:,0]/w[0,:] core attraction, overlap blocks
//   pair_sum_kernel          per molecule  sum of a per-pair quantity
//   atom_gradient_kernel     per atom      deterministic +/- gather of the pair gradients
//   elec_energy_xl_kernel, xl_propagate_kernel   XL-BOMD shadow energy and field propagation
#pragma once
#include "common.cuh"

SEQM_GLOBAL void atom_multipoles_kernel(seqm_batch_t b) {
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < b.nat; a += gridDim.x * blockDim.x) {
    AtomMultipole m = atom_multipole(b.atom_Z[a], par(b, SEQM_P_QN, a), par(b, SEQM_P_ZS, a), par(b, SEQM_P_ZP, a),
                                     par(b, SEQM_P_GSS, a), par(b, SEQM_P_GPP, a), par(b, SEQM_P_GP2, a),
                                     par(b, SEQM_P_HSP, a));
    b.atom_par[(long long)SEQM_P_DD * b.nat + a] = m.dd;
    b.atom_par[(long long)SEQM_P_QQ * b.nat + a] = m.qq;
    b.atom_par[(long long)SEQM_P_RHO0 * b.nat + a] = m.rho0;
    b.atom_par[(long long)SEQM_P_RHO1 * b.nat + a] = m.rho1;
    b.atom_par[(long long)SEQM_P_RHO2 * b.nat + a] = m.rho2;
  }
}

// `slices` CTAs per molecule (1 for the shared-memory sized molecules, more for the large ones whose single CTA would
// walk hundreds of thousands of pairs alone): packed, fully symmetric Hcore.
SEQM_GLOBAL void hcore_kernel(seqm_batch_t b, const double* __restrict__ w, const double* __restrict__ hab,
                              double* __restrict__ H, int slices) {
  const MolView v = mol_view(b, b.mol_order[blockIdx.x / slices]);
  double* Hm = H + v.mat0;
  const int n = v.n;
  const int tid0 = (blockIdx.x % slices) * blockDim.x + threadIdx.x, tstep = slices * blockDim.x;
  // off-diagonal blocks (both triangles)
  for (int t = tid0; t < v.npair * 16; t += tstep) {
    const int pl = t >> 4, mu = (t >> 2) & 3, nu = t & 3;
    const int p = v.p0 + pl;
    const int i = b.pair_i[p] - v.a0, j = b.pair_j[p] - v.a0;
    if (mu >= orb_cnt(v, i) || nu >= orb_cnt(v, j)) continue;
    const double h = hab[(long long)p * 16 + mu * 4 + nu];
    const int r = orb_off(v, i) + mu, c = orb_off(v, j) + nu;
    Hm[r * n + c] = h;
    Hm[c * n + r] = h;
  }
  // diagonal blocks: U + sum_B core attraction, -tore_B (kl|ss_B)
  for (int t = tid0; t < v.na * 10; t += tstep) {
    const int a = t / 10, kl = t % 10;
    if (a >= v.nheavy && kl > 0) continue;
    int mu = 0, nu = 0;  // kl = pack2(mu, nu), mu >= nu
    while ((mu + 1) * (mu + 2) / 2 <= kl) ++mu;
    nu = kl - mu * (mu + 1) / 2;
    double acc = (mu == nu) ? (mu == 0 ? par(b, SEQM_P_USS, v.a0 + a) : par(b, SEQM_P_UPP, v.a0 + a)) : 0.0;
    for (int o = 0; o < v.na; ++o) {
      if (o == a) continue;
      const double to = par(b, SEQM_P_TORE, v.a0 + o);
      if (a < o)
        acc -= to * w[(long long)(v.p0 + pair_local(v, a, o)) * 100 + kl * 10];
      else
        acc -= to * w[(long long)(v.p0 + pair_local(v, o, a)) * 100 + kl];
    }
    const int oa = orb_off(v, a);
    Hm[(oa + mu) * n + oa + nu] = acc;
    Hm[(oa + nu) * n + oa + mu] = acc;
  }
}

// per-molecule sum of a per-pair quantity (pairs of a molecule are contiguous): deterministic, no atomics
SEQM_GLOBAL void pair_sum_kernel(seqm_batch_t b, const double* __restrict__ vals, double* __restrict__ out) {
  __shared__ double red[33];
  const MolView v = mol_view(b, blockIdx.x);
  double s = 0.0;
  for (int t = threadIdx.x; t < v.npair; t += blockDim.x) s += vals[v.p0 + t];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[v.m] = s;
}

// grad[a] = sum_{b>a} g(a,b) - sum_{b<a} g(b,a)   (anal_grad.py:213-221 without atomics)
SEQM_GLOBAL void atom_gradient_kernel(seqm_batch_t b, const double* __restrict__ gpair, double* __restrict__ grad) {
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < b.nat; a += gridDim.x * blockDim.x) {
    const MolView v = mol_view(b, b.atom_mol[a]);
    const int la = a - v.a0;
    double gx = 0.0, gy = 0.0, gz = 0.0;
    for (int o = 0; o < v.na; ++o) {
      if (o == la) continue;
      const long long p = v.p0 + (la < o ? pair_local(v, la, o) : pair_local(v, o, la));
      const double s = (la < o) ? 1.0 : -1.0;
      gx += s * gpair[3 * p];
      gy += s * gpair[3 * p + 1];
      gz += s * gpair[3 * p + 2];
    }
    grad[3 * a] = gx;
    grad[3 * a + 1] = gy;
    grad[3 * a + 2] = gz;
  }
}

// elec_energy_xl (energy.py:76-88): sum D o F - 1/2 (F - h) o P, one CTA per molecule
SEQM_GLOBAL void elec_energy_xl_kernel(seqm_batch_t b, const double* __restrict__ D, const double* __restrict__ P,
                                       const double* __restrict__ F, const double* __restrict__ H, double* __restrict__ E) {
  __shared__ double red[33];
  const MolView v = mol_view(b, b.mol_order[blockIdx.x]);
  const int nn = v.n * v.n;
  double s = 0.0;
  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    const double f = F[v.mat0 + t];
    s += D[v.mat0 + t] * f - 0.5 * (f - H[v.mat0 + t]) * P[v.mat0 + t];
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) E[v.m] = s;
}

// XL-BOMD field propagation (MolecularDynamics.py:1418-1435 `_propagate_P`), one pass over the packed buffers:
//   P(n+1) = kappa [c D + (1 - c) P(n)] + sum_j coef_j Pt_j ;  Pt[slot] <- P(n+1)
#define SEQM_XL_MAXHIST 16
SEQM_GLOBAL void xl_propagate_kernel(long long total, double kappa, double c, const double* __restrict__ D,
                                     const double* __restrict__ Pin, double* __restrict__ Pt, const double* __restrict__ coef,
                                     int m, int slot, double* __restrict__ Pout) {
  double cf[SEQM_XL_MAXHIST];
  for (int j = 0; j < SEQM_XL_MAXHIST; ++j) cf[j] = (j < m) ? coef[j] : 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    double s = kappa * (c * D[t] + (1.0 - c) * Pin[t]);
    for (int j = 0; j < m; ++j) s += cf[j] * Pt[(long long)j * total + t];
    Pout[t] = s;
    Pt[(long long)slot * total + t] = s;
  }
}
