// hestenes_kernels.cuh -- eigensolver + density for molecules between the shared-memory resident Jacobi kernel
// (n <= 118 orbitals, eig_kernels.cuh) and SEQM_MID_ORB = 256 orbitals: sym_eig_trunc (diag.py:110-241) for mid-size
// molecules such as a stacked coronene dimer (216 orbitals), north_star "one-sided Jacobi ... up to about 256 orbitals".
//
// One-sided (Hestenes) Jacobi, one CTA per molecule: F is shifted below its Gershgorin bound so that G = F - sigma I is
// positive definite; plane rotations applied from the left orthogonalise the ROWS of G (contiguous, coalesced); at
// convergence row k equals lambda_k v_k^t, i.e. the eigenvalue is the row norm (+ sigma) and the eigenvector the
// normalised row -- no separate eigenvector accumulation.  Row pairs follow the round-robin tournament (n/2 disjoint
// pairs per round, one warp per pair, one block barrier per round).  G lives in shared memory when n^2 doubles fit the
// opt-in limit (n <= ~165) and in the molecule's slot of the caller's packed eigenvector buffer (L2 resident) otherwise.
// The rotation count is O(n^3) per sweep with 6-9 sweeps: this path is for the handful of mid-size molecules a batch may
// hold, not a throughput kernel.
#pragma once
#include "common.cuh"

#define SEQM_MID_ORB 256
#define HEST_THREADS 256

SEQM_D double hest_warp_sum(double v, int wsz) {
#ifndef SEQM_HOSTEMU
  for (int o = wsz >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#else
  (void)wsz;
#endif
  return v;
}

// shared: [G n*n doubles if in_smem] | lam[n] | red[32] | rank[n] ints
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS(HEST_THREADS) hestenes_kernel(seqm_batch_t b, const double* __restrict__ F, double* __restrict__ P,
                                                                  double* __restrict__ evals, double* __restrict__ C,
                                                                  const int32_t* __restrict__ active, int in_smem) {
  const int mol = b.mol_order[blockIdx.x];
  if (active && !active[mol]) return;
  const MolView v = mol_view(b, mol);
  const int n = v.n, nn = n * n;
  SEQM_DYN_SMEM(double, sm);
  double* G = in_smem ? sm : C + v.mat0;
  double* lam = sm + (in_smem ? nn : 0);
  double* red = lam + n;
  int* rank = reinterpret_cast<int*>(red + 34);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int wsz = nthr < 32 ? nthr : 32, lane = tid % wsz, warp = tid / wsz, nw = nthr / wsz;
  const double* Fm = F + v.mat0;
  // Gershgorin bounds -> shift
  double lo = 1.0e300, hi = -1.0e300;
  for (int r = tid; r < n; r += nthr) {
    double off = 0.0;
    for (int c = 0; c < n; ++c) off += (c == r) ? 0.0 : fabs(Fm[r * n + c]);
    lo = fmin(lo, Fm[r * n + r] - off);
    hi = fmax(hi, Fm[r * n + r] + off);
  }
  lo = -block_max(-lo, red);
  hi = block_max(hi, red);
  const double sigma = lo - 0.05 * (hi - lo) - 1.0e-6;
  for (int t = tid; t < nn; t += nthr) G[t] = Fm[t] - ((t / n == t % n) ? sigma : 0.0);
  SEQM_SYNC();
  const int np_ = n + (n & 1);  // players of the tournament (a dummy when n is odd)
  for (int sweep = 0; sweep < 40; ++sweep) {
    double worst = 0.0;
    for (int round = 0; round < np_ - 1; ++round) {
      for (int k = warp; k < np_ / 2; k += nw) {
        int p, q;
        if (k == 0) {
          p = np_ - 1;
          q = round;
        } else {
          p = (round + k) % (np_ - 1);
          q = (round - k + (np_ - 1)) % (np_ - 1);
        }
        if (p >= n || q >= n) continue;
        double* gp = G + p * n;
        double* gq = G + q * n;
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int c = lane; c < n; c += wsz) {
          const double x = gp[c], y = gq[c];
          al += x * x;
          be += y * y;
          ga += x * y;
        }
        al = hest_warp_sum(al, wsz);
        be = hest_warp_sum(be, wsz);
        ga = hest_warp_sum(ga, wsz);
        const double scale = sqrt(al * be);
        const double rel = (scale > 0.0) ? fabs(ga) / scale : 0.0;
        worst = fmax(worst, rel);
        if (rel > 1.0e-16) {
          const double zeta = (be - al) / (2.0 * ga);
          const double t = ((zeta >= 0.0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          for (int c = lane; c < n; c += wsz) {
            const double x = gp[c], y = gq[c];
            gp[c] = cs * x - sn * y;
            gq[c] = sn * x + cs * y;
          }
        }
      }
      SEQM_SYNC();
    }
    worst = block_max(worst, red);
    if (worst < 2.0e-15) break;
  }
  // eigenvalues = row norms; ascending rank by counting
  for (int r = warp; r < n; r += nw) {
    double s = 0.0;
    for (int c = lane; c < n; c += wsz) s += G[r * n + c] * G[r * n + c];
    s = hest_warp_sum(s, wsz);
    if (lane == 0) lam[r] = sqrt(s);
  }
  SEQM_SYNC();
  for (int r = tid; r < n; r += nthr) {
    int rk = 0;
    const double lr = lam[r];
    for (int c = 0; c < n; ++c) rk += (lam[c] < lr || (lam[c] == lr && c < r)) ? 1 : 0;
    rank[r] = rk;
    if (evals) evals[(long long)mol * b.nmax + rk] = lr + sigma;
  }
  SEQM_SYNC();
  // eigenvectors as columns in ascending order.  G in shared memory: straight into C.  G in C's own slot: through P's slot.
  double* Cm = C + v.mat0;
  double* Pm = P + v.mat0;
  double* Vdst = in_smem ? Cm : Pm;
  for (int t = tid; t < nn; t += nthr) {
    const int k = t / n, r = t % n;
    Vdst[r * n + rank[k]] = G[k * n + r] / lam[k];
  }
  SEQM_SYNC();
  if (!in_smem) {
    for (int t = tid; t < nn; t += nthr) Cm[t] = Pm[t];
    SEQM_SYNC();
  }
  // density P = 2 C_occ C_occ^t
  const int nocc = v.nocc;
  for (int t = tid; t < nn; t += nthr) {
    const int r = t / n, c = t % n;
    if (c < r) continue;
    double s = 0.0;
    for (int k = 0; k < nocc; ++k) s += Cm[r * n + k] * Cm[c * n + k];
    Pm[r * n + c] = 2.0 * s;
    Pm[c * n + r] = 2.0 * s;
  }
}
