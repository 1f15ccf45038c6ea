// eig_kernels.cuh -- batched small-matrix symmetric eigensolver and density builders, one CTA per molecule
// with the matrices resident in shared memory.
//   jacobi_density_kernel   two-sided cyclic Jacobi (round-robin parallel ordering, 2x2 tile updates,
//                           optional warm start from a previous eigenbasis) + P = 2 C_occ C_occ^t
//                           -> replaces sym_eig_trunc (diag.py:110-241: pack, padded eigh, per-molecule
//                              matmul map, unpack)
//   sp2_kernel              SP2 purification at native size (SP2.py:9-85)
// Per rotation step all n/2 disjoint pairs are annihilated at once: thread (k,l) owns the 2x2 tile
// rows {p_k,q_k} x cols {p_l,q_l} and applies R_k^t . tile . R_l in place, so one barrier per step.
#pragma once
#include "common.cuh"

#define SEQM_JACOBI_MAX_SWEEPS 40

// circle-method pairing: round s of m-1, pair k of m/2  (m even)
SEQM_HD void rr_pair(int m, int s, int k, int& p, int& q) {
  if (k == 0) {
    p = m - 1;
    q = s;
  } else {
    p = (s + k) % (m - 1);
    q = (s - k + (m - 1)) % (m - 1);
  }
  if (p > q) { int t = p; p = q; q = t; }
}

// shared layout: A[n*n] | V[n*n] | cs[2*(m/2)] | scratch
SEQM_GLOBAL void jacobi_density_kernel(seqm_batch_t b, const double* __restrict__ F, double* __restrict__ Pout,
                                       double* __restrict__ evals, double* __restrict__ Cout,
                                       const double* __restrict__ Cguess, const int32_t* __restrict__ active) {
  const int mol = b.mol_order[blockIdx.x];
  if (active && !active[mol]) return;
  const MolView v = mol_view(b, mol);
  const int n = v.n;
  const int m = (n + 1) & ~1;  // even number of tournament slots; slot n (if any) is a bye
  const int np = m / 2;
  SEQM_DYN_SMEM(double, sm);
  double* A = sm;
  double* V = A + n * n;
  double* cs = V + n * n;      // c[k], s[k]
  double* scr = cs + 2 * np;   // 40 doubles of scratch
  int* perm = reinterpret_cast<int*>(scr + 40);  // n ints
  const double* Fm = F + v.mat0;

  if (Cguess) {
    // warm start: V = C0, A = C0^t F C0 (A used as scratch for T = F C0 first would need a third matrix;
    // instead form A column block by column block through global F, which is L1/L2 resident)
    const double* C0 = Cguess + v.mat0;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) V[t] = C0[t];
    SEQM_SYNC();
    // T = F V  -> A
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int i = t / n, j = t % n;
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += Fm[i * n + k] * V[k * n + j];
      A[t] = s;
    }
    SEQM_SYNC();
    // A <- V^t T, done in place row-block-wise is not possible; use Pout's global slot as scratch
    double* G = Pout + v.mat0;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int i = t / n, j = t % n;
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += V[k * n + i] * A[k * n + j];
      G[t] = s;
    }
    SEQM_SYNC();
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int i = t / n, j = t % n;
      A[t] = 0.5 * (G[t] + G[j * n + i]);
    }
  } else {
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      A[t] = Fm[t];
      V[t] = ((t / n) == (t % n)) ? 1.0 : 0.0;
    }
  }
  SEQM_SYNC();

  // scale for the convergence test
  double dmax = 0.0;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) dmax = fmax(dmax, fabs(A[t]));
  dmax = block_max(dmax, scr);
  const double tol = 1.0e-15 * fmax(dmax, 1.0e-300);

  for (int sweep = 0; sweep < SEQM_JACOBI_MAX_SWEEPS; ++sweep) {
    int rotated = 0;
    for (int s = 0; s < m - 1; ++s) {
      // rotation parameters for the np disjoint pairs
      int any = 0;
      for (int k = threadIdx.x; k < np; k += blockDim.x) {
        int p, q;
        rr_pair(m, s, k, p, q);
        double c = 1.0, sn = 0.0;
        if (q < n) {
          const double apq = A[p * n + q];
          if (fabs(apq) > tol) {
            const double app = A[p * n + p], aqq = A[q * n + q];
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            sn = t * c;
            any = 1;
          }
        }
        cs[2 * k] = c;
        cs[2 * k + 1] = sn;
      }
      any = seqm_sync_or(any);
      if (!any) continue;
      rotated = 1;
      // A <- J^t A J : tile (k,l) = rows {pk,qk} x cols {pl,ql}
      for (int t = threadIdx.x; t < np * np; t += blockDim.x) {
        const int k = t / np, l = t % np;
        int pk, qk, pl, ql;
        rr_pair(m, s, k, pk, qk);
        rr_pair(m, s, l, pl, ql);
        if (qk >= n || ql >= n) {
          // a bye slot: only the real row/column of the other pair rotates
          if (qk >= n && ql >= n) continue;
          if (qk >= n) {  // row pk untouched by rows; rotate its columns pl,ql
            const double c = cs[2 * l], sn = cs[2 * l + 1];
            const double x = A[pk * n + pl], y = A[pk * n + ql];
            A[pk * n + pl] = c * x - sn * y;
            A[pk * n + ql] = sn * x + c * y;
          } else {  // column pl untouched by columns; rotate rows pk,qk
            const double c = cs[2 * k], sn = cs[2 * k + 1];
            const double x = A[pk * n + pl], y = A[qk * n + pl];
            A[pk * n + pl] = c * x - sn * y;
            A[qk * n + pl] = sn * x + c * y;
          }
          continue;
        }
        const double ck = cs[2 * k], sk = cs[2 * k + 1], cl = cs[2 * l], sl = cs[2 * l + 1];
        const double a00 = A[pk * n + pl], a01 = A[pk * n + ql], a10 = A[qk * n + pl], a11 = A[qk * n + ql];
        // columns: [x y] -> [c x - s y, s x + c y]
        const double b00 = cl * a00 - sl * a01, b01 = sl * a00 + cl * a01;
        const double b10 = cl * a10 - sl * a11, b11 = sl * a10 + cl * a11;
        // rows
        A[pk * n + pl] = ck * b00 - sk * b10;
        A[qk * n + pl] = sk * b00 + ck * b10;
        A[pk * n + ql] = ck * b01 - sk * b11;
        A[qk * n + ql] = sk * b01 + ck * b11;
      }
      // V <- V J
      for (int t = threadIdx.x; t < n * np; t += blockDim.x) {
        const int i = t / np, l = t % np;
        int pl, ql;
        rr_pair(m, s, l, pl, ql);
        if (ql >= n) continue;
        const double c = cs[2 * l], sn = cs[2 * l + 1];
        const double x = V[i * n + pl], y = V[i * n + ql];
        V[i * n + pl] = c * x - sn * y;
        V[i * n + ql] = sn * x + c * y;
      }
      SEQM_SYNC();
    }
    if (!rotated) break;
  }

  // rank eigenvalues (ascending, ties by index): perm[rank] = column
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double ei = A[i * n + i];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const double ej = A[j * n + j];
      r += (ej < ei) || (ej == ei && j < i);
    }
    perm[r] = i;
  }
  SEQM_SYNC();
  if (evals) {
    for (int r = threadIdx.x; r < b.nmax; r += blockDim.x)
      evals[(long long)mol * b.nmax + r] = (r < n) ? A[perm[r] * n + perm[r]] : 0.0;
  }
  if (Cout) {
    double* Cm = Cout + v.mat0;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) Cm[t] = V[(t / n) * n + perm[t % n]];
  }
  if (Pout) {
    double* Pm = Pout + v.mat0;
    const int nocc = v.nocc;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int i = t / n, j = t % n;
      if (j < i) continue;
      double s = 0.0;
      for (int r = 0; r < nocc; ++r) {
        const int c = perm[r];
        s += V[i * n + c] * V[j * n + c];
      }
      s *= 2.0;
      Pm[i * n + j] = s;
      Pm[j * n + i] = s;
    }
  }
}
static inline size_t jacobi_smem_bytes(int n) {
  const int m = (n + 1) & ~1;
  return sizeof(double) * ((size_t)2 * n * n + m + 40) + sizeof(int) * (n + 2);
}

// SP2 purification of one molecule per CTA: X and X^2 in shared memory.
SEQM_GLOBAL void sp2_kernel(seqm_batch_t b, const double* __restrict__ F, double* __restrict__ Pout, double eps,
                            int32_t* __restrict__ niter, const int32_t* __restrict__ active) {
  const int mol = b.mol_order[blockIdx.x];
  if (active && !active[mol]) return;
  const MolView v = mol_view(b, mol);
  const int n = v.n;
  SEQM_DYN_SMEM(double, sm);
  double* X = sm;
  double* X2 = X + n * n;
  double* scr = X2 + n * n;  // 40 doubles
  const double* Fm = F + v.mat0;
  if (eps > 1.0e-3) eps = 1.0e-3;
  if (eps < 1.0e-7) eps = 1.0e-7;
  // Gershgorin bounds
  double lo = 1.0e300, hi = -1.0e300;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double r = 0.0;
    for (int j = 0; j < n; ++j) r += fabs(Fm[i * n + j]);
    const double aii = Fm[i * n + i];
    r -= fabs(aii);
    lo = fmin(lo, aii - r);
    hi = fmax(hi, aii + r);
  }
  const double hN = block_max(hi, scr);
  const double h1 = -block_max(-lo, scr);
  for (int t = threadIdx.x; t < n * n; t += blockDim.x)
    X[t] = ((((t / n) == (t % n)) ? hN : 0.0) - Fm[t]) / (hN - h1);
  SEQM_SYNC();
  const double nocc = (double)v.nocc;
  double tr = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) tr += X[i * n + i];
  tr = block_sum(tr, scr);
  double errm0 = fabs(tr - nocc), errm1 = errm0;
  int k = 0;
  for (;;) {
    double t2 = 0.0;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int i = t / n, j = t % n;
      double s = 0.0;
      for (int q = 0; q < n; ++q) s += X[i * n + q] * X[q * n + j];
      X2[t] = s;
      if (i == j) t2 += s;
    }
    t2 = block_sum(t2, scr);
    const bool take_sq = fabs(t2 - nocc) < fabs(2.0 * tr - t2 - nocc);
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) X[t] = take_sq ? X2[t] : 2.0 * X[t] - X2[t];
    SEQM_SYNC();
    tr = take_sq ? t2 : 2.0 * tr - t2;
    // the reference re-sums the diagonal of the updated matrix; do the same for identical rounding
    double tr2 = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tr2 += X[i * n + i];
    tr = block_sum(tr2, scr);
    errm1 = errm0;
    errm0 = fabs(tr - nocc);
    ++k;
    if ((errm0 < eps && errm1 < eps) || k >= 10000) break;
  }
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) Pout[v.mat0 + t] = 2.0 * X[t];
  if (niter && threadIdx.x == 0) niter[mol] = k;
}
