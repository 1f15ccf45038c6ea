// scf_driver.cuh -- the SCF fixed-point iteration kept on the device: mixing, Pulay DIIS, convergence
// bookkeeping; the host only launches and polls two integers per iteration.
// Replaces seqm/seqm_functions/scf_loop.py: get_error 106-147, scf_forward0 164-347, adaptive_mix 350-420,
// scf_forward1 424-635, scf_forward2 639-1132 (nFock = 10, batch-global DIIS reset), MAX_ITER = 1000.
#pragma once
#include "eig_kernels.cuh"
#include "fock_kernels.cuh"
#include "large_kernels.cuh"

#define SEQM_NFOCK 10
#define SEQM_EM (SEQM_NFOCK + 1)

struct ScfCtrl {  // device-resident control block (one per half-batch on the pipelined DIIS path)
  int nnot;       // molecules still not converged after the last get_error
  int reset;      // DIIS: some active molecule had cond > 1e7 this iteration (synchronous path)
  int counter;    // DIIS ring slot of the current iteration (pipelined path: advanced by diis_begin_kernel)
  int cF;         // DIIS history depth of the current iteration
  int pad[4];
};

struct ScfWork {  // carved out of the caller's workspace
  double *Pold, *Pnew, *C;
  double *FOCK, *RES, *EMAT, *coeff, *diis_err;  // converger 2
  double *old2, *dnew, *fac, *sum0, *sum3;       // converger 1
  double *err, *dm_err, *dm_elem, *Eel_new, *Eel_run;
  int32_t* active;
  ScfCtrl* ctrl;
  int32_t* order2;  // pipelined path: processing order of half A followed by half B
  int* rflag;       // pipelined path: DIIS reset flags [half][iteration parity]
  int has_C;
  // large-molecule path (n > SEQM_MAX_ORB): SP2 / commutator scratch for one molecule at a time
  double *Xl, *X2l, *rmax, *part;
  Sp2State* sp2st;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t scf_carve(const seqm_batch_t* b, const seqm_scf_opts_t* o, unsigned char* base, ScfWork* W) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = base ? base + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  const size_t mat = sizeof(double) * (size_t)b->mat_total, nm = sizeof(double) * (size_t)b->nmol;
  ScfWork w;
  memset(&w, 0, sizeof(w));
  w.Pold = (double*)take(mat);
  w.Pnew = (double*)take(mat);
  w.C = (double*)take(mat);
  if (o->converger == 2) {
    w.FOCK = (double*)take(mat * SEQM_NFOCK);
    w.RES = (double*)take(mat * SEQM_NFOCK);
    w.EMAT = (double*)take(nm * SEQM_EM * SEQM_EM);
    w.coeff = (double*)take(nm * SEQM_NFOCK);
    w.diis_err = (double*)take(nm);
  }
  if (o->converger == 1) {
    w.old2 = (double*)take(nm * b->nmax);
    w.dnew = (double*)take(nm * b->nmax);
    w.fac = (double*)take(nm);
    w.sum0 = (double*)take(nm);
    w.sum3 = (double*)take(nm);
  }
  w.err = (double*)take(nm);
  w.dm_err = (double*)take(nm);
  w.dm_elem = (double*)take(nm);
  w.Eel_new = (double*)take(nm);
  w.Eel_run = (double*)take(nm);
  w.active = (int32_t*)take(sizeof(int32_t) * (size_t)b->nmol);
  w.ctrl = (ScfCtrl*)take(2 * sizeof(ScfCtrl));
  w.order2 = (int32_t*)take(sizeof(int32_t) * (size_t)b->nmol);
  w.rflag = (int*)take(4 * sizeof(int));
  if (b->nmax > SEQM_MAX_ORB) {
    const size_t big = sizeof(double) * (size_t)b->nmax * b->nmax;
    w.Xl = (double*)take(big);
    w.X2l = (double*)take(big);
    w.rmax = (double*)take(sizeof(double));
    w.part = (double*)take(sizeof(double) * 3 * 4096);
    w.sp2st = (Sp2State*)take(sizeof(Sp2State));
  }
  if (W) *W = w;
  return off;
}

SEQM_GLOBAL void scf_init_kernel(seqm_batch_t b, ScfWork W, int converger) {
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < b.nmol; m += gridDim.x * blockDim.x) {
    W.err[m] = 1.0;
    W.dm_err[m] = 1.0;
    W.dm_elem[m] = 1.0;
    W.Eel_new[m] = 0.0;
    W.active[m] = 1;
    if (converger == 2) {
      W.diis_err[m] = 1.79769313486231570e308;
      double* E = W.EMAT + (long long)m * SEQM_EM * SEQM_EM;
      for (int i = 0; i < SEQM_EM; ++i)
        for (int j = 0; j < SEQM_EM; ++j) E[i * SEQM_EM + j] = (j < i) ? -1.0 : 0.0;
    }
    if (converger == 1)
      for (int k = 0; k < b.nmax; ++k) W.old2[(long long)m * b.nmax + k] = 0.0;
    if (m == 0) {
      for (int h = 0; h < 2; ++h) {
        W.ctrl[h].nnot = b.nmol;
        W.ctrl[h].reset = 0;
        W.ctrl[h].counter = -1;
        W.ctrl[h].cF = 0;
        for (int q = 0; q < 4; ++q) W.ctrl[h].pad[q] = 0;  // pad[0..1]: pass counters of the adaptive mixing
      }
      for (int i = 0; i < 4; ++i) W.rflag[i] = 0;
    }
    // interleaved split of the descending-size processing order: both halves see the same size mix
    const int nA = (b.nmol + 1) / 2;
    W.order2[(m & 1) ? nA + (m >> 1) : (m >> 1)] = b.mol_order[m];
  }
}
// Pipelined DIIS path, start of iteration k of one half-batch (b.mol_order / b.nmol describe the half, W.ctrl is
// its control block): apply the batch-global reset decided in iteration k-1 (scf_loop.py:1117-1130: any active
// molecule of EITHER half with cond > 1e7 empties everyone's history), advance the ring (scf_loop.py:993-996),
// clear this iteration's reset flag and the not-converged counter.
SEQM_GLOBAL void diis_begin_kernel(seqm_batch_t b, ScfWork W, int k, int half) {
  const int prev = (k - 1) & 1;
  const bool rst = (k > 0) && ((W.rflag[prev] | W.rflag[2 + prev]) != 0);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (rst)
    for (int i = tid; i < b.nmol; i += gridDim.x * blockDim.x) {
      double* E = W.EMAT + (long long)b.mol_order[i] * SEQM_EM * SEQM_EM;
      for (int r = 0; r < SEQM_EM; ++r)
        for (int c = 0; c < SEQM_EM; ++c) E[r * SEQM_EM + c] = (c < r) ? -1.0 : 0.0;
    }
  if (tid == 0) {
    int cF = W.ctrl->cF, counter = W.ctrl->counter;
    if (rst) {
      counter = -1;
      cF = 0;
    }
    W.ctrl->cF = (cF < SEQM_NFOCK) ? cF + 1 : SEQM_NFOCK;
    W.ctrl->counter = (counter + 1) % SEQM_NFOCK;
    W.ctrl->nnot = 0;
    W.rflag[2 * half + (k & 1)] = 0;
  }
}
SEQM_GLOBAL void emat_reset_kernel(seqm_batch_t b, ScfWork W) {
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < b.nmol; m += gridDim.x * blockDim.x) {
    double* E = W.EMAT + (long long)m * SEQM_EM * SEQM_EM;
    for (int i = 0; i < SEQM_EM; ++i)
      for (int j = 0; j < SEQM_EM; ++j) E[i * SEQM_EM + j] = (j < i) ? -1.0 : 0.0;
    if (m == 0) W.ctrl->reset = 0;
  }
}

// History slots of the shared-memory path hold the UPPER TRIANGLE of the stored Fock matrix (symmetric) and residual
// (antisymmetric) row-packed -- half the bytes that diis_store writes / reads back for the EMAT row and that
// diis_extrapolate streams; the slot stride stays n*n (the large-molecule path keeps full matrices in the same buffers).
SEQM_HD int diis_tri(int a, int c, int n) { return a * n - a * (a - 1) / 2 + (c - a); }  // a <= c

// DIIS step 1 (scf_loop.py:994-1007): store F, residual R = F P - P F, max |R|, EMAT row `counter`.
SEQM_GLOBAL void diis_store_kernel(seqm_batch_t b, ScfWork W, const double* __restrict__ F, const double* __restrict__ P,
                                   int counter, int cF) {
  __shared__ double red[33];
  const int mol = b.mol_order[blockIdx.x];
  if (!W.active[mol]) return;
  if (counter < 0) {  // pipelined path: the ring state lives on the device
    counter = W.ctrl->counter;
    cF = W.ctrl->cF;
  }
  const MolView v = mol_view(b, mol);
  const int n = v.n, nn = n * n;
  SEQM_DYN_SMEM(double, sm);
  const long long h0 = v.mat0 * SEQM_NFOCK;  // this molecule's history block: [slot][n*n]
  double* Fh = W.FOCK + h0 + (long long)counter * nn;
  double* Rh = W.RES + h0 + (long long)counter * nn;
  double dots[SEQM_NFOCK];
  for (int j = 0; j < SEQM_NFOCK; ++j) dots[j] = 0.0;
  double rmax = 0.0;
#ifndef SEQM_HOSTEMU
  // R = F P - P F on the FP64 tensor cores: F and P zero-padded to a multiple of 8 in shared memory (row stride
  // 4 mod 16), one warp per upper 8x8 tile with two accumulators (F P and P F); R is antisymmetric
  // row stride np8 + 4 is 4 or 12 mod 16 (conflict-free fragments); the largest size class (np8 = 120) only fits unpadded
  const int np8 = (n + 7) & ~7, nt8 = np8 >> 3, ld = (np8 > 112) ? np8 : np8 + 4;
  double* sF = sm;
  double* sP = sm + np8 * ld;
  for (int t = threadIdx.x; t < np8 * ld; t += blockDim.x) {
    const int r = t / ld, c = t - r * ld;
    const bool in = (r < n && c < n);
    const double f = in ? F[v.mat0 + r * n + c] : 0.0;
    sF[t] = f;
    sP[t] = in ? P[v.mat0 + r * n + c] : 0.0;
    if (in && c >= r) Fh[diis_tri(r, c, n)] = f;
  }
  SEQM_SYNC();
  {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int ntri = nt8 * (nt8 + 1) / 2;
    for (int tile = warp; tile < ntri; tile += nwarps) {
      int mi = 0, rem = tile;
      while (rem >= nt8 - mi) { rem -= nt8 - mi; ++mi; }
      const int i0 = mi * 8, j0 = (mi + rem) * 8;
      double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
      for (int k0 = 0; k0 < np8; k0 += 4) {
        const int ka = (i0 + g) * ld + k0 + t4, kb = (k0 + t4) * ld + j0 + g;
        seqm_dmma(x0, x1, sF[ka], sP[kb]);
        seqm_dmma(y0, y1, sP[ka], sF[kb]);
      }
      const int a = i0 + g;
      const double rv[2] = {x0 - y0, x1 - y1};
      for (int e = 0; e < 2; ++e) {
        const int c = j0 + 2 * t4 + e;
        if (a >= n || c >= n || c <= a) continue;
        const double sv = rv[e];
        const int at = diis_tri(a, c, n);
        Rh[at] = sv;
        rmax = fmax(rmax, fabs(sv));
        for (int q = 0; q < cF; ++q)
          dots[q] += sv * ((q == counter) ? sv : W.RES[h0 + (long long)q * nn + at]);
      }
    }
  }
#else
  double* sF = sm;
  double* sP = sm + nn;

  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    const double f = F[v.mat0 + t];
    sF[t] = f;
    sP[t] = P[v.mat0 + t];
    if (t % n >= t / n) Fh[diis_tri(t / n, t % n, n)] = f;
  }
  SEQM_SYNC();
  // R = F P - P F on the strict upper triangle (R is antisymmetric), 2x2 register blocks
  const int nb = (n + 1) >> 1;
  for (int t = threadIdx.x; t < nb * nb; t += blockDim.x) {
    const int bi = t / nb, bj = t % nb;
    if (bj < bi) continue;
    const int i = 2 * bi, j = 2 * bj;
    const int i1 = (i + 1 < n) ? i + 1 : i, j1 = (j + 1 < n) ? j + 1 : j;
    double r00 = 0.0, r01 = 0.0, r10 = 0.0, r11 = 0.0;
    for (int k = 0; k < n; ++k) {
      const double fi0 = sF[i * n + k], fi1 = sF[i1 * n + k], pi0 = sP[i * n + k], pi1 = sP[i1 * n + k];
      const double pj0 = sP[k * n + j], pj1 = sP[k * n + j1], fj0 = sF[k * n + j], fj1 = sF[k * n + j1];
      r00 += fi0 * pj0 - pi0 * fj0;
      r01 += fi0 * pj1 - pi0 * fj1;
      r10 += fi1 * pj0 - pi1 * fj0;
      r11 += fi1 * pj1 - pi1 * fj1;
    }
    // the (up to) four elements of the block that lie strictly above the diagonal
    const int ri[4] = {i, i, i1, i1}, rj[4] = {j, j1, j, j1};
    const double rv[4] = {r00, r01, r10, r11};
    for (int e = 0; e < 4; ++e) {
      const int a = ri[e], c = rj[e];
      if (c <= a) continue;
      if ((e == 1 || e == 3) && j1 == j) continue;  // clamped duplicate column
      if ((e == 2 || e == 3) && i1 == i) continue;  // clamped duplicate row
      const double sv = rv[e];
      const int at = diis_tri(a, c, n);
      Rh[at] = sv;
      rmax = fmax(rmax, fabs(sv));
      for (int q = 0; q < cF; ++q)
        dots[q] += sv * ((q == counter) ? sv : W.RES[h0 + (long long)q * nn + at]);
    }
  }
#endif
  for (int t = threadIdx.x; t < n; t += blockDim.x) Rh[diis_tri(t, t, n)] = 0.0;
  rmax = block_max(rmax, red);
  double* E = W.EMAT + (long long)mol * SEQM_EM * SEQM_EM;
  for (int q = 0; q < cF; ++q) {
    const double d = block_sum(dots[q], red);
    if (threadIdx.x == 0) E[counter * SEQM_EM + q] = d;
  }
  if (threadIdx.x == 0) W.diis_err[mol] = rmax;
}

#define SEQM_DIIS_WARPS 4

// DIIS step 2 (scf_loop.py:1009-1031): pseudo-inverse solve of the (cF+1)x(cF+1) Pulay system, one WARP per
// molecule (lanes own rows/columns of the tiny matrix in shared memory; cyclic Jacobi, the lower triangle of
// EMAT is authoritative exactly as torch.linalg.eigh(UPLO='L') reads it), condition-number reset flag.
SEQM_GLOBAL void diis_solve_kernel(seqm_batch_t b, ScfWork W, int counter, int cF, int* reset_flag) {
  __shared__ double sA[SEQM_DIIS_WARPS][SEQM_EM * SEQM_EM];
  __shared__ double sQ[SEQM_DIIS_WARPS][SEQM_EM * SEQM_EM];
  const int L = (blockDim.x >= 32) ? 32 : 1;
  const int lane = threadIdx.x % L, wib = threadIdx.x / L, wpb = blockDim.x / L;
  const int slot = blockIdx.x * wpb + wib;
  if (slot >= b.nmol) return;
  const int mol = b.mol_order[slot];
  if (!W.active[mol]) return;
  if (counter < 0) {
    counter = W.ctrl->counter;
    cF = W.ctrl->cF;
    if (cF < 2) return;  // scf_loop.py:1016
  }
  const int n = cF + 1;
  double* A = sA[wib];
  double* Q = sQ[wib];
  const double* E = W.EMAT + (long long)mol * SEQM_EM * SEQM_EM;
  double denom = E[counter * SEQM_EM + counter];
  if (denom < 1.0e-15) denom = 1.0e-15;
  for (int t = lane; t < n * n; t += L) {
    const int i = t / n, j = t % n;
    const int r = (i >= j) ? i : j, c = (i >= j) ? j : i;  // read the lower triangle only
    double e = E[r * SEQM_EM + c];
    if (r < cF && c < cF) e /= denom;
    A[t] = e;
    Q[t] = (i == j) ? 1.0 : 0.0;
  }
  SEQM_SYNCWARP();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < n; ++i) {
      dg = fmax(dg, fabs(A[i * n + i]));
      for (int j = 0; j < i; ++j) off = fmax(off, fabs(A[i * n + j]));
    }
    if (off <= 2.0e-16 * dg || off == 0.0) break;  // at the rounding floor; identical in every lane
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (fabs(apq) < 1.0e-300) continue;
        const double tau = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
        SEQM_SYNCWARP();
        for (int k = lane; k < n; k += L) {  // columns p,q of A and Q
          const double x = A[k * n + p], y = A[k * n + q];
          A[k * n + p] = c * x - s * y;
          A[k * n + q] = s * x + c * y;
          const double u = Q[k * n + p], w2 = Q[k * n + q];
          Q[k * n + p] = c * u - s * w2;
          Q[k * n + q] = s * u + c * w2;
        }
        SEQM_SYNCWARP();
        for (int k = lane; k < n; k += L) {  // rows p,q of A
          const double x = A[p * n + k], y = A[q * n + k];
          A[p * n + k] = c * x - s * y;
          A[q * n + k] = s * x + c * y;
        }
        SEQM_SYNCWARP();
      }
  }
  SEQM_SYNCWARP();
  double amax = 0.0, amin = 1.0e300;
  for (int i = 0; i < n; ++i) {
    const double a = fabs(A[i * n + i]);
    amax = fmax(amax, a);
    amin = fmin(amin, a);
  }
  if (lane == 0 && amax / amin > 1.0e7) seqm_atomic_or(reset_flag, 1);
  for (int k = lane; k < cF; k += L) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
      const double l = A[i * n + i];
      if (fabs(l) > 1.0e-13) s += Q[k * n + i] * Q[(n - 1) * n + i] / l;
    }
    W.coeff[(long long)mol * SEQM_NFOCK + k] = -s;
  }
}

// DIIS step 3 (scf_loop.py:1033-1035): F = sum_k coeff_k FOCK_k
SEQM_GLOBAL void diis_extrapolate_kernel(seqm_batch_t b, ScfWork W, double* __restrict__ F, int cF) {
  const int mol = b.mol_order[blockIdx.x];
  if (!W.active[mol]) return;
  if (cF < 0) {
    cF = W.ctrl->cF;
    if (cF < 2) return;
  }
  const MolView v = mol_view(b, mol);
  const int nn = v.n * v.n;
  const long long h0 = v.mat0 * SEQM_NFOCK;
  double c[SEQM_NFOCK];
  for (int k = 0; k < cF; ++k) c[k] = W.coeff[(long long)mol * SEQM_NFOCK + k];
  const int n = v.n, ntri = n * (n + 1) / 2;
  for (int t = threadIdx.x; t < ntri; t += blockDim.x) {  // packed index -> (r, cc >= r); both triangles get the same sum
    int r = (int)((2.0 * n + 1.0 - sqrt((2.0 * n + 1.0) * (2.0 * n + 1.0) - 8.0 * t)) * 0.5);
    while (diis_tri(r, r, n) > t) --r;                    // guard the floating-point row estimate
    while (r + 1 < n && diis_tri(r + 1, r + 1, n) <= t) ++r;
    const int cc = r + (t - diis_tri(r, r, n));
    double s = 0.0;
    for (int k = 0; k < cF; ++k) s += c[k] * W.FOCK[h0 + (long long)k * nn + t];
    F[v.mat0 + r * n + cc] = s;
    F[v.mat0 + cc * n + r] = s;
  }
}

// Pold <- P ; P <- a P + (1-a) Pnew on active molecules (scf_loop.py:264-267, 1045-1056)
SEQM_GLOBAL void mix_linear_kernel(seqm_batch_t b, ScfWork W, double* __restrict__ P, double a) {
  const int mol = b.mol_order[blockIdx.x];
  if (!W.active[mol]) return;
  if (a < 0.0) a = (W.ctrl->cF < 2) ? 0.5 : 0.0;  // pipelined DIIS: alpha_direct until two Fock matrices are stored
  const MolView v = mol_view(b, mol);
  const int nn = v.n * v.n;
  const double oma = 1.0 - a;
  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    const double p = P[v.mat0 + t];
    W.Pold[v.mat0 + t] = p;
    P[v.mat0 + t] = (a == 0.0) ? W.Pnew[v.mat0 + t] : a * p + oma * W.Pnew[v.mat0 + t];
  }
}

// adaptive mixing, diagonal part (scf_loop.py:350-415).  The renormalisation loop of the reference stops only when
// EVERY active molecule is normalised (`if torch.all(done): break`, scf_loop.py:404) and until then it rescales all of
// them, so the number of rescaling passes K is a batch-global quantity.  Two kernels, one thread per molecule:
//   adaptive_diag_a_kernel  FAC, the capped / extrapolated diagonal, SUM0, and the passes this molecule needs on its own
//                           (they are applied in place); K = max over the batch by atomicMax into ctrl->pad[k & 1]
//   adaptive_diag_b_kernel  the K - own further passes every molecule still owes (same arithmetic, same order)
// (round 1 ran this as ONE 1024-thread CTA for the whole batch: 0.12 ms per SCF iteration at 2048 molecules.)
SEQM_D bool adaptive_pass(double* dn, int n, double* sum0) {  // one pass of scf_loop.py:396-412; true = this molecule is done
  double sum2 = 0.0;
  for (int i = 0; i < n; ++i) sum2 += dn[i];
  const bool large = sum2 > 1.0e-3;
  const double sum3 = large ? *sum0 / sum2 : 0.0;
  if (!(large && !(fabs(sum3 - 1.0) <= 1.0e-5))) return true;
  int nfull = 0;
  for (int i = 0; i < n; ++i) {
    double s = fmax(dn[i] * sum3, 0.0);
    if (s > 2.0) {
      s = 2.0;
      ++nfull;
    }
    dn[i] = s;
  }
  *sum0 -= 2.0 * nfull;
  return false;
}
// a pass applied to a molecule that is already done on its own but rides along with the batch (rescales unconditionally)
SEQM_D void adaptive_pass_forced(double* dn, int n, double* sum0) {
  double sum2 = 0.0;
  for (int i = 0; i < n; ++i) sum2 += dn[i];
  const double sum3 = (sum2 > 1.0e-3) ? *sum0 / sum2 : 0.0;
  int nfull = 0;
  for (int i = 0; i < n; ++i) {
    double s = fmax(dn[i] * sum3, 0.0);
    if (s > 2.0) {
      s = 2.0;
      ++nfull;
    }
    dn[i] = s;
  }
  *sum0 -= 2.0 * nfull;
}
SEQM_GLOBAL void adaptive_diag_a_kernel(seqm_batch_t b, ScfWork W, const double* __restrict__ P, int k) {
  const bool third = (k % 3) == 0;
  const double DAMP = (k > 4) ? 0.05 : 1.0e10;
  for (int mol = blockIdx.x * blockDim.x + threadIdx.x; mol < b.nmol; mol += gridDim.x * blockDim.x) {
    if (!W.active[mol]) continue;
    const MolView v = mol_view(b, mol);
    const int n = v.n;
    double* dn = W.dnew + (long long)mol * b.nmax;
    const double* o2 = W.old2 + (long long)mol * b.nmax;
    const double* Pc = W.Pnew + v.mat0;  // "current": the freshly diagonalised density
    const double* Pp = P + v.mat0;       // "previous": the density F was built from
    double fac = 0.0;
    if (third) {
      double num = 0.0, den = 0.0;
      for (int i = 0; i < n; ++i) {
        const double dc = Pc[i * n + i], dp = Pp[i * n + i];
        const double d1 = dc - dp, d2 = dc - 2.0 * dp + o2[i];
        num += d1 * d1;
        den += d2 * d2;
      }
      if (den > 0.0 && num < 100.0 * den) fac = sqrt(num / den);
    }
    W.fac[mol] = fac;
    double sum0 = 0.0;
    for (int i = 0; i < n; ++i) {
      const double dc = Pc[i * n + i], dp = Pp[i * n + i];
      const double delta = dc - dp;
      double d;
      if (fabs(delta) > DAMP)
        d = dp + (delta > 0.0 ? DAMP : (delta < 0.0 ? -DAMP : 0.0));
      else
        d = dc + fac * delta;
      dn[i] = fmin(fmax(d, 0.0), 2.0);
      sum0 += dc;
    }
    int own = 0;
    while (own < 20 && !adaptive_pass(dn, n, &sum0)) ++own;
    W.sum0[mol] = sum0;
    W.sum3[mol] = (double)own;  // passes already applied to this molecule
    seqm_atomic_max(&W.ctrl->pad[k & 1], own);
  }
}
SEQM_GLOBAL void adaptive_diag_b_kernel(seqm_batch_t b, ScfWork W, int k) {
  const int K = W.ctrl->pad[k & 1];
  for (int mol = blockIdx.x * blockDim.x + threadIdx.x; mol < b.nmol; mol += gridDim.x * blockDim.x) {
    if (!W.active[mol]) continue;
    const int own = (int)W.sum3[mol];
    if (own >= K) continue;
    const int n = mol_view(b, mol).n;
    double* dn = W.dnew + (long long)mol * b.nmax;
    double sum0 = W.sum0[mol];
    for (int q = own; q < K; ++q) adaptive_pass_forced(dn, n, &sum0);
    W.sum0[mol] = sum0;
  }
}
// clears the pass counter of the NEXT iteration (runs after adaptive_diag_b_kernel on the same stream)
SEQM_GLOBAL void adaptive_diag_reset_kernel(ScfWork W, int k) { W.ctrl->pad[(k + 1) & 1] = 0; }
// adaptive mixing, matrix part (scf_loop.py:373-383, 417-420, 546-554)
SEQM_GLOBAL void adaptive_apply_kernel(seqm_batch_t b, ScfWork W, double* __restrict__ P) {
  const int mol = b.mol_order[blockIdx.x];
  if (!W.active[mol]) return;
  const MolView v = mol_view(b, mol);
  const int n = v.n, nn = n * n;
  const double f = W.fac[mol];
  const double* dn = W.dnew + (long long)mol * b.nmax;
  double* o2 = W.old2 + (long long)mol * b.nmax;
  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    const int i = t / n, j = t % n;
    const double pp = P[v.mat0 + t], pc = W.Pnew[v.mat0 + t];
    W.Pold[v.mat0 + t] = pp;
    if (i == j) {
      o2[i] = pp;
      P[v.mat0 + t] = dn[i];
    } else {
      P[v.mat0 + t] = (f != 0.0) ? (1.0 + f) * pc - f * pp : pc;
    }
  }
}

// get_error bookkeeping of one molecule (scf_loop.py:106-147), executed by a single thread
SEQM_D void finalize_error(const seqm_batch_t& b, const ScfWork& W, int mol, double e, double d2, double dmax,
                           int32_t* notconv, double eps, int use_diis) {
  const MolView v = mol_view(b, mol);
  const double err = e - W.Eel_run[mol];
  W.err[mol] = err;
  bool bad = fabs(err) > eps;
  if (use_diis) bad = bad || (W.diis_err[mol] > 50.0 * eps);
  if (!bad) {
    W.dm_err[mol] = sqrt(d2) / (double)(5 * v.nsh + 4 * v.nheavy + 4 * v.nhyd);  // scf_loop.py:245: 9 nSH + 4 nHeavy + 4 nHydro
    W.dm_elem[mol] = dmax;
  }
  const bool nc = bad || (W.dm_err[mol] > 2.0 * eps) || (W.dm_elem[mol] > 15.0 * eps);
  W.Eel_new[mol] = e;
  notconv[mol] = nc ? 1 : 0;
  if (nc) {
    W.Eel_run[mol] = e;
    seqm_atomic_add(&W.ctrl->nnot, 1);
  }
}

// get_error (scf_loop.py:106-147) fused with elec_energy of the new density; one CTA per molecule.
SEQM_GLOBAL void energy_error_kernel(seqm_batch_t b, ScfWork W, const double* __restrict__ P, const double* __restrict__ H,
                                     const double* __restrict__ F, int32_t* __restrict__ notconv, double eps,
                                     int use_diis) {
  __shared__ double red[33];
  const int mol = b.mol_order[blockIdx.x];
  if (!W.active[mol]) return;
  const MolView v = mol_view(b, mol);
  const int nn = v.n * v.n;
  double e = 0.0, d2 = 0.0, dmax = 0.0;
  for (int t = threadIdx.x; t < nn; t += blockDim.x) {
    const double p = P[v.mat0 + t];
    e += p * (H[v.mat0 + t] + F[v.mat0 + t]);
    const double d = p - W.Pold[v.mat0 + t];
    d2 += d * d;
    dmax = fmax(dmax, fabs(d));
  }
  e = 0.5 * block_sum(e, red);
  d2 = block_sum(d2, red);
  dmax = block_max(dmax, red);
  if (threadIdx.x == 0) finalize_error(b, W, mol, e, d2, dmax, notconv, eps, use_diis);
}

// ---- large molecules: the same three element-wise SCF pieces with a grid per molecule ----------------------
SEQM_GLOBAL void extrapolate_large_kernel(long long nn, const double* __restrict__ coeff, const double* __restrict__ hist,
                                          double* __restrict__ F, int cF) {
  double c[SEQM_NFOCK];
  for (int k = 0; k < cF; ++k) c[k] = coeff[k];
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nn; t += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < cF; ++k) s += c[k] * hist[(long long)k * nn + t];
    F[t] = s;
  }
}
SEQM_GLOBAL void mix_large_kernel(long long nn, double* __restrict__ P, double* __restrict__ Pold,
                                  const double* __restrict__ Pnew, double a) {
  const double oma = 1.0 - a;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nn; t += (long long)gridDim.x * blockDim.x) {
    const double p = P[t];
    Pold[t] = p;
    P[t] = (a == 0.0) ? Pnew[t] : a * p + oma * Pnew[t];
  }
}
// per-CTA partial sums (fixed order, no atomics): part[cta] = {sum P(H+F), sum dP^2, max |dP|}
SEQM_GLOBAL void energy_partial_kernel(long long nn, const double* __restrict__ P, const double* __restrict__ H,
                                       const double* __restrict__ F, const double* __restrict__ Pold, double* __restrict__ part) {
  __shared__ double red[33];
  double e = 0.0, d2 = 0.0, dmax = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nn; t += (long long)gridDim.x * blockDim.x) {
    const double p = P[t];
    e += p * (H[t] + F[t]);
    const double d = p - Pold[t];
    d2 += d * d;
    dmax = fmax(dmax, fabs(d));
  }
  e = block_sum(e, red);
  d2 = block_sum(d2, red);
  dmax = block_max(dmax, red);
  if (threadIdx.x == 0) {
    part[3 * blockIdx.x] = e;
    part[3 * blockIdx.x + 1] = d2;
    part[3 * blockIdx.x + 2] = dmax;
  }
}
SEQM_GLOBAL void energy_finalize_kernel(seqm_batch_t b, ScfWork W, int mol, const double* __restrict__ part, int nparts,
                                        int32_t* __restrict__ notconv, double eps, int use_diis) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double e = 0.0, d2 = 0.0, dmax = 0.0;
  for (int k = 0; k < nparts; ++k) {
    e += part[3 * k];
    d2 += part[3 * k + 1];
    dmax = fmax(dmax, part[3 * k + 2]);
  }
  finalize_error(b, W, mol, 0.5 * e, d2, dmax, notconv, eps, use_diis);
}
SEQM_GLOBAL void commit_active_kernel(seqm_batch_t b, ScfWork W, const int32_t* __restrict__ notconv) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < b.nmol; i += gridDim.x * blockDim.x) {
    const int m = b.mol_order[i];
    W.active[m] = notconv[m];
  }
}
