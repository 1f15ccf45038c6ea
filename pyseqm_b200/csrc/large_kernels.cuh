// large_kernels.cuh -- path for molecules whose matrices do not fit one SM's shared memory (n > SEQM_MAX_ORB),
// e.g. the C380 fullerene (n = 1520, BASELINE configs[3]).  The density comes from SP2 purification
// (SP2.py:9-85), which is a chain of dense symmetric X^2 products: a register-tiled FP64 GEMM on the CUDA cores.
// (tcgen05 has no FP64 kind and B200's FP64 DMMA rate equals its FMA rate, so the tensor path buys nothing.)
//   dgemm_kernel            C = A B, row-major, 128x128x8 CTA tile, 8x8 per thread, double-buffered shared tiles
//   fock_large_*            Fock build with P, H, F in global memory (grid over all pairs / atoms of the batch)
//   sp2_* / commutator_*    element-wise pieces of SP2 and of the DIIS residual around the GEMM
#pragma once
#include "common.cuh"

#define SEQM_GEMM_BM 128
#define SEQM_GEMM_BN 128
#define SEQM_GEMM_BK 8

// C[M x N] = A[M x K] * B[K x N], all row-major with leading dimensions lda/ldb/ldc.
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS(256) dgemm_kernel(int M, int N, int K, const double* __restrict__ A, int lda,
                                                       const double* __restrict__ B, int ldb, double* __restrict__ C, int ldc) {
#ifndef SEQM_HOSTEMU
  __shared__ double sA[2][SEQM_GEMM_BK][SEQM_GEMM_BM + 4];  // A tile stored k-major (transposed) for conflict-free reads
  __shared__ double sB[2][SEQM_GEMM_BK][SEQM_GEMM_BN + 4];
  const int tid = threadIdx.x;
  const int nbx = (N + SEQM_GEMM_BN - 1) / SEQM_GEMM_BN;
  const int bm = (blockIdx.x / nbx) * SEQM_GEMM_BM, bn = (blockIdx.x % nbx) * SEQM_GEMM_BN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8 x 8 outputs (rows ty + 16 i, cols tx + 16 j)
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  // global -> register staging: A tile 128 x 8 (1024 doubles, 4 per thread), B tile 8 x 128 (4 per thread)
  double ra[4], rb[4];
  const int a_row = tid >> 1, a_col = (tid & 1) * 4;   // thread loads 4 consecutive k of one A row
  const int b_row = tid >> 5, b_col = (tid & 31) * 4;  // thread loads 4 consecutive columns of one B row
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = bm + a_row, c = k0 + a_col + q;
      ra[q] = (r < M && c < K) ? A[(long long)r * lda + c] : 0.0;
      const int rr = k0 + b_row, cc = bn + b_col + q;
      rb[q] = (rr < K && cc < N) ? B[(long long)rr * ldb + cc] : 0.0;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sA[buf][a_col + q][a_row] = ra[q];
      sB[buf][b_row][b_col + q] = rb[q];
    }
  };
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += SEQM_GEMM_BK) {
    const bool more = (k0 + SEQM_GEMM_BK) < K;
    if (more) load_tiles(k0 + SEQM_GEMM_BK);
#pragma unroll
    for (int kk = 0; kk < SEQM_GEMM_BK; ++kk) {
      double av[8], bv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = sA[buf][kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) bv[j] = sB[buf][kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += av[i] * bv[j];
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = bm + ty + 16 * i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = bn + tx + 16 * j;
      if (c < N) C[(long long)r * ldc + c] = acc[i][j];
    }
  }
#else
  // host emulation: one "thread" per CTA tile
  const int nbx = (N + SEQM_GEMM_BN - 1) / SEQM_GEMM_BN;
  const int bm = (blockIdx.x / nbx) * SEQM_GEMM_BM, bn = (blockIdx.x % nbx) * SEQM_GEMM_BN;
  for (int r = bm; r < bm + SEQM_GEMM_BM && r < M; ++r)
    for (int c = bn; c < bn + SEQM_GEMM_BN && c < N; ++c) {
      double s = 0.0;
      for (int k = 0; k < K; ++k) s += A[(long long)r * lda + k] * B[(long long)k * ldb + c];
      C[(long long)r * ldc + c] = s;
    }
#endif
}

// FP64 tensor-core variant: C = A B with mma.sync.m8n8k4.f64 (DMMA).  B200's DMMA rate equals its DFMA rate (37 vs
// 34-36 TFLOP/s measured, tools/probes/dmma_peak.cu), but one DMMA retires 256 FMAs per issue slot instead of 32, so
// the issue ports are free for the shared-memory fragment loads and the pipe can actually be kept full.
// 128x128x16 CTA tile, 8 warps as 2 (m) x 4 (n), each warp 64x32 = 8x4 DMMA tiles with the accumulators in registers;
// fragments are read straight from padded shared tiles (row strides 20 and 132 doubles: the 16 lanes of a wavefront
// hit 16 distinct 8-byte banks); global -> register -> shared double buffering as in dgemm_kernel (74.8 KB dynamic).
#define SEQM_DMMA_BK 16
#define SEQM_DMMA_SMEM (2 * (128 * (SEQM_DMMA_BK + 4) + SEQM_DMMA_BK * 132) * sizeof(double))
#ifndef SEQM_HOSTEMU
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS(256) dgemm_dmma_kernel(int M, int N, int K, const double* __restrict__ A, int lda,
                                                            const double* __restrict__ B, int ldb, double* __restrict__ C, int ldc) {
  constexpr int BM = 128, BN = 128, BK = SEQM_DMMA_BK, LDA = BK + 4, LDB = BN + 4;
  SEQM_DYN_SMEM(double, dsm);
  double* const sA0 = dsm;                 // [2][m][k]
  double* const sB0 = dsm + 2 * BM * LDA;  // [2][k][n]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
  const int nbx = (N + BN - 1) / BN;
  const int bm = (blockIdx.x / nbx) * BM, bn = (blockIdx.x % nbx) * BN;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  constexpr int NQ = BK / 2;  // doubles per thread per matrix and stage
  double ra[NQ], rb[NQ];
  const int a_row = tid >> 1, a_col = (tid & 1) * NQ;  // NQ consecutive k of one A row
  const int b_row0 = tid >> 5, b_col = (tid & 31) * 4;  // 4 consecutive columns of rows b_row0, b_row0 + 8, ...
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int r = bm + a_row, c = k0 + a_col + q;
      ra[q] = (r < M && c < K) ? A[(long long)r * lda + c] : 0.0;
      const int rr = k0 + b_row0 + 8 * (q >> 2), cc = bn + b_col + (q & 3);
      rb[q] = (rr < K && cc < N) ? B[(long long)rr * ldb + cc] : 0.0;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      sA0[buf * BM * LDA + a_row * LDA + a_col + q] = ra[q];
      sB0[buf * BK * LDB + (b_row0 + 8 * (q >> 2)) * LDB + b_col + (q & 3)] = rb[q];
    }
  };
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = (k0 + BK) < K;
    if (more) load_tiles(k0 + BK);
    const double* pa = sA0 + buf * BM * LDA + (wm + g) * LDA + t;
    const double* pb = sB0 + buf * BK * LDB + t * LDB + wn + g;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = pa[i * 8 * LDA + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = pb[kk * LDB + j * 8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(af[i]), "d"(bf[j]));
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = bm + wm + i * 8 + g;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = bn + wn + j * 8 + 2 * t;
      if (c + 1 < N) {
        double2 v2;
        v2.x = acc[i][j][0];
        v2.y = acc[i][j][1];
        if (((((long long)r * ldc + c) & 1) == 0)) *reinterpret_cast<double2*>(C + (long long)r * ldc + c) = v2;
        else { C[(long long)r * ldc + c] = v2.x; C[(long long)r * ldc + c + 1] = v2.y; }
      } else if (c < N) {
        C[(long long)r * ldc + c] = acc[i][j][0];
      }
    }
  }
}
#endif

// C = X X for a SYMMETRIC X (n x n, leading dimension n), the product SP2 spends its time in (SP2.py:55): only the tiles
// on and above the diagonal are computed and every off-diagonal tile is also stored transposed, i.e. 0.53 of the FLOPs
// of the general product for n = 1520.  96x96x16 CTA tiles (16 x 17 / 2 = 136 tiles for n = 1520: one wave of the 148
// SMs), 8 warps as 2 (m) x 4 (n), each 48x24 = 6x3 DMMA tiles; fragment layout and shared-memory padding as in
// dgemm_dmma_kernel (row strides 20 and 100 doubles), global loads with unit lane stride in both tiles.
// skip: optional device flag; a non-zero value turns the launch into a no-op (SP2 already converged, see sp2_large_one).
#define SEQM_SYM_TB 96
#define SEQM_SYM_SMEM (2 * (SEQM_SYM_TB * (SEQM_DMMA_BK + 4) + SEQM_DMMA_BK * (SEQM_SYM_TB + 4)) * sizeof(double))
SEQM_HD void sym_tile_of(int q, int nb, int* bi, int* bj) {  // q-th tile of the upper triangle, row by row
  int i = 0;
  while (q >= nb - i) {
    q -= nb - i;
    ++i;
  }
  *bi = i;
  *bj = i + q;
}
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS(256) dgemm_sym_kernel(int n, const double* __restrict__ X, double* __restrict__ C,
                                                           const int* __restrict__ skip) {
  if (skip && *skip) return;
  constexpr int TB = SEQM_SYM_TB, BK = SEQM_DMMA_BK;
  const int nb = (n + TB - 1) / TB;
  int bi, bj;
  sym_tile_of(blockIdx.x, nb, &bi, &bj);
  const int bm = bi * TB, bn = bj * TB;
#ifndef SEQM_HOSTEMU
  constexpr int LDA = BK + 4, LDB = TB + 4, NQ = TB * BK / 256;  // 6 doubles per thread, matrix and stage
  SEQM_DYN_SMEM(double, dsm);
  double* const sA0 = dsm;                 // [2][m][k]
  double* const sB0 = dsm + 2 * TB * LDA;  // [2][k][n]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 2) * 48, wn = (warp & 3) * 24;
  double acc[6][3][2];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double ra[NQ], rb[NQ];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int e = tid + 256 * q;
      const int r = bm + (e >> 4), c = k0 + (e & 15);
      ra[q] = (r < n && c < n) ? X[(long long)r * n + c] : 0.0;
      const int rr = k0 + e / TB, cc = bn + e % TB;
      rb[q] = (rr < n && cc < n) ? X[(long long)rr * n + cc] : 0.0;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int e = tid + 256 * q;
      sA0[buf * TB * LDA + (e >> 4) * LDA + (e & 15)] = ra[q];
      sB0[buf * BK * LDB + (e / TB) * LDB + e % TB] = rb[q];
    }
  };
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < n; k0 += BK) {
    const bool more = (k0 + BK) < n;
    if (more) load_tiles(k0 + BK);
    const double* pa = sA0 + buf * TB * LDA + (wm + g) * LDA + t;
    const double* pb = sB0 + buf * BK * LDB + t * LDB + wn + g;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[6], bf[3];
#pragma unroll
      for (int i = 0; i < 6; ++i) af[i] = pa[i * 8 * LDA + kk];
#pragma unroll
      for (int j = 0; j < 3; ++j) bf[j] = pb[kk * LDB + j * 8];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(af[i]), "d"(bf[j]));
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
  const bool mirror = bi != bj;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int r = bm + wm + i * 8 + g;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int c = bn + wn + j * 8 + 2 * t;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (c + h >= n) continue;
        C[(long long)r * n + c + h] = acc[i][j][h];
        if (mirror) C[(long long)(c + h) * n + r] = acc[i][j][h];  // 8 lanes (g) write 8 consecutive doubles
      }
    }
  }
#else
  for (int r = bm; r < bm + TB && r < n; ++r)
    for (int c = bn; c < bn + TB && c < n; ++c) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += X[(long long)r * n + k] * X[(long long)k * n + c];
      C[(long long)r * n + c] = s;
      if (bi != bj) C[(long long)c * n + r] = s;
    }
#endif
}

// ---- Fock build with everything in global memory -------------------------------------------------------
// off-diagonal blocks: one work item per (pair, mu, lambda)
SEQM_GLOBAL void fock_large_offdiag_kernel(seqm_batch_t b, const double* __restrict__ P, const double* __restrict__ H,
                                           const double* __restrict__ w, double* __restrict__ F,
                                           const int32_t* __restrict__ active) {
  const long long total = (long long)b.npairs * 16;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t >> 4), mu = (int)((t >> 2) & 3), la = (int)(t & 3);
    const int gi = b.pair_i[p], gj = b.pair_j[p];
    const int m = b.atom_mol[gi];
    if (active && !active[m]) continue;
    const MolView v = mol_view(b, m);
    const int i = gi - v.a0, j = gj - v.a0;
    const int ni = orb_cnt(v, i), nj = orb_cnt(v, j);
    if (mu >= ni || la >= nj) continue;
    const int n = v.n, oi = orb_off(v, i), oj = orb_off(v, j);
    const double* Pm = P + v.mat0;
    const double* wp = w + (long long)p * 100;
    double k = 0.0;
    for (int nu = 0; nu < ni; ++nu)
      for (int sg = 0; sg < nj; ++sg)
        k += Pm[(long long)(oi + nu) * n + oj + sg] * wp[pack2(mu, nu) * 10 + pack2(la, sg)];
    const long long r = oi + mu, c = oj + la;
    const double f = H[v.mat0 + r * n + c] - 0.5 * k;
    F[v.mat0 + r * n + c] = f;
    F[v.mat0 + c * n + r] = f;
  }
}
// diagonal blocks: one WARP-sized group per (atom, packed kl) would be ideal; a thread per item loops over the
// partner atoms (380 for C380), which is ample parallelism (nat*10 items) for the large path
SEQM_GLOBAL void fock_large_diag_kernel(seqm_batch_t b, const double* __restrict__ P, const double* __restrict__ H,
                                        const double* __restrict__ w, double* __restrict__ F,
                                        const int32_t* __restrict__ active) {
  const long long total = (long long)b.nat * 10;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int ga = (int)(t / 10), kl = (int)(t % 10);
    const int m = b.atom_mol[ga];
    if (active && !active[m]) continue;
    const MolView v = mol_view(b, m);
    const int a = ga - v.a0;
    if (a >= v.nheavy && kl > 0) continue;
    int mu = 0;
    while ((mu + 1) * (mu + 2) / 2 <= kl) ++mu;
    const int nu = kl - mu * (mu + 1) / 2;
    const int n = v.n, oa = orb_off(v, a);
    const double* Pm = P + v.mat0;
#define PM(r, c) Pm[(long long)(r) * n + (c)]
    const double gss = par(b, SEQM_P_GSS, ga), gsp = par(b, SEQM_P_GSP, ga), gpp = par(b, SEQM_P_GPP, ga);
    const double gp2 = par(b, SEQM_P_GP2, ga), hsp = par(b, SEQM_P_HSP, ga);
    const double Pss = PM(oa, oa);
    double Ppt = 0.0;
    if (a < v.nheavy) Ppt = PM(oa + 1, oa + 1) + PM(oa + 2, oa + 2) + PM(oa + 3, oa + 3);
    double g;
    if (mu == 0)
      g = 0.5 * Pss * gss + Ppt * (gsp - 0.5 * hsp);
    else if (nu == 0)
      g = PM(oa, oa + mu) * (1.5 * hsp - 0.5 * gsp);
    else if (mu == nu) {
      const double Pk = PM(oa + mu, oa + mu);
      g = Pss * (gsp - 0.5 * hsp) + 0.5 * Pk * gpp + (Ppt - Pk) * (1.25 * gp2 - 0.25 * gpp);
    } else
      g = PM(oa + nu, oa + mu) * (0.75 * gpp - 1.25 * gp2);
    for (int o = 0; o < v.na; ++o) {
      if (o == a) continue;
      const int oo = orb_off(v, o), no = orb_cnt(v, o);
      const bool first = a < o;
      const double* wp = w + (long long)(v.p0 + (first ? pair_local(v, a, o) : pair_local(v, o, a))) * 100;
      const int sk = first ? 10 : 1, sm = first ? 1 : 10;
      double j = PM(oo, oo) * wp[kl * sk];
      if (no == 4) {
        for (int x = 1; x < 4; ++x) {
          j += 2.0 * PM(oo, oo + x) * wp[kl * sk + pack2(x, 0) * sm];
          for (int y = 1; y <= x; ++y) j += (x == y ? 1.0 : 2.0) * PM(oo + y, oo + x) * wp[kl * sk + pack2(x, y) * sm];
        }
      }
      g += j;
    }
#undef PM
    const long long r = oa + mu, c = oa + nu;
    const double f = H[v.mat0 + r * n + c] + g;
    F[v.mat0 + r * n + c] = f;
    F[v.mat0 + c * n + r] = f;
  }
}

// ---- SP2 pieces (one molecule at a time; state in a small device struct) ---------------------------------
// done: set by the decision step of the iteration that converged (its update still runs); finished: set by the decision
// step after that -- from then on every SP2 kernel of the stream is a no-op, so the host can queue iterations ahead
// without reading the state back
struct Sp2State {
  double h1, hN, tr, tr2, errm0, errm1, nocc;
  int take_sq, done, iters, finished;
};
// Gershgorin bounds: one CTA, rows strided over threads
SEQM_GLOBAL void sp2_bounds_kernel(int n, const double* __restrict__ Fm, double nocc, Sp2State* st) {
  __shared__ double red[33];
  double lo = 1.0e300, hi = -1.0e300;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double r = 0.0;
    for (int j = 0; j < n; ++j) r += fabs(Fm[(long long)i * n + j]);
    const double aii = Fm[(long long)i * n + i];
    r -= fabs(aii);
    lo = fmin(lo, aii - r);
    hi = fmax(hi, aii + r);
  }
  const double hN = block_max(hi, red);
  const double h1 = -block_max(-lo, red);
  if (threadIdx.x == 0) {
    st->h1 = h1;
    st->hN = hN;
    st->nocc = nocc;
    st->done = 0;
    st->iters = 0;
    st->finished = 0;
  }
}
SEQM_GLOBAL void sp2_init_kernel(int n, const double* __restrict__ Fm, double* __restrict__ X, const Sp2State* st) {
  const double h1 = st->h1, hN = st->hN;
  const long long nn = (long long)n * n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nn; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / n), j = (int)(t % n);
    X[t] = (((i == j) ? hN : 0.0) - Fm[t]) / (hN - h1);
  }
}
// traces of X and X2 (one CTA), then the SP2 branch decision and convergence bookkeeping (SP2.py:55-83)
SEQM_GLOBAL void sp2_trace_kernel(int n, const double* __restrict__ X, const double* __restrict__ X2, Sp2State* st,
                                  double eps, int first) {
  __shared__ double red[33];
  double a = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    a += X[(long long)i * n + i];
    if (X2) c += X2[(long long)i * n + i];
  }
  a = block_sum(a, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    if (first) {  // trace of X0
      st->tr = a;
      st->errm0 = fabs(a - st->nocc);
      st->errm1 = st->errm0;
    } else if (X2) {  // decide the branch from tr X and tr X^2
      st->tr2 = c;
      st->take_sq = (fabs(c - st->nocc) < fabs(2.0 * st->tr - c - st->nocc)) ? 1 : 0;
    } else {  // after the update: re-summed trace of the new X, error history, stop test
      st->tr = a;
      st->errm1 = st->errm0;
      st->errm0 = fabs(a - st->nocc);
      st->iters += 1;
      if ((st->errm0 < eps && st->errm1 < eps) || st->iters >= 10000) st->done = 1;
    }
  }
}
// One decision step per SP2 iteration (one CTA): traces of X and X^2 from their diagonals, the branch (SP2.py:55-83), the
// re-summed trace of the matrix the update kernel is about to write (same element expressions, same summation order as
// sp2_trace_kernel on the updated matrix), error history and stop test.
SEQM_GLOBAL void sp2_decide_kernel(int n, const double* __restrict__ X, const double* __restrict__ X2, Sp2State* st, double eps) {
  __shared__ double red[33];
  __shared__ int s_sq;
  if (st->done) {
    if (threadIdx.x == 0) st->finished = 1;
    return;
  }
  double c = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += X2[(long long)i * n + i];
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    st->tr2 = c;
    s_sq = (fabs(c - st->nocc) < fabs(2.0 * st->tr - c - st->nocc)) ? 1 : 0;
    st->take_sq = s_sq;
  }
  SEQM_SYNC();
  const int sq = s_sq;
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x2 = X2[(long long)i * n + i];
    a += sq ? x2 : 2.0 * X[(long long)i * n + i] - x2;
  }
  a = block_sum(a, red);
  if (threadIdx.x == 0) {
    st->tr = a;
    st->errm1 = st->errm0;
    st->errm0 = fabs(a - st->nocc);
    st->iters += 1;
    if ((st->errm0 < eps && st->errm1 < eps) || st->iters >= 10000) st->done = 1;
  }
}
SEQM_GLOBAL void sp2_update_kernel(int n, double* __restrict__ X, const double* __restrict__ X2, const Sp2State* st) {
  if (st->finished) return;
  const int sq = st->take_sq;
  const long long nn = (long long)n * n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nn; t += (long long)gridDim.x * blockDim.x)
    X[t] = sq ? X2[t] : 2.0 * X[t] - X2[t];
}
SEQM_GLOBAL void scale_copy_kernel(long long nn, const double* __restrict__ X, double* __restrict__ out, double s) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nn; t += (long long)gridDim.x * blockDim.x)
    out[t] = s * X[t];
}

// ---- DIIS residual for a large molecule: R = FP - (FP)^t from the GEMM result G = F P ----------------------
SEQM_GLOBAL void commutator_kernel(int n, const double* __restrict__ G, double* __restrict__ R, double* __restrict__ rmax_out) {
  __shared__ double red[33];
  double rmax = 0.0;
  const long long nn = (long long)n * n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nn; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / n), j = (int)(t % n);
    const double r = G[t] - G[(long long)j * n + i];
    R[t] = r;
    rmax = fmax(rmax, fabs(r));
  }
  rmax = block_max(rmax, red);
  if (threadIdx.x == 0) {
    // max over CTAs through an ordered-int atomic max (values are non-negative doubles)
#ifndef SEQM_HOSTEMU
    atomicMax(reinterpret_cast<unsigned long long*>(rmax_out), (unsigned long long)__double_as_longlong(rmax));
#else
    if (rmax > *rmax_out) *rmax_out = rmax;
#endif
  }
}
// dots[q] = sum_{i<j} R[i][j] * Rq[i][j] for the cF stored residuals: SEQM_DOT_PARTS CTAs per q write partial sums
// (rows strided over the parts), residual_dots_finish_kernel adds them in a fixed order (deterministic, no atomics)
#define SEQM_DOT_PARTS 64
SEQM_GLOBAL void residual_dots_kernel(int n, const double* __restrict__ R, const double* __restrict__ hist, long long stride,
                                      double* __restrict__ part) {
  __shared__ double red[33];
  const int q = blockIdx.x / SEQM_DOT_PARTS, slice = blockIdx.x % SEQM_DOT_PARTS;
  const double* Rq = hist + (long long)q * stride;
  double s = 0.0;
  for (int i = slice; i < n; i += SEQM_DOT_PARTS)
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) s += R[(long long)i * n + j] * Rq[(long long)i * n + j];
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
SEQM_GLOBAL void residual_dots_finish_kernel(const double* __restrict__ part, int cF, double* __restrict__ emat_row) {
  for (int q = threadIdx.x; q < cF; q += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < SEQM_DOT_PARTS; ++k) s += part[q * SEQM_DOT_PARTS + k];
    emat_row[q] = s;
  }
}
