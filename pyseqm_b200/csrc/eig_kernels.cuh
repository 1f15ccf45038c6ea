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
#include <utility>

#include "common.cuh"

#define SEQM_JACOBI_MAX_SWEEPS 60
// statistics counters: [0] molecules solved, [1] sweeps, [2] solves finished by the first-order correction,
// [3] solves that needed no sweep at all, [4..7] SM clock cycles (thread 0 of every CTA) spent in: warm-start
// transform, sweeps + convergence checks, epilogue (eigenvector store, correction, density), total
#ifndef SEQM_HOSTEMU
__device__ unsigned long long g_jacobi_stats[8];
#define SEQM_CLOCK() clock64()
#else
static unsigned long long g_jacobi_stats[8];
#define SEQM_CLOCK() 0LL
#endif
SEQM_D void stat_add(int k, unsigned long long v) {
#ifndef SEQM_HOSTEMU
  atomicAdd(&g_jacobi_stats[k], v);
#else
  g_jacobi_stats[k] += v;
#endif
}

struct alignas(16) seqm_d2 { double x, y; };

// ---------------------------------------------------------------------------------------------------
// Two-sided Jacobi on m = 2*NP FIXED slots (Brent-Luk odd-even ordering): even steps pair (0,1)(2,3)...,
// odd steps pair (1,2)(3,4)...(m-1,0); every rotation is followed by a swap of the two slots, so after m
// steps every pair has met exactly once.  Static positions mean no index tables, no integer division by
// runtime values and compile-time register indices:
//   * the upper triangle of A lives in shared memory tile-major (jacobi_aidx): one plane per element of the 2x2 tiles,
//     indexed by the owner slot of the tile, so the even step is a unit-stride 64-bit access and the odd step is unit
//     stride along every tile row (the kernel is bound by shared-memory wavefronts: ncu).
//   * V never touches shared memory during the sweeps: a thread keeps RB rows x m/(SR RB) consecutive columns in
//     registers (SR = 4 threads per row, 8 for the classes with NP >= 40; RB = 2 rows where m is a multiple of 16, so that
//     every rotation pair read from shared memory serves two rows); the pairs that straddle two segments in odd steps
//     are exchanged with 64-bit shuffles inside the segment group of lanes.
//   * per pair l the transform is x' = a x + b y, y' = b x - a y with (a,b) = (sin, cos) [rotate + swap],
//     (0,1) [swap only, |a_pq| below threshold] or (1,0) for the wrap pair (m-1,0) of odd steps (identity up
//     to the sign of slot 0, which is irrelevant for an eigenbasis).
// Slots n..m-1 are decoupled dummies whose diagonal lies above the Gershgorin bound; they rank last.
// blockDim.x must be SR*m (V ownership); tiles are strided over all threads.
// shared: A[AREG] | cs[2][NP] (double2) | scr[40] | dg[m] | perm[m] (int) | occm[m] (int)
//
// Inside the SCF only the density is consumed, and it depends on the occupied SUBSPACE alone.  Before every sweep
// the occupied-virtual block of the current A is inspected (slots ranked by their diagonal); once its largest
// element is below 1e-7 |A| and 1e-6 of the HOMO-LUMO gap the sweeps stop and the occupied vectors get the
// first-order correction  c_i += sum_a c_a A_ai / (d_i - d_a)  (second-order error < 1e-10 in P; rotations inside
// the occupied or the virtual space never mattered).  This replaces the last, all-tiny-rotations sweep, and in late
// SCF iterations -- where the warm-started A is already that close -- every sweep.  The eigenvectors handed to the
// next warm start stay the uncorrected, exactly orthogonal product of rotations.
// ---------------------------------------------------------------------------------------------------
#ifndef SEQM_SR8_FROM
#define SEQM_SR8_FROM 40
#endif
template <int NP>
struct JacobiCfg {
  static constexpr int M = 2 * NP;
  static constexpr int SR = (NP >= SEQM_SR8_FROM && NP % 8 == 0) ? 8 : 4;  // threads per row of V (more for the big classes: 1 CTA/SM there,
                                                 // so the CTA itself must bring enough warps to hide latency)
  static constexpr int SEG = M / SR;             // V entries per thread
  static constexpr int THREADS = SR * M;
#ifndef SEQM_JB16  // measured on B200 (4096 QM9-size molecules): 6/5/5/4/4 CTAs per SM beat the unconstrained allocation
#define SEQM_JB16 6
#define SEQM_JB20 5
#define SEQM_JB24 5
#define SEQM_JB28 4
#define SEQM_JB32 4
#endif
  // resident CTAs per SM the register allocation is asked to allow (occupancy of a barrier/latency-bound kernel)
  static constexpr int MINBLOCKS = NP == 16 ? SEQM_JB16 : NP == 20 ? SEQM_JB20 : NP == 24 ? SEQM_JB24
                                   : NP == 28 ? SEQM_JB28 : NP == 32 ? SEQM_JB32 : 0;
  static constexpr int LDT = M + 4;  // staging stride of the tensor-core products: 4 or 12 mod 16, conflict-free
  // tile ownership (see the kernel): NT upper-triangular 2x2 tiles, the NA "part A" tiles on threads 0..NA-1, the NR others
  // dealt TPX per thread over the remaining NO threads
  static constexpr int NT = NP * (NP + 1) / 2, NA = 2 * NP, NR = NT - NA;
  static constexpr int NO = (THREADS > NA) ? THREADS - NA : 1;
  static constexpr int TPX = ((NR + NO - 1) / NO > 1) ? (NR + NO - 1) / NO : 1;
  static constexpr int PL = TPX * THREADS;  // slots of one element plane of A (tile-major storage, jacobi_aidx)
  static constexpr int AREG = (4 * PL > M * LDT) ? 4 * PL : M * LDT;  // doubles reserved for A (also holds the staging tile)
  static constexpr size_t SMEM = sizeof(double) * ((size_t)AREG + 4 * NP + 40 + M) + sizeof(int) * (2 * M + 4);
};

// Storage of the upper triangle of A (r <= c), TILE-MAJOR: element e = 2 (r & 1) + (c & 1) of the 2x2 tile (k, l) = (r/2, c/2)
// lives in plane e at the slot of the tile's owner, slot = qt * THREADS + tid.  In an even step thread tid reads and writes
// plane[e][qt * THREADS + tid] -- unit lane stride, no bank conflicts; in an odd step its tile is made of element 3 of even
// tile (k, l), 2 of (k, l+1), 1 of (k+1, l) and 0 of (k+1, l+1), whose slots differ from its own by amounts that are
// constant along a tile row, so conflicts are confined to the lanes where a warp crosses a tile row (1.15-1.35 wavefronts
// per ideal one; the row-major, plane-split layout it replaces had 1.7-2.0, ncu and tools/probes/jacobi_banks.py).
template <int NP>
SEQM_HD int jacobi_slot(int k, int l) {  // owner slot of tile (k <= l)
  typedef JacobiCfg<NP> K;
  if (k == l) return k;
  if (l == k + 1) return NP + k;
  if (k == 0 && l == NP - 1) return 2 * NP - 1;
  const int before = (k == 0) ? 0 : (NP - 3) + (k - 1) * (NP - 2) - (k - 1) * k / 2;  // rest tiles of the rows above
  const int r = before + (l - k - 2);
  return (r / K::NO) * K::THREADS + K::NA + (r % K::NO);
}
template <int NP>
SEQM_HD int jacobi_aidx(int r, int c) {  // r <= c
  return (2 * (r & 1) + (c & 1)) * JacobiCfg<NP>::PL + jacobi_slot<NP>(r >> 1, c >> 1);
}
#define SEQM_AIDX(r, c) jacobi_aidx<NP>((r), (c))

// shared-memory offsets of tile (k <= l): (p_k,p_l) (p_k,q_l) (q_k,p_l) (q_k,q_l) in even (oe) and odd (oo) steps; the
// third element of a diagonal tile lies below the diagonal and is never accessed (its offset repeats the second)
template <int NP>
SEQM_HD void jacobi_tile_offsets(int k, int l, int* oe, int* oo) {
  constexpr int m = 2 * NP;
  const bool diag = (k == l);
  oe[0] = jacobi_aidx<NP>(2 * k, 2 * l);
  oe[1] = jacobi_aidx<NP>(2 * k, 2 * l + 1);
  oe[2] = diag ? oe[1] : jacobi_aidx<NP>(2 * k + 1, 2 * l);
  oe[3] = jacobi_aidx<NP>(2 * k + 1, 2 * l + 1);
  if (l < NP - 1) {
    oo[0] = jacobi_aidx<NP>(2 * k + 1, 2 * l + 1);
    oo[1] = jacobi_aidx<NP>(2 * k + 1, 2 * l + 2);
    oo[2] = diag ? oo[1] : jacobi_aidx<NP>(2 * k + 2, 2 * l + 1);
    oo[3] = jacobi_aidx<NP>(2 * k + 2, 2 * l + 2);
  } else if (k < NP - 1) {  // the wrap pair (m-1, 0): column 0 is read at its mirrored location (0, r)
    oo[0] = jacobi_aidx<NP>(2 * k + 1, m - 1);
    oo[1] = jacobi_aidx<NP>(0, 2 * k + 1);
    oo[2] = jacobi_aidx<NP>(2 * k + 2, m - 1);
    oo[3] = jacobi_aidx<NP>(0, 2 * k + 2);
  } else {
    oo[0] = jacobi_aidx<NP>(m - 1, m - 1);
    oo[1] = jacobi_aidx<NP>(0, m - 1);
    oo[2] = oo[1];
    oo[3] = jacobi_aidx<NP>(0, 0);
  }
}

// part-A tile a (0 <= a < 2 NP): diagonal (a,a), super-diagonal (a-NP, a-NP+1), corner (0, NP-1)
template <int NP>
SEQM_HD void jacobi_part_a_tile(int a, int& k, int& l) {
  if (a < NP) { k = a; l = a; }
  else if (a < 2 * NP - 1) { k = a - NP; l = k + 1; }
  else { k = 0; l = NP - 1; }
}
// remaining tile r: rows k = 0.. with l = k+2..NP-1, the corner (0, NP-1) left out
template <int NP>
SEQM_HD void jacobi_rest_tile(int r, int& k, int& l) {
  k = 0;
  int cnt = NP - 3;  // row 0: l = 2..NP-2
  while (r >= cnt) {
    r -= cnt;
    ++k;
    cnt = NP - k - 2;
  }
  l = k + 2 + r;
}
// rotation (a, b) = (sin, cos) [rotate + swap] of the pair at slots (p, p+1), or (0, 1) [swap only] below tol;
// (1, 0) for the wrap pair of odd steps
template <int NP>
SEQM_D seqm_d2 jacobi_pair_rotation(const double* A, int k, int ph, double tol, double tol_big, int* flag) {
  seqm_d2 ab;
  ab.x = 0.0;
  ab.y = 1.0;
  if (ph && k == NP - 1) {
    ab.x = 1.0;
    ab.y = 0.0;
    return ab;
  }
  // (p, q) = (2k + ph, 2k + ph + 1): even steps read the diagonal tile k, odd steps the corners of the diagonal tiles k, k+1
  // and the third element of the super-diagonal tile (k, k+1)
  constexpr int PL = JacobiCfg<NP>::PL;
  const int ipq = ph ? 2 * PL + NP + k : PL + k, ipp = ph ? 3 * PL + k : k, iqq = ph ? k + 1 : 3 * PL + k;
  const double apq = A[ipq];
  if (fabs(apq) > tol) {
    // |theta| <= pi/4 from two reciprocal square roots (no division on the critical path):
    // cos 2t = |d|/h, sin 2t = sgn(d) 2 a_pq / h, c = sqrt((1 + cos 2t)/2), s = sin 2t / (2c)
    const double d = A[iqq] - A[ipp], b2 = 2.0 * apq;
    const double rh = seqm_rsqrt(d * d + b2 * b2);
    const double c2 = 0.5 + 0.5 * fabs(d) * rh;
    const double ic = seqm_rsqrt(c2);
    ab.y = c2 * ic;                                 // cos
    ab.x = (d >= 0.0 ? 0.5 : -0.5) * b2 * rh * ic;  // sin
    flag[0] = 1;  // benign races: every writer stores 1
    if (fabs(apq) > tol_big) flag[1] = 1;
  }
  return ab;
}

// Optional fused DIIS mixing (scf_loop.py:1045-1056) at the end of the density solve: Pold <- P ;
// P <- a P + (1 - a) Pnew with a = 0.5 until two Fock matrices are stored (*cF < 2), else 0.
struct JacobiMix { double* P; double* Pold; const int* cF; };

#ifndef SEQM_HOSTEMU
// shared-memory accesses through 32-bit shared-window addresses (no generic-pointer arithmetic in the sweep loop);
// volatile: never merged or moved across the block barriers
SEQM_D double seqm_lds(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
SEQM_D void seqm_sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
template <int OFF>
SEQM_D double seqm_lds_at(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF>
SEQM_D void seqm_sts_at(unsigned a, double v) {
  asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory");
}
template <int OFF>
SEQM_D seqm_d2 seqm_lds2(unsigned a) {
  seqm_d2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF));
  return v;
}
template <int PH>
struct JacobiPhase { static constexpr int value = PH; };
// f(JacobiPhase<0>{}), f(JacobiPhase<1>{}), ... : a loop whose index is a compile-time constant inside the body
template <int... Js, class F>
SEQM_D void seqm_static_for(std::integer_sequence<int, Js...>, F&& f) {
  (f(JacobiPhase<Js>{}), ...);
}
#endif

#ifndef SEQM_HOSTEMU
// one FP64 tensor-core step: the 8x8 accumulator tile (c0, c1 = row lane/4, columns 2 (lane%4) + {0,1}) gains
// A(8x4) B(4x8) with a = A[lane/4][lane%4], b = B[lane%4][lane/4]
SEQM_D void seqm_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
#endif

template <int NP>
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS2(JacobiCfg<NP>::THREADS, JacobiCfg<NP>::MINBLOCKS) jacobi_fixed_kernel(seqm_batch_t b, int first, const double* __restrict__ F, double* __restrict__ Pout,
                                     double* __restrict__ evals, double* __restrict__ Cout,
                                     const double* __restrict__ Cguess, const int32_t* __restrict__ active,
                                     JacobiMix mix) {
  typedef JacobiCfg<NP> K;
  constexpr int M = K::M, SR = K::SR;
  const int mol = b.mol_order[first + blockIdx.x];
  if (active && !active[mol]) return;
  const MolView v = mol_view(b, mol);
  const int n = v.n;
  SEQM_DYN_SMEM(double, sm);
  double* A = sm;
  seqm_d2* cs = reinterpret_cast<seqm_d2*>(A + K::AREG);
  double* scr = reinterpret_cast<double*>(cs + 2 * NP);  // cs is double buffered: [2][NP]
  double* dg = scr + 40;
  int* perm = reinterpret_cast<int*>(dg + M);
  int* occm = perm + M;
  const double* Fm = F + v.mat0;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const bool warm = (Cguess != nullptr);
  const long long clk0 = SEQM_CLOCK();
#ifndef SEQM_HOSTEMU
  // V ownership: thread (row block, segment) keeps RB rows x SEG consecutive columns in registers.  RB = 2 where the
  // segment length stays even (m a multiple of 16): every rotation pair read from shared memory then serves two rows,
  // which halves the V update's share of the shared-memory traffic the sweep loop is bound by.
  constexpr int RB = (SR == 4 && M % 16 == 0) ? 2 : 1;
  constexpr int SRN = SR * RB;       // segments per row
  constexpr int SEG = M / SRN;       // columns per segment (shadows K::SEG, which is the RB = 1 value)
  double vr[RB][SEG];
  const int vrow = (tid / SRN) * RB, vseg = tid % SRN;
#else
  static double Vh[128 * 128];  // host emulation keeps V in memory (one "thread" plays all owners)
#endif

  if (warm) {
    // V = C0 ; A = C0^t F C0 through two global scratch slots (the caller's P and C slots, L2 resident)
    const double* C0 = Cguess + v.mat0;
    double* G = Pout + v.mat0;
    double* G2 = Cout + v.mat0;
#ifndef SEQM_HOSTEMU
    // FP64 tensor cores (DMMA 8x8x4 tiles, one warp per output tile): C0 zero-padded to a multiple of 8 in shared
    // memory with a row stride of 4 mod 16 (conflict-free fragment loads), F fragments straight from global / L1
    {
      constexpr int LDT = K::LDT;
      const int np8 = (n + 7) & ~7, nt8 = np8 >> 3;
      double* S = A;
      for (int t = tid; t < np8 * LDT; t += nthr) {
        const int r = t / LDT, c = t - r * LDT;
        S[t] = (r < n && c < n) ? C0[r * n + c] : 0.0;
      }
      SEQM_SYNC();
#pragma unroll
      for (int rb = 0; rb < RB; ++rb)
#pragma unroll
        for (int e = 0; e < SEG; ++e) {
          const int rr = vrow + rb, c = vseg * SEG + e;
          vr[rb][e] = (rr < n && c < n) ? S[rr * LDT + c] : ((rr == c) ? 1.0 : 0.0);
        }
      const int warp = tid >> 5, nwarps = nthr >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
      // T = F C0 -> G
      for (int tile = warp; tile < nt8 * nt8; tile += nwarps) {
        const int m0 = (tile / nt8) * 8, n0 = (tile % nt8) * 8;
        const int fr = m0 + g;
        const double* frow = Fm + fr * n;
        double c0 = 0.0, c1 = 0.0;
        for (int k0 = 0; k0 < np8; k0 += 16) {
          double af[4], bf[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int kc = k0 + 4 * u + t4;
            af[u] = (fr < n && kc < n) ? frow[kc] : 0.0;
            bf[u] = (kc < np8) ? S[kc * LDT + n0 + g] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) seqm_dmma(c0, c1, af[u], bf[u]);
        }
        const int r = m0 + g, c = n0 + 2 * t4;
        if (r < n) {
          if (c < n) G[r * n + c] = c0;
          if (c + 1 < n) G[r * n + c + 1] = c1;
        }
      }
      SEQM_SYNC();
      // upper tiles of C0^t T -> G2 (tile (mi, nj >= mi); a diagonal tile is written in full)
      const int ntri = nt8 * (nt8 + 1) / 2;
      for (int tile = warp; tile < ntri; tile += nwarps) {
        int mi = 0, rem = tile;
        while (rem >= nt8 - mi) { rem -= nt8 - mi; ++mi; }
        const int m0 = mi * 8, n0 = (mi + rem) * 8;
        double c0 = 0.0, c1 = 0.0;
        for (int k0 = 0; k0 < np8; k0 += 16) {
          double af[4], bf[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int kc = k0 + 4 * u + t4;
            af[u] = (kc < np8) ? S[kc * LDT + m0 + g] : 0.0;  // C0^t[m][k] = C0[k][m]
            bf[u] = (kc < n && n0 + g < n) ? G[kc * n + n0 + g] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) seqm_dmma(c0, c1, af[u], bf[u]);
        }
        const int r = m0 + g, c = n0 + 2 * t4;
        if (r < n) {
          if (c < n) G2[r * n + c] = c0;
          if (c + 1 < n) G2[r * n + c + 1] = c1;
        }
      }
    }
#else
    double* S = A;  // C0 staged in shared memory, standard layout, row stride n
    for (int t = tid; t < n * n; t += nthr) S[t] = C0[t];
    SEQM_SYNC();
    {  // T = F C0 in 2x2 register blocks (halves the loads per FMA)
      const int nb = (n + 1) >> 1;
      for (int t = tid; t < nb * nb; t += nthr) {
        const int i = 2 * (t / nb), j = 2 * (t % nb);
        const int i1 = (i + 1 < n) ? i + 1 : i, j1 = (j + 1 < n) ? j + 1 : j;
        double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
        for (int k = 0; k < n; ++k) {
          const double f0 = Fm[i * n + k], f1 = Fm[i1 * n + k];
          const double s0 = S[k * n + j], s1 = S[k * n + j1];
          a00 += f0 * s0;
          a01 += f0 * s1;
          a10 += f1 * s0;
          a11 += f1 * s1;
        }
        G[i * n + j] = a00;
        if (j1 != j) G[i * n + j1] = a01;
        if (i1 != i) {
          G[i1 * n + j] = a10;
          if (j1 != j) G[i1 * n + j1] = a11;
        }
      }
    }
#ifndef SEQM_HOSTEMU
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
      for (int e = 0; e < SEG; ++e) {
        const int rr = vrow + rb, c = vseg * SEG + e;
        vr[rb][e] = (rr < n && c < n) ? S[rr * n + c] : ((rr == c) ? 1.0 : 0.0);
      }
#else
    for (int i = 0; i < M; ++i)
      for (int c = 0; c < M; ++c) Vh[i * M + c] = (i < n && c < n) ? S[i * n + c] : ((i == c) ? 1.0 : 0.0);
#endif
    SEQM_SYNC();
    {  // upper triangle of C0^t T, 2x2 register blocks
      const int nb = (n + 1) >> 1;
      for (int t = tid; t < nb * nb; t += nthr) {
        const int bi = t / nb, bj = t % nb;
        if (bj < bi) continue;
        const int i = 2 * bi, j = 2 * bj;
        const int i1 = (i + 1 < n) ? i + 1 : i, j1 = (j + 1 < n) ? j + 1 : j;
        double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
        for (int k = 0; k < n; ++k) {
          const double c0 = S[k * n + i], c1 = S[k * n + i1];
          const double g0 = G[k * n + j], g1 = G[k * n + j1];
          a00 += c0 * g0;
          a01 += c0 * g1;
          a10 += c1 * g0;
          a11 += c1 * g1;
        }
        G2[i * n + j] = a00;
        if (j1 != j) G2[i * n + j1] = a01;
        if (i1 != i && j1 != j) G2[i1 * n + j1] = a11;
        if (i1 != i && bj > bi) G2[i1 * n + j] = a10;
      }
    }
#endif
    SEQM_SYNC();
    for (int t = tid; t < 4 * K::PL; t += nthr) A[t] = 0.0;  // unowned slots stay zero (the scan for max |A| reads them)
    SEQM_SYNC();
    for (int t = tid; t < M * M; t += nthr) {
      const int i = t / M, j = t - i * M;
      if (j >= i) A[SEQM_AIDX(i, j)] = (j < n) ? G2[i * n + j] : 0.0;
    }
  } else {
    for (int t = tid; t < 4 * K::PL; t += nthr) A[t] = 0.0;
    SEQM_SYNC();
    for (int t = tid; t < M * M; t += nthr) {
      const int i = t / M, j = t - i * M;
      if (j >= i) A[SEQM_AIDX(i, j)] = (j < n) ? Fm[i * n + j] : 0.0;
    }
#ifndef SEQM_HOSTEMU
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
      for (int e = 0; e < SEG; ++e) vr[rb][e] = (vrow + rb == vseg * SEG + e) ? 1.0 : 0.0;
#else
    for (int i = 0; i < M; ++i)
      for (int c = 0; c < M; ++c) Vh[i * M + c] = (i == c) ? 1.0 : 0.0;
#endif
  }
  SEQM_SYNC();
  const long long clk1 = SEQM_CLOCK();
  double dmax = 0.0;
  for (int t = tid; t < 4 * K::PL; t += nthr) dmax = fmax(dmax, fabs(A[t]));
  dmax = block_max(dmax, scr);
  const double tol = 1.0e-14 * fmax(dmax, 1.0e-300);
  // eigenvalues / eigenvectors requested for output (final solve): stricter early-finish threshold
  const double tol_big = ((evals || !Pout) ? 1.0e-8 : 1.0e-6) * fmax(dmax, 1.0e-300);
  for (int d = n + tid; d < M; d += nthr) A[SEQM_AIDX(d, d)] = (d + 2.0) * dmax + 1.0 + d;  // dummies above the spectrum
  SEQM_SYNC();

  // Tile ownership.  The NP diagonal tiles (k,k), the NP-1 super-diagonal tiles (k,k+1) and the corner tile
  // (0,NP-1) hold every element the NEXT step's rotations are computed from; they belong to threads 0..2NP-1 (the
  // first warps), are updated first, and those warps then work out the next step's rotations while all the other
  // threads update the remaining tiles and V: one block barrier per step, the rsqrt chain off the critical path.
  // The four shared-memory offsets of every owned tile in even and in odd steps are hoisted into registers.
  constexpr int NT = K::NT;
  constexpr int NA = K::NA;                           // part-A tiles
  constexpr int GA = ((NA + 31) / 32) * 32;           // threads of the warps that own them
  constexpr int NR = K::NR;                           // remaining tiles: l >= k + 2 without the corner
  constexpr int NO = K::NO;
  constexpr int TPX = K::TPX;                         // remaining tiles per non-part-A thread
#ifndef SEQM_HOSTEMU
  // per owned tile: does it exist / is it diagonal, the shared-window byte address of its slot in element plane 0 (even
  // steps touch the four planes at that slot; the first element of an odd step is the slot's plane-3 entry), the
  // addresses of the other three elements of its odd-step tile (ao) and of its two rotation pairs in the cs buffer of
  // even steps (odd: + NP * 16 bytes)
  bool thas[TPX], tdiag[TPX];
  unsigned ab[TPX], ao[TPX][3], ck_a[TPX], cl_a[TPX];
  const unsigned A_u32 = seqm_smem_u32(A), cs_u32 = seqm_smem_u32(cs);
#pragma unroll
  for (int qt = 0; qt < TPX; ++qt) {
    int k = -1, l = 0;
    if (tid < NA) {
      if (qt == 0) jacobi_part_a_tile<NP>(tid, k, l);
    } else {
      const int r = (tid - NA) + qt * NO;
      if (r < NR) jacobi_rest_tile<NP>(r, k, l);
    }
    thas[qt] = k >= 0;
    tdiag[qt] = k == l;
    int oe[4], oo[4];
    jacobi_tile_offsets<NP>(k < 0 ? 0 : k, l, oe, oo);
    ab[qt] = A_u32 + 8u * (unsigned)oe[0];  // oe[e] = e * PL + slot, oo[0] = oe[3]
#pragma unroll
    for (int e = 0; e < 3; ++e) ao[qt][e] = A_u32 + 8u * (unsigned)oo[e + 1];
    ck_a[qt] = cs_u32 + 16u * (unsigned)(k < 0 ? 0 : k);
    cl_a[qt] = cs_u32 + 16u * (unsigned)l;
  }
  const unsigned vcs_a = cs_u32 + 16u * (unsigned)(vseg * (SEG / 2));  // first rotation pair of this thread's V segment
  const unsigned vcp_a = cs_u32 + 16u * (unsigned)((vseg * (SEG / 2) + NP - 1) % NP);  // (previous segment's last, my first)
#endif
  __shared__ int s_flag[2][2];  // [sweep parity][0: some rotation, 1: some rotation above tol_big]
  if (tid == 0) { s_flag[0][0] = s_flag[0][1] = s_flag[1][0] = s_flag[1][1] = 0; }
  SEQM_SYNC();
  int nsweep = 0;
  const bool scf_mode = (Pout != nullptr) && (evals == nullptr) && v.nocc > 0 && v.nocc < n;
  const double tol_pert = 1.0e-7 * fmax(dmax, 1.0e-300);
  bool pert = false;
  for (int sweep = 0; sweep < SEQM_JACOBI_MAX_SWEEPS; ++sweep) {
    if (scf_mode && (warm || sweep > 0)) {
      // occupied / virtual slots by rank of the current diagonal, then the largest coupling between them
      for (int i = tid; i < M; i += nthr) dg[i] = A[SEQM_AIDX(i, i)];
      SEQM_SYNC();
      for (int i = tid; i < M; i += nthr) {
        const double ei = dg[i];
        int r = 0;
        for (int j = 0; j < M; ++j) r += (dg[j] < ei) || (dg[j] == ei && j < i);
        perm[r] = i;
        occm[i] = (r < v.nocc) ? 1 : 0;
      }
      SEQM_SYNC();
      double ov = 0.0;
      for (int t = tid; t < M * M; t += nthr) {
        const int r = t / M, c = t - r * M;
        if (c > r && occm[r] != occm[c]) ov = fmax(ov, fabs(A[SEQM_AIDX(r, c)]));
      }
      ov = block_max(ov, scr);
      const double gap = dg[perm[v.nocc]] - dg[perm[v.nocc - 1]];
      if (gap > 0.0 && ov <= tol_pert && ov <= 1.0e-6 * gap) {  // gap > 0: no 0/0 for a degenerate open shell
        pert = true;
        break;
      }
    }
    ++nsweep;
    int* flag = s_flag[sweep & 1];
    // rotations of step 0 (the later ones are computed inside the previous step)
    for (int k = tid; k < NP; k += nthr) cs[k] = jacobi_pair_rotation<NP>(A, k, 0, tol, tol_big, flag);
    SEQM_SYNC();
#ifndef SEQM_HOSTEMU
    // One rotation step with the phase (even / odd pairing) as a compile-time constant: the tile addresses of the
    // phase are plain registers and the rotation pairs are read at immediate offsets from per-tile base addresses.
    auto step_body = [&](auto phc, int step) {
      constexpr int PH = decltype(phc)::value;
      constexpr int CUR = PH * NP * 16;              // byte offset of this step's rotations in the cs double buffer
      seqm_d2* csn = cs + (PH ^ 1) * NP;             // next step's
      if (PH == 1 && step == 1 && tid == 0) { s_flag[(sweep + 1) & 1][0] = 0; s_flag[(sweep + 1) & 1][1] = 0; }
      // ---- A <- M_k^t A M_l on the upper-triangular tiles k <= l only (A is symmetric; elements (r,c) with
      //      r <= c are authoritative, the wrap pair (m-1,0) of odd steps uses the mirrored location (0,r))
#pragma unroll
      for (int qt = 0; qt < TPX; ++qt) {
        if (thas[qt]) {
          constexpr int PB = 8 * K::PL;  // bytes between element planes
          const seqm_d2 ck = seqm_lds2<CUR>(ck_a[qt]), cl = seqm_lds2<CUR>(cl_a[qt]);
          const bool diag = tdiag[qt];
          const unsigned a0 = ab[qt];
          double x0, y0, x1, y1;
          if (PH == 0) {
            x0 = seqm_lds_at<0>(a0), y0 = seqm_lds_at<PB>(a0), y1 = seqm_lds_at<3 * PB>(a0);
            x1 = diag ? y0 : seqm_lds_at<2 * PB>(a0);
          } else {
            x0 = seqm_lds_at<3 * PB>(a0), y0 = seqm_lds(ao[qt][0]), y1 = seqm_lds(ao[qt][2]);
            x1 = diag ? y0 : seqm_lds(ao[qt][1]);
          }
          const double bx0 = cl.x * x0 + cl.y * y0, by0 = cl.y * x0 - cl.x * y0;
          const double bx1 = cl.x * x1 + cl.y * y1, by1 = cl.y * x1 - cl.x * y1;
          const double n00 = ck.x * bx0 + ck.y * bx1, n01 = ck.x * by0 + ck.y * by1;
          const double n11 = ck.y * by0 - ck.x * by1, n10 = ck.y * bx0 - ck.x * bx1;
          if (PH == 0) {
            seqm_sts_at<0>(a0, n00);
            seqm_sts_at<PB>(a0, n01);
            seqm_sts_at<3 * PB>(a0, n11);
            if (!diag) seqm_sts_at<2 * PB>(a0, n10);
          } else {
            seqm_sts_at<3 * PB>(a0, n00);
            seqm_sts(ao[qt][0], n01);
            seqm_sts(ao[qt][2], n11);
            if (!diag) seqm_sts(ao[qt][1], n10);
          }
        }
        if (qt == 0 && tid < GA) {
          // the part-A tiles are done: their warps (only) meet on named barrier 1 and compute the next rotations
          if (GA == K::THREADS) __syncthreads();
          else asm volatile("bar.sync 1, %0;" ::"r"(GA) : "memory");
          if (step + 1 < M && tid < NP) csn[tid] = jacobi_pair_rotation<NP>(A, tid, PH ^ 1, tol, tol_big, flag);
        }
      }
      // ---- V <- V M
#ifndef SEQM_EXP_NOV  // SEQM_EXP_NOV: timing experiment only (tools/bench_eig.py), sweeps without the eigenvector update
#ifdef SEQM_EXP_NOV_WARP0  // ... or without it in the warp that computes the rotation parameters only
      if (tid < 32) {
      } else
#endif
      if (PH == 0) {
        seqm_static_for(std::make_integer_sequence<int, SEG / 2>{}, [&](auto jc) {
          constexpr int j = decltype(jc)::value;
          const seqm_d2 c = seqm_lds2<CUR + 16 * j>(vcs_a);
#pragma unroll
          for (int rb = 0; rb < RB; ++rb) {
            const double x = vr[rb][2 * j], y = vr[rb][2 * j + 1];
            vr[rb][2 * j] = c.x * x + c.y * y;
            vr[rb][2 * j + 1] = c.y * x - c.x * y;
          }
        });
      } else {
        const int lane = tid & 31;
        double first_old[RB], last_old[RB], y_next[RB], x_prev[RB];
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          first_old[rb] = vr[rb][0];
          last_old[rb] = vr[rb][SEG - 1];
          y_next[rb] = __shfl_sync(0xffffffffu, first_old[rb], (lane & ~(SRN - 1)) | ((lane + 1) & (SRN - 1)));
          x_prev[rb] = __shfl_sync(0xffffffffu, last_old[rb], (lane & ~(SRN - 1)) | ((lane + SRN - 1) & (SRN - 1)));
        }
        seqm_static_for(std::make_integer_sequence<int, SEG / 2 - 1>{}, [&](auto jc) {
          constexpr int j = decltype(jc)::value;
          const seqm_d2 c = seqm_lds2<CUR + 16 * j>(vcs_a);
#pragma unroll
          for (int rb = 0; rb < RB; ++rb) {
            const double x = vr[rb][2 * j + 1], y = vr[rb][2 * j + 2];
            vr[rb][2 * j + 1] = c.x * x + c.y * y;
            vr[rb][2 * j + 2] = c.y * x - c.x * y;
          }
        });
        const seqm_d2 cn = seqm_lds2<CUR + 16 * (SEG / 2 - 1)>(vcs_a);  // pair (my last, next segment's first)
        const seqm_d2 cp = seqm_lds2<CUR>(vcp_a);                       // pair (previous segment's last, my first)
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          vr[rb][SEG - 1] = cn.x * last_old[rb] + cn.y * y_next[rb];
          vr[rb][0] = cp.y * x_prev[rb] - cp.x * first_old[rb];
        }
      }
#endif
      SEQM_SYNC();
    };
    for (int step = 0; step < M; step += 2) {
      step_body(JacobiPhase<0>{}, step);
      step_body(JacobiPhase<1>{}, step + 1);
    }
#else
    for (int step = 0; step < M; ++step) {
      const int ph = step & 1;
      const seqm_d2* csc = cs + (step & 1) * NP;       // this step's rotations
      seqm_d2* csn = cs + ((step + 1) & 1) * NP;       // next step's
      if (step == 1 && tid == 0) { s_flag[(sweep + 1) & 1][0] = 0; s_flag[(sweep + 1) & 1][1] = 0; }
      for (int t = 0; t < NT; ++t) {  // the single emulation thread plays every tile owner, part A first
        int k, l;
        if (t < NA) jacobi_part_a_tile<NP>(t, k, l);
        else jacobi_rest_tile<NP>(t - NA, k, l);
        int oe1[4], oo1[4];
        jacobi_tile_offsets<NP>(k, l, oe1, oo1);
        const int a00 = ph ? oo1[0] : oe1[0], a01 = ph ? oo1[1] : oe1[1];
        const int a10 = ph ? oo1[2] : oe1[2], a11 = ph ? oo1[3] : oe1[3];
        const seqm_d2 ck = csc[k], cl = csc[l];
        const bool diag = (k == l);
        const double x0 = A[a00], y0 = A[a01], y1 = A[a11];
        const double x1 = diag ? y0 : A[a10];
        const double bx0 = cl.x * x0 + cl.y * y0, by0 = cl.y * x0 - cl.x * y0;
        const double bx1 = cl.x * x1 + cl.y * y1, by1 = cl.y * x1 - cl.x * y1;
        A[a00] = ck.x * bx0 + ck.y * bx1;
        A[a01] = ck.x * by0 + ck.y * by1;
        A[a11] = ck.y * by0 - ck.x * by1;
        if (!diag) A[a10] = ck.y * bx0 - ck.x * bx1;
        if (t == NA - 1 && step + 1 < M)
          for (int kk = 0; kk < NP; ++kk) csn[kk] = jacobi_pair_rotation<NP>(A, kk, ph ^ 1, tol, tol_big, flag);
      }
      // ---- V <- V M
      for (int i = 0; i < M; ++i)
        for (int l = 0; l < NP; ++l) {
          const int p = 2 * l + ph, q = (p + 1) % M;
          const double x = Vh[i * M + p], y = Vh[i * M + q];
          Vh[i * M + p] = csc[l].x * x + csc[l].y * y;
          Vh[i * M + q] = csc[l].y * x - csc[l].x * y;
        }
      SEQM_SYNC();
    }
#endif
#ifdef SEQM_EXP_NOV
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
      for (int e = 0; e < SEG; ++e) vr[rb][e] = (vrow + rb == vseg * SEG + e) ? 1.0 : 0.0;
#endif
    // quadratic convergence: once every rotation of a sweep was below tol_big (1e-6 |A| inside the SCF, 1e-8 |A|
    // when eigenpairs are returned) the off-diagonal left behind is O(tol_big^2 / gap): an occupied-virtual
    // coupling below 1e-9 eV, i.e. a density error below 1e-10, so no check sweep is needed
    if (!flag[0] || !flag[1]) break;
  }
  const long long clk2 = SEQM_CLOCK();
  if (tid == 0) {
    stat_add(0, 1);
    stat_add(1, nsweep);
    if (pert) stat_add(2, 1);
    if (nsweep == 0) stat_add(3, 1);
  }
  // eigenvalues -> dg and their ranking, then reuse the A storage for V in standard layout (row stride M)
  SEQM_SYNC();  // the convergence check above may still be reading dg / perm in other warps
  for (int i = tid; i < M; i += nthr) dg[i] = A[SEQM_AIDX(i, i)];
  SEQM_SYNC();
  for (int i = tid; i < M; i += nthr) {
    const double ei = dg[i];
    int r = 0;
    for (int j = 0; j < M; ++j) r += (dg[j] < ei) || (dg[j] == ei && j < i);
    perm[r] = i;
  }
  SEQM_SYNC();
  const int nv = n - v.nocc;
  double* Xg = Pout ? Pout + v.mat0 : nullptr;  // first-order mixing coefficients, [virtual q][occupied r]
  if (pert) {
    for (int t = tid; t < nv * v.nocc; t += nthr) {
      const int q = t / v.nocc, r = t - q * v.nocc;
      const int a = perm[v.nocc + q], i = perm[r];
      const double e = (a < i) ? A[SEQM_AIDX(a, i)] : A[SEQM_AIDX(i, a)];
      Xg[t] = e / (dg[i] - dg[a]);
    }
    SEQM_SYNC();
  }
  double* V = A;
#ifndef SEQM_HOSTEMU
  constexpr int LV = K::LDT;  // row stride 4 mod 16: conflict-free tensor-core fragment loads
#pragma unroll
  for (int rb = 0; rb < RB; ++rb)
#pragma unroll
    for (int e = 0; e < SEG; ++e) V[(vrow + rb) * LV + vseg * SEG + e] = vr[rb][e];
#else
  constexpr int LV = M;
  for (int t = 0; t < M * M; ++t) V[t] = Vh[t];
#endif
  SEQM_SYNC();
  if (evals)
    for (int r = tid; r < b.nmax; r += nthr) evals[(long long)mol * b.nmax + r] = (r < n) ? dg[perm[r]] : 0.0;
  if (Cout) {
    double* Cm = Cout + v.mat0;
    for (int t = tid; t < n * n; t += nthr) Cm[t] = V[(t / n) * LV + perm[t % n]];
  }
#ifndef SEQM_HOSTEMU
  const int np8 = (n + 7) & ~7, nt8 = np8 >> 3;
  const int warp = tid >> 5, nwarps = nthr >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  if (pert) {
    SEQM_SYNC();  // Cout took the uncorrected vectors
    // C_occ += C_virt X on the tensor cores: tile (8 rows) x (8 occupied columns), k over the virtual columns
    const int no = v.nocc, nto = (no + 7) >> 3;
    for (int tile = warp; tile < nt8 * nto; tile += nwarps) {
      const int i0 = (tile / nto) * 8, r0 = (tile % nto) * 8;
      double c0 = 0.0, c1 = 0.0;
      for (int k0 = 0; k0 < nv; k0 += 4) {
        const int q = k0 + t4;
        const double a = (q < nv) ? V[(i0 + g) * LV + perm[no + q]] : 0.0;
        const double bq = (q < nv && r0 + g < no) ? Xg[q * no + r0 + g] : 0.0;
        seqm_dmma(c0, c1, a, bq);
      }
      const int row = i0 + g, r = r0 + 2 * t4;
      if (row < n) {
        if (r < no) V[row * LV + perm[r]] += c0;
        if (r + 1 < no) V[row * LV + perm[r + 1]] += c1;
      }
    }
    SEQM_SYNC();
  }
  if (Pout) {
    // P = 2 C_occ C_occ^t on the tensor cores, upper tiles mirrored
    double* Pm = Pout + v.mat0;
    const int no = v.nocc;
    const int ntri = nt8 * (nt8 + 1) / 2;
    for (int tile = warp; tile < ntri; tile += nwarps) {
      int mi = 0, rem = tile;
      while (rem >= nt8 - mi) { rem -= nt8 - mi; ++mi; }
      const int i0 = mi * 8, j0 = (mi + rem) * 8;
      const double* va = V + (i0 + g) * LV;
      const double* vb = V + (j0 + g) * LV;
      double c0 = 0.0, c1 = 0.0;
      for (int k0 = 0; k0 < no; k0 += 4) {
        const int r = k0 + t4;
        const int col = (r < no) ? perm[r] : 0;
        const double a = (r < no) ? va[col] : 0.0;
        const double bq = (r < no) ? vb[col] : 0.0;
        seqm_dmma(c0, c1, a, bq);
      }
      c0 *= 2.0;
      c1 *= 2.0;
      const int row = i0 + g, c = j0 + 2 * t4;
      if (row < n) {
        if (c < n) { Pm[row * n + c] = c0; Pm[c * n + row] = c0; }
        if (c + 1 < n) { Pm[row * n + c + 1] = c1; Pm[(c + 1) * n + row] = c1; }
      }
    }
  }
#else
  if (pert) {
    SEQM_SYNC();  // Cout took the uncorrected vectors
    for (int t = tid; t < n * v.nocc; t += nthr) {
      const int row = t / v.nocc, r = t - row * v.nocc;
      double sacc = 0.0;
      for (int q = 0; q < nv; ++q) sacc += V[row * LV + perm[v.nocc + q]] * Xg[q * v.nocc + r];
      V[row * LV + perm[r]] += sacc;
    }
    SEQM_SYNC();
  }
  if (Pout) {
    double* Pm = Pout + v.mat0;
    const int nocc = v.nocc;
    const int nb = (n + 1) >> 1;  // 2x2 register blocks over the upper block triangle
    for (int t = tid; t < nb * nb; t += nthr) {
      const int bi = t / nb, bj = t % nb;
      if (bj < bi) continue;
      const int i = 2 * bi, j = 2 * bj;
      const int i1 = (i + 1 < n) ? i + 1 : i, j1 = (j + 1 < n) ? j + 1 : j;
      double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
      for (int r = 0; r < nocc; ++r) {
        const int c = perm[r];
        const double x0 = V[i * LV + c], x1 = V[i1 * LV + c], y0 = V[j * LV + c], y1 = V[j1 * LV + c];
        a00 += x0 * y0;
        a01 += x0 * y1;
        a10 += x1 * y0;
        a11 += x1 * y1;
      }
      a00 *= 2.0; a01 *= 2.0; a10 *= 2.0; a11 *= 2.0;
      Pm[i * n + j] = a00;
      Pm[j * n + i] = a00;
      if (j1 != j) { Pm[i * n + j1] = a01; Pm[j1 * n + i] = a01; }
      if (i1 != i) {
        Pm[i1 * n + j] = a10;
        Pm[j * n + i1] = a10;
        if (j1 != j) { Pm[i1 * n + j1] = a11; Pm[j1 * n + i1] = a11; }
      }
    }
  }
#endif
  if (mix.P && Pout) {
    SEQM_SYNC();  // the whole new density of this molecule is in Pout (global, L1/L2 resident)
    const double a = (*mix.cF < 2) ? 0.5 : 0.0, oma = 1.0 - a;
    const double* Pn = Pout + v.mat0;
    double* Pc = mix.P + v.mat0;
    double* Po = mix.Pold + v.mat0;
    for (int t = tid; t < n * n; t += nthr) {
      const double p = Pc[t];
      Po[t] = p;
      Pc[t] = (a == 0.0) ? Pn[t] : a * p + oma * Pn[t];
    }
  }
  if (tid == 0) {
    const long long clk3 = SEQM_CLOCK();
    stat_add(4, (unsigned long long)(clk1 - clk0));
    stat_add(5, (unsigned long long)(clk2 - clk1));
    stat_add(6, (unsigned long long)(clk3 - clk2));
    stat_add(7, (unsigned long long)(clk3 - clk0));
  }
}

// size classes: a molecule with n orbitals runs in the smallest NP with 2*NP >= n
#define SEQM_JACOBI_CLASSES(X) X(4) X(8) X(12) X(16) X(20) X(24) X(28) X(32) X(40) X(48) X(56) X(64)
static const int g_jacobi_np[] = {4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64};
static const int g_jacobi_ncls = 12;
static inline int jacobi_class_of(int n) {
  for (int c = 0; c < g_jacobi_ncls; ++c)
    if (2 * g_jacobi_np[c] >= n) return c;
  return -1;
}

// SP2 purification of one molecule per CTA: X and X^2 in shared memory.
SEQM_GLOBAL void sp2_kernel(seqm_batch_t b, const double* __restrict__ F, double* __restrict__ Pout, double eps,
                            int32_t* __restrict__ niter, const int32_t* __restrict__ active) {
  const int mol = b.mol_order[blockIdx.x];
  if (active && !active[mol]) return;
  const MolView v = mol_view(b, mol);
  const int n = v.n;
  SEQM_DYN_SMEM(double, sm);
  const double* Fm = F + v.mat0;
  if (eps > 1.0e-3) eps = 1.0e-3;
  if (eps < 1.0e-7) eps = 1.0e-7;
#ifndef SEQM_HOSTEMU
  // X^2 on the FP64 tensor cores.  X lives zero-padded to a multiple of 8 in shared memory (row stride 4 mod 16);
  // every warp owns up to 8 upper 8x8 tiles of X^2, keeps them in its DMMA accumulators until all warps are done
  // reading X, then writes X^2 or 2X - X^2 back in place (mirrored): one matrix buffer, no X^2 buffer.
  const int np8 = (n + 7) & ~7, nt8 = np8 >> 3, ld = np8 + 4;  // 4 or 12 mod 16: conflict-free fragment loads
  double* X = sm;
  double* scr = X + np8 * ld;  // 40 doubles
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, nwarps = nthr >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  double lo = 1.0e300, hi = -1.0e300;
  for (int i = tid; i < n; i += nthr) {
    double r = 0.0;
    for (int j = 0; j < n; ++j) r += fabs(Fm[i * n + j]);
    const double aii = Fm[i * n + i];
    r -= fabs(aii);
    lo = fmin(lo, aii - r);
    hi = fmax(hi, aii + r);
  }
  const double hN = block_max(hi, scr);
  const double h1 = -block_max(-lo, scr);
  for (int t = tid; t < np8 * ld; t += nthr) {
    const int r = t / ld, c = t - r * ld;
    X[t] = (r < n && c < n) ? (((r == c) ? hN : 0.0) - Fm[r * n + c]) / (hN - h1) : 0.0;
  }
  SEQM_SYNC();
  const double nocc = (double)v.nocc;
  double tr = 0.0;
  for (int i = tid; i < n; i += nthr) tr += X[i * ld + i];
  tr = block_sum(tr, scr);
  double errm0 = fabs(tr - nocc), errm1 = errm0;
  const int ntri = nt8 * (nt8 + 1) / 2;
  int k = 0;
  for (;;) {
    double acc[8][2];
    double t2 = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      acc[q][0] = acc[q][1] = 0.0;
      const int tile = warp + q * nwarps;
      if (tile < ntri) {
        int mi = 0, rem = tile;
        while (rem >= nt8 - mi) { rem -= nt8 - mi; ++mi; }
        const int i0 = mi * 8, j0 = (mi + rem) * 8;
        const double* xa = X + (i0 + g) * ld + t4;
        const double* xb = X + t4 * ld + j0 + g;
        for (int k0 = 0; k0 < np8; k0 += 4) seqm_dmma(acc[q][0], acc[q][1], xa[k0], xb[k0 * ld]);
        if (i0 == j0) {
          if (g == 2 * t4) t2 += acc[q][0];
          if (g == 2 * t4 + 1) t2 += acc[q][1];
        }
      }
    }
    t2 = block_sum(t2, scr);  // its barriers also guarantee that every warp has finished reading X
    const bool take_sq = fabs(t2 - nocc) < fabs(2.0 * tr - t2 - nocc);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int tile = warp + q * nwarps;
      if (tile < ntri) {
        int mi = 0, rem = tile;
        while (rem >= nt8 - mi) { rem -= nt8 - mi; ++mi; }
        const int i0 = mi * 8, j0 = (mi + rem) * 8;
        const int r = i0 + g;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = j0 + 2 * t4 + e;
          const double xn = take_sq ? acc[q][e] : 2.0 * X[r * ld + c] - acc[q][e];
          X[r * ld + c] = xn;
          if (i0 != j0) X[c * ld + r] = xn;  // a diagonal tile owns both (r,c) and (c,r)
        }
      }
    }
    SEQM_SYNC();
    // the reference re-sums the diagonal of the updated matrix; do the same
    double tr2 = 0.0;
    for (int i = tid; i < n; i += nthr) tr2 += X[i * ld + i];
    tr = block_sum(tr2, scr);
    errm1 = errm0;
    errm0 = fabs(tr - nocc);
    ++k;
    if ((errm0 < eps && errm1 < eps) || k >= 10000) break;
  }
  for (int t = tid; t < n * n; t += nthr) Pout[v.mat0 + t] = 2.0 * X[(t / n) * ld + (t % n)];
#else
  double* X = sm;
  double* X2 = X + n * n;
  double* scr = X2 + n * n;  // 40 doubles
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
#endif
  if (niter && threadIdx.x == 0) niter[mol] = k;
}
