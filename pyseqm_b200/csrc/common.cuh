// common.cuh -- batch accessors, error plumbing, overlap polynomial tables.
#pragma once
#include <cstdarg>

#include "../../include/seqm_b200.h"
#include "pair_core.cuh"

#define SEQM_MAX_ORB 118  // two n x n fp64 matrices must fit the 227 KB shared memory of one SM

// The library is built from several translation units compiled in parallel (seqm_b200.cu is the primary one; seqm_pair.cu,
// seqm_spd.cu, seqm_eigh.cu, seqm_post.cu define SEQM_SECONDARY_TU and only see declarations of the process-wide state).
// seqm_pair.cu additionally defines SEQM_PAIR_TU: it owns the Slater-overlap polynomial tables.
#ifndef SEQM_SECONDARY_TU
static char g_seqm_err[512] = "";
long long g_seqm_launches = 0;
void seqm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_seqm_err, sizeof(g_seqm_err), fmt, ap);
  va_end(ap);
}
int seqm_check_launch(const char* what) {
#ifndef SEQM_HOSTEMU
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    seqm_set_error("%s: %s", what, cudaGetErrorString(e));
    return SEQM_ERR_CUDA;
  }
#endif
  (void)what;
  return SEQM_OK;
}

#ifdef SEQM_HOSTEMU
thread_local seqm_dim3 threadIdx, blockIdx, blockDim, gridDim;
unsigned char* seqm_hostemu_smem = nullptr;
static size_t g_hostemu_smem_bytes = 0;
void seqm_hostemu_ensure_smem(size_t bytes) {
  if (bytes > g_hostemu_smem_bytes) {
    free(seqm_hostemu_smem);
    seqm_hostemu_smem = (unsigned char*)malloc(bytes);
    g_hostemu_smem_bytes = bytes;
  }
}
#endif
#endif  // SEQM_SECONDARY_TU

// ---- per-molecule geometry of the packed layout ---------------------------------------------------
struct MolView {
  int m, a0, na, nheavy, nhyd, n, nocc, p0, npair;
  int nsh;  // PM6 d-shell atoms (9 orbitals): the first nsh of the nheavy heavy atoms; 0 for the sp methods
  long long mat0;
};
SEQM_HD MolView mol_view(const seqm_batch_t& b, int m) {
  MolView v;
  v.m = m;
  v.a0 = b.mol_atom0[m];
  v.na = b.mol_atom0[m + 1] - v.a0;
  v.nheavy = b.mol_nheavy[m];
  v.nhyd = b.mol_nhyd[m];
  v.nsh = b.mol_nsh ? b.mol_nsh[m] : 0;
  v.n = 5 * v.nsh + 4 * v.nheavy + v.nhyd;
  v.nocc = b.mol_nocc[m];
  v.p0 = b.mol_pair0[m];
  v.npair = b.mol_pair0[m + 1] - v.p0;
  v.mat0 = b.mol_mat0[m];
  return v;
}
// local atom index a (0..na-1, heavy atoms first) -> first packed orbital / number of orbitals
SEQM_HD int orb_off(const MolView& v, int a) {
  return a < v.nsh ? 9 * a : (a < v.nheavy ? 5 * v.nsh + 4 * a : 5 * v.nsh + 4 * v.nheavy + (a - v.nheavy));
}
SEQM_HD int orb_cnt(const MolView& v, int a) { return a < v.nsh ? 9 : (a < v.nheavy ? 4 : 1); }
SEQM_HD int prod_cnt(const MolView& v, int a) { return a < v.nsh ? 45 : (a < v.nheavy ? 10 : 1); }  // orbital products
// index of pair (a<b) inside the molecule's dense triangular pair list
SEQM_HD int pair_local(const MolView& v, int a, int b) { return a * (2 * v.na - a - 1) / 2 + (b - a - 1); }
SEQM_HD double par(const seqm_batch_t& b, int row, int atom) { return b.atom_par[(long long)row * b.nat + atom]; }

#ifdef SEQM_PAIR_TU
// ---- overlap polynomial tables (host-built, device constant) --------------------------------------
SEQM_CONSTANT OverlapTables c_ovl;

static void poly_mul(const int a[16][16], const int b[16][16], int out[16][16]) {
  int t[16][16];
  memset(t, 0, sizeof(t));
  for (int i = 0; i < 16; ++i)
    for (int j = 0; j < 16; ++j)
      if (a[i][j])
        for (int k = 0; i + k < 16; ++k)
          for (int l = 0; j + l < 16; ++l) t[i + k][j + l] += a[i][j] * b[k][l];
  memcpy(out, t, sizeof(t));
}
static void build_overlap_tables(OverlapTables* T) {
  memset(T, 0, sizeof(*T));
  // polynomials in (xi, eta): index [power of xi][power of eta]
  int one[16][16], xpe[16][16], xme[16][16], onep[16][16], monep[16][16], x2m1[16][16], ome2[16][16];
  memset(one, 0, sizeof(one)); one[0][0] = 1;
  memset(xpe, 0, sizeof(xpe)); xpe[1][0] = 1; xpe[0][1] = 1;      // xi + eta
  memset(xme, 0, sizeof(xme)); xme[1][0] = 1; xme[0][1] = -1;     // xi - eta
  memset(onep, 0, sizeof(onep)); onep[0][0] = 1; onep[1][1] = 1;  // 1 + xi eta
  memset(monep, 0, sizeof(monep)); monep[0][0] = -1; monep[1][1] = 1;  // xi eta - 1
  memset(x2m1, 0, sizeof(x2m1)); x2m1[2][0] = 1; x2m1[0][0] = -1;      // xi^2 - 1
  memset(ome2, 0, sizeof(ome2)); ome2[0][0] = 1; ome2[0][2] = -1;      // 1 - eta^2
  double fact[8] = {1, 1, 2, 6, 24, 120, 720, 5040};
  for (int na = 1; na <= 3; ++na)
    for (int nb = 1; nb <= 3; ++nb) {
      T->norm[na - 1][nb - 1] = 1.0 / sqrt(fact[2 * na] * fact[2 * nb]);
      for (int kind = 0; kind < 5; ++kind) {
        int pa = (kind == 0 || kind == 2) ? na : na - 1;  // power of (xi+eta)
        int pb = (kind == 0 || kind == 1) ? nb : nb - 1;  // power of (xi-eta)
        int acc[16][16];
        memcpy(acc, one, sizeof(acc));
        for (int i = 0; i < pa; ++i) poly_mul(acc, xpe, acc);
        for (int i = 0; i < pb; ++i) poly_mul(acc, xme, acc);
        if (kind == 1) poly_mul(acc, onep, acc);
        if (kind == 2) poly_mul(acc, monep, acc);
        if (kind == 3) { poly_mul(acc, onep, acc); poly_mul(acc, monep, acc); }
        if (kind == 4) { poly_mul(acc, x2m1, acc); poly_mul(acc, ome2, acc); }
        for (int k = 0; k <= SEQM_KMAX; ++k)
          for (int l = 0; l <= SEQM_KMAX; ++l) T->poly[na - 1][nb - 1][kind][k][l] = (signed char)acc[k][l];
      }
    }
}
static int g_tables_ready = 0;
int pairtu_ensure_tables() {
  if (g_tables_ready) return SEQM_OK;
  OverlapTables h;
  build_overlap_tables(&h);
#ifndef SEQM_HOSTEMU
  cudaError_t e = cudaMemcpyToSymbol(c_ovl, &h, sizeof(h));
  if (e != cudaSuccess) {
    seqm_set_error("cudaMemcpyToSymbol(overlap tables): %s", cudaGetErrorString(e));
    return SEQM_ERR_CUDA;
  }
#else
  c_ovl = h;
#endif
  g_tables_ready = 1;
  return SEQM_OK;
}
#else
int pairtu_ensure_tables();
#endif  // SEQM_PAIR_TU

// ---- bulk asynchronous copies (TMA engine, 1-D) and their mbarriers ------------------------------------------------------
// cp.async.bulk.shared::cluster.global (SASS UBLKCP) moves a contiguous, 16-byte aligned block whose size is a multiple
// of 16 bytes from global to shared memory without passing through registers; completion is counted in bytes on an
// mbarrier in shared memory.  One thread arms the barrier (arrive.expect_tx) and issues the copy, consumers spin on
// try_wait with the phase parity.  (The host emulation has no equivalent: its kernels keep the plain copy loops.)
#ifndef SEQM_HOSTEMU
SEQM_D unsigned seqm_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SEQM_D void seqm_mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(seqm_smem_u32(bar)), "r"(count) : "memory");
}
SEQM_D void seqm_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SEQM_D void seqm_bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(seqm_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   seqm_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(seqm_smem_u32(bar))
               : "memory");
}
SEQM_D void seqm_mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok = 0;
  const unsigned a = seqm_smem_u32(bar);
  while (!ok) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity)
                 : "memory");
  }
}
#endif

// block-wide sum of one double per thread (blockDim.x <= 1024, multiple of 32); result valid in all threads
SEQM_D double block_sum(double v, double* scratch /* >= 33 doubles */) {
#ifndef SEQM_HOSTEMU
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = (lane < nw) ? scratch[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
#else
  (void)scratch;
  return v;
#endif
}
SEQM_D double block_max(double v, double* scratch) {
#ifndef SEQM_HOSTEMU
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = (lane < nw) ? scratch[lane] : -1.0e300;
    for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
#else
  (void)scratch;
  return v;
#endif
}
