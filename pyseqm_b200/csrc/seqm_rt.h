// seqm_rt.h -- runtime shim.
//
// The product is compiled by nvcc for sm_100a (SEQM_HOSTEMU undefined): everything below maps 1:1 onto
// CUDA.  With -DSEQM_HOSTEMU the very same kernel sources compile with g++ into a *test-only* library
// (tests/_hostemu) in which a "kernel launch" runs every CTA sequentially with blockDim = 1 on host
// memory.  That build exists so the kernel LOGIC (index maps, phase structure, DIIS bookkeeping) can be
// checked against the oracle on GPU-less CI boxes; it is never loaded by pyseqm_b200 and is not a CPU
// fallback: pyseqm_b200/_lib.py loads libseqm_b200.so only and raises if CUDA is absent.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#ifndef SEQM_HOSTEMU
#include <cuda_runtime.h>
#define SEQM_HD __host__ __device__ __forceinline__
#define SEQM_D __device__ __forceinline__
#define SEQM_GLOBAL __global__
#define SEQM_LAUNCH_BOUNDS(n) __launch_bounds__(n)
#define SEQM_LAUNCH_BOUNDS2(n, b) __launch_bounds__(n, b)
#define SEQM_CONSTANT __constant__
#define SEQM_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char seqm_dyn_smem_[];   \
  type* name = reinterpret_cast<type*>(seqm_dyn_smem_)
extern long long g_seqm_launches;
#define SEQM_LAUNCH(kern, grid, block, smem, stream, ...) \
  do { ++g_seqm_launches; kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); } while (0)
#define SEQM_SYNC() __syncthreads()
SEQM_D int seqm_sync_or(int p) { return __syncthreads_or(p); }
SEQM_D double seqm_rsqrt(double x) { return rsqrt(x); }
SEQM_D void seqm_atomic_or(int* a, int v) { atomicOr(a, v); }
SEQM_D int seqm_atomic_add(int* a, int v) { return atomicAdd(a, v); }
SEQM_D void seqm_atomic_max_u32(unsigned* a, unsigned v) { atomicMax(a, v); }
SEQM_D void seqm_atomic_max(int* a, int v) { atomicMax(a, v); }
#else
typedef void* cudaStream_t;
typedef int cudaError_t;
struct seqm_dim3 { unsigned x, y, z; };
extern thread_local seqm_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern unsigned char* seqm_hostemu_smem;
void seqm_hostemu_ensure_smem(size_t bytes);
#define SEQM_HD inline
#define SEQM_D inline
#define SEQM_GLOBAL static
#define SEQM_LAUNCH_BOUNDS(n)
#define SEQM_LAUNCH_BOUNDS2(n, b)
#define SEQM_CONSTANT static
#define __restrict__
#define __shared__ static
#define SEQM_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(seqm_hostemu_smem)
extern long long g_seqm_launches;
#define SEQM_LAUNCH(kern, grid, block, smem, stream, ...)              \
  do {                                                                 \
    ++g_seqm_launches;                                                 \
    seqm_hostemu_ensure_smem((size_t)(smem) + 64);                     \
    unsigned g_ = (unsigned)(grid);                                    \
    gridDim = {g_, 1, 1};                                              \
    blockDim = {1, 1, 1};                                              \
    threadIdx = {0, 0, 0};                                             \
    for (unsigned b_ = 0; b_ < g_; ++b_) {                             \
      blockIdx = {b_, 0, 0};                                           \
      kern(__VA_ARGS__);                                               \
    }                                                                  \
  } while (0)
#define SEQM_SYNC() do { } while (0)
inline int seqm_sync_or(int p) { return p != 0; }
inline double seqm_rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline void seqm_atomic_or(int* a, int v) { *a |= v; }
inline int seqm_atomic_add(int* a, int v) { int o = *a; *a += v; return o; }
inline void seqm_atomic_max_u32(unsigned* a, unsigned v) { if (v > *a) *a = v; }
inline void seqm_atomic_max(int* a, int v) { if (v > *a) *a = v; }
using std::exp; using std::fabs; using std::sqrt; using std::pow; using std::fmax; using std::fmin;
template <class T> inline T min(T a, T b) { return a < b ? a : b; }
template <class T> inline T max(T a, T b) { return a > b ? a : b; }
#endif

// status codes returned by every extern "C" entry point (include/seqm_b200.h)
#define SEQM_OK 0
#define SEQM_ERR_CUDA (-1)
#define SEQM_ERR_ARG (-2)
#define SEQM_ERR_UNSUPPORTED (-3)
#define SEQM_ERR_TOO_LARGE (-4)

void seqm_set_error(const char* fmt, ...);
int seqm_check_launch(const char* what);
