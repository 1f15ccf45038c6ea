// seqm_eigh.cu -- third translation unit of libseqm_b200.so: the mid-size (119..256 orbitals) eigensolver.
#define SEQM_SECONDARY_TU
#include "hestenes_kernels.cuh"

int hestenes_set_attributes(int smem_optin) {
#ifndef SEQM_HOSTEMU
  cudaError_t e = cudaFuncSetAttribute(hestenes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 1024);
  if (e != cudaSuccess) {
    seqm_set_error("cudaFuncSetAttribute(hestenes_kernel): %s", cudaGetErrorString(e));
    cudaGetLastError();
    return SEQM_ERR_CUDA;
  }
#else
  (void)smem_optin;
#endif
  return SEQM_OK;
}
int hestenes_max_orbitals(void) { return SEQM_MID_ORB; }
// F, P, C packed; evals (nmol, nmax) or NULL; P and C are both required (C doubles as the working matrix)
int hestenes_launch(const seqm_batch_t* b, const double* F, double* P, double* evals, double* C, const int32_t* active,
                    int smem_optin, cudaStream_t st) {
  if (!P || !C) {
    seqm_set_error("mid-size eigensolver: both the density and the eigenvector buffer are required");
    return SEQM_ERR_ARG;
  }
  if (b->nmax > SEQM_MID_ORB) {
    seqm_set_error("a molecule with %d orbitals exceeds the one-sided Jacobi eigensolver (%d orbitals): use the SP2 density, "
                   "sp2=[True, eps]", b->nmax, SEQM_MID_ORB);
    return SEQM_ERR_TOO_LARGE;
  }
  const size_t tail = sizeof(double) * ((size_t)b->nmax + 34) + sizeof(int) * (size_t)b->nmax + 16;
  const size_t full = sizeof(double) * (size_t)b->nmax * b->nmax + tail;
  const int in_smem = full <= (size_t)(smem_optin - 1024) ? 1 : 0;
  SEQM_LAUNCH(hestenes_kernel, b->nmol, HEST_THREADS, in_smem ? full : tail, st, *b, F, P, evals, C, active, in_smem);
  return seqm_check_launch("hestenes_kernel");
}
