// seqm_spd.cu -- second translation unit of libseqm_b200.so: the PM6 d-orbital kernels (spd_kernels.cuh) and their
// launchers.  Compiled in parallel with seqm_b200.cu, which owns the C ABI and the process-wide state.
#define SEQM_SECONDARY_TU
#include "spd_kernels.cuh"

static int spd_threads_for(int nmax) { return nmax <= 24 ? 128 : (nmax <= 64 ? 256 : 512); }

int spd_set_attributes(int smem_optin) {
#ifndef SEQM_HOSTEMU
  const int pair_smem = (int)(sizeof(double) * SPD_SMEM_DOUBLES);
  cudaError_t e = cudaFuncSetAttribute(spd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pair_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(spd_fock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 2048);
  if (e != cudaSuccess) {
    seqm_set_error("cudaFuncSetAttribute(spd kernels): %s", cudaGetErrorString(e));
    cudaGetLastError();
    return SEQM_ERR_CUDA;
  }
#else
  (void)smem_optin;
#endif
  return SEQM_OK;
}
int spd_launch_pair(const seqm_batch_t* b, const double* xyz, const double* w, double* wd, double* hab_d, cudaStream_t st) {
  SEQM_LAUNCH(spd_pair_kernel, b->n_ypairs, SPD_THREADS, sizeof(double) * SPD_SMEM_DOUBLES, st, *b, xyz, w, wd, hab_d);
  return seqm_check_launch("spd_pair_kernel");
}
int spd_launch_hcore(const seqm_batch_t* b, const double* w, const double* hab, double* H, cudaStream_t st) {
  SEQM_LAUNCH(spd_hcore_kernel, b->nmol, 128, 0, st, *b, w, hab, H);
  return seqm_check_launch("spd_hcore_kernel");
}
int spd_launch_fock(const seqm_batch_t* b, const double* P, const double* H, const double* w, double* F,
                    const int32_t* active, int smem_limit, cudaStream_t st) {
  const size_t smem = sizeof(double) * ((size_t)b->nmax * b->nmax + (size_t)b->molsize * 45);
  if (smem > (size_t)smem_limit) {
    seqm_set_error("PM6 with d orbitals: a molecule with %d orbitals / %d atoms exceeds the shared-memory resident Fock build",
                   b->nmax, b->molsize);
    return SEQM_ERR_TOO_LARGE;
  }
  SEQM_LAUNCH(spd_fock_kernel, b->nmol, spd_threads_for(b->nmax), smem, st, *b, P, H, w, F, active);
  return seqm_check_launch("spd_fock_kernel");
}
int spd_launch_gradient(const seqm_batch_t* b, const double* xyz, const double* P, double* gp, cudaStream_t st) {
  SEQM_LAUNCH(spd_pair_gradient_kernel, b->n_ypairs, SPD_THREADS, sizeof(double) * SPG_SMEM_DOUBLES, st, *b, xyz, P, gp);
  return seqm_check_launch("spd_pair_gradient_kernel");
}
