// seqm_pair.cu -- translation unit of libseqm_b200.so for the kernels that run the full pair code (pair_kernels.cuh):
// two-centre integrals + overlap blocks, core-core repulsion and the pair gradients.  It also owns the Slater-overlap
// polynomial tables (pairtu_ensure_tables).  This is the slow unit to compile (about 25 minutes of nvcc for the dual-number
// instantiations); the primary unit seqm_b200.cu reaches it through the launchers below and rebuilds in about a minute.
#define SEQM_SECONDARY_TU
#define SEQM_PAIR_TU
#include "pair_kernels.cuh"

// cls: 0 H-H, 1 X-H, 2 X-X (the kernels walk the class's pair list, pair_cls_off)
int pairtu_launch_integrals(const seqm_batch_t* b, int cls, int grid, int block, const double* xyz, double* w, double* hab,
                            cudaStream_t st) {
  switch (cls) {
    case 0: SEQM_LAUNCH(pair_integrals_kernel<0>, grid, block, 0, st, *b, xyz, w, hab); break;
    case 1: SEQM_LAUNCH(pair_integrals_kernel<1>, grid, block, 0, st, *b, xyz, w, hab); break;
    default: SEQM_LAUNCH(pair_integrals_kernel<2>, grid, block, 0, st, *b, xyz, w, hab); break;
  }
  return seqm_check_launch("pair_integrals_kernel");
}
int pairtu_launch_gradient(const seqm_batch_t* b, int cls, int grid, int block, const double* xyz, const double* D,
                           const double* P, double* gp, cudaStream_t st) {
  switch (cls) {
    case 0: SEQM_LAUNCH(pair_gradient_kernel<0>, grid, block, 0, st, *b, xyz, D, P, gp); break;
    case 1: SEQM_LAUNCH(pair_gradient_kernel<1>, grid, block, 0, st, *b, xyz, D, P, gp); break;
    default: SEQM_LAUNCH(pair_gradient_kernel<2>, grid, block, 0, st, *b, xyz, D, P, gp); break;
  }
  return seqm_check_launch("pair_gradient_kernel");
}
int pairtu_launch_gradient_forward(const seqm_batch_t* b, int grid, int block, const double* xyz, const double* P, double* gp,
                                   cudaStream_t st) {
  SEQM_LAUNCH(pair_gradient_forward_kernel, grid, block, 0, st, *b, xyz, P, gp);
  return seqm_check_launch("pair_gradient_forward_kernel");
}
int pairtu_launch_nuclear(const seqm_batch_t* b, int grid, int block, const double* xyz, const double* w, double* EnucAB,
                          cudaStream_t st) {
  SEQM_LAUNCH(nuclear_energy_kernel, grid, block, 0, st, *b, xyz, w, EnucAB);
  return seqm_check_launch("nuclear_energy_kernel");
}
