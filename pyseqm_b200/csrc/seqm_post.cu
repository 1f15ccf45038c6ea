// seqm_post.cu -- fourth translation unit of libseqm_b200.so: post-SCF by-products in one launch.
//   seqm_post_scf   Mulliken charges q (ElectronicStructure.py:104-127), ground-state dipole (dipole.py:85-107) and the
//                   scatter of the per-atom gradient into the padded force tensor, one CTA per molecule on the packed density.
// Replaces ~40 eager PyTorch launches per forward (diagonal gathers, masked sums, index scatters) by one kernel.
#define SEQM_SECONDARY_TU
#include "common.cuh"

// q: (nmol, molsize) = tore - Mulliken population (0 on padding); dipole: (nmol, 3) or NULL; force: (nmol, molsize, 3) or
// NULL with grad (nat, 3) = dE/dR per real atom; scale = to_debye * debye_to_AU; a0 = bohr in Angstrom
SEQM_GLOBAL void post_scf_kernel(seqm_batch_t b, const double* __restrict__ P, const double* __restrict__ xyz,
                                 const double* __restrict__ grad, double* __restrict__ q, double* __restrict__ dipole,
                                 double* __restrict__ force, double a0, double scale) {
  __shared__ double red[33];
  const MolView v = mol_view(b, blockIdx.x);
  const double* Pm = P + v.mat0;
  const int n = v.n;
  double dx = 0.0, dy = 0.0, dz = 0.0;
  for (int a = threadIdx.x; a < b.molsize; a += blockDim.x) {
    double qa = 0.0;
    if (a < v.na) {
      const int ga = v.a0 + a, oa = orb_off(v, a), no = orb_cnt(v, a);
      double pop = 0.0;
      for (int k = 0; k < no; ++k) pop += Pm[(oa + k) * n + oa + k];
      const double tore = par(b, SEQM_P_TORE, ga);
      qa = tore - pop;
      const double x = xyz[3 * (long long)ga], y = xyz[3 * (long long)ga + 1], z = xyz[3 * (long long)ga + 2];
      // (tore - pop) R  - 2 dd a0 P[s, p_k]   (sp hybridisation term of the heavy atoms)
      dx += qa * x;
      dy += qa * y;
      dz += qa * z;
      if (a < v.nheavy) {
        const double h = 2.0 * par(b, SEQM_P_DD, ga) * a0;
        dx -= h * Pm[oa * n + oa + 1];
        dy -= h * Pm[oa * n + oa + 2];
        dz -= h * Pm[oa * n + oa + 3];
      }
      if (force) {
        double* f = force + ((long long)v.m * b.molsize + a) * 3;
        f[0] = -grad[3 * (long long)ga];
        f[1] = -grad[3 * (long long)ga + 1];
        f[2] = -grad[3 * (long long)ga + 2];
      }
    } else if (force) {
      double* f = force + ((long long)v.m * b.molsize + a) * 3;
      f[0] = f[1] = f[2] = 0.0;
    }
    q[(long long)v.m * b.molsize + a] = qa;
  }
  if (dipole) {
    dx = block_sum(dx, red);
    dy = block_sum(dy, red);
    dz = block_sum(dz, red);
    if (threadIdx.x == 0) {
      dipole[3 * (long long)v.m] = dx * scale;
      dipole[3 * (long long)v.m + 1] = dy * scale;
      dipole[3 * (long long)v.m + 2] = dz * scale;
    }
  }
}

extern "C" int seqm_post_scf(const seqm_batch_t* b, const double* P, const double* xyz, const double* grad, double* q,
                             double* dipole, double* force, double a0, double scale, void* stream) {
  if (!b || !P || !xyz || !q || (force && !grad)) {
    seqm_set_error("seqm_post_scf: null pointer");
    return SEQM_ERR_ARG;
  }
#ifndef SEQM_HOSTEMU
  cudaStream_t st = (cudaStream_t)stream;
#else
  cudaStream_t st = stream;
#endif
  SEQM_LAUNCH(post_scf_kernel, b->nmol, 64, 0, st, *b, P, xyz, grad, q, dipole, force, a0, scale);
  return seqm_check_launch("post_scf_kernel");
}
