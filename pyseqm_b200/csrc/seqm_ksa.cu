// seqm_ksa.cu -- translation unit of libseqm_b200.so for KSA-XL-BOMD: dense per-molecule algebra on
// the packed layout around the Fock-contraction and eigensolver kernels of the SCF path.
//   packed_gemm_kernel        C_m = op(A_m) op(B_m) for every molecule (congruence transforms Q^t F1 Q, Q X Q^t; D = Q f Q^t)
//   scale_columns_kernel      out_m[i][k] = s * f[m][k] * C_m[i][k]          (Fermi_Q: D0 = 2 (Q f) Q^t, fermi_q.py:59-61)
//   canon_prt_kernel          recursive Fermi-operator expansion of the first-order density response in the eigenbasis and the
//                             chemical-potential correction (Canon_DM_PRT, canon_dm_prt.py:17-34; Alg. 2 of JCTC 16, 3628)
//   packed_dot / axpy / scale per-molecule Frobenius products and updates of the Krylov vectors (xlbomd.py:253-333)
#define SEQM_SECONDARY_TU
#include "common.cuh"

#define KSA_TM 64
#define KSA_TK 16
// one CTA per molecule; 64x64 output tiles, 4x4 outputs per work item, operands staged through shared memory
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS(256) packed_gemm_kernel(seqm_batch_t b, const double* __restrict__ A,
                                                            const double* __restrict__ B, double* __restrict__ C, int ta, int tb) {
  __shared__ double sA[KSA_TK][KSA_TM + 1];  // [k][i]
  __shared__ double sB[KSA_TK][KSA_TM + 1];  // [k][j]
  const MolView v = mol_view(b, blockIdx.x);
  const int n = v.n, tid = threadIdx.x, nthr = blockDim.x;
  const double* Am = A + v.mat0;
  const double* Bm = B + v.mat0;
  double* Cm = C + v.mat0;
  for (int i0 = 0; i0 < n; i0 += KSA_TM)
    for (int j0 = 0; j0 < n; j0 += KSA_TM) {
#ifndef SEQM_HOSTEMU
      double acc[4][4];
      const int oi = (tid >> 4) * 4, oj = (tid & 15) * 4;
      for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
      for (int k0 = 0; k0 < n; k0 += KSA_TK) {
        for (int t = tid; t < KSA_TK * KSA_TM; t += nthr) {
          const int kk = t / KSA_TM, ii = t % KSA_TM;
          const int gi = i0 + ii, gk = k0 + kk, gj = j0 + ii;
          sA[kk][ii] = (gi < n && gk < n) ? (ta ? Am[gk * n + gi] : Am[gi * n + gk]) : 0.0;
          sB[kk][ii] = (gj < n && gk < n) ? (tb ? Bm[gj * n + gk] : Bm[gk * n + gj]) : 0.0;
        }
        SEQM_SYNC();
        for (int kk = 0; kk < KSA_TK; ++kk) {
          double a[4], c[4];
          for (int x = 0; x < 4; ++x) a[x] = sA[kk][oi + x];
          for (int y = 0; y < 4; ++y) c[y] = sB[kk][oj + y];
          for (int x = 0; x < 4; ++x)
            for (int y = 0; y < 4; ++y) acc[x][y] += a[x] * c[y];
        }
        SEQM_SYNC();
      }
      for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 4; ++y)
          if (i0 + oi + x < n && j0 + oj + y < n) Cm[(i0 + oi + x) * n + j0 + oj + y] = acc[x][y];
#else  // host emulation (one thread per CTA): plain loops over the tile
      (void)sA;
      (void)sB;
      (void)tid;
      (void)nthr;
      for (int gi = i0; gi < i0 + KSA_TM && gi < n; ++gi)
        for (int gj = j0; gj < j0 + KSA_TM && gj < n; ++gj) {
          double s = 0.0;
          for (int k = 0; k < n; ++k) s += (ta ? Am[k * n + gi] : Am[gi * n + k]) * (tb ? Bm[gj * n + k] : Bm[k * n + gj]);
          Cm[gi * n + gj] = s;
        }
#endif
    }
}

#ifndef SEQM_HOSTEMU
// FP64 tensor-core version for molecules whose op(B) fits the SM's shared memory (n <= ~160): op(B) zero-padded to a multiple
// of 8 with a row stride of 4 mod 16 (conflict-free fragment loads), op(A) fragments straight from global memory / L1, one
// warp per 8x8 output tile, mma.sync.m8n8k4.f64 (DMMA) -- the scheme of the eigensolver's warm-start transform.
SEQM_GLOBAL void SEQM_LAUNCH_BOUNDS(256) packed_gemm_dmma_kernel(seqm_batch_t b, const double* __restrict__ A,
                                                                 const double* __restrict__ B, double* __restrict__ C, int ta, int tb) {
  SEQM_DYN_SMEM(double, S);
  const MolView v = mol_view(b, blockIdx.x);
  const int n = v.n, tid = threadIdx.x, nthr = blockDim.x;
  const double* Am = A + v.mat0;
  const double* Bm = B + v.mat0;
  double* Cm = C + v.mat0;
  const int np8 = (n + 7) & ~7, nt8 = np8 >> 3, LDT = np8 + 4;
  for (int t = tid; t < np8 * LDT; t += nthr) {
    const int k = t / LDT, j = t - k * LDT;
    S[t] = (k < n && j < n) ? (tb ? Bm[j * n + k] : Bm[k * n + j]) : 0.0;
  }
  __syncthreads();
  const int warp = tid >> 5, nwarps = nthr >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  for (int tile = warp; tile < nt8 * nt8; tile += nwarps) {
    const int m0 = (tile / nt8) * 8, n0 = (tile % nt8) * 8;
    const int r = m0 + g;
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 0; k0 < np8; k0 += 16) {
      double af[4], bf[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kc = k0 + 4 * u + t4;
        af[u] = (r < n && kc < n) ? (ta ? Am[kc * n + r] : Am[r * n + kc]) : 0.0;
        bf[u] = (kc < np8) ? S[kc * LDT + n0 + g] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1)
                     : "d"(af[u]), "d"(bf[u]));
    }
    const int c = n0 + 2 * t4;
    if (r < n) {
      if (c < n) Cm[r * n + c] = c0;
      if (c + 1 < n) Cm[r * n + c + 1] = c1;
    }
  }
}
#endif

SEQM_GLOBAL void scale_columns_kernel(seqm_batch_t b, const double* __restrict__ C, const double* __restrict__ f, double s,
                                      double* __restrict__ out) {
  const MolView v = mol_view(b, blockIdx.x);
  const int n = v.n;
  const double* fm = f + (long long)v.m * b.nmax;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) out[v.mat0 + t] = s * fm[t % n] * C[v.mat0 + t];
}

// X: first-order perturbation in the eigenbasis of the unperturbed Fock matrix (packed, in place); e (nmol, nmax), mu (nmol)
SEQM_GLOBAL void canon_prt_kernel(seqm_batch_t b, const double* __restrict__ e, const double* __restrict__ mu,
                                  double* __restrict__ X, double beta, int m_iter) {
  __shared__ double red[33];
  const MolView v = mol_view(b, blockIdx.x);
  const int n = v.n;
  const double* em = e + (long long)v.m * b.nmax;
  const double mu0 = mu[v.m];
  double* Xm = X + v.mat0;
  const double cnst = exp2((double)(-2 - m_iter)) * beta;
  double tr = 0.0;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
    const int i = t / n, j = t - i * n;
    double pi = 0.5 - cnst * (em[i] - mu0), pj = 0.5 - cnst * (em[j] - mu0);
    double x = -cnst * Xm[t];
    for (int it = 0; it < m_iter; ++it) {
      const double pi2 = pi * pi, pj2 = pj * pj;
      const double dx = pi * x + x * pj;
      const double idi = 1.0 / (2.0 * (pi2 - pi) + 1.0), idj = 1.0 / (2.0 * (pj2 - pj) + 1.0);
      pi = idi * pi2;
      pj = idj * pj2;
      x = idi * (dx + 2.0 * (x - dx) * pj);
    }
    Xm[t] = x;
    if (i == j) tr += x;
  }
  double dsum = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double p = 0.5 - cnst * (em[i] - mu0);
    for (int it = 0; it < m_iter; ++it) {
      const double p2 = p * p;
      p = p2 / (2.0 * (p2 - p) + 1.0);
    }
    dsum += beta * p * (1.0 - p);
  }
  tr = block_sum(tr, red);
  dsum = block_sum(dsum, red);
  const double dmu1 = -tr / dsum;
  SEQM_SYNC();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double p = 0.5 - cnst * (em[i] - mu0);
    for (int it = 0; it < m_iter; ++it) {
      const double p2 = p * p;
      p = p2 / (2.0 * (p2 - p) + 1.0);
    }
    Xm[i * n + i] += beta * p * (1.0 - p) * dmu1;
  }
}

SEQM_GLOBAL void packed_dot_kernel(seqm_batch_t b, const double* __restrict__ X, const double* __restrict__ Y,
                                   double* __restrict__ out) {
  __shared__ double red[33];
  const MolView v = mol_view(b, blockIdx.x);
  double s = 0.0;
  for (int t = threadIdx.x; t < v.n * v.n; t += blockDim.x) s += X[v.mat0 + t] * Y[v.mat0 + t];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[v.m] = s;
}
// Y_m = a_m X_m + c_m Y_m   (a == NULL: a_m = 1; c == NULL: c_m = 1; X == NULL: Y_m = c_m Y_m)
SEQM_GLOBAL void packed_axpby_kernel(seqm_batch_t b, const double* __restrict__ a, const double* __restrict__ X,
                                     const double* __restrict__ c, double* __restrict__ Y) {
  const MolView v = mol_view(b, blockIdx.x);
  const double am = a ? a[v.m] : 1.0, cm = c ? c[v.m] : 1.0;
  for (int t = threadIdx.x; t < v.n * v.n; t += blockDim.x)
    Y[v.mat0 + t] = (X ? am * X[v.mat0 + t] : 0.0) + cm * Y[v.mat0 + t];
}

// Response of the two-electron part of the Fock operator to an ANTISYMMETRIC density Pa (the part of a CIS transition density
// the symmetric Fock build does not see; makeA_pi_batched, rcis_batch.py:336-381): only exchange survives,
//   Fa_AB[mu][la] = -1/2 sum_{nu in A, sg in B} Pa[nu][sg] (mu nu | la sg),  Fa_BA = -Fa_AB^t,
// plus the one-centre exchange terms (s,p): Pa (hsp - gsp)/2, (p,p'): Pa (gpp/4 - 3 gp2/4).  One CTA per molecule.
SEQM_GLOBAL void fock_antisym_kernel(seqm_batch_t b, const double* __restrict__ Pa, const double* __restrict__ w,
                                     double* __restrict__ Fa) {
  const MolView v = mol_view(b, blockIdx.x);
  const int n = v.n;
  const double* Pm = Pa + v.mat0;
  double* Fm = Fa + v.mat0;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) Fm[t] = 0.0;
  SEQM_SYNC();
  for (int t = threadIdx.x; t < v.npair * 16; t += blockDim.x) {
    const int pl = t >> 4, mu = (t >> 2) & 3, la = t & 3;
    const int p = v.p0 + pl;
    const int i = b.pair_i[p] - v.a0, j = b.pair_j[p] - v.a0;
    const int ni = orb_cnt(v, i), nj = orb_cnt(v, j);
    if (mu >= ni || la >= nj) continue;
    const int oi = orb_off(v, i), oj = orb_off(v, j);
    const double* wp = w + (long long)p * 100;
    double k = 0.0;
    for (int nu = 0; nu < ni; ++nu)
      for (int sg = 0; sg < nj; ++sg) k += Pm[(oi + nu) * n + oj + sg] * wp[pack2(mu, nu) * 10 + pack2(la, sg)];
    Fm[(oi + mu) * n + oj + la] = -0.5 * k;
    Fm[(oj + la) * n + oi + mu] = 0.5 * k;
  }
  for (int t = threadIdx.x; t < v.nheavy * 6; t += blockDim.x) {
    const int a = t / 6, e = t % 6;  // upper-triangle element of the 4 x 4 block: (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
    const int mu = (e < 3) ? 0 : ((e < 5) ? 1 : 2);
    const int nu = (e < 3) ? e + 1 : ((e < 5) ? e - 1 : 3);
    const int oa = orb_off(v, a), ga = v.a0 + a;
    const double c = (mu == 0) ? 0.5 * (par(b, SEQM_P_HSP, ga) - par(b, SEQM_P_GSP, ga))
                               : 0.25 * par(b, SEQM_P_GPP, ga) - 0.75 * par(b, SEQM_P_GP2, ga);
    const double f = Pm[(oa + mu) * n + oa + nu] * c;
    Fm[(oa + mu) * n + oa + nu] = f;
    Fm[(oa + nu) * n + oa + mu] = -f;
  }
}

SEQM_GLOBAL void packed_transpose_kernel(seqm_batch_t b, const double* __restrict__ X, double* __restrict__ XT) {
  const MolView v = mol_view(b, blockIdx.x);
  const int n = v.n;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) XT[v.mat0 + t] = X[v.mat0 + (t % n) * n + t / n];
}

static int ksa_check(const seqm_batch_t* b, const void* p, const void* q, const char* what) {
  if (!b || !p || !q) {
    seqm_set_error("%s: null pointer", what);
    return SEQM_ERR_ARG;
  }
  return SEQM_OK;
}
#ifndef SEQM_HOSTEMU
#define KSA_STREAM(s) ((cudaStream_t)(s))
#else
#define KSA_STREAM(s) (s)
#endif

extern "C" {
int seqm_packed_gemm(const seqm_batch_t* b, const double* A, const double* B, double* C, int transA, int transB, void* stream) {
  int rc = ksa_check(b, A, B, "seqm_packed_gemm");
  if (rc) return rc;
  if (!C || C == A || C == B) {
    seqm_set_error("seqm_packed_gemm: the result must be a distinct buffer");
    return SEQM_ERR_ARG;
  }
#ifndef SEQM_HOSTEMU
  {
    static int smem_max = -1;  // one device per process (seqm_b200.cu: ensure_device)
    if (smem_max < 0) {
      int dev = 0, optin = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      smem_max = optin - 1024;
      if (cudaFuncSetAttribute(packed_gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max) != cudaSuccess) {
        cudaGetLastError();
        smem_max = 0;
      }
    }
    const int np8 = (b->nmax + 7) & ~7;
    const size_t need = sizeof(double) * (size_t)np8 * (np8 + 4);
    if (need <= (size_t)smem_max) {
      SEQM_LAUNCH(packed_gemm_dmma_kernel, b->nmol, 256, need, KSA_STREAM(stream), *b, A, B, C, transA, transB);
      return seqm_check_launch("packed_gemm_dmma_kernel");
    }
  }
#endif
  SEQM_LAUNCH(packed_gemm_kernel, b->nmol, 256, 0, KSA_STREAM(stream), *b, A, B, C, transA, transB);
  return seqm_check_launch("packed_gemm_kernel");
}
int seqm_scale_columns(const seqm_batch_t* b, const double* C, const double* f, double s, double* out, void* stream) {
  int rc = ksa_check(b, C, f, "seqm_scale_columns");
  if (rc) return rc;
  SEQM_LAUNCH(scale_columns_kernel, b->nmol, 256, 0, KSA_STREAM(stream), *b, C, f, s, out);
  return seqm_check_launch("scale_columns_kernel");
}
int seqm_canon_prt(const seqm_batch_t* b, const double* e, const double* mu, double* X, double beta, int m_iter, void* stream) {
  int rc = ksa_check(b, e, mu, "seqm_canon_prt");
  if (rc) return rc;
  SEQM_LAUNCH(canon_prt_kernel, b->nmol, 256, 0, KSA_STREAM(stream), *b, e, mu, X, beta, m_iter);
  return seqm_check_launch("canon_prt_kernel");
}
int seqm_packed_dot(const seqm_batch_t* b, const double* X, const double* Y, double* out, void* stream) {
  int rc = ksa_check(b, X, Y, "seqm_packed_dot");
  if (rc) return rc;
  SEQM_LAUNCH(packed_dot_kernel, b->nmol, 256, 0, KSA_STREAM(stream), *b, X, Y, out);
  return seqm_check_launch("packed_dot_kernel");
}
int seqm_packed_transpose(const seqm_batch_t* b, const double* X, double* XT, void* stream) {
  int rc = ksa_check(b, X, XT, "seqm_packed_transpose");
  if (rc) return rc;
  if (X == XT) {
    seqm_set_error("seqm_packed_transpose: out of place only");
    return SEQM_ERR_ARG;
  }
  SEQM_LAUNCH(packed_transpose_kernel, b->nmol, 256, 0, KSA_STREAM(stream), *b, X, XT);
  return seqm_check_launch("packed_transpose_kernel");
}
int seqm_fock_antisym(const seqm_batch_t* b, const double* Pa, const double* w, double* Fa, void* stream) {
  int rc = ksa_check(b, Pa, Fa, "seqm_fock_antisym");
  if (rc) return rc;
  if (b->method == SEQM_PM6_D) {
    seqm_set_error("seqm_fock_antisym: sp methods only");
    return SEQM_ERR_UNSUPPORTED;
  }
  if (b->npairs > 0 && !w) {
    seqm_set_error("seqm_fock_antisym: w is NULL");
    return SEQM_ERR_ARG;
  }
  SEQM_LAUNCH(fock_antisym_kernel, b->nmol, 256, 0, KSA_STREAM(stream), *b, Pa, w, Fa);
  return seqm_check_launch("fock_antisym_kernel");
}
int seqm_packed_axpby(const seqm_batch_t* b, const double* a, const double* X, const double* c, double* Y, void* stream) {
  if (!b || !Y) {
    seqm_set_error("seqm_packed_axpby: null pointer");
    return SEQM_ERR_ARG;
  }
  SEQM_LAUNCH(packed_axpby_kernel, b->nmol, 256, 0, KSA_STREAM(stream), *b, a, X, c, Y);
  return seqm_check_launch("packed_axpby_kernel");
}
}
