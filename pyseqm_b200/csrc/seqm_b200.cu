// seqm_b200.cu -- C-ABI entry points of libseqm_b200.so (see include/seqm_b200.h) and the SCF host loop.
// Single translation unit: all kernels live in the .cuh files included below.
#include <mutex>
#include <vector>

#include "atom_kernels.cuh"
#include "scf_driver.cuh"
#include "plan_kernels.cuh"

// PM6 d-orbital kernels: second translation unit (seqm_spd.cu)
// seqm_pair.cu (the pair-code kernels; pairtu_ensure_tables is declared in common.cuh)
int pairtu_launch_integrals(const seqm_batch_t* b, int cls, int grid, int block, const double* xyz, double* w, double* hab,
                            cudaStream_t st);
int pairtu_launch_gradient(const seqm_batch_t* b, int cls, int grid, int block, const double* xyz, const double* D,
                           const double* P, double* gp, cudaStream_t st);
int pairtu_launch_gradient_forward(const seqm_batch_t* b, int grid, int block, const double* xyz, const double* P, double* gp,
                                   cudaStream_t st);
int pairtu_launch_nuclear(const seqm_batch_t* b, int grid, int block, const double* xyz, const double* w, double* EnucAB,
                          cudaStream_t st);
int spd_set_attributes(int smem_optin);
int spd_launch_pair(const seqm_batch_t* b, const double* xyz, const double* w, double* wd, double* hab_d, cudaStream_t st);
int spd_launch_hcore(const seqm_batch_t* b, const double* w, const double* hab, double* H, cudaStream_t st);
int spd_launch_fock(const seqm_batch_t* b, const double* P, const double* H, const double* w, double* F,
                    const int32_t* active, int smem_limit, cudaStream_t st);
int spd_launch_gradient(const seqm_batch_t* b, const double* xyz, const double* P, double* gp, cudaStream_t st);
// mid-size eigensolver (one-sided Jacobi, 119..256 orbitals): third translation unit (seqm_eigh.cu)
int hestenes_set_attributes(int smem_optin);
int hestenes_max_orbitals(void);
int hestenes_launch(const seqm_batch_t* b, const double* F, double* P, double* evals, double* C, const int32_t* active,
                    int smem_optin, cudaStream_t st);

#ifndef SEQM_HOSTEMU
#define SEQM_STREAM(s) ((cudaStream_t)(s))
#else
#define SEQM_STREAM(s) (s)
#endif


// ---- optional per-kernel timing (CUDA events on the launch stream; used by bench.py for the roofline) ----
enum { PK_PAIR = 0, PK_HCORE, PK_FOCK, PK_JACOBI, PK_SP2, PK_DIIS_STORE, PK_DIIS_SOLVE, PK_DIIS_EXTRAP, PK_MIX,
       PK_ENERGY_ERR, PK_NUC, PK_GRAD, PK_OTHER, PK_GEMM, PK_COUNT };
static const char* g_pk_names[PK_COUNT] = {"pair_integrals", "hcore", "fock", "jacobi_density", "sp2", "diis_store",
                                           "diis_solve", "diis_extrapolate", "mix", "energy_error", "nuclear_energy",
                                           "gradient", "other", "dgemm"};
static int g_prof_on = 0;
#ifndef SEQM_HOSTEMU
#define SEQM_PROF_MAX 32768
static cudaEvent_t g_ev0[SEQM_PROF_MAX], g_ev1[SEQM_PROF_MAX];
static int g_ev_kind[SEQM_PROF_MAX];
static int g_ev_n = 0, g_ev_created = 0;
static void prof_begin(int kind, cudaStream_t st) {
  if (!g_prof_on || g_ev_n >= SEQM_PROF_MAX) return;
  if (g_ev_n >= g_ev_created) {
    cudaEventCreate(&g_ev0[g_ev_n]);
    cudaEventCreate(&g_ev1[g_ev_n]);
    g_ev_created = g_ev_n + 1;
  }
  g_ev_kind[g_ev_n] = kind;
  cudaEventRecord(g_ev0[g_ev_n], st);
}
static void prof_end(cudaStream_t st) {
  if (!g_prof_on || g_ev_n >= SEQM_PROF_MAX) return;
  cudaEventRecord(g_ev1[g_ev_n], st);
  ++g_ev_n;
}
#else
static void prof_begin(int, cudaStream_t) {}
static void prof_end(cudaStream_t) {}
#endif
#define PROF(kind, st, stmt) do { prof_begin(kind, st); stmt; prof_end(st); } while (0)

static int g_num_sms = 148;
static int g_smem_optin = 227 * 1024;
static int g_dev_ready = 0;
static int g_device = -1;  // the ONE device this process-wide library state (function attributes, streams, events) is for
// Process-wide state (opt-in shared-memory attributes, eigensolver class streams, pipeline streams/events, the pinned
// convergence mailbox) belongs to one device and one caller at a time: one process per GPU.  A call
// with another current device fails loudly; concurrent calls from several host threads are serialised by g_api_mutex.
static std::mutex g_api_mutex;
#define SEQM_SERIAL std::lock_guard<std::mutex> seqm_serial_guard(g_api_mutex)
static int ensure_device() {
#ifndef SEQM_HOSTEMU
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (g_dev_ready) {
    if (e == cudaSuccess && dev != g_device) {
      seqm_set_error("libseqm_b200 holds per-device state for cuda:%d but the current device is cuda:%d: run one process "
                     "per GPU (torch.cuda.set_device(LOCAL_RANK) before the first call)", g_device, dev);
      return SEQM_ERR_UNSUPPORTED;
    }
    return SEQM_OK;
  }
  g_device = dev;
#else
  if (g_dev_ready) return SEQM_OK;
#endif
#ifndef SEQM_HOSTEMU
  if (e != cudaSuccess) {
    seqm_set_error("cudaGetDevice: %s (libseqm_b200 needs a CUDA device; there is no CPU fallback)", cudaGetErrorString(e));
    return SEQM_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int dyn = g_smem_optin - 2048;  // leave room for the kernels' small static shared arrays
  e = cudaSuccess;
#define SEQM_ATTR(NPV)                                                                                                 \
  if (e == cudaSuccess && JacobiCfg<NPV>::SMEM > 48 * 1024)                                                             \
    e = cudaFuncSetAttribute(jacobi_fixed_kernel<NPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JacobiCfg<NPV>::SMEM);
  SEQM_JACOBI_CLASSES(SEQM_ATTR)
#undef SEQM_ATTR
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEQM_DMMA_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(sp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(dgemm_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEQM_SYM_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fock_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(diis_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  if (e == cudaSuccess && spd_set_attributes(g_smem_optin) != SEQM_OK) return SEQM_ERR_CUDA;
  if (e == cudaSuccess && hestenes_set_attributes(g_smem_optin) != SEQM_OK) return SEQM_ERR_CUDA;
  if (e != cudaSuccess) {
    seqm_set_error("cudaFuncSetAttribute(max dynamic shared memory %d): %s", dyn, cudaGetErrorString(e));
    cudaGetLastError();
    return SEQM_ERR_CUDA;
  }
#endif
  g_dev_ready = 1;
  return pairtu_ensure_tables();
}
static int check_batch(const seqm_batch_t* b) {
  if (!b || b->nmol <= 0 || b->nat <= 0) {
    seqm_set_error("empty batch");
    return SEQM_ERR_ARG;
  }
  if (b->method < 0 || b->method > 4 || (b->method >= 3 && !b->pw_alpha) ||
      (b->method == SEQM_PM6_D && (!b->mol_nsh || !b->pair_wd0 || !b->mp_coef || !b->ovl_poly || !b->onecenter_d))) {
    seqm_set_error("method %d not supported by this build (MNDO=0, AM1=1, PM3=2, PM6_SP=3 and PM6_D=4 with pairwise tables; "
                   "PM6_D also needs the d-orbital plan arrays)", b->method);
    return SEQM_ERR_UNSUPPORTED;
  }
  return ensure_device();
}
static int check_small(const seqm_batch_t* b, const char* what) {
  if (b->nmax > SEQM_MAX_ORB) {
    seqm_set_error("%s: a molecule with %d orbitals exceeds the shared-memory resident limit of %d (large molecules: "
                   "density by SP2 only, sp2=[True, eps])", what, b->nmax, SEQM_MAX_ORB);
    return SEQM_ERR_TOO_LARGE;
  }
  return SEQM_OK;
}

static int grid1d(long long n, int block);
// ---- large-molecule helpers (host side) ------------------------------------------------------------------
struct HostMol { long long mat0; int n, nocc; };
static int fetch_host_mols(const seqm_batch_t* b, HostMol* hm, cudaStream_t st) {
  std::vector<long long> m0v(b->nmol + 1);
  std::vector<int> nhv(b->nmol), nyv(b->nmol), nov(b->nmol);
  long long* m0 = m0v.data();
  int *nh = nhv.data(), *ny = nyv.data(), *no = nov.data();
#ifndef SEQM_HOSTEMU
  cudaError_t e = cudaMemcpyAsync(m0, b->mol_mat0, sizeof(long long) * (b->nmol + 1), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(nh, b->mol_nheavy, sizeof(int) * b->nmol, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ny, b->mol_nhyd, sizeof(int) * b->nmol, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(no, b->mol_nocc, sizeof(int) * b->nmol, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { seqm_set_error("fetch_host_mols: %s", cudaGetErrorString(e)); return SEQM_ERR_CUDA; }
#else
  (void)st;
  memcpy(m0, b->mol_mat0, sizeof(long long) * (b->nmol + 1));
  memcpy(nh, b->mol_nheavy, sizeof(int) * b->nmol);
  memcpy(ny, b->mol_nhyd, sizeof(int) * b->nmol);
  memcpy(no, b->mol_nocc, sizeof(int) * b->nmol);
#endif
  for (int m = 0; m < b->nmol; ++m) { hm[m].mat0 = m0[m]; hm[m].n = 4 * nh[m] + ny[m]; hm[m].nocc = no[m]; }  // large path: sp methods only
  return SEQM_OK;
}
static int fetch_host_ints(const int32_t* dev, int32_t* host, int n, cudaStream_t st) {
#ifndef SEQM_HOSTEMU
  cudaError_t e = cudaMemcpyAsync(host, dev, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { seqm_set_error("fetch_host_ints: %s", cudaGetErrorString(e)); return SEQM_ERR_CUDA; }
#else
  (void)st;
  memcpy(host, dev, sizeof(int32_t) * n);
#endif
  return SEQM_OK;
}
static int launch_gemm(int n, const double* A, const double* B, double* C, cudaStream_t st) {
  const int nb = (n + SEQM_GEMM_BM - 1) / SEQM_GEMM_BM;
#ifndef SEQM_HOSTEMU
  static const int use_fma = getenv("SEQM_B200_GEMM_FMA") ? 1 : 0;  // experiments: the register-tiled DFMA kernel
  if (!use_fma) {
    PROF(PK_GEMM, st, SEQM_LAUNCH(dgemm_dmma_kernel, nb * nb, 256, SEQM_DMMA_SMEM, st, n, n, n, A, n, B, n, C, n));
    return seqm_check_launch("dgemm_dmma_kernel");
  }
#endif
  PROF(PK_GEMM, st, SEQM_LAUNCH(dgemm_kernel, nb * nb, 256, 0, st, n, n, n, A, n, B, n, C, n));
  return seqm_check_launch("dgemm_kernel");
}
// C = X X for a symmetric X: upper-triangle tiles only (dgemm_sym_kernel); skip: device flag that voids the launch
static int launch_gemm_sym(int n, const double* X, double* C, const int* skip, cudaStream_t st) {
  const int nb = (n + SEQM_SYM_TB - 1) / SEQM_SYM_TB;
  PROF(PK_GEMM, st, SEQM_LAUNCH(dgemm_sym_kernel, nb * (nb + 1) / 2, 256, SEQM_SYM_SMEM, st, n, X, C, skip));
  return seqm_check_launch("dgemm_sym_kernel");
}
// SP2 purification of one large molecule: X, X2 scratch of n*n doubles, state on the device; P = 2 X.
// The iterations are queued in chunks without reading the state back: once the decision step has seen convergence the
// remaining kernels of the chunk return at once (Sp2State.done / finished).  The first chunk is as long as the previous
// solve needed (consecutive SCF iterations need the same number of purification steps to within one or two).
static int g_sp2_guess = 0;
static int sp2_large_one(int n, int nocc, const double* Fm, double* Pm, double eps, double* X, double* X2, Sp2State* stt,
                         int* iters_out, cudaStream_t st) {
  if (eps > 1.0e-3) eps = 1.0e-3;
  if (eps < 1.0e-7) eps = 1.0e-7;
  const long long nn = (long long)n * n;
  const int ge = grid1d(nn, 256);
  SEQM_LAUNCH(sp2_bounds_kernel, 1, 1024, 0, st, n, Fm, (double)nocc, stt);
  SEQM_LAUNCH(sp2_init_kernel, ge, 256, 0, st, n, Fm, X, (const Sp2State*)stt);
  SEQM_LAUNCH(sp2_trace_kernel, 1, 1024, 0, st, n, (const double*)X, (const double*)nullptr, stt, eps, 1);
  int rc = seqm_check_launch("sp2 setup");
  if (rc) return rc;
  int queued = 0;
  for (;;) {
    const int chunk = (queued == 0 && g_sp2_guess > 0) ? g_sp2_guess : 4;
    for (int it = 0; it < chunk; ++it) {
      rc = launch_gemm_sym(n, X, X2, &stt->done, st);
      if (rc) return rc;
      PROF(PK_SP2, st, SEQM_LAUNCH(sp2_decide_kernel, 1, 1024, 0, st, n, (const double*)X, (const double*)X2, stt, eps));
      PROF(PK_SP2, st, SEQM_LAUNCH(sp2_update_kernel, ge, 256, 0, st, n, X, (const double*)X2, (const Sp2State*)stt));
    }
    queued += chunk;
    rc = seqm_check_launch("sp2 iteration");
    if (rc) return rc;
    Sp2State h;
#ifndef SEQM_HOSTEMU
    cudaError_t e = cudaMemcpyAsync(&h, stt, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { seqm_set_error("sp2 state read-back: %s", cudaGetErrorString(e)); return SEQM_ERR_CUDA; }
#else
    h = *stt;
#endif
    if (h.done) {
      if (iters_out) *iters_out = h.iters;
      g_sp2_guess = h.iters;
      break;
    }
  }
  SEQM_LAUNCH(scale_copy_kernel, ge, 256, 0, st, nn, (const double*)X, Pm, 2.0);
  return seqm_check_launch("scale_copy_kernel");
}

static int threads_for(int nmax);
// fe: optional get_error tail; *fused tells the caller whether the kernel that ran could carry it
static int launch_fock(const seqm_batch_t* b, const double* P, const double* H, const double* w, double* F,
                       const int32_t* active, cudaStream_t st, const FockErr* fe = nullptr, bool* fused = nullptr) {
  if (fused) *fused = false;
  FockErr none;
  memset(&none, 0, sizeof(none));
  static const int fock_bulk = (getenv("SEQM_B200_FOCK_BULK") && atoi(getenv("SEQM_B200_FOCK_BULK")) == 0) ? 0 : 1;
  none.bulk = fock_bulk;
  if (b->method == SEQM_PM6_D) {  // 9 x 9 atom blocks, ragged integral blocks of the pairs with a d atom
    if (!b->wd && b->n_ypairs > 0) {
      seqm_set_error("PM6 with d orbitals: b->wd is not set (seqm_pair_integrals_d comes first)");
      return SEQM_ERR_ARG;
    }
    int rcf = SEQM_OK;
    PROF(PK_FOCK, st, rcf = spd_launch_fock(b, P, H, w, F, active, g_smem_optin - 2048, st));
    return rcf;
  }
  if (b->nmax > SEQM_MAX_ORB) {  // matrices in global memory, grid over all pairs / atoms
    if (b->npairs > 0)
      PROF(PK_FOCK, st, SEQM_LAUNCH(fock_large_offdiag_kernel, grid1d((long long)b->npairs * 16, 256), 256, 0, st, *b, P, H, w, F, active));
    PROF(PK_FOCK, st, SEQM_LAUNCH(fock_large_diag_kernel, grid1d((long long)b->nat * 10, 128), 128, 0, st, *b, P, H, w, F, active));
    return seqm_check_launch("fock_large kernels");
  }
  const int nt = threads_for(b->nmax);
  const size_t sm_pair = fock_pair_smem_bytes(b->nmax, b->fock_scratch, nt);
  if (b->fock_scratch > 0 && sm_pair <= (size_t)(g_smem_optin - 2048)) {  // pair-centric: w read once
    FockErr fe2 = fe ? *fe : none;
    fe2.bulk = fock_bulk;
    PROF(PK_FOCK, st, SEQM_LAUNCH(fock_pair_kernel, b->nmol, nt, sm_pair, st, *b, P, H, w, F, active, fe2));
    if (fused) *fused = (fe != nullptr);
    return seqm_check_launch("fock_pair_kernel");
  }
  const size_t smem = sizeof(double) * (size_t)b->nmax * b->nmax;
  PROF(PK_FOCK, st, SEQM_LAUNCH(fock_kernel, b->nmol, nt, smem, st, *b, P, H, w, F, active));
  return seqm_check_launch("fock_kernel");
}
static int launch_pair_gradient(const seqm_batch_t* b, const double* xyz, const double* D, const double* P, double* gp,
                                cudaStream_t st) {
  const int n0 = b->pair_cls_off[1] - b->pair_cls_off[0], n1 = b->pair_cls_off[2] - b->pair_cls_off[1],
            n2 = b->pair_cls_off[3] - b->pair_cls_off[2];
  int rc = SEQM_OK;
  if (n0 > 0) PROF(PK_GRAD, st, rc = pairtu_launch_gradient(b, 0, grid1d(n0, 128), 128, xyz, D, P, gp, st));
  if (n1 > 0 && !rc) PROF(PK_GRAD, st, rc = pairtu_launch_gradient(b, 1, grid1d(n1, 64), 64, xyz, D, P, gp, st));
  if (n2 > 0 && !rc) PROF(PK_GRAD, st, rc = pairtu_launch_gradient(b, 2, grid1d(n2, 64), 64, xyz, D, P, gp, st));
  return rc;
}
static int diis_grid(int nmol) {
#ifdef SEQM_HOSTEMU
  return nmol;  // one single-thread "warp" per emulated CTA
#else
  return (nmol + SEQM_DIIS_WARPS - 1) / SEQM_DIIS_WARPS;
#endif
}
// Launch the fixed-slot Jacobi kernel once per populated size class (right-sized shared memory / threads).
// One molecule is a long dependent chain (sweeps x steps x barrier latency), so the classes are forked onto
// their own streams and run concurrently; the caller's stream joins them afterwards.
#ifndef SEQM_HOSTEMU
// two independent sets ("lanes"): the pipelined SCF runs the eigensolver of its two half-batches concurrently
static cudaStream_t g_cls_stream[2][16];
static cudaEvent_t g_cls_fork[2], g_cls_join[2][16];
static int g_cls_streams_ready = 0;
static int ensure_class_streams() {
  if (g_cls_streams_ready) return SEQM_OK;
  for (int l = 0; l < 2; ++l) {
    for (int c = 0; c < g_jacobi_ncls; ++c) {
      if (cudaStreamCreateWithFlags(&g_cls_stream[l][c], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&g_cls_join[l][c], cudaEventDisableTiming) != cudaSuccess) {
        seqm_set_error("could not create eigensolver class streams");
        return SEQM_ERR_CUDA;
      }
    }
    if (cudaEventCreateWithFlags(&g_cls_fork[l], cudaEventDisableTiming) != cudaSuccess) return SEQM_ERR_CUDA;
  }
  g_cls_streams_ready = 1;
  return SEQM_OK;
}
#endif
static int launch_jacobi(const seqm_batch_t* b, const double* F, double* P, double* evals, double* C, const double* Cguess,
                         const int32_t* active, cudaStream_t st, int lane = 0, JacobiMix mix = JacobiMix{nullptr, nullptr, nullptr}) {
  const double* cg = (Cguess && P && C) ? Cguess : nullptr;
  int npop = 0;
  for (int c = 0; c < g_jacobi_ncls; ++c) npop += (b->cls_count[c] > 0);
#ifndef SEQM_HOSTEMU
  const bool fork = npop > 1;
  if (fork) {
    int rc = ensure_class_streams();
    if (rc) return rc;
    cudaEventRecord(g_cls_fork[lane], st);
  }
#else
  const bool fork = false;
  (void)lane;
#endif
  for (int c = g_jacobi_ncls - 1; c >= 0; --c) {
    const int cnt = b->cls_count[c], first = b->cls_begin[c];
    if (cnt <= 0) continue;
    cudaStream_t cst = st;
#ifndef SEQM_HOSTEMU
    if (fork) {
      cst = g_cls_stream[lane][c];
      cudaStreamWaitEvent(cst, g_cls_fork[lane], 0);
    }
#endif
    switch (g_jacobi_np[c]) {
#define SEQM_CASE(NPV)                                                                                               \
  case NPV:                                                                                                          \
    SEQM_LAUNCH(jacobi_fixed_kernel<NPV>, cnt, JacobiCfg<NPV>::THREADS, JacobiCfg<NPV>::SMEM, cst, *b, first, F, P, evals, \
                C, cg, active, mix);                                                                                 \
    break;
      SEQM_JACOBI_CLASSES(SEQM_CASE)
#undef SEQM_CASE
    }
    int rc = seqm_check_launch("jacobi_fixed_kernel");
    if (rc) return rc;
#ifndef SEQM_HOSTEMU
    if (fork) {
      cudaEventRecord(g_cls_join[lane][c], cst);
      cudaStreamWaitEvent(st, g_cls_join[lane][c], 0);
    }
#endif
  }
  return SEQM_OK;
}
static int threads_for(int nmax) { return nmax <= 24 ? 128 : (nmax <= 64 ? 256 : 512); }
static int grid1d(long long n, int block) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  const long long cap = (long long)g_num_sms * 16;
  return (int)(g > cap ? cap : g);
}

// ---- layout conversion kernels ---------------------------------------------------------------------
// dense (nmol, S*molsize, S*molsize) <-> packed, S = 4 orbital slots per atom (9 for method="PM6" with d orbitals).
// Real atoms occupy the first na positions of a molecule.
SEQM_HD int dense_stride(const seqm_batch_t& b) { return b.method == SEQM_PM6_D ? 9 : 4; }
SEQM_HD int orb_atom(const MolView& v, int orb, int* slot) {  // packed orbital -> (local atom, orbital slot on it)
  const int nd = 9 * v.nsh, np_ = nd + 4 * (v.nheavy - v.nsh);
  if (orb < nd) { *slot = orb % 9; return orb / 9; }
  if (orb < np_) { *slot = (orb - nd) & 3; return v.nsh + ((orb - nd) >> 2); }
  *slot = 0;
  return v.nheavy + (orb - np_);
}
SEQM_HD int dense_index(const MolView& v, int orb, int S) {  // packed orbital -> row in the dense padded matrix
  int k;
  const int a = orb_atom(v, orb, &k);
  return S * a + k;
}
SEQM_GLOBAL void pack_kernel(seqm_batch_t b, const double* __restrict__ dense, double* __restrict__ packed) {
  const MolView v = mol_view(b, blockIdx.x);
  const int S = dense_stride(b), n = v.n, N = S * b.molsize;
  const double* D = dense + (long long)v.m * N * N;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x)
    packed[v.mat0 + t] = D[(long long)dense_index(v, t / n, S) * N + dense_index(v, t % n, S)];
}
SEQM_GLOBAL void unpack_kernel(seqm_batch_t b, const double* __restrict__ packed, double* __restrict__ dense) {
  const MolView v = mol_view(b, blockIdx.x);
  const int S = dense_stride(b), n = v.n, N = S * b.molsize;
  double* D = dense + (long long)v.m * N * N;
  for (int t = threadIdx.x; t < N * N; t += blockDim.x) D[t] = 0.0;
  SEQM_SYNC();
  for (int t = threadIdx.x; t < n * n; t += blockDim.x)
    D[(long long)dense_index(v, t / n, S) * N + dense_index(v, t % n, S)] = packed[v.mat0 + t];
}
// packed eigenvector matrices -> (nmol, nmax, nmax) with the identity on the padding (diag.py:110-241 `v`)
SEQM_GLOBAL void orbitals_dense_kernel(seqm_batch_t b, const double* __restrict__ C, double* __restrict__ V) {
  const MolView v = mol_view(b, blockIdx.x);
  const int n = v.n, N = b.nmax;
  double* Vm = V + (long long)v.m * N * N;
  for (int t = threadIdx.x; t < N * N; t += blockDim.x) {
    const int i = t / N, j = t - i * N;
    Vm[t] = (i < n && j < n) ? C[v.mat0 + i * n + j] : ((i == j) ? 1.0 : 0.0);
  }
}
// ---- MO crossing matcher (Energy._crossing_match_molecular_orbitals[_grouped], basics.py:596-719) ---------------
// Orbitals are dense (nmol, N, N), N = nmax, column k = MO k (the layout of `molecular_orbitals`).
// Step 1: signed overlaps S[k][l] = sum_r Vold[r][k] Vnew[r][l], old k and new l both occupied or both virtual, in a
// packed-size scratch buffer (nocc^2 + nvirt^2 <= n^2 doubles per molecule): occupied block first.
SEQM_GLOBAL void mo_overlap_kernel(seqm_batch_t b, int parts, const double* __restrict__ Vold,
                                   const double* __restrict__ Vnew, double* __restrict__ S) {
  const MolView v = mol_view(b, blockIdx.x / parts);
  const int part = blockIdx.x % parts;
  const int n = v.n, N = b.nmax, no = v.nocc, nv = n - no;
  const double* Co = Vold + (long long)v.m * N * N;
  const double* Cn = Vnew + (long long)v.m * N * N;
  const int total = no * no + nv * nv;
  for (int t = part * blockDim.x + threadIdx.x; t < total; t += parts * blockDim.x) {
    int k, l;
    if (t < no * no) {
      k = t / no;
      l = t - k * no;
    } else {
      const int u = t - no * no;
      k = u / nv;
      l = no + u - k * nv;
      k += no;
    }
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += Co[(long long)r * N + k] * Cn[(long long)r * N + l];
    S[v.mat0 + t] = s;
  }
}
// Step 2, one CTA per molecule and block (occupied, virtual): every old orbital takes the new orbital of largest
// |overlap| (basics.py:651-653); when that is not a permutation the rows are served greedily in the order of
// their margin top1 - top2, each taking its best still unused column (greedy_unique_perm, basics.py:606-626).
// Then V_out[:, k] = sign(S[k][p_k]) V_new[:, p_k] (sign 0 -> +1, basics.py:631-635) and e_out[k] = e[p_k].
SEQM_GLOBAL void mo_match_kernel(seqm_batch_t b, const double* __restrict__ Vnew, const double* __restrict__ S,
                                 const double* __restrict__ e_in, int32_t* __restrict__ perm, int32_t* __restrict__ used,
                                 double* __restrict__ prio, double* __restrict__ Vout, double* __restrict__ e_out) {
  const MolView v = mol_view(b, blockIdx.x);
  const int n = v.n, N = b.nmax;
  int32_t* p = perm + (long long)v.m * N;
  int32_t* u = used + (long long)v.m * N;
  double* pr = prio + (long long)v.m * N;
  for (int blk = 0; blk < 2; ++blk) {
    const int o = blk ? v.nocc : 0, r = blk ? n - v.nocc : v.nocc;
    if (r == 0) continue;
    const double* Sb = S + v.mat0 + (blk ? v.nocc * v.nocc : 0);
    for (int t = threadIdx.x; t < r; t += blockDim.x) u[o + t] = 0;
    SEQM_SYNC();
    for (int t = threadIdx.x; t < r; t += blockDim.x) {
      double b1 = -1.0, b2 = -1.0;
      int i1 = 0;
      for (int c = 0; c < r; ++c) {
        const double a = fabs(Sb[(long long)t * r + c]);
        if (a > b1) {
          b2 = b1;
          b1 = a;
          i1 = c;
        } else if (a > b2) {
          b2 = a;
        }
      }
      p[o + t] = i1;
      pr[o + t] = (r > 1) ? b1 - b2 : b1;
      seqm_atomic_add(&u[o + i1], 1);
    }
    SEQM_SYNC();
    int bad = 0;
    for (int t = threadIdx.x; t < r; t += blockDim.x) bad |= (u[o + t] != 1);
    bad = seqm_sync_or(bad);
    if (bad) {
      if (threadIdx.x == 0) {
        for (int c = 0; c < r; ++c) u[o + c] = 0;
        for (int it = 0; it < r; ++it) {
          int row = 0;
          double best = -1.0;
          for (int t = 0; t < r; ++t)
            if (pr[o + t] > best) {
              best = pr[o + t];
              row = t;
            }
          pr[o + row] = -2.0;  // served (margins are >= 0)
          int col = 0;
          double bv = -1.0;
          for (int c = 0; c < r; ++c) {
            const double a = fabs(Sb[(long long)row * r + c]);
            if (!u[o + c] && a > bv) {
              bv = a;
              col = c;
            }
          }
          p[o + row] = col;
          u[o + col] = 1;
        }
      }
      SEQM_SYNC();
    }
  }
  SEQM_SYNC();
  const double* Cn = Vnew + (long long)v.m * N * N;
  double* Cout = Vout + (long long)v.m * N * N;
  const int no = v.nocc;
  for (int t = threadIdx.x; t < N * N; t += blockDim.x) {
    const int i = t / N, k = t - i * N;
    double x;
    if (i < n && k < n) {
      const int blk = k >= no, o = blk ? no : 0, r = blk ? n - no : no;
      const int c = p[k];  // p is indexed by absolute MO, its value is block-local
      const double sg = S[v.mat0 + (blk ? no * no : 0) + (long long)(k - o) * r + c];
      x = (sg < 0.0 ? -1.0 : 1.0) * Cn[(long long)i * N + o + c];
    } else {
      x = Cn[t];
    }
    Cout[t] = x;
  }
  for (int k = threadIdx.x; k < N; k += blockDim.x)
    e_out[(long long)v.m * N + k] = (k < n) ? e_in[(long long)v.m * N + (k >= no ? no : 0) + p[k]] : e_in[(long long)v.m * N + k];
}
SEQM_GLOBAL void initial_density_kernel(seqm_batch_t b, double* __restrict__ P) {
  const MolView v = mol_view(b, blockIdx.x);
  const int n = v.n;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
    const int i = t / n, j = t % n;
    double x = 0.0;
    if (i == j) {  // tore/4 on s and p of heavy atoms (d shells start empty), 1 on hydrogen (scf_loop.py:2066-2081)
      int k;
      const int a = orb_atom(v, i, &k);
      x = (a < v.nheavy) ? (k < 4 ? par(b, SEQM_P_TORE, v.a0 + a) / 4.0 : 0.0) : 1.0;
    }
    P[v.mat0 + t] = x;
  }
}

// FP64 FMA throughput probe: 8 independent dependent-chains per thread, 2 flops per FMA.
SEQM_GLOBAL void fp64_peak_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
    a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 123.456) out[0] = a0;
}

// ---- C ABI ---------------------------------------------------------------------------------------------
extern "C" {

int seqm_abi_version(void) { return SEQM_ABI_VERSION; }
const char* seqm_last_error(void) { return g_seqm_err; }
int seqm_max_orbitals(void) { return SEQM_MAX_ORB; }
int seqm_max_orbitals_eig(void) { return hestenes_max_orbitals(); }


long long seqm_launch_count(void) { return g_seqm_launches; }

/* eigensolver statistics since the last call: [0] molecules solved, [1] Jacobi sweeps, [2] rotation steps */
int seqm_jacobi_stats(unsigned long long* out, int reset) {
#ifndef SEQM_HOSTEMU
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, g_jacobi_stats, 8 * sizeof(unsigned long long)) != cudaSuccess) return SEQM_ERR_CUDA;
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_jacobi_stats, z, sizeof(z));
  }
#else
  for (int i = 0; i < 8; ++i) out[i] = g_jacobi_stats[i];
  if (reset) for (int i = 0; i < 8; ++i) g_jacobi_stats[i] = 0;
#endif
  return SEQM_OK;
}

/* measured FP64 FMA peak of the device in TFLOP/s (all SMs, 8 chains/thread); blocks the host */
int seqm_square_product(int n, const double* A, const double* B, double* C, void* stream) {
  if (n <= 0 || !A || !C) {
    seqm_set_error("seqm_square_product: bad argument");
    return SEQM_ERR_ARG;
  }
  int rc = ensure_device();
  if (rc) return rc;
  return B ? launch_gemm(n, A, B, C, SEQM_STREAM(stream)) : launch_gemm_sym(n, A, C, (const int*)nullptr, SEQM_STREAM(stream));
}
double seqm_fp64_peak_tflops(void) {
#ifndef SEQM_HOSTEMU
  if (ensure_device()) return -1.0;
  double* d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess) return -1.0;
  const int iters = 20000, block = 256, grid = g_num_sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, 0);
    fp64_peak_kernel<<<grid, block>>>(d, iters);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * iters * (double)block * grid / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
#else
  return 0.0;
#endif
}

int seqm_profile_enable(int on) {
  g_prof_on = on;
#ifndef SEQM_HOSTEMU
  g_ev_n = 0;
#endif
  return SEQM_OK;
}
int seqm_profile_kinds(void) { return PK_COUNT; }
const char* seqm_profile_name(int kind) { return (kind >= 0 && kind < PK_COUNT) ? g_pk_names[kind] : ""; }
/* sums the recorded intervals per kernel kind (ms) and their launch counts; synchronises the device */
int seqm_profile_collect(double* ms, int32_t* counts) {
  for (int k = 0; k < PK_COUNT; ++k) { ms[k] = 0.0; counts[k] = 0; }
#ifndef SEQM_HOSTEMU
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { seqm_set_error("profile collect: %s", cudaGetErrorString(e)); return SEQM_ERR_CUDA; }
  for (int i = 0; i < g_ev_n; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_ev0[i], g_ev1[i]) == cudaSuccess) { ms[g_ev_kind[i]] += t; counts[g_ev_kind[i]]++; }
  }
  g_ev_n = 0;
#endif
  return SEQM_OK;
}

int seqm_plan_count(const int64_t* species, int32_t nmol, int32_t molsize, const int64_t* charges, const double* elem_rows,
                    int32_t nz, int32_t* mol_atom0, int32_t* mol_pair0, int64_t* mol_mat0, int32_t* mol_nheavy,
                    int32_t* mol_nhyd, int32_t* mol_nocc, int32_t* mol_order, int32_t* mol_cls_pair0,
                    seqm_plan_counts_t* counts_dev, seqm_plan_counts_t* counts_host, int32_t* mol_nsh, void* stream) {
  int rc = ensure_device();
  if (rc) return rc;
  const int d_mode = mol_nsh ? 1 : 0;  // method="PM6": count the d-shell atoms, n = 5 nsh + 4 nheavy + nhyd
  if (nmol <= 0 || molsize <= 0 || !species || !counts_dev || !counts_host) {
    seqm_set_error("seqm_plan_count: empty batch or null pointer");
    return SEQM_ERR_ARG;
  }
  PlanBounds pb;
  for (int c = 0; c < SEQM_PLAN_NCLS; ++c) pb.np2[c] = 2 * g_jacobi_np[c];
  const size_t dyn = sizeof(int) * 2 * (size_t)((d_mode ? 9 : 4) * molsize + 2);
  cudaStream_t st = SEQM_STREAM(stream);
  SEQM_LAUNCH(plan_count_kernel, 1, SEQM_PLAN_THREADS, dyn, st, (const long long*)species, nmol, molsize, (const long long*)charges,
              elem_rows + (size_t)SEQM_P_TORE * nz, nz, pb, mol_atom0, mol_pair0, (long long*)mol_mat0, mol_nheavy, mol_nhyd,
              mol_nocc, mol_order, mol_cls_pair0, counts_dev, d_mode, mol_nsh);
  rc = seqm_check_launch("plan_count_kernel");
  if (rc) return rc;
#ifndef SEQM_HOSTEMU
  cudaError_t e = cudaMemcpyAsync(counts_host, counts_dev, sizeof(seqm_plan_counts_t), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    seqm_set_error("seqm_plan_count: %s", cudaGetErrorString(e));
    return SEQM_ERR_CUDA;
  }
#else
  *counts_host = *counts_dev;
#endif
  if (counts_host->pairs_overflow) {
    seqm_set_error("seqm_plan_count: more than 2^31-1 atom pairs in one batch");
    return SEQM_ERR_TOO_LARGE;
  }
  return SEQM_OK;
}

int seqm_plan_fill(const int64_t* species, int32_t nmol, int32_t molsize, const seqm_plan_counts_t* counts_host,
                   const int32_t* mol_atom0, const int32_t* mol_pair0, const int32_t* mol_nheavy,
                   const int32_t* mol_cls_pair0, const double* elem_rows, int32_t nrows, int32_t nz, int32_t* atom_Z,
                   int32_t* atom_mol, int64_t* real_atoms, double* atom_par, int32_t* pair_i, int32_t* pair_j,
                   int32_t* pair_perm, void* stream) {
  int rc = ensure_device();
  if (rc) return rc;
  const int off_xh = counts_host->pair_cls_cnt[0], off_xx = off_xh + counts_host->pair_cls_cnt[1];
  SEQM_LAUNCH(plan_fill_kernel, nmol, 128, 0, SEQM_STREAM(stream), (const long long*)species, nmol, molsize, counts_host->nat,
              mol_atom0, mol_pair0, mol_nheavy, mol_cls_pair0, elem_rows, nrows, nz, off_xh, off_xx, atom_Z, atom_mol,
              (long long*)real_atoms, atom_par, pair_i, pair_j, pair_perm);
  return seqm_check_launch("plan_fill_kernel");
}

int seqm_atom_multipoles(const seqm_batch_t* b, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  SEQM_LAUNCH(atom_multipoles_kernel, grid1d(b->nat, 128), 128, 0, SEQM_STREAM(stream), *b);
  return seqm_check_launch("atom_multipoles_kernel");
}

int seqm_pair_integrals(const seqm_batch_t* b, const double* xyz, double* w, double* hab, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  if (b->npairs == 0) return SEQM_OK;
  cudaStream_t st = SEQM_STREAM(stream);
  const int n0 = b->pair_cls_off[1] - b->pair_cls_off[0], n1 = b->pair_cls_off[2] - b->pair_cls_off[1],
            n2 = b->pair_cls_off[3] - b->pair_cls_off[2];
  if (n0 > 0) PROF(PK_PAIR, st, rc = pairtu_launch_integrals(b, 0, grid1d(n0, 128), 128, xyz, w, hab, st));
  if (n1 > 0 && !rc) PROF(PK_PAIR, st, rc = pairtu_launch_integrals(b, 1, grid1d(n1, 64), 64, xyz, w, hab, st));
  if (n2 > 0 && !rc) PROF(PK_PAIR, st, rc = pairtu_launch_integrals(b, 2, grid1d(n2, 64), 64, xyz, w, hab, st));
  return rc;
}

int seqm_pair_integrals_d(const seqm_batch_t* b, const double* xyz, const double* w, double* wd, double* hab_d,
                          void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  if (b->method != SEQM_PM6_D) {
    seqm_set_error("seqm_pair_integrals_d: the batch plan is not a PM6 d-orbital plan");
    return SEQM_ERR_ARG;
  }
  if (b->n_ypairs == 0) return SEQM_OK;
  cudaStream_t st = SEQM_STREAM(stream);
  PROF(PK_PAIR, st, rc = spd_launch_pair(b, xyz, w, wd, hab_d, st));
  return rc;
}

int seqm_hcore(const seqm_batch_t* b, const double* w, const double* hab, double* H, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  if (b->method == SEQM_PM6_D) {
    if ((!b->wd || !b->hab_d) && b->n_ypairs > 0) {
      seqm_set_error("PM6 with d orbitals: b->wd / b->hab_d are not set (seqm_pair_integrals_d comes first)");
      return SEQM_ERR_ARG;
    }
    PROF(PK_HCORE, SEQM_STREAM(stream), rc = spd_launch_hcore(b, w, hab, H, SEQM_STREAM(stream)));
    return rc;
  }
  const int slices = (b->nmax > SEQM_MAX_ORB) ? 32 : 1;
  PROF(PK_HCORE, SEQM_STREAM(stream), SEQM_LAUNCH(hcore_kernel, b->nmol * slices, 128, 0, SEQM_STREAM(stream), *b, w, hab, H, slices));
  return seqm_check_launch("hcore_kernel");
}

int seqm_fock(const seqm_batch_t* b, const double* P, const double* H, const double* w, double* F,
              const int32_t* active, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  return launch_fock(b, P, H, w, F, active, SEQM_STREAM(stream));
}

int seqm_eig_density(const seqm_batch_t* b, const double* F, double* P, double* evals, double* C, const double* Cguess,
                     const int32_t* active, void* stream) {
  SEQM_SERIAL;
  int rc = check_batch(b);
  if (rc) return rc;
  if (b->nmax > SEQM_MAX_ORB) {  // mid-size molecules: one-sided Jacobi in its own kernel (always a cold solve)
    PROF(PK_JACOBI, SEQM_STREAM(stream), rc = hestenes_launch(b, F, P, evals, C, active, g_smem_optin, SEQM_STREAM(stream)));
    return rc;
  }
  PROF(PK_JACOBI, SEQM_STREAM(stream), rc = launch_jacobi(b, F, P, evals, C, Cguess, active, SEQM_STREAM(stream)));
  return rc;
}

static size_t sp2_smem(int nmax);
int seqm_sp2_density(const seqm_batch_t* b, const double* F, double* P, double eps, int32_t* niter, const int32_t* active,
                     void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  rc = check_small(b, "seqm_sp2_density");
  if (rc) return rc;
  const size_t smem = sp2_smem(b->nmax);
  PROF(PK_SP2, SEQM_STREAM(stream), SEQM_LAUNCH(sp2_kernel, b->nmol, threads_for(b->nmax), smem, SEQM_STREAM(stream), *b, F, P, eps, niter, active));
  return seqm_check_launch("sp2_kernel");
}

int64_t seqm_sp2_large_workspace_bytes(const seqm_batch_t* b) {
  return (int64_t)(2 * sizeof(double) * (size_t)b->nmax * b->nmax + 4096);
}
/* SP2 for molecules of any size: X^2 by the FP64 GEMM, one molecule at a time (blocks the host per iteration) */
int seqm_sp2_density_large(const seqm_batch_t* b, const double* F, double* P, double eps, int32_t* niter_host,
                           void* workspace, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  cudaStream_t st = SEQM_STREAM(stream);
  std::vector<HostMol> hmv(b->nmol);
  HostMol* hm = hmv.data();
  rc = fetch_host_mols(b, hm, st);
  unsigned char* base = (unsigned char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  const size_t big = (sizeof(double) * (size_t)b->nmax * b->nmax + 255) & ~(size_t)255;
  double* X = (double*)base;
  double* X2 = (double*)(base + big);
  Sp2State* stt = (Sp2State*)(base + 2 * big);
  for (int m = 0; m < b->nmol && !rc; ++m) {
    int it = 0;
    rc = sp2_large_one(hm[m].n, hm[m].nocc, F + hm[m].mat0, P + hm[m].mat0, eps, X, X2, stt, &it, st);
    if (niter_host) niter_host[m] = it;
  }
  return rc;
}

int seqm_elec_energy(const seqm_batch_t* b, const double* P, const double* H, const double* F, double* Eelec,
                     const int32_t* active, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  PROF(PK_OTHER, SEQM_STREAM(stream), SEQM_LAUNCH(elec_energy_kernel, b->nmol, 128, 0, SEQM_STREAM(stream), *b, P, H, F, Eelec, active));
  return seqm_check_launch("elec_energy_kernel");
}

int seqm_nuclear_energy(const seqm_batch_t* b, const double* xyz, const double* w, double* EnucAB, double* Enuc,
                        void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  if (b->npairs > 0) {
    PROF(PK_NUC, SEQM_STREAM(stream), rc = pairtu_launch_nuclear(b, grid1d(b->npairs, 128), 128, xyz, w, EnucAB, SEQM_STREAM(stream)));
    if (rc) return rc;
  }
  SEQM_LAUNCH(pair_sum_kernel, b->nmol, 64, 0, SEQM_STREAM(stream), *b, EnucAB, Enuc);
  return seqm_check_launch("pair_sum_kernel");
}

int seqm_gradient(const seqm_batch_t* b, const double* xyz, const double* P, double* pair_scratch, double* grad,
                  void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  if (b->npairs > 0) {
    rc = launch_pair_gradient(b, xyz, P, P, pair_scratch, SEQM_STREAM(stream));
    if (rc) return rc;
    if (b->method == SEQM_PM6_D && b->n_ypairs > 0) {
      cudaStream_t st = SEQM_STREAM(stream);
      PROF(PK_GRAD, st, rc = spd_launch_gradient(b, xyz, P, pair_scratch, st));
      if (rc) return rc;
    }
  }
  PROF(PK_GRAD, SEQM_STREAM(stream), SEQM_LAUNCH(atom_gradient_kernel, grid1d(b->nat, 128), 128, 0, SEQM_STREAM(stream), *b, pair_scratch, grad));
  return seqm_check_launch("atom_gradient_kernel");
}

int seqm_pack(const seqm_batch_t* b, const double* dense, double* packed, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  SEQM_LAUNCH(pack_kernel, b->nmol, 256, 0, SEQM_STREAM(stream), *b, dense, packed);
  return seqm_check_launch("pack_kernel");
}
int seqm_unpack(const seqm_batch_t* b, const double* packed, double* dense, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  SEQM_LAUNCH(unpack_kernel, b->nmol, 256, 0, SEQM_STREAM(stream), *b, packed, dense);
  return seqm_check_launch("unpack_kernel");
}
int seqm_orbitals_dense(const seqm_batch_t* b, const double* C, double* V, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  SEQM_LAUNCH(orbitals_dense_kernel, b->nmol, 256, 0, SEQM_STREAM(stream), *b, C, V);
  return seqm_check_launch("orbitals_dense_kernel");
}
int seqm_mo_match(const seqm_batch_t* b, const double* V_new, const double* V_old, const double* e_in, double* S_scratch,
                  int32_t* perm, int32_t* used, double* prio, double* V_out, double* e_out, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  if (!V_new || !V_old || !e_in || !S_scratch || !perm || !used || !prio || !V_out || !e_out || V_out == V_new) {
    seqm_set_error("seqm_mo_match: null buffer or V_out aliases V_new");
    return SEQM_ERR_ARG;
  }
  long long per = (long long)b->nmax * b->nmax;
  int parts = (int)((per + 8191) / 8192);  // ~32 overlap elements per thread
  if (parts < 1) parts = 1;
  if (parts > 1024) parts = 1024;
  SEQM_LAUNCH(mo_overlap_kernel, b->nmol * parts, 256, 0, SEQM_STREAM(stream), *b, parts, V_old, V_new, S_scratch);
  rc = seqm_check_launch("mo_overlap_kernel");
  if (rc) return rc;
  SEQM_LAUNCH(mo_match_kernel, b->nmol, 256, 0, SEQM_STREAM(stream), *b, V_new, (const double*)S_scratch, e_in, perm, used,
              prio, V_out, e_out);
  return seqm_check_launch("mo_match_kernel");
}
int seqm_gradient_xl(const seqm_batch_t* b, const double* xyz, const double* D, const double* P, double* pair_scratch,
                     double* grad, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  if (b->npairs > 0) {
    rc = launch_pair_gradient(b, xyz, D, P, pair_scratch, SEQM_STREAM(stream));
    if (rc) return rc;
  }
  PROF(PK_GRAD, SEQM_STREAM(stream), SEQM_LAUNCH(atom_gradient_kernel, grid1d(b->nat, 128), 128, 0, SEQM_STREAM(stream), *b, pair_scratch, grad));
  return seqm_check_launch("atom_gradient_kernel");
}
int seqm_xl_propagate(int64_t total, double kappa, double c, const double* D, const double* P_in, double* Pt,
                       const double* coef, int32_t m, int32_t slot, double* P_out, void* stream) {
  int rc = ensure_device();
  if (rc) return rc;
  if (m < 1 || m > SEQM_XL_MAXHIST || slot < 0 || slot >= m || total <= 0) {
    seqm_set_error("seqm_xl_propagate: bad history depth %d / slot %d", m, slot);
    return SEQM_ERR_ARG;
  }
  PROF(PK_OTHER, SEQM_STREAM(stream), SEQM_LAUNCH(xl_propagate_kernel, grid1d(total, 256), 256, 0, SEQM_STREAM(stream), (long long)total,
                                                  kappa, c, D, P_in, Pt, coef, m, slot, P_out));
  return seqm_check_launch("xl_propagate_kernel");
}
int seqm_elec_energy_xl(const seqm_batch_t* b, const double* D, const double* P, const double* F, const double* H,
                        double* Eelec, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  PROF(PK_OTHER, SEQM_STREAM(stream), SEQM_LAUNCH(elec_energy_xl_kernel, b->nmol, 128, 0, SEQM_STREAM(stream), *b, D, P, F, H, Eelec));
  return seqm_check_launch("elec_energy_xl_kernel");
}
int seqm_gradient_forward(const seqm_batch_t* b, const double* xyz, const double* P, double* pair_scratch, double* grad,
                          void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  if (b->npairs > 0) {
    rc = pairtu_launch_gradient_forward(b, grid1d(b->npairs, 64), 64, xyz, P, pair_scratch, SEQM_STREAM(stream));
    if (rc) return rc;
  }
  SEQM_LAUNCH(atom_gradient_kernel, grid1d(b->nat, 128), 128, 0, SEQM_STREAM(stream), *b, pair_scratch, grad);
  return seqm_check_launch("atom_gradient_kernel");
}
int seqm_initial_density(const seqm_batch_t* b, double* P, void* stream) {
  int rc = check_batch(b);
  if (rc) return rc;
  SEQM_LAUNCH(initial_density_kernel, b->nmol, 128, 0, SEQM_STREAM(stream), *b, P);
  return seqm_check_launch("initial_density_kernel");
}

int64_t seqm_scf_workspace_bytes(const seqm_batch_t* b, const seqm_scf_opts_t* o) {
  if (!b || !o) return -1;
  return (int64_t)scf_carve(b, o, nullptr, nullptr) + 256;
}

// Read the control block back; blocks until the stream has drained.
static int read_ctrl(const ScfWork& W, void* stream, ScfCtrl* h) {
#ifndef SEQM_HOSTEMU
  cudaError_t e = cudaMemcpyAsync(h, W.ctrl, sizeof(ScfCtrl), cudaMemcpyDeviceToHost, SEQM_STREAM(stream));
  if (e == cudaSuccess) e = cudaStreamSynchronize(SEQM_STREAM(stream));
  if (e != cudaSuccess) {
    seqm_set_error("SCF control read-back: %s", cudaGetErrorString(e));
    return SEQM_ERR_CUDA;
  }
#else
  (void)stream;
  *h = *W.ctrl;
#endif
  return SEQM_OK;
}
static int zero_nnot(const ScfWork& W, void* stream) {
#ifndef SEQM_HOSTEMU
  cudaError_t e = cudaMemsetAsync(&W.ctrl->nnot, 0, sizeof(int), SEQM_STREAM(stream));
  if (e != cudaSuccess) {
    seqm_set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
    return SEQM_ERR_CUDA;
  }
#else
  (void)stream;
  W.ctrl->nnot = 0;
#endif
  return SEQM_OK;
}

// ---- Pulay DIIS, pipelined (converger 2, shared-memory-resident molecules) ---------------------------------------
// The batch is split into two half-batches with the same size mix.  Each half runs the iteration
//   begin -> store -> solve -> extrapolate -> eigensolver -> mix -> Fock -> get_error
// on its own high-priority stream, half an iteration out of phase with the other, so the HBM-bound kernels of one
// half run under the FP64-bound eigensolver of the other.  Everything that decides iteration counts stays
// batch-global exactly as in the reference: the DIIS ring state lives on the device (diis_begin_kernel), a reset
// raised by either half in iteration k empties both histories before iteration k+1, and the loop ends when no
// molecule of either half is left.  The host never waits for the iteration it has just enqueued: it reads the
// not-converged counters of iteration k-1 while iteration k runs (one trailing iteration is therefore enqueued
// with every molecule inactive; its kernels exit immediately).
#ifndef SEQM_HOSTEMU
static cudaStream_t g_half_stream[2];
static cudaEvent_t g_ev_fork, g_ev_join[2], g_ev_solve[2][2], g_ev_enter[2], g_ev_done[2][2];
static int* g_h_nnot = nullptr;  // pinned: [half][iteration parity]
static int g_pipe_ready = 0;
static int ensure_pipeline() {
  if (g_pipe_ready) return SEQM_OK;
  int least = 0, greatest = 0;
  cudaDeviceGetStreamPriorityRange(&least, &greatest);
  if (getenv("SEQM_PIPE_NOPRIO")) greatest = least;
  bool ok = cudaHostAlloc((void**)&g_h_nnot, 4 * sizeof(int), cudaHostAllocDefault) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int h = 0; h < 2 && ok; ++h) {
    ok = ok && cudaStreamCreateWithPriority(&g_half_stream[h], cudaStreamNonBlocking, greatest) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&g_ev_join[h], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&g_ev_enter[h], cudaEventDisableTiming) == cudaSuccess;
    for (int q = 0; q < 2 && ok; ++q) {
      ok = ok && cudaEventCreateWithFlags(&g_ev_solve[h][q], cudaEventDisableTiming) == cudaSuccess;
      ok = ok && cudaEventCreateWithFlags(&g_ev_done[h][q], cudaEventDisableTiming) == cudaSuccess;
    }
  }
  if (!ok) {
    seqm_set_error("could not create the SCF pipeline streams/events");
    return SEQM_ERR_CUDA;
  }
  g_pipe_ready = 1;
  return SEQM_OK;
}
#else
static int g_h_nnot_emu[4];
static int* g_h_nnot = g_h_nnot_emu;
#endif

// in-SM SP2: one zero-padded matrix (row stride 4 mod 16) + 40 scratch doubles; the host emulation keeps X and X^2
static size_t sp2_smem(int nmax) {
  const size_t np8 = ((size_t)nmax + 7) & ~(size_t)7;
#ifndef SEQM_HOSTEMU
  return sizeof(double) * (np8 * (np8 + 4) + 40);
#else
  return sizeof(double) * ((size_t)2 * nmax * nmax + np8 * 4 + 40);
#endif
}
// F and P of one molecule, zero-padded to a multiple of 8 with a row stride of 4 mod 16 (tensor-core fragments)
static size_t diis_store_smem(int nmax) {
  const size_t np8 = ((size_t)nmax + 7) & ~(size_t)7;
  return sizeof(double) * 2 * np8 * ((np8 > 112) ? np8 : np8 + 4);
}

static int scf_diis_pipelined(const seqm_batch_t* b, const seqm_scf_opts_t* o, const ScfWork& W0, const double* H,
                              const double* w, double* P, double* F, int32_t* notconverged, cudaStream_t st,
                              int* n_iter, bool have_guess) {
  int rc = SEQM_OK;
  const int max_iter = o->max_iter > 0 ? o->max_iter : 1000;
  int nh = (o->pipeline == 1) ? 1 : ((o->pipeline == 2 || b->nmol >= 256) ? 2 : 1);
  if (b->nmol < 2) nh = 1;
#ifdef SEQM_HOSTEMU
  cudaStream_t hs[2] = {st, st};
#else
  cudaStream_t hs[2] = {st, st};
  rc = ensure_pipeline();
  if (rc) return rc;
  if (nh == 2) {
    rc = ensure_class_streams();
    if (rc) return rc;
    hs[0] = g_half_stream[0];
    hs[1] = g_half_stream[1];
    cudaEventRecord(g_ev_fork, st);
    cudaStreamWaitEvent(hs[0], g_ev_fork, 0);
    cudaStreamWaitEvent(hs[1], g_ev_fork, 0);
  }
  // every exit path (errors included) re-joins the caller's stream behind whatever is queued on the half streams
  struct PipeJoin {
    cudaStream_t st, *hs;
    bool on;
    ~PipeJoin() {
      if (!on) return;
      for (int h = 0; h < 2; ++h) {
        cudaEventRecord(g_ev_join[h], hs[h]);
        cudaStreamWaitEvent(st, g_ev_join[h], 0);
      }
    }
  } pipe_join{st, hs, nh == 2};
#endif
  // half-batch views: same arrays, own processing order / size-class ranges / control block
  seqm_batch_t bh[2] = {*b, *b};
  ScfWork Wh[2] = {W0, W0};
  if (nh == 2) {
    const int nA = (b->nmol + 1) / 2;
    bh[0].nmol = nA;
    bh[0].mol_order = W0.order2;
    bh[1].nmol = b->nmol - nA;
    bh[1].mol_order = W0.order2 + nA;
    Wh[1].ctrl = W0.ctrl + 1;
    for (int c = 0; c < 12; ++c) {
      const int b0 = b->cls_begin[c], e0 = b0 + b->cls_count[c];
      const int pe = b0 + (b0 & 1), po = b0 + ((b0 & 1) ? 0 : 1);  // first even / odd position of the class
      bh[0].cls_begin[c] = pe / 2;
      bh[0].cls_count[c] = (pe < e0) ? (e0 - pe + 1) / 2 : 0;
      bh[1].cls_begin[c] = (po - 1) / 2;
      bh[1].cls_count[c] = (po < e0) ? (e0 - po + 1) / 2 : 0;
    }
  }
#define CHKP(name)              \
  rc = seqm_check_launch(name); \
  if (rc) return rc
  int k = 0, done_at = -1;
  for (; k <= max_iter; ++k) {
    for (int h = 0; h < nh; ++h) {
      cudaStream_t s = hs[h];
      const seqm_batch_t& B = bh[h];
      const ScfWork& W = Wh[h];
      const int nt = threads_for(b->nmax);
      const size_t sm1 = sizeof(double) * (size_t)b->nmax * b->nmax;
      const int gm = grid1d(B.nmol, 128);
#ifndef SEQM_HOSTEMU
      if (nh == 2) {
        static const int nostagger = getenv("SEQM_PIPE_NOSTAGGER") ? atoi(getenv("SEQM_PIPE_NOSTAGGER")) : 0;
        if (nostagger == 0 ? (k > 0 || h == 1) : (nostagger == 1 ? (k == 0 && h == 1) : false))
          cudaStreamWaitEvent(s, g_ev_enter[1 - h], 0);  // half an iteration out of phase
        if (k > 0) cudaStreamWaitEvent(s, g_ev_solve[1 - h][(k - 1) & 1], 0);  // the other half's reset flag of k-1
      }
#endif
      SEQM_LAUNCH(diis_begin_kernel, gm, 128, 0, s, B, W, k, h);
      CHKP("diis_begin_kernel");
      PROF(PK_DIIS_STORE, s, SEQM_LAUNCH(diis_store_kernel, B.nmol, nt, diis_store_smem(b->nmax), s, B, W, (const double*)F, (const double*)P, -1, -1));
      CHKP("diis_store_kernel");
      PROF(PK_DIIS_SOLVE, s, SEQM_LAUNCH(diis_solve_kernel, diis_grid(B.nmol), 32 * SEQM_DIIS_WARPS, 0, s, B, W, -1, -1,
                                         W0.rflag + 2 * h + (k & 1)));
      CHKP("diis_solve_kernel");
#ifndef SEQM_HOSTEMU
      if (nh == 2) cudaEventRecord(g_ev_solve[h][k & 1], s);
#endif
      PROF(PK_DIIS_EXTRAP, s, SEQM_LAUNCH(diis_extrapolate_kernel, B.nmol, 256, 0, s, B, W, F, -1));
      CHKP("diis_extrapolate_kernel");
#ifndef SEQM_HOSTEMU
      if (nh == 2) cudaEventRecord(g_ev_enter[h], s);
#endif
      if (o->use_sp2) {
        const size_t smsp2 = sp2_smem(b->nmax);
        PROF(PK_SP2, s, SEQM_LAUNCH(sp2_kernel, B.nmol, nt, smsp2, s, B, (const double*)F, W.Pnew, o->sp2_eps, (int32_t*)nullptr, (const int32_t*)W.active));
        CHKP("sp2_kernel");
      } else {
        // the DIIS mixing (Pold <- P, P <- mix(P, Pnew)) is the tail of the density kernel
        PROF(PK_JACOBI, s, rc = launch_jacobi(&B, F, W.Pnew, (double*)nullptr, W.C,
                                              (o->warm_start && (k > 0 || have_guess)) ? (const double*)W.C : (const double*)nullptr, W.active, s, h,
                                              JacobiMix{P, W.Pold, &W.ctrl->cF}));
        if (rc) return rc;
      }
      if (o->use_sp2) {
        PROF(PK_MIX, s, SEQM_LAUNCH(mix_linear_kernel, B.nmol, 256, 0, s, B, W, P, -1.0));
        CHKP("mix_linear_kernel");
      }
      // elec_energy + get_error + active-mask update ride on the Fock kernel when the pair-centric one runs
      FockErr fe;
      fe.on = 1;
      fe.use_diis = 1;
      fe.eps = o->eps;
      fe.Pold = W.Pold;
      fe.diis_err = W.diis_err;
      fe.Eel_run = W.Eel_run;
      fe.Eel_new = W.Eel_new;
      fe.err = W.err;
      fe.dm_err = W.dm_err;
      fe.dm_elem = W.dm_elem;
      fe.notconv = notconverged;
      fe.active_out = W.active;
      fe.nnot = &W.ctrl->nnot;
      bool fused = false;
      rc = launch_fock(&B, P, H, w, F, (const int32_t*)W.active, s, &fe, &fused);
      if (rc) return rc;
      if (!fused) {
        PROF(PK_ENERGY_ERR, s, SEQM_LAUNCH(energy_error_kernel, B.nmol, 128, 0, s, B, W, (const double*)P, H, (const double*)F, notconverged, o->eps, 1));
        CHKP("energy_error_kernel");
        SEQM_LAUNCH(commit_active_kernel, gm, 128, 0, s, B, W, (const int32_t*)notconverged);
        CHKP("commit_active_kernel");
      }
#ifndef SEQM_HOSTEMU
      cudaMemcpyAsync(g_h_nnot + 2 * h + (k & 1), &W.ctrl->nnot, sizeof(int), cudaMemcpyDeviceToHost, s);
      cudaEventRecord(g_ev_done[h][k & 1], s);
#else
      g_h_nnot[2 * h + (k & 1)] = W.ctrl->nnot;
#endif
    }
    // look one iteration behind: the GPU is busy with iteration k while the host learns about k-1
    const int kk = k - 1;
    if (kk >= 0) {
      int nnot = 0;
      for (int h = 0; h < nh; ++h) {
#ifndef SEQM_HOSTEMU
        cudaError_t e = cudaEventSynchronize(g_ev_done[h][kk & 1]);
        if (e != cudaSuccess) {
          seqm_set_error("SCF pipeline: %s", cudaGetErrorString(e));
          return SEQM_ERR_CUDA;
        }
#endif
        nnot += g_h_nnot[2 * h + (kk & 1)];
      }
      if (nnot == 0) {
        done_at = kk + 1;
        break;
      }
    }
  }
#undef CHKP
  *n_iter = (done_at >= 0) ? done_at : max_iter + 1;
  return SEQM_OK;
}

int seqm_scf(const seqm_batch_t* b, const seqm_scf_opts_t* o, const double* H, const double* w, double* P, double* F,
             double* Eelec, int32_t* notconverged, void* workspace, int32_t* n_iter_out, double* C_last, void* stream) {
  SEQM_SERIAL;
  int rc = check_batch(b);
  if (rc) return rc;
  if (o->converger < 0 || o->converger > 2) {
    seqm_set_error("scf_converger %d not supported (0: constant mixing, 1: adaptive, 2: Pulay DIIS)", o->converger);
    return SEQM_ERR_UNSUPPORTED;
  }
  ScfWork W;
  unsigned char* base = (unsigned char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  scf_carve(b, o, base, &W);
  cudaStream_t st = SEQM_STREAM(stream);
  const int nt = threads_for(b->nmax);
  const size_t sm1 = sizeof(double) * (size_t)b->nmax * b->nmax;
  const size_t smsp2 = sp2_smem(b->nmax);
  const int gm = grid1d(b->nmol, 128);
  const int max_iter = o->max_iter > 0 ? o->max_iter : 1000;
#define CHK(name)                     \
  rc = seqm_check_launch(name);       \
  if (rc) return rc
  SEQM_LAUNCH(scf_init_kernel, gm, 128, 0, st, *b, W, o->converger);
  CHK("scf_init_kernel");
  // F(P0), Eelec(P0)
  const bool large = b->nmax > SEQM_MAX_ORB;
  std::vector<HostMol> hmv;
  std::vector<int32_t> h_activev;
  HostMol* hm = nullptr;
  int32_t* h_active = nullptr;
  if (large) {
    if (!o->use_sp2 && b->nmax > hestenes_max_orbitals()) {
      seqm_set_error("molecules above %d orbitals need the SP2 density (sp2=[True, eps]): the eigensolver route ends at %d "
                     "orbitals (one-sided Jacobi)", hestenes_max_orbitals(), hestenes_max_orbitals());
      return SEQM_ERR_TOO_LARGE;
    }
    hmv.resize(b->nmol);
    h_activev.resize(b->nmol);
    hm = hmv.data();
    h_active = h_activev.data();
    rc = fetch_host_mols(b, hm, st);
    if (rc) return rc;
    for (int m = 0; m < b->nmol; ++m) h_active[m] = 1;
  }
  rc = launch_fock(b, P, H, w, F, (const int32_t*)nullptr, st);
  if (rc) return rc;
  PROF(PK_OTHER, SEQM_STREAM(stream), SEQM_LAUNCH(elec_energy_kernel, b->nmol, 128, 0, st, *b, P, H, F, W.Eel_run, (const int32_t*)nullptr));
  CHK("elec_energy_kernel");
  int have_C = 0;
  int counter = -1, cF = 0;
  int nnot = b->nmol;
  int printed = 0;
  // warm_start == 2: C_last holds, on entry, eigenvectors to start the FIRST density solve from (a restart from a
  // nearby geometry: MD steps, geometry optimisation); they must belong to the same batch layout
  if (o->warm_start == 2 && C_last && !o->use_sp2 && !large) {
#ifndef SEQM_HOSTEMU
    if (cudaMemcpyAsync(W.C, C_last, sizeof(double) * (size_t)b->mat_total, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      seqm_set_error("seqm_scf: initial eigenvector copy failed");
      return SEQM_ERR_CUDA;
    }
#else
    memcpy(W.C, C_last, sizeof(double) * (size_t)b->mat_total);
#endif
    have_C = 1;
  }
  const bool pipelined = (o->converger == 2) && !large;
  if (pipelined) {
    rc = scf_diis_pipelined(b, o, W, H, w, P, F, notconverged, st, &printed, have_C != 0);
    if (rc) return rc;
    have_C = o->use_sp2 ? 0 : 1;
  }
  // iteration index conventions of the reference: converger 0 counts from 0, converger 1 from 1,
  // converger 2 reports the number of completed iterations.
  const int k_first = (o->converger == 1) ? 1 : 0;
  const int k_last = max_iter;
  int k = k_first;
  for (; !pipelined && k <= k_last; ++k) {
    if (o->converger == 2) {
      if (nnot == 0) break;
      cF = (cF < SEQM_NFOCK) ? cF + 1 : SEQM_NFOCK;
      counter = (counter + 1) % SEQM_NFOCK;
      if (!large) {
        PROF(PK_DIIS_STORE, SEQM_STREAM(stream), SEQM_LAUNCH(diis_store_kernel, b->nmol, nt, diis_store_smem(b->nmax), st, *b, W, F, P, counter, cF));
        CHK("diis_store_kernel");
      } else {
        for (int m = 0; m < b->nmol; ++m) {
          if (!h_active[m]) continue;
          const int n = hm[m].n;
          const long long nn = (long long)n * n, h0 = hm[m].mat0 * SEQM_NFOCK;
          double* Fh = W.FOCK + h0 + (long long)counter * nn;
          double* Rh = W.RES + h0 + (long long)counter * nn;
#ifndef SEQM_HOSTEMU
          cudaMemcpyAsync(Fh, F + hm[m].mat0, sizeof(double) * nn, cudaMemcpyDeviceToDevice, st);
          cudaMemsetAsync(W.rmax, 0, sizeof(double), st);
#else
          memcpy(Fh, F + hm[m].mat0, sizeof(double) * nn);
          *W.rmax = 0.0;
#endif
          rc = launch_gemm(n, F + hm[m].mat0, P + hm[m].mat0, W.Xl, st);  // G = F P ; R = G - G^t
          if (rc) return rc;
          PROF(PK_DIIS_STORE, st, SEQM_LAUNCH(commutator_kernel, grid1d(nn, 256), 256, 0, st, n, (const double*)W.Xl, Rh, W.rmax));
          PROF(PK_DIIS_STORE, st, SEQM_LAUNCH(residual_dots_kernel, cF * SEQM_DOT_PARTS, 256, 0, st, n, (const double*)Rh,
                                              (const double*)(W.RES + h0), nn, W.part));
          PROF(PK_DIIS_STORE, st, SEQM_LAUNCH(residual_dots_finish_kernel, 1, 32, 0, st, (const double*)W.part, cF,
                                              W.EMAT + (long long)m * SEQM_EM * SEQM_EM + counter * SEQM_EM));
#ifndef SEQM_HOSTEMU
          cudaMemcpyAsync(W.diis_err + m, W.rmax, sizeof(double), cudaMemcpyDeviceToDevice, st);
#else
          W.diis_err[m] = *W.rmax;
#endif
          CHK("large diis_store");
        }
      }
      if (cF >= 2) {
        PROF(PK_DIIS_SOLVE, SEQM_STREAM(stream), SEQM_LAUNCH(diis_solve_kernel, diis_grid(b->nmol), 32 * SEQM_DIIS_WARPS, 0, st, *b, W, counter, cF, &W.ctrl->reset));
        CHK("diis_solve_kernel");
        if (!large) {
          PROF(PK_DIIS_EXTRAP, SEQM_STREAM(stream), SEQM_LAUNCH(diis_extrapolate_kernel, b->nmol, 256, 0, st, *b, W, F, cF));
        } else {
          for (int m = 0; m < b->nmol; ++m) {
            if (!h_active[m]) continue;
            const long long nn = (long long)hm[m].n * hm[m].n;
            PROF(PK_DIIS_EXTRAP, st, SEQM_LAUNCH(extrapolate_large_kernel, grid1d(nn, 256), 256, 0, st, nn,
                                                 (const double*)(W.coeff + (long long)m * SEQM_NFOCK),
                                                 (const double*)(W.FOCK + hm[m].mat0 * SEQM_NFOCK), F + hm[m].mat0, cF));
          }
        }
        CHK("diis_extrapolate_kernel");
      }
    }
    // Pnew from F on the active molecules
    if (large && !o->use_sp2) {  // eigensolver route of mid-size molecules (<= 256 orbitals): all active molecules, one launch
      PROF(PK_JACOBI, st, rc = hestenes_launch(b, F, W.Pnew, (double*)nullptr, W.C, W.active, g_smem_optin, st));
      if (rc) return rc;
      have_C = 1;
    } else if (large) {
      for (int m = 0; m < b->nmol; ++m) {
        if (!h_active[m]) continue;
        rc = sp2_large_one(hm[m].n, hm[m].nocc, F + hm[m].mat0, W.Pnew + hm[m].mat0, o->sp2_eps, W.Xl, W.X2l, W.sp2st,
                           (int*)nullptr, st);
        if (rc) return rc;
      }
    } else if (o->use_sp2) {
      PROF(PK_SP2, SEQM_STREAM(stream), SEQM_LAUNCH(sp2_kernel, b->nmol, nt, smsp2, st, *b, F, W.Pnew, o->sp2_eps, (int32_t*)nullptr, W.active));
      CHK("sp2_kernel");
    } else {
      PROF(PK_JACOBI, SEQM_STREAM(stream), rc = launch_jacobi(b, F, W.Pnew, (double*)nullptr, W.C,
                (o->warm_start && have_C) ? (const double*)W.C : (const double*)nullptr, W.active, st));
      if (rc) return rc;
      have_C = 1;
    }
    // mixing
    if (o->converger == 0) {
      PROF(PK_MIX, SEQM_STREAM(stream), SEQM_LAUNCH(mix_linear_kernel, b->nmol, 256, 0, st, *b, W, P, o->alpha));
      CHK("mix_linear_kernel");
    } else if (o->converger == 1) {
      PROF(PK_MIX, SEQM_STREAM(stream), SEQM_LAUNCH(adaptive_diag_a_kernel, grid1d(b->nmol, 64), 64, 0, st, *b, W, (const double*)P, k));
      PROF(PK_MIX, SEQM_STREAM(stream), SEQM_LAUNCH(adaptive_diag_b_kernel, grid1d(b->nmol, 64), 64, 0, st, *b, W, k));
      SEQM_LAUNCH(adaptive_diag_reset_kernel, 1, 1, 0, st, W, k);
      CHK("adaptive_diag kernels");
      PROF(PK_MIX, SEQM_STREAM(stream), SEQM_LAUNCH(adaptive_apply_kernel, b->nmol, 256, 0, st, *b, W, P));
      CHK("adaptive_apply_kernel");
    } else if (!large) {
      PROF(PK_MIX, SEQM_STREAM(stream), SEQM_LAUNCH(mix_linear_kernel, b->nmol, 256, 0, st, *b, W, P, (cF < 2) ? 0.5 : 0.0));
      CHK("mix_linear_kernel");
    } else {
      for (int m = 0; m < b->nmol; ++m) {
        if (!h_active[m]) continue;
        const long long nn = (long long)hm[m].n * hm[m].n;
        PROF(PK_MIX, st, SEQM_LAUNCH(mix_large_kernel, grid1d(nn, 256), 256, 0, st, nn, P + hm[m].mat0, W.Pold + hm[m].mat0,
                                     (const double*)(W.Pnew + hm[m].mat0), (cF < 2) ? 0.5 : 0.0));
      }
      CHK("mix_large_kernel");
    }
    rc = launch_fock(b, P, H, w, F, (const int32_t*)W.active, st);
    if (rc) return rc;
    rc = zero_nnot(W, stream);
    if (rc) return rc;
    if (!large) {
      PROF(PK_ENERGY_ERR, SEQM_STREAM(stream), SEQM_LAUNCH(energy_error_kernel, b->nmol, 128, 0, st, *b, W, P, H, F, notconverged, o->eps, o->converger == 2));
    } else {
      for (int m = 0; m < b->nmol; ++m) {
        if (!h_active[m]) continue;
        const long long nn = (long long)hm[m].n * hm[m].n, m0 = hm[m].mat0;
        int np_ = grid1d(nn, 256);
        if (np_ > 4096) np_ = 4096;
        PROF(PK_ENERGY_ERR, st, SEQM_LAUNCH(energy_partial_kernel, np_, 256, 0, st, nn, (const double*)(P + m0), H + m0,
                                            (const double*)(F + m0), (const double*)(W.Pold + m0), W.part));
        PROF(PK_ENERGY_ERR, st, SEQM_LAUNCH(energy_finalize_kernel, 1, 32, 0, st, *b, W, m, (const double*)W.part, np_,
                                            notconverged, o->eps, o->converger == 2));
      }
    }
    CHK("energy_error_kernel");
    SEQM_LAUNCH(commit_active_kernel, gm, 128, 0, st, *b, W, notconverged);
    CHK("commit_active_kernel");
    ScfCtrl h;
    rc = read_ctrl(W, stream, &h);
    if (rc) return rc;
    nnot = h.nnot;
    if (large) {
      rc = fetch_host_ints(W.active, h_active, b->nmol, st);
      if (rc) return rc;
    }
    if (o->converger == 2) {
      if (h.reset) {
        counter = -1;
        cF = 0;
        SEQM_LAUNCH(emat_reset_kernel, gm, 128, 0, st, *b, W);
        CHK("emat_reset_kernel");
      }
    } else if (nnot == 0) {
      printed = k;
      break;
    }
    printed = k;
  }
  if (o->converger == 2 && !pipelined) printed = (k > k_last) ? k_last + 1 : k;
  if (n_iter_out) *n_iter_out = printed;
#ifndef SEQM_HOSTEMU
  cudaError_t e = cudaMemcpyAsync(Eelec, W.Eel_new, sizeof(double) * b->nmol, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && C_last && have_C)
    e = cudaMemcpyAsync(C_last, W.C, sizeof(double) * (size_t)b->mat_total, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) {
    seqm_set_error("result copy: %s", cudaGetErrorString(e));
    return SEQM_ERR_CUDA;
  }
#else
  memcpy(Eelec, W.Eel_new, sizeof(double) * b->nmol);
  if (C_last && have_C) memcpy(C_last, W.C, sizeof(double) * (size_t)b->mat_total);
#endif
#undef CHK
  return SEQM_OK;
}

}  // extern "C"
