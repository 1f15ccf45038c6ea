/* seqm_b200.h -- C ABI of libseqm_b200.so: the B200-native (sm_100a) kernels behind PYSEQM's batched
 * ground-state SCF path (seqm.Molecule / Electronic_Structure(seqm_parameters).forward).
 *
 * The reference (lanl/PYSEQM v2.0.0) is pure Python/PyTorch and has no FFI; the boundary this library
 * replaces is the operator level of seqm/seqm_functions (SURVEY.md 8(b) "level B").  Each entry point
 * names the reference function it stands in for.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - extern "C"; every function returns 0 (SEQM_OK) or a negative error code, never throws, never
 *    allocates device memory, and enqueues its work on the cudaStream_t passed as `stream`
 *    (a void* here so that the header needs no CUDA include).  seqm_last_error() gives the message.
 *  - all pointers inside seqm_batch_t and all array arguments are DEVICE pointers owned by the caller
 *    (torch allocations in pyseqm_b200); fp64 data, int32 indices, int64 matrix offsets.
 *  - "packed matrix" = per-molecule n x n row-major fp64 block at offset mol_mat0[m] of one flat buffer,
 *    n = 4*nheavy + nhyd, orbital order [4 AOs (s,px,py,pz) per heavy atom][one s AO per hydrogen]
 *    (the reference's pack() layout, seqm/seqm_functions/pack.py:8-16, without padding to the batch max).
 *  - pairs are all i<j atom pairs inside each molecule, ordered (molecule, i, j)  (seqm/basics.py:306-343).
 *  - w is the dense (npairs,10,10) tensor of the reference (kl on atom i | mn on atom j), packed pair index
 *    0:(ss) 1:(x s) 2:(x x) 3:(y s) 4:(y x) 5:(y y) 6:(z s) 7:(z x) 8:(z y) 9:(z z).
 */
#ifndef SEQM_B200_H
#define SEQM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEQM_ABI_VERSION 3

/* rows of the per-atom parameter table atom_par[row * nat + atom] */
enum seqm_par_row {
  SEQM_P_USS = 0, SEQM_P_UPP, SEQM_P_ZS, SEQM_P_ZP, SEQM_P_BS, SEQM_P_BP, SEQM_P_GSS, SEQM_P_GSP, SEQM_P_GPP,
  SEQM_P_GP2, SEQM_P_HSP, SEQM_P_ALPHA,
  SEQM_P_K1, SEQM_P_K2, SEQM_P_K3, SEQM_P_K4, SEQM_P_L1, SEQM_P_L2, SEQM_P_L3, SEQM_P_L4,
  SEQM_P_M1, SEQM_P_M2, SEQM_P_M3, SEQM_P_M4,
  SEQM_P_TORE, SEQM_P_QN, SEQM_P_RHOCORE, SEQM_P_ATNUM,
  /* filled by seqm_atom_multipoles(): */
  SEQM_P_DD, SEQM_P_QQ, SEQM_P_RHO0, SEQM_P_RHO1, SEQM_P_RHO2,
  /* PM6 d-shell rows, all from the per-element table (zero for sp-only elements).  UDD, ZD, BD: U_dd, zeta_d, beta_d of
   * the parameter file; QND: principal quantum number of the d shell (constants.py qnD_int); DP, DS, DDQ: charge
   * separations of the p-d dipole, s-d and d-d quadrupoles; RHO3..RHO6: additive terms of the d-d monopole, p-d
   * dipole, s-d and d-d quadrupoles; RHO2D: the p-p quadrupole term used wherever a d orbital takes part
   * (two_elec_two_center_int.py:31-97, 204-247; host-side preparation in pyseqm_b200/seqm_functions/pm6d_tables.py) */
  SEQM_P_UDD, SEQM_P_ZD, SEQM_P_BD, SEQM_P_QND, SEQM_P_DP, SEQM_P_DS, SEQM_P_DDQ, SEQM_P_RHO3, SEQM_P_RHO4, SEQM_P_RHO5,
  SEQM_P_RHO6, SEQM_P_RHO2D,
  SEQM_NPAR
};

/* SEQM_PM6_D: method="PM6" on a batch that contains d-shell elements (9 orbitals on those atoms) */
enum seqm_method { SEQM_MNDO = 0, SEQM_AM1 = 1, SEQM_PM3 = 2, SEQM_PM6_SP = 3, SEQM_PM6_D = 4 };

typedef struct seqm_batch {
  int32_t nmol, nat, npairs, method;
  int32_t nmax;        /* largest orbital count n in the batch */
  int32_t molsize;     /* padded atoms per molecule at the API boundary */
  int64_t mat_total;   /* doubles in one packed-matrix buffer = mol_mat0[nmol] */
  const int32_t* mol_atom0; /* [nmol+1] first real atom of each molecule */
  const int32_t* mol_pair0; /* [nmol+1] first pair of each molecule */
  const int64_t* mol_mat0;  /* [nmol+1] packed-matrix offsets (doubles, even) */
  const int32_t* mol_nheavy;/* [nmol] */
  const int32_t* mol_nhyd;  /* [nmol] */
  const int32_t* mol_nocc;  /* [nmol] doubly occupied orbitals */
  const int32_t* mol_order; /* [nmol] CTA -> molecule map (largest first) */
  const int32_t* atom_Z;    /* [nat] */
  const int32_t* atom_mol;  /* [nat] */
  const int32_t* pair_i;    /* [npairs] global real-atom index of the first atom */
  const int32_t* pair_j;    /* [npairs] */
  double* atom_par;         /* [SEQM_NPAR * nat] */
  /* HOST arrays: eigensolver size classes.  Class c holds the molecules with n <= 2*np_c orbitals
   * (np_c = 4,8,...,32,40,48,56,64) that do not fit class c-1; they occupy the contiguous range
   * mol_order[cls_begin[c] .. cls_begin[c]+cls_count[c]) because mol_order is sorted by descending n. */
  int32_t cls_begin[12];
  int32_t cls_count[12];
  /* PM6_SP pairwise core-core parameters alpha[Zi*pw_dim+Zj], chi[...] (PWCCT_PM6_SP_MOPAC.csv; parameters.py:49-88);
   * NULL for the other methods */
  const double* pw_alpha;
  const double* pw_chi;
  int32_t pw_dim;
  /* pair indices grouped by class (H-H | X-H | X-X) so that the pair kernels run divergence-free with compile-time
   * block sizes: pair_perm[pair_cls_off[c] .. pair_cls_off[c+1]) are the pairs of class c (HOST offsets) */
  int32_t pair_cls_off[4];
  const int32_t* pair_perm;
  /* doubles of shared-memory Coulomb scratch the pair-centric Fock kernel needs for the largest molecule:
   * max over molecules of 20 nXX + 11 nXH + 2 nHH (pairs by class); 0 = use the two-pass kernel */
  int32_t fock_scratch;
  /* seqm_parameters["pair_outer_cutoff"] in Angstrom (Parser, basics.py:209, 326: pairs with |R_i - R_j| >= cutoff
   * are dropped from the pair list).  Here the dense triangular pair list is kept and such a pair contributes
   * exactly nothing: w, its overlap block, its core-core energy and its gradient are zero.  <= 0: no cutoff. */
  double pair_outer_cutoff;
  /* ---- PM6 with d orbitals (method SEQM_PM6_D; every pointer NULL / count 0 otherwise) -----------------------------
   * Atoms are sorted by descending Z and the reference's packd() (packd.py:195-218) requires the d-shell atoms to be
   * the FIRST mol_nsh atoms of their molecule; mol_nheavy keeps counting every atom with Z > 1.  Packed orbital
   * order: [9 per d atom][4 per sp heavy atom][1 per hydrogen], n = 5 mol_nsh + 4 mol_nheavy + mol_nhyd.
   * Pairs with a d atom ("Y pairs") carry a ragged block wd[pair_wd0[p] .. ) of (np_i x np_j) integrals
   * (kl on i | mn on j), np = 45 / 10 / 1 orbital products for a d / sp heavy / hydrogen atom; their sp x sp
   * sub-block repeats the dense w.  hab_d: (n_ypairs, 9, 9) beta-scaled overlap blocks of the Y pairs. */
  const int32_t* mol_nsh;     /* [nmol] */
  const int64_t* pair_wd0;    /* [npairs+1] */
  const int32_t* ypairs;      /* [n_ypairs] pair ids of the Y pairs, ascending */
  const int32_t* ypair_slot;  /* [npairs] index into ypairs / hab_d, -1 for pairs without a d atom */
  int32_t n_ypairs;
  int32_t oc_dim;             /* rows of onecenter_d (zmax + 1) */
  const double* onecenter_d;  /* [oc_dim][45*45] one-centre (kl|mn) containing a d orbital, per element */
  const double* mp_coef;      /* [45][7][5] multipole coefficients of the 45 local orbital products */
  const double* mp_coef_yx;   /* the same for the d atom of a (d, sp-heavy) pair */
  const double* ovl_poly;     /* [4][4][14][9][9] overlap polynomials in (xi, eta) by (n_a, n_b, kind) */
  /* set per geometry by the caller after seqm_pair_integrals_d(): */
  const double* wd;
  const double* hab_d;
} seqm_batch_t;

int seqm_abi_version(void);
const char* seqm_last_error(void);
/* largest orbital count for the shared-memory resident solvers (Jacobi eigensolver, in-SM SP2, in-SM DIIS);
 * larger molecules run the global-memory Fock / GEMM-SP2 / GEMM-DIIS path and need sp2=[True, eps] */
int seqm_max_orbitals(void);
/* largest orbital count of the eigensolver route: above seqm_max_orbitals() the density / eigenpairs of sym_eig_trunc
 * (diag.py:110-241) come from a one-sided Jacobi kernel (one CTA per molecule, matrix in shared memory or L2) up to this
 * size (256); beyond it only the SP2 density is available */
int seqm_max_orbitals_eig(void);

/* ---- batch plan construction on the device ---------------------------------------------------------------------
 * Replaces Parser.forward (seqm/basics.py:219-403: real-atom compaction, pair list i<j, molecule ids, nHeavy /
 * nHydro / nocc, closed-shell check) and Pack_Parameters.forward (basics.py:442-448: per-atom parameter gather).
 * Two calls, because the atom / pair array sizes are results of the first:
 *   1. seqm_plan_count : per-molecule arrays + every scalar the host needs (blocks until they are on the host)
 *   2. seqm_plan_fill  : atom lists, per-atom parameter rows, pair list, class-sorted pair ids
 * species: (nmol, molsize) int64 on the device, rows sorted by descending Z, zero padded (Molecule.py:188-206);
 * charges: NULL or (nmol) int64; elem_rows: (nrows, nz) per-element table, row SEQM_P_TORE = valence electrons. */
typedef struct {
  int32_t nat, npairs, nmax, zmax;
  int64_t mat_total;
  int32_t odd_electrons;      /* some molecule has an odd electron count (basics.py:298-299 raises) */
  int32_t unsorted;           /* some species row is not non-increasing (Molecule.py:198-206 raises) */
  int32_t pairs_overflow;     /* more than 2^31-1 pairs */
  int32_t fock_scratch;
  int32_t pair_cls_cnt[3];    /* H-H, X-H, X-X */
  int32_t jacobi_cls_cnt[13]; /* molecules per eigensolver size class; [12] = beyond the last class (large path) */
  int32_t elements[128];      /* 1 where element Z occurs */
} seqm_plan_counts_t;
int seqm_plan_count(const int64_t* species, int32_t nmol, int32_t molsize, const int64_t* charges,
                    const double* elem_rows, int32_t nz, int32_t* mol_atom0, int32_t* mol_pair0, int64_t* mol_mat0,
                    int32_t* mol_nheavy, int32_t* mol_nhyd, int32_t* mol_nocc, int32_t* mol_order,
                    int32_t* mol_cls_pair0 /* 3*nmol */, seqm_plan_counts_t* counts_dev, seqm_plan_counts_t* counts_host,
                    int32_t* mol_nsh /* [nmol] out, or NULL: non-NULL selects method="PM6" counting (d-shell atoms carry 9
                    orbitals; `unsorted` is also raised when a d-shell atom follows an sp-only heavy atom) */,
                    void* stream);
/* atom_par: (SEQM_NPAR, nat), rows [0, nrows) are written; real_atoms: (nat) int64 flat index mol*molsize + pos */
int seqm_plan_fill(const int64_t* species, int32_t nmol, int32_t molsize, const seqm_plan_counts_t* counts_host,
                   const int32_t* mol_atom0, const int32_t* mol_pair0, const int32_t* mol_nheavy,
                   const int32_t* mol_cls_pair0, const double* elem_rows, int32_t nrows, int32_t nz, int32_t* atom_Z,
                   int32_t* atom_mol, int64_t* real_atoms, double* atom_par, int32_t* pair_i, int32_t* pair_j,
                   int32_t* pair_perm, void* stream);

/* cal_par.py:11-28,112-169,198-257 + two_elec_two_center_int.py:116-247: dd, qq, rho0, rho1, rho2 per atom */
int seqm_atom_multipoles(const seqm_batch_t* b, void* stream);

/* hcore() pair part -- two_elec_two_center_int.py:98-283 (w), diat_overlap_PM6_SP.py:6-444 (di) and the
 * beta scaling of hcore.py:155-173.  xyz [nat*3] Angstrom; w [npairs*100]; hab [npairs*16] = di*(beta_i+beta_j)/2 */
int seqm_pair_integrals(const seqm_batch_t* b, const double* xyz, double* w, double* hab, void* stream);

/* PM6 d-orbital pair integrals (method SEQM_PM6_D), one CTA per Y pair: 45 x 45 local-frame integrals as point-charge
 * multipole interactions (two_elec_two_center_int_local_frame_d_orbitals.py:23-4164), rotation to the molecular frame
 * (RotationMatrixD.py:5-310), spd Slater overlaps (diat_overlapD.py:4-5148) with the beta scaling of hcore.py:155-173.
 * w: the dense (npairs,10,10) tensor of seqm_pair_integrals (source of the sp x sp sub-blocks);
 * wd [pair_wd0[npairs]] and hab_d [n_ypairs*81] are written; the caller then stores them in b->wd / b->hab_d. */
int seqm_pair_integrals_d(const seqm_batch_t* b, const double* xyz, const double* w, double* wd, double* hab_d,
                          void* stream);

/* hcore() assembly -- hcore.py:124-173: packed symmetric Hcore (U_ss/U_pp + core attraction, beta*S blocks) */
int seqm_hcore(const seqm_batch_t* b, const double* w, const double* hab, double* H, void* stream);

/* fock() -- fock.py:132-347.  active: optional [nmol] int32 mask (NULL = all) */
int seqm_fock(const seqm_batch_t* b, const double* P, const double* H, const double* w, double* F,
              const int32_t* active, void* stream);

/* sym_eig_trunc() -- diag.py:110-241: batched Jacobi eigensolver + density P = 2 C_occ C_occ^T.
 * evals [nmol*nmax] ascending (0 beyond n), C optional packed eigenvector matrices (columns), may be NULL.
 * Cguess: optional packed orthogonal warm-start basis (e.g. last iteration's C), may be NULL. */
int seqm_eig_density(const seqm_batch_t* b, const double* F, double* P, double* evals, double* C,
                     const double* Cguess, const int32_t* active, void* stream);

/* SP2() -- SP2.py:9-85 at each molecule's own size; P packed; niter optional [nmol] */
int seqm_sp2_density(const seqm_batch_t* b, const double* F, double* P, double eps, int32_t* niter,
                     const int32_t* active, void* stream);

/* SP2 for molecules of any size (n > seqm_max_orbitals(), e.g. C380 with 1520 orbitals): X^2 by the library's
 * register-tiled FP64 GEMM, one molecule at a time; workspace of seqm_sp2_large_workspace_bytes() bytes;
 * niter_host: optional HOST array [nmol].  Blocks the host (one 64-byte read-back per SP2 iteration). */
int64_t seqm_sp2_large_workspace_bytes(const seqm_batch_t* b);
int seqm_sp2_density_large(const seqm_batch_t* b, const double* F, double* P, double eps, int32_t* niter_host,
                           void* workspace, void* stream);

/* elec_energy() -- energy.py:26-53 */
int seqm_elec_energy(const seqm_batch_t* b, const double* P, const double* H, const double* F, double* Eelec,
                     const int32_t* active, void* stream);

/* pair_nuclear_energy() + total_energy() sums -- energy.py:91-139,177-192: EnucAB [npairs], Enuc [nmol] */
int seqm_nuclear_energy(const seqm_batch_t* b, const double* xyz, const double* w, double* EnucAB, double* Enuc,
                        void* stream);

/* scf_analytic_grad() -- anal_grad.py:16-225: grad [nat*3] eV/Angstrom (dE/dR per real atom);
 * pair_scratch [npairs*3] */
int seqm_gradient(const seqm_batch_t* b, const double* xyz, const double* P, double* pair_scratch, double* grad,
                  void* stream);

/* XL-BOMD field propagation (MolecularDynamics.py:1418-1435): P_out = kappa [c D + (1-c) P_in] + sum_j coef[j] Pt[j],
 * and Pt[slot] <- P_out, in one pass.  Pt: (m, total) history of field densities, coef: m device doubles, m <= 16.
 * P_out may alias P_in. */
int seqm_xl_propagate(int64_t total, double kappa, double c, const double* D, const double* P_in, double* Pt,
                      const double* coef, int32_t m, int32_t slot, double* P_out, void* stream);
/* XL-BOMD (seqm/dynamics/xlbomd.py:73-570, non-KSA branch): shadow electronic energy
 * sum D o F(P) - 1/2 (F(P) - Hcore) o P  (elec_energy_xl, energy.py:76-88) and its nuclear gradient at fixed
 * density D and field P (what ForceXL obtains by autograd through hcore + fock, xlbomd.py:536-551). */
int seqm_elec_energy_xl(const seqm_batch_t* b, const double* D, const double* P, const double* F, const double* H,
                        double* Eelec, void* stream);
int seqm_gradient_xl(const seqm_batch_t* b, const double* xyz, const double* D, const double* P, double* pair_scratch,
                     double* grad, void* stream);

/* the same gradient by forward-mode differentiation of the whole pair code (slower; cross-check of seqm_gradient) */
int seqm_gradient_forward(const seqm_batch_t* b, const double* xyz, const double* P, double* pair_scratch, double* grad,
                          void* stream);

/* packed eigenvector matrices -> dense (nmol, nmax, nmax), identity on the padding: the `v` of diag.py:110-241 */
int seqm_orbitals_dense(const seqm_batch_t* b, const double* C, double* V, void* stream);

/* Post-SCF by-products in one launch on the packed density: Mulliken charges q = tore - population
 * (ElectronicStructure.py:104-127; (nmol, molsize), 0 on padding), ground-state dipole (calc_ground_dipole, dipole.py:85-107;
 * (nmol, 3) or NULL; scale = to_debye * debye_to_AU, a0 = bohr in Angstrom) and force = -grad scattered into the padded
 * (nmol, molsize, 3) layout (force NULL: skipped; grad [nat*3] from seqm_gradient). */
int seqm_post_scf(const seqm_batch_t* b, const double* P, const double* xyz, const double* grad, double* q, double* dipole,
                  double* force, double a0, double scale, void* stream);

/* ---- KSA-XL-BOMD building blocks on the packed layout (seqm_ksa.cu; reference: xlbomd.py:201-341, fermi_q.py:8-72,
 * canon_dm_prt.py:6-39).  One CTA per molecule; every matrix argument is a packed [mat_total] buffer.
 *   seqm_packed_gemm     C_m = op(A_m) op(B_m), op = transpose when the flag is non-zero; C must not alias A or B
 *   seqm_scale_columns   out_m[i][k] = s * f[m*nmax + k] * C_m[i][k]   (D0 = (2 Q f) Q^t with seqm_packed_gemm)
 *   seqm_canon_prt       in place: X_m (first-order Fock perturbation in the eigenbasis) -> first-order density response in
 *                        the eigenbasis incl. the chemical-potential correction; e [nmol*nmax] eigenvalues, mu [nmol],
 *                        beta = 1 / (kB T_el), m_iter recursion steps (the reference uses 10)
 *   seqm_packed_dot      out[m] = sum_ij X_m[i][j] Y_m[i][j]
 *   seqm_packed_axpby    Y_m = a[m] X_m + c[m] Y_m  (a NULL: 1, c NULL: 1, X NULL: Y_m = c[m] Y_m) */
int seqm_packed_gemm(const seqm_batch_t* b, const double* A, const double* B, double* C, int transA, int transB, void* stream);
int seqm_scale_columns(const seqm_batch_t* b, const double* C, const double* f, double s, double* out, void* stream);
int seqm_canon_prt(const seqm_batch_t* b, const double* e, const double* mu, double* X, double beta, int m_iter, void* stream);
int seqm_packed_dot(const seqm_batch_t* b, const double* X, const double* Y, double* out, void* stream);
int seqm_packed_axpby(const seqm_batch_t* b, const double* a, const double* X, const double* c, double* Y, void* stream);
/* CIS sigma-vector, antisymmetric half (makeA_pi_batched, rcis_batch.py:336-381): Fa = response of the two-electron part of
 * the Fock operator to the ANTISYMMETRIC density Pa (exchange only; Fa is antisymmetric).  The symmetric half is seqm_fock on
 * the symmetric part of the transition density with Hcore = 0.  sp methods only. */
int seqm_fock_antisym(const seqm_batch_t* b, const double* Pa, const double* w, double* Fa, void* stream);
/* XT_m = X_m^t for every molecule of a packed buffer (out of place) */
int seqm_packed_transpose(const seqm_batch_t* b, const double* X, double* XT, void* stream);

/* MO crossing matcher -- Energy._crossing_match_molecular_orbitals / _grouped, seqm/basics.py:596-719 (called on every
 * forward after the first one on the same Molecule, basics.py:846-857): the new orbitals are permuted inside the
 * occupied and inside the virtual block so that orbital k continues old orbital k (largest |overlap|, greedy repair
 * when that is not a permutation), their signs are aligned with the old ones, and the eigenvalues follow.
 * V_new, V_old, V_out: (nmol, nmax, nmax) dense, column = MO (the `molecular_orbitals` layout; V_out must not alias
 * V_new); e_in, e_out: (nmol, nmax); S_scratch: mat_total doubles; perm, used: (nmol, nmax) int32; prio: (nmol, nmax). */
int seqm_mo_match(const seqm_batch_t* b, const double* V_new, const double* V_old, const double* e_in, double* S_scratch,
                  int32_t* perm, int32_t* used, double* prio, double* V_out, double* e_out, void* stream);

/* pack()/unpack() -- pack.py:64-96 between dense (nmol, 4*molsize, 4*molsize) and packed matrices */
int seqm_pack(const seqm_batch_t* b, const double* dense, double* packed, void* stream);
int seqm_unpack(const seqm_batch_t* b, const double* packed, double* dense, void* stream);
/* scf_loop.py:2066-2081 initial diagonal density, packed */
int seqm_initial_density(const seqm_batch_t* b, double* P, void* stream);

/* scf_loop() drivers -- scf_loop.py:164-347 (converger 0), 424-635 (1), 639-1132 (2).
 * P in/out (packed), F out (packed, Fock of the converged P), Eelec out [nmol], notconverged out [nmol].
 * workspace: device scratch of seqm_scf_workspace_bytes() bytes.  use_sp2 != 0 -> SP2 density with sp2_eps.
 * n_iter_out: host int, the iteration count the reference prints.  Blocks the host until converged. */
typedef struct seqm_scf_opts {
  double eps;        /* scf_eps */
  int32_t converger; /* 0, 1, 2 */
  double alpha;      /* mixing for converger 0 */
  int32_t use_sp2;
  double sp2_eps;
  int32_t max_iter;  /* reference: 1000 */
  int32_t warm_start;/* 1: eigensolver starts from the previous iteration's eigenvectors; 2: additionally the FIRST
                      * solve starts from the eigenvectors the caller left in C_last (restart from a nearby geometry) */
  int32_t pipeline;  /* Pulay DIIS only. 0: auto (two half-batches out of phase on two streams when nmol >= 256),
                      * 1: single stream, 2: always two half-batches. Results do not depend on it. */
} seqm_scf_opts_t;
int64_t seqm_scf_workspace_bytes(const seqm_batch_t* b, const seqm_scf_opts_t* o);
/* C_last: optional packed buffer receiving the eigenvectors of the last density solve (a warm start for the
 * final seqm_eig_density of the converged Fock matrix); NULL or unused when use_sp2 != 0. */
int seqm_scf(const seqm_batch_t* b, const seqm_scf_opts_t* o, const double* H, const double* w, double* P, double* F,
             double* Eelec, int32_t* notconverged, void* workspace, int32_t* n_iter_out, double* C_last, void* stream);

/* Optional per-kernel timing with CUDA events on the launch stream (bench.py roofline evidence).
 * enable(1) clears and starts recording; collect() synchronises and returns summed ms / launch counts
 * per kernel kind (seqm_profile_kinds() entries, names from seqm_profile_name()). */
/* number of kernel launches issued by this library since load (bench.py gpu_launches) */
long long seqm_launch_count(void);
/* eigensolver statistics since the last reset, out[8]: [0] molecules solved, [1] Jacobi sweeps, [2] solves finished by
 * the first-order occupied-virtual correction, [3] solves without any sweep, [4..7] SM cycles summed over CTAs:
 * warm-start transform, sweeps, epilogue, total */
int seqm_jacobi_stats(unsigned long long* out, int reset);
/* measured FP64 FMA peak (TFLOP/s) of the current device: roofline denominator of the FP64-bound kernels */
double seqm_fp64_peak_tflops(void);
/* C = A B for row-major n x n device matrices through the library's FP64 tensor-core GEMM (the product inside the
 * large-molecule SP2 loop, SP2.py:55, and the DIIS commutator); B == NULL: C = A A for a SYMMETRIC A through the
 * upper-triangle kernel (only tiles on and above the diagonal are computed, the rest is mirrored).  For tests and the
 * bench's DGEMM line. */
int seqm_square_product(int n, const double* A, const double* B, double* C, void* stream);
int seqm_profile_enable(int on);
int seqm_profile_kinds(void);
const char* seqm_profile_name(int kind);
int seqm_profile_collect(double* ms, int32_t* counts);

#ifdef __cplusplus
}
#endif
#endif
