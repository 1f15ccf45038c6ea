// plan_kernels.cuh -- batch plan construction on the device: species (nmol, molsize) -> the index arrays of
// seqm_batch_t.  Replaces Parser.forward (seqm/basics.py:219-403: real-atom compaction, pair list, molecule ids,
// nHeavy / nHydro / nocc) and Pack_Parameters.forward (basics.py:442-448: per-atom parameter gather), without the
// reference's dense (nmol * molsize^2) difference tensor and boolean compaction.
//   plan_count_kernel  one CTA: per-molecule counts, exclusive scans (atoms, pairs, packed-matrix offsets, pair
//                      classes), size-class histogram, descending-size processing order (stable counting sort),
//                      and every scalar the host needs to size the second step
//   plan_fill_kernel   one CTA per molecule: atom lists, per-atom parameter rows from the per-element table,
//                      triangular pair list in the reference's (molecule, i, j) order, class-sorted pair ids
#pragma once
#include "common.cuh"

#define SEQM_PLAN_THREADS 256
#define SEQM_PLAN_NCLS 12

struct PlanBounds { int np2[SEQM_PLAN_NCLS]; };  // 2*NP of the eigensolver size classes

struct PlanRow { int na, nhyd, nsh, zmax, sorted; long long nel2; };  // nel2 = 2 x electron count (tore is integral)

// The reference's nSuperHeavy set (basics.py:258-269): elements that carry 9 orbitals under method="PM6"
SEQM_HD bool plan_d_shell(long long z) {
  return (z > 12 && z < 18) || (z > 20 && z < 30) || (z > 32 && z < 36) || (z > 38 && z < 48) || (z > 50 && z < 54) ||
         (z > 70 && z < 80) || z == 57;
}

SEQM_D PlanRow plan_scan_row(const long long* __restrict__ sp, int molsize, const double* __restrict__ tore, int nz, int* elem_flags,
                             int d_mode) {
  PlanRow r;
  r.na = 0; r.nhyd = 0; r.nsh = 0; r.zmax = 0; r.sorted = 1; r.nel2 = 0;
  int seen_sp_heavy = 0;
  long long prev = 0x7fffffffffffffffLL;
  double nel = 0.0;
  for (int t = 0; t < molsize; ++t) {
    const long long z = sp[t];
    if (z > prev) r.sorted = 0;
    prev = z;
    if (z > 0) {
      ++r.na;
      if (z == 1) ++r.nhyd;
      if (d_mode && plan_d_shell(z)) {
        ++r.nsh;
        if (seen_sp_heavy) r.sorted = 0;  // a d-shell atom after an sp-only heavy atom: packd.py:195-218 cannot hold it
      } else if (z > 1) {
        seen_sp_heavy = 1;
      }
      if (z > r.zmax) r.zmax = (int)z;
      if (z < nz) nel += tore[z];
      if (elem_flags && z < 128) elem_flags[z] = 1;  // benign race: every writer stores 1
    }
  }
  r.nel2 = (long long)(nel + 0.5);
  return r;
}

SEQM_GLOBAL void plan_count_kernel(const long long* __restrict__ species, int nmol, int molsize,
                                   const long long* __restrict__ charges, const double* __restrict__ tore, int nz,
                                   PlanBounds bounds, int32_t* __restrict__ atom0, int32_t* __restrict__ pair0,
                                   long long* __restrict__ mat0, int32_t* __restrict__ nheavy_o, int32_t* __restrict__ nhyd_o,
                                   int32_t* __restrict__ nocc_o, int32_t* __restrict__ order, int32_t* __restrict__ cls_pair0,
                                   seqm_plan_counts_t* __restrict__ out, int d_mode, int32_t* __restrict__ nsh_o) {
  __shared__ long long part[SEQM_PLAN_THREADS][6];
  __shared__ int s_elem[128];
  __shared__ int s_cls[SEQM_PLAN_NCLS + 1];
  __shared__ int s_max[4];   // nmax, zmax, fock_scratch, (unused)
  __shared__ int s_flag[2];  // odd electrons, unsorted
  __shared__ int s_keys[SEQM_PLAN_THREADS];
  SEQM_DYN_SMEM(int, dyn);   // hist[nbins] | add[nbins]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int nbins = (d_mode ? 9 : 4) * molsize + 2;
  int* hist = dyn;
  int* add = dyn + nbins;
  for (int i = tid; i < 128; i += nthr) s_elem[i] = 0;
  for (int i = tid; i <= SEQM_PLAN_NCLS; i += nthr) s_cls[i] = 0;
  for (int i = tid; i < 2 * nbins; i += nthr) dyn[i] = 0;
  if (tid == 0) { s_max[0] = s_max[1] = s_max[2] = s_max[3] = 0; s_flag[0] = s_flag[1] = 0; }
  SEQM_SYNC();
  const int seg = (nmol + nthr - 1) / nthr;
  const int lo = tid * seg, hi = (lo + seg < nmol) ? lo + seg : nmol;
  long long s[6] = {0, 0, 0, 0, 0, 0};
  int mx_n = 0, mx_z = 0, mx_f = 0, odd = 0, uns = 0;
  for (int m = lo; m < hi; ++m) {
    const PlanRow r = plan_scan_row(species + (long long)m * molsize, molsize, tore, nz, s_elem, d_mode);
    const int nh = r.na - r.nhyd, ny = r.nhyd, n = 4 * nh + ny + 5 * r.nsh;
    if (nsh_o) nsh_o[m] = r.nsh;
    long long nel = r.nel2 - (charges ? charges[m] : 0);
    if (nel & 1) odd = 1;
    if (!r.sorted) uns = 1;
    nheavy_o[m] = nh;
    nhyd_o[m] = ny;
    nocc_o[m] = (int)(nel / 2);
    const long long nxx = (long long)nh * (nh - 1) / 2, nxh = (long long)nh * ny, nhh = (long long)ny * (ny - 1) / 2;
    long long nn = (long long)n * n;
    nn += (nn & 1);
    s[0] += r.na; s[1] += (long long)r.na * (r.na - 1) / 2; s[2] += nn; s[3] += nhh; s[4] += nxh; s[5] += nxx;
    const long long fs = 20 * nxx + 11 * nxh + 2 * nhh;
    if (n > mx_n) mx_n = n;
    if (r.zmax > mx_z) mx_z = r.zmax;
    if (fs > mx_f) mx_f = (int)(fs > 0x7fffffffLL ? 0x7fffffffLL : fs);
    int c = 0;
    while (c < SEQM_PLAN_NCLS && bounds.np2[c] < n) ++c;  // == NCLS: beyond the last class (large path)
    seqm_atomic_add(&s_cls[c], 1);
    seqm_atomic_add(&hist[n < nbins ? n : nbins - 1], 1);
  }
  for (int q = 0; q < 6; ++q) part[tid][q] = s[q];
  seqm_atomic_max(&s_max[0], mx_n);
  seqm_atomic_max(&s_max[1], mx_z);
  seqm_atomic_max(&s_max[2], mx_f);
  if (odd) s_flag[0] = 1;
  if (uns) s_flag[1] = 1;
  SEQM_SYNC();
  __shared__ long long s_tot[6];
#ifndef SEQM_HOSTEMU
  // exclusive scan of the per-thread partial sums, one warp per column: every lane scans a chunk of consecutive threads,
  // the chunk sums are scanned with shuffles (the single-thread version of round 1 was half of this kernel's 0.18 ms)
  if (tid < 6 * 32 && (nthr & 31) == 0) {
    const int q = tid >> 5, lane = tid & 31, chunk = nthr >> 5;
    long long sum = 0;
    for (int t = lane * chunk; t < (lane + 1) * chunk; ++t) sum += part[t][q];
    long long incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    long long runq = incl - sum;
    for (int t = lane * chunk; t < (lane + 1) * chunk; ++t) {
      const long long v = part[t][q];
      part[t][q] = runq;
      runq += v;
    }
    if (lane == 31) s_tot[q] = incl;
  }
  SEQM_SYNC();
  if (tid == 0) {
    long long run[6];
    for (int q = 0; q < 6; ++q) run[q] = s_tot[q];
#else
  if (tid == 0) {  // exclusive scan of the per-thread partial sums; the totals go to the host
    long long run[6] = {0, 0, 0, 0, 0, 0};
    for (int t = 0; t < nthr; ++t)
      for (int q = 0; q < 6; ++q) {
        const long long v = part[t][q];
        part[t][q] = run[q];
        run[q] += v;
      }
    (void)s_tot;
#endif
    atom0[nmol] = (int32_t)run[0];
    pair0[nmol] = (int32_t)run[1];
    mat0[nmol] = run[2];
    out->nat = (int32_t)run[0];
    out->npairs = (int32_t)run[1];
    out->mat_total = run[2];
    out->pairs_overflow = (run[1] > 0x7fffffffLL) ? 1 : 0;
    out->pair_cls_cnt[0] = (int32_t)run[3];
    out->pair_cls_cnt[1] = (int32_t)run[4];
    out->pair_cls_cnt[2] = (int32_t)run[5];
    out->nmax = s_max[0];
    out->zmax = s_max[1];
    out->fock_scratch = s_max[2];
    out->odd_electrons = s_flag[0];
    out->unsorted = s_flag[1];
    for (int c = 0; c <= SEQM_PLAN_NCLS; ++c) out->jacobi_cls_cnt[c] = s_cls[c];
    for (int z = 0; z < 128; ++z) out->elements[z] = s_elem[z];
    // processing order: descending n; cursor[key] = first slot of that key
    int start = 0;
    for (int key = nbins - 1; key >= 0; --key) {
      const int c = hist[key];
      hist[key] = start;
      start += c;
    }
  }
  SEQM_SYNC();
  long long run[6];
  for (int q = 0; q < 6; ++q) run[q] = part[tid][q];
  for (int m = lo; m < hi; ++m) {
    const int nh = nheavy_o[m], ny = nhyd_o[m], na = nh + ny, n = 4 * nh + ny + (nsh_o ? 5 * nsh_o[m] : 0);
    atom0[m] = (int32_t)run[0];
    pair0[m] = (int32_t)run[1];
    mat0[m] = run[2];
    cls_pair0[m] = (int32_t)run[3];
    cls_pair0[nmol + m] = (int32_t)run[4];
    cls_pair0[2 * nmol + m] = (int32_t)run[5];
    long long nn = (long long)n * n;
    nn += (nn & 1);
    run[0] += na; run[1] += (long long)na * (na - 1) / 2; run[2] += nn;
    run[3] += (long long)ny * (ny - 1) / 2; run[4] += (long long)nh * ny; run[5] += (long long)nh * (nh - 1) / 2;
  }
  // stable placement, nthr molecules at a time in index order
  for (int base = 0; base < nmol; base += nthr) {
    const int m = base + tid;
    int key = -1;
    if (m < nmol) {
      key = 4 * nheavy_o[m] + nhyd_o[m] + (nsh_o ? 5 * nsh_o[m] : 0);
      if (key >= nbins) key = nbins - 1;
    }
    s_keys[tid] = key;
    SEQM_SYNC();
    if (m < nmol) {
      int rank = 0;
      for (int j = 0; j < tid; ++j) rank += (s_keys[j] == key);
      order[hist[key] + rank] = m;
      seqm_atomic_add(&add[key], 1);
    }
    SEQM_SYNC();
    for (int bkt = tid; bkt < nbins; bkt += nthr) {
      hist[bkt] += add[bkt];
      add[bkt] = 0;
    }
    SEQM_SYNC();
  }
}

// pair q of a molecule with na atoms in (i, j > i) lexicographic order
SEQM_D void plan_pair_decode(int q, int na, int* i_out, int* j_out) {
  const double t = 2.0 * na - 1.0;
  int i = (int)((t - sqrt(t * t - 8.0 * (double)q)) * 0.5);
  if (i < 0) i = 0;
  while (i > 0 && (long long)i * (2 * na - i - 1) / 2 > q) --i;
  while ((long long)(i + 1) * (2 * na - i - 2) / 2 <= q) ++i;
  *i_out = i;
  *j_out = i + 1 + (q - (int)((long long)i * (2 * na - i - 1) / 2));
}

SEQM_GLOBAL void plan_fill_kernel(const long long* __restrict__ species, int nmol, int molsize, int nat,
                                  const int32_t* __restrict__ atom0, const int32_t* __restrict__ pair0,
                                  const int32_t* __restrict__ nheavy, const int32_t* __restrict__ cls_pair0,
                                  const double* __restrict__ elem_rows, int nrows, int nz, int off_xh, int off_xx,
                                  int32_t* __restrict__ atom_Z, int32_t* __restrict__ atom_mol, long long* __restrict__ real_atoms,
                                  double* __restrict__ atom_par, int32_t* __restrict__ pair_i, int32_t* __restrict__ pair_j,
                                  int32_t* __restrict__ pair_perm) {
  const int m = blockIdx.x;
  const int a0 = atom0[m], na = atom0[m + 1] - a0, nh = nheavy[m], ny = na - nh;
  const long long* sp = species + (long long)m * molsize;
  for (int t = threadIdx.x; t < na; t += blockDim.x) {
    const long long z = sp[t];
    atom_Z[a0 + t] = (int32_t)z;
    atom_mol[a0 + t] = m;
    real_atoms[a0 + t] = (long long)m * molsize + t;
  }
  for (int t = threadIdx.x; t < na * nrows; t += blockDim.x) {
    const int r = t / na, a = t - r * na;
    const long long z = sp[a];
    atom_par[(long long)r * nat + a0 + a] = (z < nz) ? elem_rows[(long long)r * nz + z] : 0.0;
  }
  const int p0 = pair0[m], np_ = na * (na - 1) / 2;
  const int b_hh = cls_pair0[m], b_xh = off_xh + cls_pair0[nmol + m], b_xx = off_xx + cls_pair0[2 * nmol + m];
  for (int q = threadIdx.x; q < np_; q += blockDim.x) {
    int i, j;
    plan_pair_decode(q, na, &i, &j);
    pair_i[p0 + q] = a0 + i;
    pair_j[p0 + q] = a0 + j;
    // rows are sorted by descending Z: atoms [0, nh) are heavy, [nh, na) hydrogen.  Class-sorted position keeps the
    // (i, j) order inside each class (what a stable sort by class produces)
    int pos;
    if (j < nh) pos = b_xx + (int)((long long)i * (2 * nh - i - 1) / 2) + (j - i - 1);
    else if (i < nh) pos = b_xh + i * ny + (j - nh);
    else pos = b_hh + (int)((long long)(i - nh) * (2 * ny - (i - nh) - 1) / 2) + (j - i - 1);
    pair_perm[pos] = p0 + q;
  }
}
