// fock_kernels.cuh -- one CTA per molecule, density staged in shared memory.
//   fock_kernel         F = Hcore + one-centre + two-centre J/K        (fock.py:132-347)
//   elec_energy_kernel  Eelec = 1/2 sum P o (H + F)                     (energy.py:26-53)
// Diagonal blocks are accumulated atom-centrically (each (atom, packed kl) work item sums over all
// partner atoms), so no atomics and a fixed summation order; each off-diagonal block belongs to one pair.
#pragma once
#include "common.cuh"

#ifndef SEQM_SYNCWARP
#ifndef SEQM_HOSTEMU
#define SEQM_SYNCWARP() __syncwarp()
#else
#define SEQM_SYNCWARP() do { } while (0)
#endif
#endif

SEQM_GLOBAL void fock_kernel(seqm_batch_t b, const double* __restrict__ P, const double* __restrict__ H,
                             const double* __restrict__ w, double* __restrict__ F, const int32_t* __restrict__ active) {
  const int m = b.mol_order[blockIdx.x];
  if (active && !active[m]) return;
  const MolView v = mol_view(b, m);
  const int n = v.n;
  SEQM_DYN_SMEM(double, sP);
  const double* Pm = P + v.mat0;
  const double* Hm = H + v.mat0;
  double* Fm = F + v.mat0;
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) sP[t] = Pm[t];
  SEQM_SYNC();
  // exchange: F_AB[mu,la] = H_AB[mu,la] - 1/2 sum_{nu in A, sg in B} P_AB[nu,sg] w[pack(mu,nu)][pack(la,sg)]
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
      for (int sg = 0; sg < nj; ++sg) k += sP[(oi + nu) * n + oj + sg] * wp[pack2(mu, nu) * 10 + pack2(la, sg)];
    const int r = oi + mu, c = oj + la;
    const double f = Hm[r * n + c] - 0.5 * k;
    Fm[r * n + c] = f;
    Fm[c * n + r] = f;
  }
  // diagonal blocks
  for (int t = threadIdx.x; t < v.na * 10; t += blockDim.x) {
    const int a = t / 10, kl = t % 10;
    if (a >= v.nheavy && kl > 0) continue;
    int mu = 0;
    while ((mu + 1) * (mu + 2) / 2 <= kl) ++mu;
    const int nu = kl - mu * (mu + 1) / 2;  // mu >= nu
    const int oa = orb_off(v, a), ga = v.a0 + a;
    const double gss = par(b, SEQM_P_GSS, ga), gsp = par(b, SEQM_P_GSP, ga), gpp = par(b, SEQM_P_GPP, ga);
    const double gp2 = par(b, SEQM_P_GP2, ga), hsp = par(b, SEQM_P_HSP, ga);
    const double Pss = sP[oa * n + oa];
    double Ppt = 0.0;
    if (a < v.nheavy) Ppt = sP[(oa + 1) * n + oa + 1] + sP[(oa + 2) * n + oa + 2] + sP[(oa + 3) * n + oa + 3];
    double g;
    if (mu == 0)
      g = 0.5 * Pss * gss + Ppt * (gsp - 0.5 * hsp);
    else if (nu == 0)
      g = sP[oa * n + oa + mu] * (1.5 * hsp - 0.5 * gsp);
    else if (mu == nu) {
      const double Pk = sP[(oa + mu) * n + oa + mu];
      g = Pss * (gsp - 0.5 * hsp) + 0.5 * Pk * gpp + (Ppt - Pk) * (1.25 * gp2 - 0.25 * gpp);
    } else
      g = sP[(oa + nu) * n + oa + mu] * (0.75 * gpp - 1.25 * gp2);
    // Coulomb from every other atom o:  sum_mn wt_mn P_oo[mn] (kl on a | mn on o)
    for (int o = 0; o < v.na; ++o) {
      if (o == a) continue;
      const int oo = orb_off(v, o), no = orb_cnt(v, o);
      const bool first = a < o;
      const double* wp = w + (long long)(v.p0 + (first ? pair_local(v, a, o) : pair_local(v, o, a))) * 100;
      const int sk = first ? 10 : 1, sm = first ? 1 : 10;  // strides of (kl, mn) in this pair's w
      double j = sP[oo * n + oo] * wp[kl * sk];
      if (no == 4) {
        for (int x = 1; x < 4; ++x) {
          j += 2.0 * sP[oo * n + oo + x] * wp[kl * sk + pack2(x, 0) * sm];
          for (int y = 1; y <= x; ++y)
            j += (x == y ? 1.0 : 2.0) * sP[(oo + y) * n + oo + x] * wp[kl * sk + pack2(x, y) * sm];
        }
      }
      g += j;
    }
    const double f = Hm[(oa + mu) * n + oa + nu] + g;
    Fm[(oa + mu) * n + oa + nu] = f;
    Fm[(oa + nu) * n + oa + mu] = f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pair-centric Fock build: every pair's w is read from HBM exactly once (and only the entries its class has:
// 1 for H-H, column 0 for X-H, all 100 for X-X).  Pass 1 walks the pairs by class -- X-X pairs one WARP per pair
// (w staged in a per-warp shared tile, lanes own the 16 exchange + 10 + 10 Coulomb outputs), X-H and H-H pairs one
// thread per pair -- writes the exchange blocks F_AB and parks the Coulomb vectors J_A, J_B in shared memory;
// pass 2 sums them per atom in a fixed order (no atomics) together with the one-centre terms.
// On the device the density of the molecule and the w blocks of the X-X pairs arrive by bulk asynchronous copies
// (cp.async.bulk / UBLKCP): P in one transfer, the 800-byte w blocks through a per-warp double buffer whose next block is
// in flight while the current one is contracted (the kernel was bound by the latency of those dependent loads: long
// scoreboard 8.9 stalled warps per issue, profiles/others_r01_final.txt).
// shared: sP[n*n] | sJ[fock_scratch] | wbuf[warps][2][100] | mbarriers[warps][2] + 1
SEQM_D void tri_decode(int t, int m, int& i, int& j) {  // t-th pair (i<j) of m items, row-major
  i = 0;
  while (t >= m - 1 - i) {
    t -= m - 1 - i;
    ++i;
  }
  j = i + 1 + t;
}
SEQM_D int tri_index(int i, int j, int m) { return i * (2 * m - i - 1) / 2 + (j - i - 1); }

// Optional tail of the Fock kernel inside the DIIS loop: elec_energy of the new density + get_error
// (scf_loop.py:106-147) + the update of the active mask, for the molecule this CTA has just finished.
struct FockErr {
  int on, use_diis;
  int bulk;  // stage P and the X-X w blocks by bulk asynchronous copies (default; SEQM_B200_FOCK_BULK=0 keeps the plain loads)
  double eps;
  const double* Pold;
  const double* diis_err;
  double *Eel_run, *Eel_new, *err, *dm_err, *dm_elem;
  int32_t *notconv, *active_out;
  int* nnot;
};

SEQM_GLOBAL void fock_pair_kernel(seqm_batch_t b, const double* __restrict__ P, const double* __restrict__ H,
                                  const double* __restrict__ w, double* __restrict__ F, const int32_t* __restrict__ active,
                                  FockErr fe) {
  __shared__ double red[33];
  const int m = b.mol_order[blockIdx.x];
  if (active && !active[m]) return;
  const MolView v = mol_view(b, m);
  const int n = v.n, nh = v.nheavy, ny = v.nhyd;
  const int nXX = nh * (nh - 1) / 2, nXH = nh * ny, nHH = ny * (ny - 1) / 2;
  SEQM_DYN_SMEM(double, sP);
  double* JXA = sP + ((n * n + 1) & ~1);  // [nXX][10] onto the first atom
  double* JXB = JXA + 10 * nXX;           // [nXX][10] onto the second atom
  double* JHA = JXB + 10 * nXX;           // [nXH][10] onto the heavy atom
  double* JHB = JHA + 10 * nXH;           // [nXH]     onto the hydrogen
  double* JHH = JHB + nXH;                // [nHH][2]
  double* wbuf = sP + ((n * n + 1) & ~1) + ((b.fock_scratch + 1) & ~1);  // 16-byte aligned: destination of bulk copies
  const double* Pm = P + v.mat0;
  const double* Hm = H + v.mat0;
  double* Fm = F + v.mat0;
  const int tid = threadIdx.x, nthr = blockDim.x;
#ifndef SEQM_HOSTEMU
  const bool bulk = fe.bulk != 0;
  const int nwarp = nthr / 32;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(wbuf + (nwarp + 1) * 200);  // [warp][2], then the P barrier
  unsigned long long* barP = bars + 2 * (nwarp + 1);
  if (bulk) {
    if (tid < 2 * nwarp) seqm_mbar_init(&bars[tid], 1);
    if (tid == 0) seqm_mbar_init(barP, 1);
    seqm_mbar_fence_init();
    SEQM_SYNC();
    if (tid == 0) seqm_bulk_load(sP, Pm, (unsigned)(sizeof(double) * ((n * n + 1) & ~1)), barP);  // the slot is padded to even
    seqm_mbar_wait(barP, 0);
  } else {
    for (int t = tid; t < n * n; t += nthr) sP[t] = Pm[t];
    SEQM_SYNC();
  }
#else
  for (int t = tid; t < n * n; t += nthr) sP[t] = Pm[t];
  SEQM_SYNC();
#endif
  // ---- pass 1a: X-X pairs, one warp per pair
  {
    const int L = (nthr >= 32) ? 32 : 1;
    const int lane = tid % L, wid = tid / L, nw = nthr / L;
    double* wb2 = wbuf + wid * 200;  // this warp's two w tiles
#ifndef SEQM_HOSTEMU
    unsigned long long* bar = bars + 2 * wid;
    if (bulk && lane == 0 && wid < nXX) {
      int i0, j0;
      tri_decode(wid, nh, i0, j0);
      seqm_bulk_load(wb2, w + (long long)(v.p0 + pair_local(v, i0, j0)) * 100, 800u, &bar[0]);
    }
#endif
    int it = 0;
    for (int t = wid; t < nXX; t += nw, ++it) {
      int i, j;
      tri_decode(t, nh, i, j);
      const double* wb = wb2;
#ifndef SEQM_HOSTEMU
      if (bulk) {
        const int cur = it & 1;
        if (lane == 0 && t + nw < nXX) {  // next block of this warp into the other tile (its readers passed the syncwarp below)
          int i1, j1;
          tri_decode(t + nw, nh, i1, j1);
          seqm_bulk_load(wb2 + (cur ^ 1) * 100, w + (long long)(v.p0 + pair_local(v, i1, j1)) * 100, 800u, &bar[cur ^ 1]);
        }
        seqm_mbar_wait(&bar[cur], (unsigned)((it >> 1) & 1));
        wb = wb2 + cur * 100;
      } else
#endif
      {
        const double* wp = w + (long long)(v.p0 + pair_local(v, i, j)) * 100;
        for (int q = lane; q < 100; q += L) wb2[q] = wp[q];
        SEQM_SYNCWARP();
      }
      const int oi = 4 * i, oj = 4 * j;
      for (int o = lane; o < 36; o += L) {
        if (o < 16) {  // exchange
          const int mu = o >> 2, la = o & 3;
          double k = 0.0;
          for (int nu = 0; nu < 4; ++nu)
            for (int sg = 0; sg < 4; ++sg) k += sP[(oi + nu) * n + oj + sg] * wb[pack2(mu, nu) * 10 + pack2(la, sg)];
          const int r = oi + mu, c = oj + la;
          const double f = Hm[r * n + c] - 0.5 * k;
          Fm[r * n + c] = f;
          Fm[c * n + r] = f;
        } else {
          const bool ontoA = o < 26;
          const int q = ontoA ? o - 16 : o - 26;      // packed index of the output
          const int oo = ontoA ? oj : oi;             // the density comes from the OTHER atom
          double s = 0.0;
          for (int x = 0; x < 4; ++x)
            for (int y = 0; y <= x; ++y) {
              const int pk = pack2(x, y);
              const double d = (x == y ? 1.0 : 2.0) * sP[(oo + y) * n + oo + x];
              s += d * (ontoA ? wb[q * 10 + pk] : wb[pk * 10 + q]);
            }
          (ontoA ? JXA : JXB)[t * 10 + q] = s;
        }
      }
      SEQM_SYNCWARP();
    }
  }
  // ---- pass 1b: X-H pairs (column 0 of w only), one thread per pair
  for (int t = tid; t < nXH; t += nthr) {
    const int i = t / ny, hj = t - i * ny, j = nh + hj;
    const double* wp = w + (long long)(v.p0 + pair_local(v, i, j)) * 100;
    const int oi = 4 * i, oj = 4 * nh + hj;
    double wc[10];
    for (int kl = 0; kl < 10; ++kl) wc[kl] = wp[kl * 10];
    const double pjj = sP[oj * n + oj];
    double jb = 0.0;
    for (int x = 0; x < 4; ++x)
      for (int y = 0; y <= x; ++y) {
        const int pk = pack2(x, y);
        jb += (x == y ? 1.0 : 2.0) * sP[(oi + y) * n + oi + x] * wc[pk];
        JHA[t * 10 + pk] = wc[pk] * pjj;
      }
    JHB[t] = jb;
    for (int mu = 0; mu < 4; ++mu) {
      double k = 0.0;
      for (int nu = 0; nu < 4; ++nu) k += sP[(oi + nu) * n + oj] * wc[pack2(mu, nu)];
      const int r = oi + mu;
      const double f = Hm[r * n + oj] - 0.5 * k;
      Fm[r * n + oj] = f;
      Fm[oj * n + r] = f;
    }
  }
  // ---- pass 1c: H-H pairs
  for (int t = tid; t < nHH; t += nthr) {
    int hi, hj;
    tri_decode(t, ny, hi, hj);
    const int i = nh + hi, j = nh + hj;
    const double w00 = w[(long long)(v.p0 + pair_local(v, i, j)) * 100];
    const int oi = 4 * nh + hi, oj = 4 * nh + hj;
    JHH[2 * t] = w00 * sP[oj * n + oj];
    JHH[2 * t + 1] = w00 * sP[oi * n + oi];
    const double f = Hm[oi * n + oj] - 0.5 * sP[oi * n + oj] * w00;
    Fm[oi * n + oj] = f;
    Fm[oj * n + oi] = f;
  }
  SEQM_SYNC();
  // ---- pass 2: diagonal blocks = Hcore + one-centre + sum of the parked Coulomb vectors (fixed order)
  for (int t = tid; t < v.na * 10; t += nthr) {
    const int a = t / 10, kl = t % 10;
    if (a >= nh && kl > 0) continue;
    int mu = 0;
    while ((mu + 1) * (mu + 2) / 2 <= kl) ++mu;
    const int nu = kl - mu * (mu + 1) / 2;
    const int oa = orb_off(v, a), ga = v.a0 + a;
    const double gss = par(b, SEQM_P_GSS, ga), gsp = par(b, SEQM_P_GSP, ga), gpp = par(b, SEQM_P_GPP, ga);
    const double gp2 = par(b, SEQM_P_GP2, ga), hsp = par(b, SEQM_P_HSP, ga);
    const double Pss = sP[oa * n + oa];
    double Ppt = 0.0;
    if (a < nh) Ppt = sP[(oa + 1) * n + oa + 1] + sP[(oa + 2) * n + oa + 2] + sP[(oa + 3) * n + oa + 3];
    double g;
    if (mu == 0)
      g = 0.5 * Pss * gss + Ppt * (gsp - 0.5 * hsp);
    else if (nu == 0)
      g = sP[oa * n + oa + mu] * (1.5 * hsp - 0.5 * gsp);
    else if (mu == nu) {
      const double Pk = sP[(oa + mu) * n + oa + mu];
      g = Pss * (gsp - 0.5 * hsp) + 0.5 * Pk * gpp + (Ppt - Pk) * (1.25 * gp2 - 0.25 * gpp);
    } else
      g = sP[(oa + nu) * n + oa + mu] * (0.75 * gpp - 1.25 * gp2);
    if (a < nh) {
      for (int o = 0; o < a; ++o) g += JXB[tri_index(o, a, nh) * 10 + kl];
      for (int o = a + 1; o < nh; ++o) g += JXA[tri_index(a, o, nh) * 10 + kl];
      for (int hh = 0; hh < ny; ++hh) g += JHA[(a * ny + hh) * 10 + kl];
    } else {
      const int ha = a - nh;
      for (int o = 0; o < nh; ++o) g += JHB[o * ny + ha];
      for (int o = 0; o < ha; ++o) g += JHH[2 * tri_index(o, ha, ny) + 1];
      for (int o = ha + 1; o < ny; ++o) g += JHH[2 * tri_index(ha, o, ny)];
    }
    const double f = Hm[(oa + mu) * n + oa + nu] + g;
    Fm[(oa + mu) * n + oa + nu] = f;
    Fm[(oa + nu) * n + oa + mu] = f;
  }
  if (fe.on) {
    SEQM_SYNC();  // every element of this molecule's F has been written by this CTA
    const double* Po = fe.Pold + v.mat0;
    double e = 0.0, d2 = 0.0, dmax = 0.0;
    for (int t = tid; t < n * n; t += nthr) {
      const double p = sP[t];
      e += p * (Hm[t] + Fm[t]);
      const double d = p - Po[t];
      d2 += d * d;
      dmax = fmax(dmax, fabs(d));
    }
    e = 0.5 * block_sum(e, red);
    d2 = block_sum(d2, red);
    dmax = block_max(dmax, red);
    if (tid == 0) {
      const double err = e - fe.Eel_run[m];
      fe.err[m] = err;
      bool bad = fabs(err) > fe.eps;
      if (fe.use_diis) bad = bad || (fe.diis_err[m] > 50.0 * fe.eps);
      if (!bad) {
        fe.dm_err[m] = sqrt(d2) / (double)(4 * nh + 4 * ny);
        fe.dm_elem[m] = dmax;
      }
      const bool nc = bad || (fe.dm_err[m] > 2.0 * fe.eps) || (fe.dm_elem[m] > 15.0 * fe.eps);
      fe.Eel_new[m] = e;
      fe.notconv[m] = nc ? 1 : 0;
      fe.active_out[m] = nc ? 1 : 0;
      if (nc) {
        fe.Eel_run[m] = e;
        seqm_atomic_add(fe.nnot, 1);
      }
    }
  }
}
static inline size_t fock_pair_smem_bytes(int nmax, int scratch, int threads) {
  const size_t nw1 = (size_t)(threads / 32 + 1);  // per warp: two 100-double w tiles and two mbarriers; + the P barrier
  return sizeof(double) * ((size_t)((nmax * nmax + 1) & ~1) + (size_t)((scratch + 1) & ~1) + nw1 * 200 + nw1 * 2 + 2);
}

SEQM_GLOBAL void elec_energy_kernel(seqm_batch_t b, const double* __restrict__ P, const double* __restrict__ H,
                                    const double* __restrict__ F, double* __restrict__ E,
                                    const int32_t* __restrict__ active) {
  __shared__ double red[33];
  const int m = b.mol_order[blockIdx.x];
  if (active && !active[m]) return;
  const MolView v = mol_view(b, m);
  const int nn = v.n * v.n;
  double s = 0.0;
  for (int t = threadIdx.x; t < nn; t += blockDim.x) s += P[v.mat0 + t] * (H[v.mat0 + t] + F[v.mat0 + t]);
  s = block_sum(s, red);
  if (threadIdx.x == 0) E[m] = 0.5 * s;
}
