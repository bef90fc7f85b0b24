// rxg_lists_qeq.cuh -- neighbour lists (A2, A3), QEq matrix assembly (D1) and the CG kernels (D2-D4).
// Reference: src/main.F90:321-477 (NEIGHBORLIST, GetNonbondingPairList), src/qeq.F90 (QEq).
#pragma once
#include "rxg_halo_cells.cuh"

namespace rxg {

__device__ __forceinline__ int rec_index(double w) { return (int)(__double_as_longlong(w) & 0xffffffffLL); }
__device__ __forceinline__ int rec_type(double w) { return (int)(__double_as_longlong(w) >> 32); }

// ---------------------------------------------------------------------------------------------------
// A2: bonded neighbour list.  One thread per atom (cell order, so a warp walks the same 27 cells).
// Row order == the reference's: cells c4,c5,c6 in -1..1 with c6 fastest, in-cell descending index.
__global__ void k_nbrlist(DevGrid g, const DevFF *__restrict__ ffp, int ntot, int nlayer, int MAXN,
                          int *__restrict__ nbrcnt, int *__restrict__ nbrlist, int *__restrict__ ovf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntot) return;
  const DevFF &ff = *ffp;
  // which cell holds sorted slot t?  recover from the record's own cell id
  double4 me = g.sorted[t];
  int m = rec_index(me.w), mty = rec_type(me.w);
  int cid = g.cell_of[m];
  int c3 = cid % g.dim[2] - g.L, c2 = (cid / g.dim[2]) % g.dim[1] - g.L, c1 = cid / (g.dim[2] * g.dim[1]) - g.L;
  // NEIGHBORLIST visits cells -nlayer .. cc-1+nlayer only (src/main.F90:343-345)
  if (c1 < -nlayer || c1 >= g.nc[0] + nlayer || c2 < -nlayer || c2 >= g.nc[1] + nlayer || c3 < -nlayer || c3 >= g.nc[2] + nlayer) {
    nbrcnt[m] = 0;
    return;
  }
  int cnt = 0;
  int *row = nbrlist + (size_t)m * MAXN;
  for (int c4 = -1; c4 <= 1; c4++)
    for (int c5 = -1; c5 <= 1; c5++) {
      int base = ((c1 + c4 + g.L) * g.dim[1] + (c2 + c5 + g.L)) * g.dim[2] + (c3 + g.L);
      int s = g.start[base - 1], e = g.start[base + 2];   // cells c3-1 .. c3+1 are contiguous (z fastest)
      for (int k = s; k < e; k++) {
        double4 o = g.sorted[k];
        int n = rec_index(o.w);
        if (n == m) continue;
        int nty = rec_type(o.w);
        int inxn = ff.inxn2[(mty - 1) + ff.nso * (nty - 1)];
        if (inxn <= 0) continue;
        double dr2 = dist2_rn(sub_rn(o.x, me.x), sub_rn(o.y, me.y), sub_rn(o.z, me.z));
        if (dr2 < ff.rc2[inxn - 1]) {
          if (cnt < MAXN) row[cnt] = n;
          cnt++;
        }
      }
    }
  nbrcnt[m] = cnt;
  if (cnt > MAXN) atomicMax(ovf, cnt);
}

// reverse index: nbrindx(i,i1) = j1 with nbrlist(j,j1) == i, src/main.F90:383-398
__global__ void k_nbrindx(int ntot, int MAXN, const int *__restrict__ nbrcnt, const int *__restrict__ nbrlist,
                          int *__restrict__ nbrindx, int *__restrict__ bad) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int i = t / MAXN, i1 = t % MAXN;
  if (i >= ntot || i1 >= nbrcnt[i]) return;
  int j = nbrlist[(size_t)i * MAXN + i1];
  int found = -1;
  int nj = nbrcnt[j];
  for (int j1 = 0; j1 < nj; j1++)
    if (nbrlist[(size_t)j * MAXN + j1] == i) found = j1;
  nbrindx[(size_t)i * MAXN + i1] = found;
  if (found < 0) atomicExch(bad, 1);
}

// ---------------------------------------------------------------------------------------------------
// A3 (+D1): 10 A pair list over the stencil runs; one warp per resident atom.
//   QEQ=false: GetNonbondingPairList, fp64 dr2 <= rctap2 (src/main.F90:456-458)
//   QEQ=true : qeq_initialize, real(4) dr2 < rctap2 and hessian = lerp of TBL_Eclmb_QEq in r^2 (src/qeq.F90:222-240)
// FILL=false counts, FILL=true writes col (and val).
template <bool QEQ, bool FILL>
__global__ void __launch_bounds__(512) k_pairlist(DevGrid g, const DevFF *__restrict__ ffp, const int *__restrict__ runs,
                                                  int nruns, int natoms, int ntot, const double *__restrict__ pos, int NB,
                                                  const int *__restrict__ itype, int *__restrict__ rowcnt,
                                                  const long long *__restrict__ rowptr, int *__restrict__ col,
                                                  double *__restrict__ val, int maxrow, int *__restrict__ ovf) {
  const int lane = threadIdx.x & 31;
  // warps walk the atoms in cell order: the 8-16 warps of a CTA then scan (nearly) the same stencil runs, so the
  // candidate records are served by L1 instead of L2
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slot >= ntot) return;
  const int i = rec_index(g.sorted[slot].w);
  if (i >= natoms) return;
  const DevFF &ff = *ffp;
  int cid = g.cell_of[i];
  int c3 = -1000, c2 = 0, c1 = 0;
  if (cid >= 0) { c3 = cid % g.dim[2] - g.L; c2 = (cid / g.dim[2]) % g.dim[1] - g.L; c1 = cid / (g.dim[2] * g.dim[1]) - g.L; }
  // only atoms in resident cells 0..nbcc-1 own a row (src/main.F90:438-440, src/qeq.F90:202-204)
  if (cid < 0 || c1 < 0 || c1 >= g.nc[0] || c2 < 0 || c2 >= g.nc[1] || c3 < 0 || c3 >= g.nc[2]) {
    if (!FILL && lane == 0) rowcnt[i] = 0;
    return;
  }
  const double xi = pos[i], yi = pos[NB + i], zi = pos[2 * NB + i];
  const int ity = itype[i];
  const float rctap2f = (float)ff.rctap2;
  long long base = FILL ? rowptr[i] : 0;
  int cnt = 0;
  for (int r = 0; r < nruns; r++) {
    int dx = runs[4 * r], dy = runs[4 * r + 1], zlo = runs[4 * r + 2], zhi = runs[4 * r + 3];
    int a1 = c1 + dx, a2 = c2 + dy;
    int z0 = c3 + zlo, z1 = c3 + zhi;
    if (a1 < -g.L || a1 >= g.nc[0] + g.L || a2 < -g.L || a2 >= g.nc[1] + g.L) continue;
    if (z0 < -g.L) z0 = -g.L;
    if (z1 >= g.nc[2] + g.L) z1 = g.nc[2] + g.L - 1;
    if (z1 < z0) continue;
    int cbase = ((a1 + g.L) * g.dim[1] + (a2 + g.L)) * g.dim[2] + g.L;
    int s = g.start[cbase + z0], e = g.start[cbase + z1 + 1];
    for (int k0 = s; k0 < e; k0 += 32) {
      int k = k0 + lane;
      bool acc = false;
      int j = -1, jty = 0;
      double dr2 = 0.0;
      if (k < e) {
        double4 o = g.sorted[k];
        j = rec_index(o.w);
        jty = rec_type(o.w);
        if (j != i) {
          dr2 = dist2_rn(sub_rn(xi, o.x), sub_rn(yi, o.y), sub_rn(zi, o.z));
          acc = QEQ ? ((float)dr2 < rctap2f) : (dr2 <= ff.rctap2);
        }
      }
      unsigned mask = __ballot_sync(0xffffffffu, acc);
      if (FILL && acc) {
        int w = cnt + __popc(mask & ((1u << lane) - 1u));
        col[base + w] = j;
        if (QEQ) {
          double d2 = (double)(float)dr2;                    // real(4) dr2 promoted back
          int itb = (int)mul_rn(d2, ff.UDRi);
          double drtb = mul_rn(sub_rn(d2, mul_rn((double)itb, ff.UDR)), ff.UDRi);
          int inxn = ff.inxn2[(ity - 1) + ff.nso * (jty - 1)];
          double h = 0.0;
          if (inxn > 0 && itb >= 1 && itb < ff.ntable) {
            const double *T = ff.TBL_Eclmb_QEq + (size_t)(inxn - 1) * ff.ntable + (itb - 1);
            h = add_rn(mul_rn(sub_rn(1.0, drtb), T[0]), mul_rn(drtb, T[1]));
          }
          val[base + w] = h;
        }
      }
      cnt += __popc(mask);
    }
  }
  if (!FILL && lane == 0) {
    rowcnt[i] = cnt;
    if (cnt > maxrow) atomicMax(ovf, cnt);
  }
}

inline int build_nbrlist(Ctx *c) {
  const int n = c->cp[6];
  RXG_CUDA(cudaMemsetAsync(c->d_flag, 0, 2 * sizeof(int), c->st));
  RXG_CUDA(cudaMemsetAsync(c->nbrcnt, 0, sizeof(int) * n, c->st));
  LAUNCH(c, k_nbrlist, cdiv(n, 128), 128, 0, c->gb, c->d_ff, n, c->cfg.nmincell, c->MAXN, c->nbrcnt, c->nbrlist, c->d_flag);
  LAUNCH(c, k_nbrindx, cdiv((long long)n * c->MAXN, 256), 256, 0, n, c->MAXN, c->nbrcnt, c->nbrlist, c->nbrindx, c->d_flag + 1);
  RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (c->h_int[0] > c->MAXN) {
    c->err = "ERROR: overflow of max # in neighbor list, " + std::to_string(c->h_int[0]);
    return RXG_ERR_MAXNEIGHBS;
  }
  if (c->h_int[1]) {
    c->err = "ERROR: inconsistency between nbrlist and nbrindx found";
    return RXG_ERR_STATE;
  }
  return RXG_OK;
}

template <bool QEQ>
int build_pairlist(Ctx *c) {
  const int n = c->natoms;
  RXG_CUDA(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->st));
  RXG_CUDA(cudaMemsetAsync(c->rowcnt, 0, sizeof(int) * (size_t)(n + 1), c->st));
  int grid = cdiv((long long)c->cp[6] * 32, 512);
  LAUNCH(c, (k_pairlist<QEQ, false>), grid, 512, 0, c->gnb, c->d_ff, c->d_runs, c->nruns, n, c->cp[6], c->pos, c->NB, c->itype, c->rowcnt,
         c->rowptr, c->col, c->val, c->cfg.maxneighbs10, c->d_flag);
  RXG_TRY(ensure_blk(c, n));
  RXG_TRY(device_scan<long long>(c, c->rowcnt, n, c->rowptr, c->d_blk64, (long long *)(c->d_acc + 32)));
  RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaMemcpyAsync(c->h_acc + 32, c->d_acc + 32, sizeof(long long), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (c->h_int[0] > c->cfg.maxneighbs10) {
    c->err = "ERROR: nbplist greater then MAXNEIGHBS10, value " + std::to_string(c->h_int[0]);
    return RXG_ERR_MAXNEIGHBS10;
  }
  long long nnz = *(long long *)(c->h_acc + 32);
  if (nnz > c->nnz_cap) {
    if (c->col) cudaFree(c->col);
    if (c->val) cudaFree(c->val);
    c->nnz_cap = nnz + nnz / 16 + 1024;
    RXG_CUDA(cudaMalloc(&c->col, sizeof(int) * c->nnz_cap));
    RXG_CUDA(cudaMalloc(&c->val, sizeof(double) * c->nnz_cap));
  }
  c->nnz = nnz;
  c->list_is_qeq = QEQ;
  LAUNCH(c, (k_pairlist<QEQ, true>), grid, 512, 0, c->gnb, c->d_ff, c->d_runs, c->nruns, n, c->cp[6], c->pos, c->NB, c->itype, c->rowcnt,
         c->rowptr, c->col, c->val, c->cfg.maxneighbs10, c->d_flag);
  return RXG_OK;
}

// ---------------------------------------------------------------------------------------------------
// QEq CG.  d_acc slots: 0 Est, 1 hshs, 2 hsht, 3 g.h (s), 4 g.h (t), 5 sum qs, 6 sum qt, 7 g.g (s) new, 8 g.g (t) new,
// 9 g.g (s) old, 10 g.g (t) old.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__device__ __forceinline__ void block_accumulate(double (&v)[NV], double *__restrict__ acc) {
  __shared__ double sh[NV][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double s = warp_sum(v[k]);
    if (lane == 0) sh[k][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; w++) s += sh[threadIdx.x][w];
    atomicAdd(&acc[threadIdx.x], s);
  }
}

// qs=q, qt=0 for every atom, hsq.z = q  (isQEq==1 branch, src/qeq.F90:39-47; ghosts get their values from MODE_COPY)
__global__ void k_qeq_init(int natoms, int ntot_prev, const double *__restrict__ q, double2 *__restrict__ qst,
                           double4 *__restrict__ hsq, double *__restrict__ qsfp, double *__restrict__ qsfv, int mode,
                           double Lex_fqs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot_prev) return;
  if (i < natoms) {
    double qi = q[i];
    if (mode == 1) {
      qsfp[i] = qi; qsfv[i] = 0.0;
      qst[i] = make_double2(qi, 0.0);
    } else {
      qst[i] = make_double2(Lex_fqs * qsfp[i] + (1.0 - Lex_fqs) * qi, 0.0);
    }
    double4 h = hsq[i];
    h.z = qi;
    hsq[i] = h;
  } else if (mode == 1) {
    qst[i] = make_double2(0.0, 0.0);   // qs(:)=0; qt(:)=0
  }
}

// D3: get_gradient, src/qeq.F90:321-363.  One warp per row; 8 rows per 256-thread CTA.
__global__ void __launch_bounds__(256) k_gradient(int natoms, const long long *__restrict__ rowptr, const int *__restrict__ col,
                                                  const double *__restrict__ val, const double2 *__restrict__ qst,
                                                  const int *__restrict__ itype, const DevFF *__restrict__ ffp,
                                                  double2 *__restrict__ gst, double *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double part[2] = {0.0, 0.0};
  if (i < natoms) {
    long long s = rowptr[i], e = rowptr[i + 1];
    double gs = 0.0, gt = 0.0;
    for (long long k = s + lane; k < e; k += 32) {
      double h = __ldcs(val + k);
      int j = __ldcs(col + k);
      double2 x = qst[j];
      gs += h * x.x;
      gt += h * x.y;
    }
    gs = warp_sum(gs);
    gt = warp_sum(gt);
    if (lane == 0) {
      int t = itype[i] - 1;
      double eta = ffp->eta[t], chi = ffp->chi[t];
      double2 x = qst[i];
      double a = sub_rn(sub_rn(-chi, mul_rn(eta, x.x)), gs);
      double b = sub_rn(sub_rn(-1.0, mul_rn(eta, x.y)), gt);
      gst[i] = make_double2(a, b);
      part[0] = a * a;
      part[1] = b * b;
    }
  }
  block_accumulate<2>(part, acc + 7);
}

// D2: get_hsh (src/qeq.F90:271-318) fused with the g.h dots of src/qeq.F90:119-124
__global__ void __launch_bounds__(256) k_hsh(int natoms, const long long *__restrict__ rowptr, const int *__restrict__ col,
                                             const double *__restrict__ val, const double4 *__restrict__ hsq,
                                             const double2 *__restrict__ gst, const int *__restrict__ itype,
                                             const DevFF *__restrict__ ffp, double *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < natoms) {
    long long s = rowptr[i], e = rowptr[i + 1];
    double4 me = hsq[i];
    double ts = 0.0, tt = 0.0, es = 0.0;
    for (long long k = s + lane; k < e; k += 32) {
      double h = __ldcs(val + k);
      int j = __ldcs(col + k);
      double4 x = hsq[j];
      ts += h * x.x;
      tt += h * x.y;
      double e1 = 0.5 * h * me.z * x.z;
      es += (j < natoms) ? (e1 + e1) : e1;   // resident pairs appear in both rows
    }
    ts = warp_sum(ts);
    tt = warp_sum(tt);
    es = warp_sum(es);
    if (lane == 0) {
      int t = itype[i] - 1;
      double eta = ffp->eta[t], chi = ffp->chi[t];
      ts += eta * me.x;
      tt += eta * me.y;
      double2 g = gst[i];
      part[0] = es + chi * me.z + 0.5 * eta * me.z * me.z;
      part[1] = ts * me.x;
      part[2] = tt * me.y;
      part[3] = g.x * me.x;
      part[4] = g.y * me.y;
    }
  }
  block_accumulate<5>(part, acc + 0);
}



// ---------------------------------------------------------------------------------------------------
// Single-pass CG (default).  The reference does two sparse products per iteration, H.(hs,ht) in get_hsh and
// H.(qs,qt) in get_gradient (src/qeq.F90:105,157).  Because qs_new = qs + lmin*hs, the second product is
// H.qs_old + lmin*H.hs, so the gradient follows from the first product: gs_new = gs - lmin_s*(eta*hs + H.hs).
// One matrix stream per iteration instead of two; identical in exact arithmetic (round-off: see DESIGN.md "QEq").
// Est (src/qeq.F90:296-306) needs sum_j w_ij H_ij q_j with w = 2 for resident j, 1 for ghost j; it is
// carried the same way in wst = resident-weighted H.(qs,qt), with q = qs - mu*qt.
template <bool INIT>
__global__ void __launch_bounds__(256) k_spmv1(int natoms, const long long *__restrict__ rowptr, const int *__restrict__ col,
                                               const double *__restrict__ val, const double2 *__restrict__ x,
                                               const double2 *__restrict__ qst, const double *__restrict__ q,
                                               double2 *__restrict__ gst, double2 *__restrict__ tst,
                                               double2 *__restrict__ ust, double2 *__restrict__ wst,
                                               const int *__restrict__ itype, const DevFF *__restrict__ ffp,
                                               double *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < natoms) {
    long long s = rowptr[i], e = rowptr[i + 1];
    double a = 0.0, b = 0.0, ga = 0.0, gb = 0.0;
    for (long long k = s + lane; k < e; k += 32) {
      double h = __ldcs(val + k);
      int j = __ldcs(col + k);
      double2 v = x[j];
      double pa = h * v.x, pb = h * v.y;
      a += pa; b += pb;
      if (j >= natoms) { ga += pa; gb += pb; }
    }
    a = warp_sum(a); b = warp_sum(b); ga = warp_sum(ga); gb = warp_sum(gb);
    if (lane == 0) {
      int t = itype[i] - 1;
      double eta = ffp->eta[t], chi = ffp->chi[t];
      double2 me = x[i];
      if (INIT) {
        double g1 = sub_rn(sub_rn(-chi, mul_rn(eta, me.x)), a);
        double g2 = sub_rn(sub_rn(-1.0, mul_rn(eta, me.y)), b);
        gst[i] = make_double2(g1, g2);
        wst[i] = make_double2(2.0 * a - ga, 2.0 * b - gb);
        part[0] = g1 * g1; part[1] = g2 * g2;
      } else {
        double ts = eta * me.x + a, tt = eta * me.y + b;
        tst[i] = make_double2(ts, tt);
        ust[i] = make_double2(2.0 * a - ga, 2.0 * b - gb);
        double2 g = gst[i], w = wst[i];
        double mu = acc[11], qi = q[i];
        part[0] = chi * qi + 0.5 * eta * qi * qi + 0.5 * qi * (w.x - mu * w.y);
        part[1] = ts * me.x; part[2] = tt * me.y; part[3] = g.x * me.x; part[4] = g.y * me.y;
      }
    }
  }
  if (INIT) { double p2[2] = {part[0], part[1]}; block_accumulate<2>(p2, acc + 7); }
  else block_accumulate<5>(part, acc + 0);
}
__global__ void k_roll_g(double *__restrict__ acc) {
  acc[9] = acc[7]; acc[10] = acc[8];
  acc[5] = 0.0; acc[6] = 0.0; acc[7] = 0.0; acc[8] = 0.0;
}
// qs,qt step (src/qeq.F90:136-137) + gradient / Est-bookkeeping recurrences + partial sums (sum qs, sum qt, g.g)
__global__ void __launch_bounds__(256) k_cg_update1(int natoms, float lmin_s, float lmin_t, const double2 *__restrict__ hst,
                                                    const double2 *__restrict__ tst, const double2 *__restrict__ ust,
                                                    double2 *__restrict__ qst, double2 *__restrict__ gst,
                                                    double2 *__restrict__ wst, double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double part[4] = {0.0, 0.0, 0.0, 0.0};
  if (i < natoms) {
    const double ls = (double)lmin_s, lt = (double)lmin_t;   // real(4) lmin promoted, SURVEY Q3
    double2 h = hst[i], x = qst[i], g = gst[i], t = tst[i], u = ust[i], w = wst[i];
    x.x = add_rn(x.x, mul_rn(ls, h.x));
    x.y = add_rn(x.y, mul_rn(lt, h.y));
    g.x = sub_rn(g.x, mul_rn(ls, t.x));
    g.y = sub_rn(g.y, mul_rn(lt, t.y));
    w.x = add_rn(w.x, mul_rn(ls, u.x));
    w.y = add_rn(w.y, mul_rn(lt, u.y));
    qst[i] = x; gst[i] = g; wst[i] = w;
    part[0] = x.x; part[1] = x.y; part[2] = g.x * g.x; part[3] = g.y * g.y;
  }
  block_accumulate<4>(part, acc + 5);
}
// mu, q = qs - mu*qt (src/qeq.F90:147-150) and the Fletcher-Reeves direction update (:160-161)
__global__ void k_cg_update2(int natoms, const double2 *__restrict__ qst, const double2 *__restrict__ gst,
                             double2 *__restrict__ hst, double *__restrict__ q, double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double mu = acc[5] / acc[6];
  if (i == 0) acc[11] = mu;
  if (i >= natoms) return;
  double bs = acc[7] / acc[9], bt = acc[8] / acc[10];
  double2 x = qst[i], g = gst[i], h = hst[i];
  q[i] = sub_rn(x.x, mul_rn(mu, x.y));
  h.x = add_rn(g.x, mul_rn(bs, h.x));
  h.y = add_rn(g.y, mul_rn(bt, h.y));
  hst[i] = h;
}
__global__ void k_h_from_g2(int natoms, const double2 *__restrict__ gst, double2 *__restrict__ hst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < natoms) hst[i] = gst[i];
}

// ---------------------------------------------------------------------------------------------------
// STRICT-ORDER validation path (RXG_STRICT_ORDER=1): the same CG with every sum taken in the reference's serial
// order and without FMA, so that the iterates are bit-identical to a serial x86-64 build of the reference.
// The reference's CG amplifies round-off (its real(4) step length keeps it in a noise-dominated regime: an FMA
// build of the same Fortran/C++ loops changes the converged charges by ~1e-5), so only this path can be compared
// at 1e-8; the production kernels above differ from it by summation order alone.  Small systems only.
__global__ void k_rows_strict_grad(int natoms, const long long *__restrict__ rowptr, const int *__restrict__ col,
                                   const double *__restrict__ val, const double2 *__restrict__ qst,
                                   const int *__restrict__ itype, const DevFF *__restrict__ ffp, double2 *__restrict__ gst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  double gs = 0.0, gt = 0.0;
  for (long long k = rowptr[i]; k < rowptr[i + 1]; k++) {
    double2 x = qst[col[k]];
    gs = add_rn(gs, mul_rn(val[k], x.x));
    gt = add_rn(gt, mul_rn(val[k], x.y));
  }
  int t = itype[i] - 1;
  double eta = ffp->eta[t], chi = ffp->chi[t];
  double2 x = qst[i];
  gst[i] = make_double2(sub_rn(sub_rn(-chi, mul_rn(eta, x.x)), gs), sub_rn(sub_rn(-1.0, mul_rn(eta, x.y)), gt));
}
__global__ void k_rows_strict_hsh(int natoms, const long long *__restrict__ rowptr, const int *__restrict__ col,
                                  const double *__restrict__ val, const double4 *__restrict__ hsq,
                                  const int *__restrict__ itype, const DevFF *__restrict__ ffp, double4 *__restrict__ rowbuf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  int t = itype[i] - 1;
  double eta = ffp->eta[t], chi = ffp->chi[t];
  double4 me = hsq[i];
  double ts = mul_rn(eta, me.x), tt = mul_rn(eta, me.y);
  double es = add_rn(mul_rn(chi, me.z), mul_rn(mul_rn(mul_rn(0.5, eta), me.z), me.z));
  for (long long k = rowptr[i]; k < rowptr[i + 1]; k++) {
    int j = col[k];
    double4 x = hsq[j];
    ts = add_rn(ts, mul_rn(val[k], x.x));
    tt = add_rn(tt, mul_rn(val[k], x.y));
    double e1 = mul_rn(mul_rn(mul_rn(0.5, val[k]), me.z), x.z);
    es = add_rn(es, e1);
    if (j < natoms) es = add_rn(es, e1);
  }
  rowbuf[i] = make_double4(ts, tt, es, 0.0);
}
// which: 0 = after hsh rows (acc 0..4), 1 = sums of qs,qt (acc 5,6), 2 = g.g (acc 7,8).  One thread, index order.
__global__ void k_seq_reduce(int which, int natoms, const double4 *__restrict__ rowbuf, const double4 *__restrict__ hsq,
                             const double2 *__restrict__ gst, const double2 *__restrict__ qst, double *__restrict__ acc) {
  if (which == 0) {
    double e = 0, a = 0, b = 0, c = 0, d = 0;
    for (int i = 0; i < natoms; i++) {
      double4 r = rowbuf[i], h = hsq[i];
      double2 g = gst[i];
      e = add_rn(e, r.z);
      a = add_rn(a, mul_rn(r.x, h.x));
      b = add_rn(b, mul_rn(r.y, h.y));
      c = add_rn(c, mul_rn(g.x, h.x));
      d = add_rn(d, mul_rn(g.y, h.y));
    }
    acc[0] = e; acc[1] = a; acc[2] = b; acc[3] = c; acc[4] = d;
  } else if (which == 1) {
    double a = 0, b = 0;
    for (int i = 0; i < natoms; i++) { double2 x = qst[i]; a = add_rn(a, x.x); b = add_rn(b, x.y); }
    acc[5] = a; acc[6] = b;
  } else {
    double a = 0, b = 0;
    for (int i = 0; i < natoms; i++) { double2 g = gst[i]; a = add_rn(a, mul_rn(g.x, g.x)); b = add_rn(b, mul_rn(g.y, g.y)); }
    acc[7] = a; acc[8] = b;
  }
}

// hs = gs, ht = gt (src/qeq.F90:90-91)
__global__ void k_h_from_g(int natoms, const double2 *__restrict__ gst, double4 *__restrict__ hsq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  double2 g = gst[i];
  double4 h = hsq[i];
  h.x = g.x; h.y = g.y;
  hsq[i] = h;
}

// qs += lmin_s*hs ; qt += lmin_t*ht ; partial sums of qs, qt  (src/qeq.F90:133-141); lmin is real(4)
__global__ void __launch_bounds__(256) k_qupdate(int natoms, float lmin_s, float lmin_t, const double4 *__restrict__ hsq,
                                                 double2 *__restrict__ qst, double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double part[2] = {0.0, 0.0};
  if (i < natoms) {
    double4 h = hsq[i];
    double2 x = qst[i];
    x.x = add_rn(x.x, mul_rn((double)lmin_s, h.x));
    x.y = add_rn(x.y, mul_rn((double)lmin_t, h.y));
    qst[i] = x;
    part[0] = x.x;
    part[1] = x.y;
  }
  block_accumulate<2>(part, acc + 5);
}
// mu = ssum/tsum ; q = qs - mu*qt (src/qeq.F90:147-150); also saves Gold and clears the g.g accumulators
__global__ void k_qfinal(int natoms, const double2 *__restrict__ qst, double *__restrict__ q, double4 *__restrict__ hsq,
                         const double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  double mu = acc[5] / acc[6];
  double2 x = qst[i];
  double qi = sub_rn(x.x, mul_rn(mu, x.y));
  q[i] = qi;
  double4 h = hsq[i];
  h.z = qi;
  hsq[i] = h;
}
__global__ void k_roll_gnew(double *__restrict__ acc) {
  acc[9] = acc[7]; acc[10] = acc[8];
  acc[7] = 0.0; acc[8] = 0.0;
}
__global__ void k_clear_iter(double *__restrict__ acc) {
  for (int k = 0; k < 7; k++) acc[k] = 0.0;
}
// hs = gs + (Gnew/Gold)*hs (src/qeq.F90:160-161)
__global__ void k_hupdate(int natoms, const double2 *__restrict__ gst, double4 *__restrict__ hsq, const double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  double bs = acc[7] / acc[9], bt = acc[8] / acc[10];
  double2 g = gst[i];
  double4 h = hsq[i];
  h.x = add_rn(g.x, mul_rn(bs, h.x));
  h.y = add_rn(g.y, mul_rn(bt, h.y));
  hsq[i] = h;
}

}   // namespace rxg
