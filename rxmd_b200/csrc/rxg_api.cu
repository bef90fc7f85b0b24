// rxg_api.cu -- C-ABI entry points of librxmd_b200.so (see include/rxmd_b200.h) and the per-call orchestration.
// No CPU fallback exists: every entry point fails with RXG_ERR_CUDA when no sm_100 device is usable.
#include "rxg_common.cuh"
#include "rxg_halo_cells.cuh"
#include "rxg_lists_qeq.cuh"
#include "rxg_bonded.cuh"

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace rxg;

namespace {

template <typename T>
int dalloc(Ctx *c, T **p, size_t n) {
  RXG_CUDA(cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)));
  c->allocs.push_back((void *)*p);
  RXG_CUDA(cudaMemsetAsync(*p, 0, sizeof(T) * (n ? n : 1), c->st));
  return RXG_OK;
}

template <typename T>
int upload(Ctx *c, const T *host, size_t n, const T **dev) {
  T *d = nullptr;
  RXG_CUDA(cudaMalloc((void **)&d, sizeof(T) * (n ? n : 1)));
  if (n) RXG_CUDA(cudaMemcpy(d, host, sizeof(T) * n, cudaMemcpyHostToDevice));
  c->ff_allocs.push_back(d);
  *dev = d;
  return RXG_OK;
}

int setup_grid(Ctx *c, DevGrid &g, const int *cc, const double *cs, int L) {
  g.L = L;
  for (int a = 0; a < 3; a++) { g.nc[a] = cc[a]; g.dim[a] = cc[a] + 2 * L; g.cs[a] = cs[a]; }
  long long ncell = (long long)g.dim[0] * g.dim[1] * g.dim[2];
  if (ncell > 2000000000LL) { c->err = "cell grid too large"; return RXG_ERR_ARG; }
  g.ncell = (int)ncell;
  RXG_TRY(dalloc(c, &g.cell_of, c->NB));
  RXG_TRY(dalloc(c, &g.start, (size_t)g.ncell + 2));
  RXG_TRY(dalloc(c, &g.fill, (size_t)g.ncell + 1));
  RXG_TRY(dalloc(c, &g.order, c->NB));
  RXG_TRY(dalloc(c, &g.slot_of, c->NB));
  RXG_TRY(dalloc(c, &g.sorted, c->NB));
  return RXG_OK;
}

// host array with leading dimension NB (planes) -> device plane array; n leading entries of each of `planes` planes
int h2d_planes(Ctx *c, double *dev, const double *host, int planes, int n) {
  c->timers_ms[20] += 8.0 * (double)planes * (double)n;
  for (int p = 0; p < planes; p++)
    RXG_CUDA(cudaMemcpyAsync(dev + (size_t)p * c->NB, host + (size_t)p * c->NB, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
  return RXG_OK;
}
int d2h_planes(Ctx *c, double *host, const double *dev, int planes, int n) {
  c->timers_ms[21] += 8.0 * (double)planes * (double)n;
  for (int p = 0; p < planes; p++)
    RXG_CUDA(cudaMemcpyAsync(host + (size_t)p * c->NB, dev + (size_t)p * c->NB, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
  return RXG_OK;
}

struct Timer {
  Ctx *c;
  int slot;
  Timer(Ctx *c_, int s) : c(c_), slot(s) { cudaEventRecord(c->ev0, c->st); }
  ~Timer() {
    cudaEventRecord(c->ev1, c->st);
    cudaEventSynchronize(c->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->timers_ms[slot] += ms;
  }
};

// ---------------------------------------------------------------------------------------------------
// MPI_ALLREDUCE(SUM) of a few doubles (src/qeq.F90:107,129,144,357) as ncclAllReduce on the compute stream
int allreduce_acc(Ctx *c, int first, int count) {
  if (!c->comm) return RXG_OK;
  if (c->peer_all && count <= PW_ARW) {   // through the peer windows, summed in rank order (k_peer_allreduce)
    PeerAll pa;
    const int nr = (int)c->peer.size();
    for (int r = 0; r < nr; r++) pa.win[r] = c->peer[r];
    LAUNCH(c, k_peer_allreduce, 1, PW_MAXR, 0, pa, nr, c->box.myid, (size_t)PW_HDR + 12 * c->pw_cap, ++c->arseq, c->d_acc + first, count, c->d_flag + 3);
    return RXG_OK;
  }
  ncclResult_t r = nccl_api().AllReduce(c->d_acc + first, c->d_acc + first, count, ncclDouble, ncclSum, c->comm, c->st);
  if (r != ncclSuccess) { c->err = std::string("NCCL error: ") + nccl_api().GetErrorString(r); return RXG_ERR_NCCL; }
  c->nccl_msgs++;
  return RXG_OK;
}

// literal two-product CG of src/qeq.F90:86-166 (qeq_mode 1) and its serial-order variant (strict)
int qeq_cg_literal(Ctx *c, int nmax, int *iters) {
  const int n = c->natoms;
  const int rgrid = cdiv((long long)c->cp[6] * 32, 256);
  const bool strict = c->strict;
  const bool pq = c->cfg.isPQEq != 0;   // PQEq exists in the serial-order form only (validation); production PQEq is qeq_cg_single
  double4 *rowbuf = (double4 *)c->tmp;
  double *fpq = nullptr;
  if (pq) {
    if (!strict) { c->err = "the literal two-product CG of PQEq exists in the serial-order validation mode only"; return RXG_ERR_STATE; }
    fpq = c->qsl;   // free until the shell relaxation packs the final charges into it
    if (n > 0) LAUNCH(c, k_fpqeq_strict, cdiv(n, 64), 64, 0, c->gnb, n, c->rowbeg, c->rowend, c->col, c->val, c->sps, c->d_ff, fpq);
  }
  RXG_TRY(halo_qcopy(c, 1));
  auto harvest_grad = [&]() {   // call only after a stream sync
    if (c->grad_pending) {
      float ms = 0;
      cudaEventElapsedTime(&ms, c->evk[2], c->evk[3]);
      c->timers_ms[12] += ms;
      c->timers_ms[13] += 1;
      c->grad_pending = false;
    }
  };
  auto gradient = [&]() -> int {
    if (!strict) {
      cudaEventRecord(c->evk[2], c->st);
      LAUNCH(c, k_gradient, rgrid, 256, 0, c->gnb.order, c->cp[6], n, c->rowbeg, c->rowend, c->col, c->val, c->qst, c->itype, c->d_ff, c->gst, c->d_acc);
      cudaEventRecord(c->evk[3], c->st);
      c->grad_pending = true;
    } else {
      LAUNCH(c, k_rows_strict_grad, cdiv(n, 64), 64, 0, c->gnb.order, n, c->rowbeg, c->rowend, c->col, c->val, c->qst, c->itype, c->d_ff, c->gst, fpq);
      LAUNCH(c, k_seq_reduce, 1, 1, 0, 2, n, rowbuf, c->hsq, c->gst, c->qst, c->d_acc);
    }
    return allreduce_acc(c, 7, 2);
  };
  RXG_TRY(gradient());
  LAUNCH(c, k_h_from_g, cdiv(n, 256), 256, 0, n, c->gst, c->hsq);
  RXG_TRY(halo_qcopy(c, 2));
  double GEst2 = 1e99;
  int it;
  for (it = 0; it < nmax; it++) {
    LAUNCH(c, k_clear_iter, 1, 1, 0, c->d_acc);
    if (!strict) {
      cudaEventRecord(c->evk[0], c->st);
      LAUNCH(c, k_hsh, rgrid, 256, 0, c->gnb.order, c->cp[6], n, c->rowbeg, c->rowend, c->col, c->val, c->hsq, c->gst, c->itype, c->d_ff, c->d_acc);
      cudaEventRecord(c->evk[1], c->st);
    } else {
      if (pq) LAUNCH(c, k_rows_strict_hsh_pqeq, cdiv(n, 64), 64, 0, c->gnb, n, c->rowbeg, c->rowend, c->col, c->val, c->hsq, c->sps, c->d_ff, rowbuf);
      else LAUNCH(c, k_rows_strict_hsh, cdiv(n, 64), 64, 0, c->gnb.order, n, c->rowbeg, c->rowend, c->col, c->val, c->hsq, c->itype, c->d_ff, rowbuf);
      LAUNCH(c, k_seq_reduce, 1, 1, 0, 0, n, rowbuf, c->hsq, c->gst, c->qst, c->d_acc);
    }
    RXG_TRY(allreduce_acc(c, 0, 5));
    RXG_CUDA(cudaMemcpyAsync(c->h_acc, c->d_acc, sizeof(double) * 5, cudaMemcpyDeviceToHost, c->st));
    RXG_CUDA(cudaStreamSynchronize(c->st));
    if (!strict) {
      float ms = 0;
      cudaEventElapsedTime(&ms, c->evk[0], c->evk[1]);
      c->timers_ms[10] += ms;
      c->timers_ms[11] += 1;
      harvest_grad();
    }
    double GEst1 = c->h_acc[0];
    if (0.5 * (std::fabs(GEst2) + std::fabs(GEst1)) < c->cfg.QEq_tol) break;                    // src/qeq.F90:114
    if (std::fabs(GEst2) > 0.0 && std::fabs(GEst1 / GEst2 - 1.0) < c->cfg.QEq_tol) break;      // src/qeq.F90:115
    GEst2 = GEst1;
    float lmin_s = (float)(c->h_acc[3] / c->h_acc[1]);   // real(4) :: lmin, src/qeq.F90:23,133
    float lmin_t = (float)(c->h_acc[4] / c->h_acc[2]);
    LAUNCH(c, k_qupdate, cdiv(n, 256), 256, 0, n, lmin_s, lmin_t, c->hsq, c->qst, c->d_acc);
    if (strict) LAUNCH(c, k_seq_reduce, 1, 1, 0, 1, n, rowbuf, c->hsq, c->gst, c->qst, c->d_acc);
    RXG_TRY(allreduce_acc(c, 5, 2));
    LAUNCH(c, k_qfinal, cdiv(n, 256), 256, 0, n, c->qst, c->q, c->hsq, c->d_acc);
    RXG_TRY(halo_qcopy(c, 1));
    LAUNCH(c, k_roll_gnew, 1, 1, 0, c->d_acc);
    RXG_TRY(gradient());
    LAUNCH(c, k_hupdate, cdiv(n, 256), 256, 0, n, c->gst, c->hsq, c->d_acc);
    RXG_TRY(halo_qcopy(c, 2));
  }
  *iters = it;
  return RXG_OK;
}

// the CG's sparse product H.(x1,x2) -> four raw row sums per cell-order slot.  Default: the row-per-sub-warp, TMA-staged
// kernel (k_spmv_rows) whose launch shape follows the average row length (RXG_SPMV_SHAPE=8x8|4x16|2x32 overrides it).
// RXG_SPMV=items selects the cell-blocked kernel (k_spmv_items: union column stream, one gather of x per column of a row
// block) -- correct and tested, but measured slower on B200, so it stays an experiment.  RXG_SPMV_STAGE=0
// makes either kernel read the matrix straight from global memory (the path oversize rows take), RXG_SPMV_RING=<bytes>
// shrinks the ring of k_spmv_items.
template <int RG>
int spmv_items_launch(Ctx *c, int slot) {
  auto kern = k_spmv_items<RG>;
  if (c->spmv_grid[slot] == 0) {
    int sms = 0, optin = 0;
    RXG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->dev));
    RXG_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->dev));
    cudaFuncAttributes fa;
    RXG_CUDA(cudaFuncGetAttributes(&fa, kern));
    c->spmv_ring = ((optin - (int)fa.sharedSizeBytes - 1024) / 128) * 128;   // one CTA per SM takes all the shared memory
    if (c->spmv_ring_env > 0) c->spmv_ring = std::min(c->spmv_ring, c->spmv_ring_env);
    RXG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, c->spmv_ring));
    c->spmv_grid[slot] = std::max(1, sms);
  }
  const int grid = std::max(1, std::min(c->spmv_grid[slot], c->nitems));
  LAUNCH(c, kern, grid, (SI_CONS + 1) * 32, c->spmv_ring, c->items, c->nitems, c->ucol, c->umask, c->val, c->xs, (double4 *)c->tmp, c->d_acc, c->spmv_ring,
         c->spmv_stage);
  return RXG_OK;
}
// ---- window SpMV (k_spmv_win, the default): group size and shared-memory budget --------------------------------------------
// Cells per group: the largest G whose expected window (stencil runs lengthened by G-1 cells, at the average number of atoms
// per cell, +10 %) fits the per-CTA budget (RXG_WIN_SMEM, default 64 KB: three CTAs per SM).  The list build reports the
// largest window it saw (win_max); if that exceeds what a CTA can hold, the next build halves G, and a list whose windows do
// not fit at all is multiplied by k_spmv_rows.
template <int NW, int U, int MINB, int LPR = 32>
int spmv_win_launch(Ctx *c, int slot, bool *took) {
  auto kern = k_spmv_win<NW, U, MINB, LPR>;
  int optin = 0;
  RXG_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->dev));
  const int limit = optin - 4096;   // static shared memory of the kernel + slack
  int wcap = ((c->win_max + c->win_max / 64 + 16) + 7) & ~7;
  if (wcap > 32768) wcap = 32768;
  if (c->win_wcap_env > 0) wcap = std::min(wcap, c->win_wcap_env);   // (tests: forces the per-CTA fallback to global gathers)
  *took = false;
  if (c->win_max > 32768 || c->win_max * 16 > limit) {   // no group of this list fits: shrink the groups of the next build
    if (c->win_g_env <= 0 && c->win_g > 1) c->win_g = std::max(1, c->win_g / 2);
    return RXG_OK;
  }
  if (wcap * 16 > limit) wcap = (limit / 16) & ~7;
  const int smem = wcap * 16;
  if (smem > c->win_smem_set[slot]) {
    RXG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, limit));
    c->win_smem_set[slot] = limit;
  }
  const DevGrid &g = c->gnb;
  const int G = c->win_g_built;
  const int grid = g.nc[0] * g.nc[1] * cdiv(g.nc[2], G);
  LAUNCH(c, kern, grid, NW * 32, smem, g, c->nruns, G, c->stencil_reach, c->win_desc, c->rowoff, c->rowlen, c->col, c->col16, c->val, c->xs, (double4 *)c->tmp, c->d_acc, wcap);
  // windows much larger than the budget: fewer cells per group from the next build on
  if (c->win_g_env <= 0 && c->win_g > 1) {
    if (c->win_max * 16 > 2 * c->win_smem_target) c->win_g = std::max(1, c->win_g / 2);
    else if (c->win_max * 16 > c->win_smem_target + c->win_smem_target / 8) c->win_g = std::max(1, c->win_g - std::max(1, c->win_g / 6));
  }
  *took = true;
  c->win_launches++;
  return RXG_OK;
}
// part: 0 = all rows, 1 = interior row groups only, 2 = boundary row groups only (c->overlap; see k_group_class)
int spmv_launch(Ctx *c, int part = 0) {
  const int n = c->natoms, nt = c->cp[6];
  double4 *rowsum = (double4 *)c->tmp;
  if (nt <= 0) return RXG_OK;
  if (c->spmv_kind == 2 && c->win_built && part == 0 && c->nruns <= 4096) {
    bool took = false;
    // entries a lane keeps in flight per batch, from the average row length: a whole 10 A row (~400 entries) in one batch of
    // 32 x 13, short rows (sparse systems: ~120 entries) in one batch of 32 x 4
    int u = c->win_u;
    if (u == 0) {
      const double avgrow = (double)c->nnz_real / (double)std::max(n, 1);
      u = avgrow <= 140.0 ? 4 : (avgrow <= 270.0 ? 8 : 13);
    }
    if (c->win_nw == 4 && u == 13) RXG_TRY((spmv_win_launch<4, 13, 6>(c, 0, &took)));
    else if (c->win_nw == 4) RXG_TRY((spmv_win_launch<4, 8, 8>(c, 1, &took)));
    else if (u == 4 && c->win_lpr == 16) RXG_TRY((spmv_win_launch<8, 8, 4, 16>(c, 5, &took)));   // short rows: two per warp, 16 lanes x 8 entries each
    else if (u == 4) RXG_TRY((spmv_win_launch<8, 4, 6>(c, 2, &took)));
    else if (u == 13) RXG_TRY((spmv_win_launch<8, 13, 3>(c, 4, &took)));
    else RXG_TRY((spmv_win_launch<8, 8, 4>(c, 3, &took)));
    if (took) return RXG_OK;
  }
  c->rows_launches++;
  if (c->spmv_kind == 0) {
    if (c->nitems < 0) {   // first product after a list build: the item count left by the fill pass
      RXG_CUDA(cudaMemcpyAsync(c->h_int + 17, c->d_flag + 17, sizeof(int), cudaMemcpyDeviceToHost, c->st));
      RXG_CUDA(cudaStreamSynchronize(c->st));
      c->nitems = c->h_int[17];
    }
    if (c->nitems == 0 || part == 2) return RXG_OK;
    return c->spmv_rg == 4 ? spmv_items_launch<4>(c, 0) : spmv_items_launch<2>(c, 1);
  }
  const double avgrow = (double)c->nnz_real / (double)std::max(n, 1);
  int shape = c->spmv_shape;
  if (shape == 0) shape = avgrow <= 160.0 ? 1 : (avgrow > 440.0 ? 3 : 2);
  const int rows = shape == 1 ? 8 : (shape == 3 ? 2 : 4);
  const int *grp = nullptr;
  int grid = cdiv(nt, rows);
  if (part != 0) {
    if (c->grp_rows != rows) { c->err = "spmv_launch: row groups were classified for another launch shape"; return RXG_ERR_STATE; }
    grp = part == 1 ? c->grp_int : c->grp_bnd;
    grid = part == 1 ? c->ngrp_int : c->ngrp - c->ngrp_int;
    if (grid <= 0) return RXG_OK;
  }
  if (shape == 1)   // short rows (sparse systems such as the SiC nanoparticles, 117 entries): 8 rows per CTA, 8 lanes per row
    LAUNCH(c, (k_spmv_rows<8, 8, 256>), grid, 64, 0, c->gnb.order, nt, n, c->rowoff, c->rowbeg, c->rowend, c->col, c->val, c->xs, rowsum, c->d_acc, c->spmv_stage, grp);
  else if (shape == 3)   // 12.5 A lists (PQEq, 1060 entries): two long rows per CTA, a full warp per row
    LAUNCH(c, (k_spmv_rows<2, 32, 1216>), grid, 64, 0, c->gnb.order, nt, n, c->rowoff, c->rowbeg, c->rowend, c->col, c->val, c->xs, rowsum, c->d_acc, c->spmv_stage, grp);
  else   // 10 A lists: 4 rows per CTA, 16 lanes per row
    LAUNCH(c, (k_spmv_rows<4, 16>), grid, 64, 0, c->gnb.order, nt, n, c->rowoff, c->rowbeg, c->rowend, c->col, c->val, c->xs, rowsum, c->d_acc, c->spmv_stage, grp);
  return RXG_OK;
}

// single-pass CG (default): one sparse product per iteration, see rxg_lists_qeq.cuh.  Control flow (stop rule, real(4) step
// lengths) runs on the device (k_cg_ctrl); the host enqueues cg_batch (4) iterations at a time and reads the stop flag once per
// batch -- iterations enqueued past the stop return at their first instruction.
// classify the row groups of the launch shape in use (once per list build) -- only when the refresh has somewhere to go
int build_row_groups(Ctx *c) {
  c->overlap = false;
  if (!c->overlap_env || !c->comm || !c->peer_ok || c->halo_self || c->spmv_kind != 1 || c->cp[6] <= 0) return RXG_OK;   // (k_spmv_rows only)
  const int n = c->natoms, nt = c->cp[6];
  const double avgrow = (double)c->nnz_real / (double)std::max(n, 1);
  int shape = c->spmv_shape;
  if (shape == 0) shape = avgrow <= 160.0 ? 1 : (avgrow > 440.0 ? 3 : 2);
  const int rows = shape == 1 ? 8 : (shape == 3 ? 2 : 4);
  // reach of the stencil in cells (src/init.F90:538-592: cells whose nearest corner is within rctap)
  const int lay = c->stencil_reach;
  const int ngrp = cdiv(nt, rows);
  LAUNCH(c, k_group_class, cdiv(ngrp, 256), 256, 0, c->gnb, nt, rows, lay, c->grp_cls);
  RXG_TRY(ensure_blk(c, ngrp));
  RXG_TRY(device_scan<int>(c, c->grp_cls, ngrp, c->grp_off, c->d_blk, c->d_flag + 18));
  LAUNCH(c, k_group_lists, cdiv(ngrp, 256), 256, 0, ngrp, c->grp_cls, c->grp_off, c->grp_int, c->grp_bnd);
  RXG_CUDA(cudaMemcpyAsync(c->h_int + 18, c->d_flag + 18, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  c->ngrp = ngrp; c->ngrp_int = c->h_int[18]; c->grp_rows = rows;
  c->overlap = c->ngrp_int > 0;
  return RXG_OK;
}
// ghost refresh of (hs,ht) on the side stream, forked after the kernel that produced them; the caller joins (ev_join) before
// the boundary rows of the next sparse product
int refresh_h(Ctx *c) {
  if (!c->overlap) return halo_refresh(c, 3, 0);
  RXG_CUDA(cudaEventRecord(c->ev_fork, c->st));
  RXG_CUDA(cudaStreamWaitEvent(c->st2, c->ev_fork, 0));
  cudaStream_t s0 = c->st;
  c->st = c->st2;
  const int rc = halo_refresh(c, 3, 0);
  c->st = s0;
  RXG_TRY(rc);
  RXG_CUDA(cudaEventRecord(c->ev_join, c->st2));
  return RXG_OK;
}
int spmv_after_refresh(Ctx *c) {
  if (!c->overlap) return spmv_launch(c);
  RXG_TRY(spmv_launch(c, 1));                                  // interior rows: no ghost column
  RXG_CUDA(cudaStreamWaitEvent(c->st, c->ev_join, 0));         // the neighbours' (hs,ht) have arrived
  return spmv_launch(c, 2);
}

// A list that was filled without a count pass (build_pairlist, `capped`) reports its traps and statistics through device flags.
// Called right after a stream synchronisation that covered copy_capped_flags().  RXG_RETRY: a row outgrew its capacity, the
// caller rebuilds the list with the count pass and starts the solve over.
constexpr int RXG_RETRY = 100;
int copy_capped_flags(Ctx *c) {
  if (!c->list_capped) return RXG_OK;
  RXG_CUDA(cudaMemcpyAsync(c->h_int + 20, c->d_flag + 20, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaMemcpyAsync(c->h_int + 26, c->d_flag + 26, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaMemcpyAsync(c->h_int + 16, c->d_flag + 16, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaMemcpyAsync(c->h_int + 21, c->d_flag + 21, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaMemcpyAsync(c->h_acc + 33, c->d_acc + 33, sizeof(long long), cudaMemcpyDeviceToHost, c->st));
  if (c->comm) RXG_CUDA(cudaMemcpyAsync(c->h_acc + 26, c->d_acc + 26, sizeof(double), cudaMemcpyDeviceToHost, c->st));
  return RXG_OK;
}
int check_capped_flags(Ctx *c) {
  if (!c->list_capped) return RXG_OK;
  c->list_capped = false;   // checked once per list
  if (c->h_int[26] > c->cfg.maxneighbs10) {
    c->err = "ERROR: nbplist greater then MAXNEIGHBS10, value " + std::to_string(c->h_int[26]);
    return RXG_ERR_MAXNEIGHBS10;
  }
  // multi-rank: every rank must take the same decision -- the flags were summed over the ranks (qeq_cg_single)
  if (c->comm ? c->h_acc[26] > 0.5 : c->h_int[20] != 0) {
    // A row's count changes by the atoms that cross its 10 A sphere in one step: ~0.1-0.5 on average, Poisson-distributed, so
    // with a slack of s entries a row overflows with probability ~ lambda^(s+1)/(s+1)! -- times 10^6-10^7 rows per rank and
    // step.  Measured at 979 776 atoms per rank (RDX, ~500 K): slack 4 overflows in ~10 % of the steps on one rank and in
    // nearly every step on eight (every rank retries when one does).  The default slack is therefore 8; an overflow raises
    // it by 4 (up to 32) and keeps the count pass for the next few builds (4, 8, ... 256): a retry costs a list build and a
    // CG batch, a count pass 2 ms.
    c->caps_overflows++;
    if (getenv("RXG_CAP_DEBUG")) {
      int dbg[4] = {0, 0, 0, 0};
      cudaMemcpy(dbg, c->d_flag + 22, sizeof(dbg), cudaMemcpyDeviceToHost);
      double p3[3] = {0, 0, 0};
      int cell = -1;
      if (dbg[3] >= 0 && dbg[3] < c->NB) {
        for (int a = 0; a < 3; a++) cudaMemcpy(&p3[a], c->pos + (size_t)a * c->NB + dbg[3], sizeof(double), cudaMemcpyDeviceToHost);
        cudaMemcpy(&cell, c->gnb.cell_of + dbg[3], sizeof(int), cudaMemcpyDeviceToHost);
      }
      const DevGrid &g = c->gnb;
      fprintf(stderr, "[rxg rank %d] capped list overflow: gid %d row count %d capacity %d (atom %d of %d residents, default capacity %d, slack %d) pos %.6f %.6f %.6f cell (%d %d %d) of (%d %d %d)\n",
              c->box.myid, dbg[0], dbg[1], dbg[2], dbg[3], c->natoms, c->maxrow + 8, c->caps_slack, p3[0], p3[1], p3[2],
              cell < 0 ? -1 : cell / (g.dim[2] * g.dim[1]) - g.L, cell < 0 ? -1 : (cell / g.dim[2]) % g.dim[1] - g.L, cell < 0 ? -1 : cell % g.dim[2] - g.L, g.nc[0], g.nc[1], g.nc[2]);
    }
    c->caps_skip = c->caps_cooldown ? (4 << std::min(c->caps_fails, 6)) : 0;
    if (!c->caps_slack_env) c->caps_slack = std::min(32, c->caps_slack + 4);
    c->caps_fails++;
    return RXG_RETRY;
  }
  c->maxrow = c->h_int[16];
  c->nnz_real = *(long long *)(c->h_acc + 33);
  if (c->win_built) c->win_max = c->h_int[21];   // the largest window of this list: sizes the next launches
  return RXG_OK;
}

// The charge-independent part of FORCE (force_device stage 2: bond orders, bonded energy terms, ForceBondedTerms) on the
// high-priority side stream, beside the CG: the sparse products are HBM-bound and leave the fp64 pipes idle, the bonded terms
// are fp64- and latency-bound and leave HBM idle.  ev_bfork was recorded after the lists of this step were built.
int launch_bonded_side(Ctx *c) {
  RXG_CUDA(cudaStreamWaitEvent(c->st2, c->ev_bfork, 0));
  cudaStream_t s0 = c->st;
  const bool ph = c->ph_on;
  c->st = c->st2; c->ph_on = false;   // (the phase clock follows one stream)
  const int rc = force_device(c, true, 2);
  c->st = s0; c->ph_on = ph;
  RXG_TRY(rc);
  RXG_CUDA(cudaEventRecord(c->ev_bjoin, c->st2));
  c->bonded_stage = 2;
  return RXG_OK;
}

int qeq_cg_single(Ctx *c, int nmax, int *iters) {
  const int n = c->natoms;
  RXG_TRY(halo_refresh(c, 1, 0));   // ghost qs,qt (MODE_QCOPY1, src/qeq.F90:86)
  LAUNCH(c, k_to_slots, cdiv(c->cp[6], 256), 256, 0, c->cp[6], c->gnb.order, c->qst, c->xs);
  double4 *rowsum = (double4 *)c->tmp;
  const int dgrid = cdiv(c->cp[6], 256);
  const bool pq = c->cfg.isPQEq != 0;   // PQEq: same CG, other gradient constant and Est (rxg_pqeq.cuh)
  RXG_CUDA(cudaMemsetAsync(c->d_acc + ACC_DONE, 0, sizeof(double), c->st));
  RXG_TRY(spmv_launch(c));
  if (pq)
    LAUNCH(c, (k_cg_dots_pqeq<true>), dgrid, 256, 0, c->gnb.order, c->cp[6], n, rowsum, c->xs, c->q, c->gst, c->tst, c->ust, c->wst, c->itype, c->d_ff,
           c->prow, c->pcs, c->sps, c->d_acc);
  else
    LAUNCH(c, (k_cg_dots<true>), dgrid, 256, 0, c->gnb.order, c->cp[6], n, rowsum, c->xs, c->q, c->gst, c->tst, c->ust, c->wst, c->itype, c->d_ff, c->d_acc);
  RXG_TRY(allreduce_acc(c, 7, 2));
  if (c->list_capped && c->comm) {   // did ANY rank's capped list overflow?  (all ranks retry together or not at all)
    LAUNCH(c, k_flag_to_acc, 1, 1, 0, c->d_flag + 20, c->d_acc + 26);
    RXG_TRY(allreduce_acc(c, 26, 1));
  }
  LAUNCH(c, k_h_from_g2, cdiv(std::max(n, 1), 256), 256, 0, n, c->gst, c->hst, c->xs, c->gnb.slot_of, c->d_acc);
  RXG_TRY(build_row_groups(c));
  RXG_TRY(refresh_h(c));   // ghost hs,ht (MODE_QCOPY2, :93)
  int launched = 0, it = 0;
  bool done = false;
  while (!done && launched < nmax) {
    const int kb = std::min(c->cg_batch, nmax - launched);
    for (int j = 0; j < kb; j++) {
      cudaEventRecord(c->evs[2 * j], c->st);
      RXG_TRY(spmv_after_refresh(c));
      cudaEventRecord(c->evs[2 * j + 1], c->st);
      if (pq)
        LAUNCH(c, (k_cg_dots_pqeq<false>), dgrid, 256, 0, c->gnb.order, c->cp[6], n, rowsum, c->xs, c->q, c->gst, c->tst, c->ust, c->wst, c->itype, c->d_ff,
               c->prow, c->pcs, c->sps, c->d_acc);
      else
        LAUNCH(c, (k_cg_dots<false>), dgrid, 256, 0, c->gnb.order, c->cp[6], n, rowsum, c->xs, c->q, c->gst, c->tst, c->ust, c->wst, c->itype, c->d_ff, c->d_acc);
      RXG_TRY(allreduce_acc(c, 0, 5));   // (PQEq's ghost-column sums acc[12..15] are per-rank quantities and stay local)
      LAUNCH(c, k_cg_ctrl, 1, 1, 0, c->d_acc, c->cfg.QEq_tol, pq ? 1 : 0);
      LAUNCH(c, k_cg_update1, cdiv(std::max(n, 1), 256), 256, 0, n, c->hst, c->tst, c->ust, c->qst, c->gst, c->wst, c->d_acc);
      RXG_TRY(allreduce_acc(c, 5, 4));
      LAUNCH(c, k_cg_update2, cdiv(std::max(n, 1), 256), 256, 0, n, c->qst, c->gst, c->hst, c->xs, c->gnb.slot_of, c->q, c->d_acc);
      RXG_TRY(refresh_h(c));
    }
    // the first batch has been verified (a list without a count pass may have to be rebuilt): with the second one queued,
    // start the bonded part of FORCE on the side stream
    if (launched > 0 && c->bonded_stage == 1) RXG_TRY(launch_bonded_side(c));
    if (c->overlap) RXG_CUDA(cudaStreamWaitEvent(c->st, c->ev_join, 0));   // the batch's last refresh runs on the side stream
    RXG_TRY(copy_capped_flags(c));
    RXG_CUDA(cudaMemcpyAsync(c->h_acc + ACC_GEST2, c->d_acc + ACC_GEST2, sizeof(double) * 5, cudaMemcpyDeviceToHost, c->st));
    if (c->peer_ok) RXG_CUDA(cudaMemcpyAsync(c->h_int + 3, c->d_flag + 3, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    RXG_CUDA(cudaStreamSynchronize(c->st));
    if (c->peer_ok && c->h_int[3]) { c->err = "peer halo: a neighbour's ghost values did not arrive (timeout)"; return RXG_ERR_NCCL; }
    RXG_TRY(check_capped_flags(c));   // first batch after a capped list build: traps, overflow (-> RXG_RETRY), statistics
    done = c->h_acc[ACC_DONE] != 0.0;
    it = (int)c->h_acc[ACC_NITER];
    for (int j = 0; j < kb; j++) {   // sparse products that did work: iterations 0..it (the one that met the stop rule included)
      if (launched + j > it) break;
      float ms = 0;
      cudaEventElapsedTime(&ms, c->evs[2 * j], c->evs[2 * j + 1]);
      c->timers_ms[10] += ms;
      c->timers_ms[11] += 1;
    }
    launched += kb;
  }
  if (c->bonded_stage == 2) RXG_CUDA(cudaStreamWaitEvent(c->st, c->ev_bjoin, 0));   // the bonded terms read pos: before the round trips
  // the reference converts positions to normalised coordinates and back in every COPYATOMS call: QCOPY1 + QCOPY2
  // before the loop and two per completed iteration (src/qeq.F90:86,93,153,164); apply them in one launch
  if (c->cp[6] > 0)
    LAUNCH(c, k_roundtrip, cdiv(c->cp[6], 256), 256, 0, c->pos, c->NB, c->cp[6], make_boxdev(c->box), 2 + 2 * it);
  *iters = it;
  return RXG_OK;
}

// FORCE after a QEq that may already have run its charge-independent part (qeq_device / launch_bonded_side)
int force_staged(Ctx *c, bool reuse) {
  const int st = reuse ? c->bonded_stage : 0;
  c->bonded_stage = 0;
  if (st == 2) return force_device(c, true, 3);
  if (st == 1) { RXG_TRY(force_device(c, true, 2)); return force_device(c, true, 3); }
  return force_device(c, reuse, 0);
}

// subroutine QEq on device-resident state, reference src/qeq.F90:2-178
// `for_force`: build the halo with FORCE's width and the 10 A list with FORCE's predicate as well, so that the FORCE call of
// the same step can reuse them (device-resident stepping only; the per-call API stays literal)
int qeq_device(Ctx *c, bool for_force = false) {
  const int isQEq = c->cfg.isQEq;
  c->bonded_stage = 0;
  if (isQEq != 1 && isQEq != 2) return RXG_OK;
  const int n = c->natoms;
  const int nmax = (isQEq == 1) ? c->cfg.NMAXQEq : 1;
  int nprev = c->cp[6] > n ? c->cp[6] : n;
  if (c->caps_skip > 0) c->caps_skip--;
  const bool may_cap = c->caps_on && c->caps_valid && c->caps_skip == 0 && !c->strict && c->spmv_kind >= 1 && c->qeq_mode == 0;
  if (may_cap && n > 0) RXG_CUDA(cudaMemcpyAsync(c->q_save, c->q, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->st));
  if (nprev > 0)
    LAUNCH(c, k_qeq_init, cdiv(nprev, 256), 256, 0, n, nprev, c->q, c->qst, c->hsq, c->qsfp, c->qsfv, isQEq, c->cfg.Lex_fqs);
  double QCopyDr[3] = {c->ff.rctap / c->box.lata, c->ff.rctap / c->box.latb, c->ff.rctap / c->box.latc};
  c->lists_shared = false;
  if (for_force && !c->strict) {
    bool wide = true;
    for (int a = 0; a < 3; a++) wide = wide && (c->cfg.nmincell * c->box.lcsize[a] >= QCopyDr[a]);
    if (wide) {
      for (int a = 0; a < 3; a++) QCopyDr[a] = c->cfg.nmincell * c->box.lcsize[a];
      c->lists_shared = true;
    }
  }
  const bool pq = c->cfg.isPQEq != 0;
  phase_mark(c, 4 | PH_QEQ);    // COPYATOMS
  RXG_TRY(halo_copy(c, QCopyDr));
  if (pq) RXG_TRY(halo_refresh(c, 5, 0));   // ghost spos travels with MODE_COPY in the reference (src/comm.F90:129-131)
  if (c->cp[6] > 0) LAUNCH(c, k_types, cdiv(c->cp[6], 256), 256, 0, c->atype, c->cp[6], c->itype, c->gid);
  phase_mark(c, 3 | PH_QEQ);    // LINKEDLIST
  RXG_TRY(bin_grid(c, c->gnb));
  const int nt = c->cp[6];
  const int rgrid = cdiv((long long)nt * 32, 256);
  int it = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    phase_mark(c, 16 | PH_QEQ);   // qeq_initialize
    // first attempt: rows laid out from last step's counts, no count pass (build_pairlist `capped`); if a row outgrew its
    // capacity the solve that was started on the truncated list is thrown away and everything from here runs again with counts
    const bool allow_capped = may_cap && attempt == 0;
    if (c->lists_shared) RXG_TRY((build_pairlist<2>(c, !pq, allow_capped)));
    else RXG_TRY((build_pairlist<1>(c, !pq, allow_capped)));
    RXG_CUDA(cudaMemsetAsync(c->d_acc, 0, sizeof(double) * 24, c->st));
    if (pq && nt > 0) {   // qeq_initialize of src/pqeq.F90:262-365 on the compacted rows
      RXG_CUDA(cudaMemsetAsync(c->pcs, 0, sizeof(double) * (size_t)nt, c->st));
      RXG_CUDA(cudaMemsetAsync(c->d_flag + 6, 0, sizeof(int), c->st));
      LAUNCH(c, k_pack_sps, cdiv(nt, 256), 256, 0, nt, c->spos, c->NB, c->itype, c->gnb.slot_of, c->d_ff, c->sps);
      LAUNCH(c, k_pqeq_rows, rgrid, 256, 0, c->gnb, nt, n, c->rowbeg, c->rowend, c->col, c->val, c->sps, c->d_ff, c->prow, c->pcs, c->d_acc, c->d_flag + 6);
    }
    // FORCE beside the CG (same step, shared halo and list): its bonded cells and bonded list now (host synchronisations),
    // its charge-independent kernels on the side stream once the CG is under way (qeq_cg_single)
    if (c->bonded_env && c->lists_shared && !c->strict && !pq && c->qeq_mode == 0 && n > 0 && nt > 0) {
      if (c->bonded_stage == 0) {
        phase_mark(c, 0);
        RXG_TRY(force_device(c, true, 1));
        c->bonded_stage = 1;
      }
      RXG_CUDA(cudaEventRecord(c->ev_bfork, c->st));   // (again after a rebuild: Ehb walks the 10 A list)
    }
    phase_mark(c, 18 | PH_QEQ);   // the CG: get_hsh (+ get_gradient, which the single-pass CG folds into the same sparse product)
    int rc;
    if (c->strict || (!pq && c->qeq_mode == 1)) rc = qeq_cg_literal(c, nmax, &it);
    else rc = qeq_cg_single(c, nmax, &it);
    if (rc == RXG_OK && c->list_capped) {   // no CG batch synchronised (NMAXQEq = 0): look at the list's flags now
      RXG_TRY(copy_capped_flags(c));
      RXG_CUDA(cudaStreamSynchronize(c->st));
      rc = check_capped_flags(c);
    }
    if (rc == RXG_RETRY && attempt == 0) {
      RXG_CUDA(cudaMemcpyAsync(c->q, c->q_save, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->st));
      if (n > 0) LAUNCH(c, k_qeq_init, cdiv(n, 256), 256, 0, n, n, c->q, c->qst, c->hsq, c->qsfp, c->qsfv, isQEq, c->cfg.Lex_fqs);
      c->caps_valid = false;
      continue;
    }
    RXG_TRY(rc);
    break;
  }
  phase_mark(c, 0);
  c->ph_sec[24] += it;   // it_timer(24): QEq iterations (src/qeq.F90:172)
  if (pq && nt > 0) {   // update_shell_positions, src/pqeq.F90:171,187-259, with the final charges of residents and ghosts
    RXG_TRY(halo_refresh(c, 4, 0));
    LAUNCH(c, k_pack_qsl, cdiv(nt, 256), 256, 0, nt, c->q, c->gnb.slot_of, c->qsl);
    LAUNCH(c, k_shell_relax, rgrid, 256, 0, c->gnb, nt, n, c->rowbeg, c->rowend, c->col, c->sps, c->qsl, c->d_ff, c->cfg.isEfield, c->cfg.eFieldDir,
           c->cfg.eFieldStrength, c->spos, c->NB, c->d_flag + 6);
    RXG_CUDA(cudaMemcpyAsync(c->h_int + 6, c->d_flag + 6, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    if (c->peer_ok) RXG_CUDA(cudaMemcpyAsync(c->h_int + 3, c->d_flag + 3, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    RXG_CUDA(cudaStreamSynchronize(c->st));
    if (c->peer_ok && c->h_int[3]) { c->err = "peer halo: a neighbour's ghost values did not arrive (timeout)"; return RXG_ERR_NCCL; }
    c->pqeq_skips += c->h_int[6];
  }
  c->nstep_qeq = it;
  c->timers_ms[14] = (double)c->nnz_real;
  c->timers_ms[18] = (double)c->nnz;
  c->timers_ms[19] = (double)c->nunion;
  c->timers_ms[15] = n;
  c->timers_ms[16] = c->cp[6];
  c->timers_ms[17] += it;
  c->timers_ms[24] = (double)c->caps_overflows;
  return RXG_OK;
}

}   // namespace

// compact bond storage: 2 int + 16 double planes of `bond_cap` slots, grown on demand (contents need not survive: every
// FORCE rebuilds them)
namespace rxg {
int win_pick_group(Ctx *c) {
  if (c->win_g_env > 0) return c->win_g_env;
  const DevGrid &g = c->gnb;
  // atoms per cell where there are atoms: residents over resident cells (the ghost layers of the grid are mostly empty)
  const double per_cell = (double)std::max(c->natoms, 1) / (double)std::max(g.nc[0] * g.nc[1] * g.nc[2], 1);
  const int nr = (int)(c->h_runs.size() / 4);
  int best = 1;
  for (int G = 1; G <= std::min(32, g.nc[2]); G++) {
    long long cells = 0;
    for (int r = 0; r < nr; r++) cells += c->h_runs[4 * r + 3] - c->h_runs[4 * r + 2] + G;
    if (1.1 * per_cell * (double)cells * 16.0 <= (double)c->win_smem_target) best = G;
  }
  // even split of the z-column: the same number of groups, none of them a small remainder
  const int ngz = cdiv(g.nc[2], best);
  return cdiv(g.nc[2], ngz);
}
int ensure_bond_capacity(Ctx *c, long long need) {
  if (need <= c->bond_cap) return RXG_OK;
  const long long cap = need + need / 8 + 1024;
  auto re = [&](auto **p) -> int {
    if (*p) cudaFree(*p);
    RXG_CUDA(cudaMalloc((void **)p, sizeof(**p) * (size_t)cap));
    RXG_CUDA(cudaMemsetAsync(*p, 0, sizeof(**p) * (size_t)cap, c->st));
    return RXG_OK;
  };
  RXG_TRY(re(&c->nbrlist)); RXG_TRY(re(&c->nbrindx)); RXG_TRY(re(&c->bown));
  for (int k = 0; k < 4; k++) RXG_TRY(re(&c->BO[k]));
  for (int k = 0; k < 3; k++) { RXG_TRY(re(&c->dln[k])); RXG_TRY(re(&c->cB[k])); }
  RXG_TRY(re(&c->dBOp)); RXG_TRY(re(&c->A0)); RXG_TRY(re(&c->A1)); RXG_TRY(re(&c->A2)); RXG_TRY(re(&c->A3)); RXG_TRY(re(&c->cdslot));
  c->bond_cap = cap;
  return RXG_OK;
}
}   // namespace rxg

// ====================================================================================================
extern "C" {

int rxg_create(const rxg_config *cfg, rxg_handle *out) {
  if (!cfg || !out) return RXG_ERR_ARG;
  *out = nullptr;
  Ctx *c = new Ctx();
  c->cfg = *cfg;
  c->NB = cfg->nbuffer;
  c->MAXN = cfg->maxneighbs;
  c->dev = cfg->device;
  const char *so = getenv("RXG_STRICT_ORDER");
  c->strict = so && so[0] == '1';
  const char *nf = getenv("RXG_NO_FUSE");
  c->fuse = !(nf && nf[0] == '1');
  const char *fa = getenv("RXG_FUSE_API");
  c->fuse_api = fa && fa[0] == '1';
  const char *ov = getenv("RXG_OVERLAP");   // opt-in: measured slower at 2 GPUs
  c->overlap_env = ov && ov[0] == '1';
  const char *hf = getenv("RXG_HESS_FUSE");
  c->hess_fuse = hf && hf[0] == '1';
  const char *eo = getenv("RXG_EVAL_OCC");
  c->eval_occ = !(eo && eo[0] == '0');   // default on: measured 15.7 -> 14.6 ms per FORCE at 979 776 RDX atoms
  const char *sk = getenv("RXG_SPMV");
  // 2: k_spmv_win (default), 1: k_spmv_rows (round 1's kernel; also what lists with oversize windows fall back to),
  // 0: k_spmv_items (experiment)
  c->spmv_kind = (sk && std::string(sk) == "items") ? 0 : ((sk && std::string(sk) == "rows") ? 1 : 2);
  {
    const char *wg = getenv("RXG_WIN_G"), *ww = getenv("RXG_WIN_WARPS"), *wsm = getenv("RXG_WIN_SMEM");
    c->win_g_env = wg ? std::max(0, atoi(wg)) : 0;
    c->win_nw = ww ? atoi(ww) : 8;
    if (c->win_nw != 4) c->win_nw = 8;
    const char *wu = getenv("RXG_WIN_U");
    c->win_u = wu ? atoi(wu) : 0;   // 0: from the average row length (spmv_launch)
    if (c->win_u != 4 && c->win_u != 8 && c->win_u != 13) c->win_u = 0;
    c->win_smem_target = wsm ? std::max(4096, atoi(wsm)) : 64 * 1024;
    const char *wl = getenv("RXG_WIN_LPR");
    c->win_lpr = (wl && atoi(wl) == 16) ? 16 : 32;   // 16: two short rows per warp (experiment: slower on the SiC nanoparticles)
    const char *wra = getenv("RXG_WIN_RALIGN");
    if (wra && (atoi(wra) == 4 || atoi(wra) == 8 || atoi(wra) == 16)) c->win_ralign = atoi(wra);
    const char *wc = getenv("RXG_WIN_WCAP");
    c->win_wcap_env = wc ? atoi(wc) : 0;
  }
  const char *ss = getenv("RXG_SPMV_SHAPE");
  c->spmv_shape = !ss ? 0 : (std::string(ss) == "8x8" ? 1 : (std::string(ss) == "4x16" ? 2 : (std::string(ss) == "2x32" ? 3 : 0)));
  const char *sg = getenv("RXG_SPMV_STAGE");
  c->spmv_stage = !(sg && sg[0] == '0');
  const char *sl = getenv("RXG_SPMV_RING");   // ring size of k_spmv_items in bytes (default: all the shared memory of an SM)
  c->spmv_ring_env = sl ? atoi(sl) : 0;
  const char *tp = getenv("RXG_QEQ_TWOPASS");
  c->qeq_mode = (tp && tp[0] == '1') ? 1 : 0;
  *out = c;   // returned even on failure so that rxg_last_error can be read
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    c->err = "no CUDA device: librxmd_b200 has no CPU fallback";
    return RXG_ERR_CUDA;
  }
  RXG_CUDA(cudaSetDevice(c->dev));
  cudaDeviceProp prop;
  RXG_CUDA(cudaGetDeviceProperties(&prop, c->dev));
  if (prop.major != 10) {
    c->err = std::string("device '") + prop.name + "' is not sm_100: this library is built for B200 only";
    return RXG_ERR_CUDA;
  }
  RXG_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
  {   // side stream of the ghost refresh: highest priority, so that its small kernels are dispatched ahead of the queued CTAs
      // of the sparse product they overlap with
    int lo = 0, hi = 0;
    RXG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    RXG_CUDA(cudaStreamCreateWithPriority(&c->st2, cudaStreamNonBlocking, hi));
  }
  RXG_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  RXG_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  RXG_CUDA(cudaEventCreateWithFlags(&c->ev_bfork, cudaEventDisableTiming));
  RXG_CUDA(cudaEventCreateWithFlags(&c->ev_bjoin, cudaEventDisableTiming));
  // opt-in (measured: zero-sum): RXG_BONDED_OVERLAP=1 runs FORCE's charge-independent part beside the QEq CG
  { const char *bo = getenv("RXG_BONDED_OVERLAP"); c->bonded_env = bo && bo[0] == '1'; }
  RXG_CUDA(cudaEventCreate(&c->ev0));
  RXG_CUDA(cudaEventCreate(&c->ev1));
  RXG_CUDA(cudaEventCreate(&c->evm0));
  RXG_CUDA(cudaEventCreate(&c->evm1));
  for (int k = 0; k < 4; k++) RXG_CUDA(cudaEventCreate(&c->evk[k]));
  { const char *cb = getenv("RXG_CG_BATCH"); if (cb) c->cg_batch = std::max(1, std::min(16, atoi(cb))); }
  for (int k = 0; k < 32; k++) RXG_CUDA(cudaEventCreate(&c->evs[k]));
  const size_t NB = c->NB, NS = NB * (size_t)c->MAXN;
  RXG_TRY(dalloc(c, &c->pos, 3 * NB)); RXG_TRY(dalloc(c, &c->v, 3 * NB)); RXG_TRY(dalloc(c, &c->f, 3 * NB)); RXG_TRY(dalloc(c, &c->fsl, 3 * NB));
  RXG_TRY(dalloc(c, &c->atype, NB)); RXG_TRY(dalloc(c, &c->q, NB)); RXG_TRY(dalloc(c, &c->qsfp, NB)); RXG_TRY(dalloc(c, &c->qsfv, NB));
  RXG_TRY(dalloc(c, &c->qst, NB)); RXG_TRY(dalloc(c, &c->hsq, NB)); RXG_TRY(dalloc(c, &c->gst, NB));
  RXG_TRY(dalloc(c, &c->hst, NB)); RXG_TRY(dalloc(c, &c->tst, NB)); RXG_TRY(dalloc(c, &c->ust, NB)); RXG_TRY(dalloc(c, &c->wst, NB));
  RXG_TRY(dalloc(c, &c->sel, NB)); c->sel_cap = (int)NB;
  RXG_TRY(dalloc(c, &c->gsrc, NB));
  RXG_TRY(dalloc(c, &c->itype, NB)); RXG_TRY(dalloc(c, &c->gid, NB)); RXG_TRY(dalloc(c, &c->frcindx, NB));
  RXG_TRY(dalloc(c, &c->tmp, 15 * NB));
  if (cfg->isPQEq) {
    RXG_TRY(dalloc(c, &c->spos, 3 * NB)); RXG_TRY(dalloc(c, &c->sps, NB)); RXG_TRY(dalloc(c, &c->prow, NB));
    RXG_TRY(dalloc(c, &c->pcs, NB)); RXG_TRY(dalloc(c, &c->qsl, NB));
  }
  RXG_TRY(dalloc(c, &c->xs, NB)); RXG_TRY(dalloc(c, &c->pqa, NB)); RXG_TRY(dalloc(c, &c->pqs, NB)); RXG_TRY(dalloc(c, &c->tgs, NB)); RXG_TRY(dalloc(c, &c->gts, NB));
  { const char *eq = getenv("RXG_ENBOND_QUEUE"); c->enbond_queue = !(eq && eq[0] == '0'); }
  RXG_TRY(dalloc(c, &c->nbrcnt, NB)); RXG_TRY(dalloc(c, &c->nbrpad, NS)); RXG_TRY(dalloc(c, &c->bptr, NB + 2));
  RXG_TRY(dalloc(c, &c->rowoff, NB + 2)); RXG_TRY(dalloc(c, &c->rowbeg, NB + 2)); RXG_TRY(dalloc(c, &c->rowend, NB + 2));
  RXG_TRY(dalloc(c, &c->rowcnt, NB + 2)); RXG_TRY(dalloc(c, &c->ucnt, NB + 2)); RXG_TRY(dalloc(c, &c->uoff, NB + 2));
  RXG_TRY(dalloc(c, &c->rowlen, NB + 2));
  RXG_TRY(dalloc(c, &c->items, NB + 2));
  {   // row counts by global atom id, for list builds without a count pass (RXG_NOCOUNT=0 keeps the count pass)
    const char *nc = getenv("RXG_NOCOUNT");
    c->caps_on = !(nc && nc[0] == '0');
    const char *sl = getenv("RXG_CAP_SLACK");
    if (sl) { c->caps_slack = std::max(0, atoi(sl)); c->caps_slack_env = true; }
    const char *cd = getenv("RXG_CAP_COOLDOWN");
    c->caps_cooldown = !(cd && cd[0] == '0');
    if (c->caps_on) {
      size_t m = 1;
      while (m < 2 * NB) m <<= 1;
      c->cnt_mask = (unsigned)(m - 1);
      RXG_TRY(dalloc(c, &c->cnt_tab, m));
      RXG_TRY(dalloc(c, &c->q_save, NB));
    }
  }
  RXG_TRY(dalloc(c, &c->grp_cls, NB / 2 + 2)); RXG_TRY(dalloc(c, &c->grp_off, NB / 2 + 2));
  RXG_TRY(dalloc(c, &c->grp_int, NB / 2 + 2)); RXG_TRY(dalloc(c, &c->grp_bnd, NB / 2 + 2));
  RXG_TRY(ensure_bond_capacity(c, 8 * (long long)NB));
  RXG_TRY(dalloc(c, &c->delta, NB)); RXG_TRY(dalloc(c, &c->deltap1, NB)); RXG_TRY(dalloc(c, &c->deltap2, NB));
  RXG_TRY(dalloc(c, &c->nlp, NB)); RXG_TRY(dalloc(c, &c->dDlp, NB)); RXG_TRY(dalloc(c, &c->deltalp, NB));
  RXG_TRY(dalloc(c, &c->ccbnd, NB)); RXG_TRY(dalloc(c, &c->cdbnd, NB));
  RXG_TRY(dalloc(c, &c->s3, 3 * NB)); RXG_TRY(dalloc(c, &c->sbo, NB));
  RXG_TRY(dalloc(c, &c->d_acc, 128)); RXG_TRY(dalloc(c, &c->d_flag, 32));
  RXG_CUDA(cudaMallocHost((void **)&c->h_acc, sizeof(double) * 128));
  RXG_CUDA(cudaMallocHost((void **)&c->h_int, sizeof(int) * 32));
  RXG_CUDA(cudaMallocHost((void **)&c->h_cnt, sizeof(int) * (4 + 4 * 64)));
  RXG_TRY(dalloc(c, &c->d_cnt, 4 + 4 * 64));
  RXG_TRY(ensure_blk(c, NB));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  return RXG_OK;
}

int rxg_set_forcefield(rxg_handle h, const rxg_ff *ff) {
  Ctx *c = (Ctx *)h;
  if (!c || !ff) return RXG_ERR_ARG;
  RXG_CUDA(cudaSetDevice(c->dev));
  for (void *p : c->ff_allocs) cudaFree(p);
  c->ff_allocs.clear();
  DevFF &d = c->ff;
  d.nso = ff->nso; d.nboty = ff->nboty; d.nvaty = ff->nvaty; d.ntoty = ff->ntoty; d.nhbty = ff->nhbty; d.ntable = ff->ntable;
  d.vpar1 = ff->vpar1; d.vpar2 = ff->vpar2; d.cutoff_vpar30 = ff->cutoff_vpar30;
  d.rctap = ff->rctap; d.rctap2 = ff->rctap2; d.UDR = ff->UDR; d.UDRi = ff->UDRi;
  const size_t ns = ff->nso, nb = ff->nboty, nv = ff->nvaty, nt = ff->ntoty, nh = ff->nhbty;
#define UP(name, n) RXG_TRY(upload(c, ff->name, (n), &d.name))
  UP(Val, ns); UP(Valval, ns); UP(Valangle, ns); UP(Vale, ns); UP(mass, ns); UP(plp1, ns); UP(plp2, ns); UP(nlpopt, ns);
  UP(povun2, ns); UP(povun3, ns); UP(povun4, ns); UP(povun5, ns); UP(povun6, ns); UP(povun7, ns); UP(povun8, ns);
  UP(pval3, ns); UP(pval5, ns); UP(chi, ns); UP(eta, ns);
  UP(cBOp1, nb); UP(cBOp3, nb); UP(cBOp5, nb); UP(pbo2h, nb); UP(pbo4h, nb); UP(pbo6h, nb); UP(pbo2, nb); UP(pbo4, nb); UP(pbo6, nb);
  UP(swtch, 3 * nb); UP(rc2, nb); UP(pboc1, nb); UP(pboc3, nb); UP(pboc4, nb); UP(pboc5, nb); UP(ovc, nb); UP(v13cor, nb);
  UP(Desig, nb); UP(Depi, nb); UP(Depipi, nb); UP(pbe1, nb); UP(pbe2, nb); UP(povun1, nb);
  UP(theta00, nv); UP(pval1, nv); UP(pval2, nv); UP(pval4, nv); UP(pval6, nv); UP(pval7, nv); UP(pval8, nv); UP(pval9, nv); UP(pval10, nv);
  UP(ppen1, nv); UP(ppen2, nv); UP(ppen3, nv); UP(ppen4, nv); UP(pcoa1, nv); UP(pcoa2, nv); UP(pcoa3, nv); UP(pcoa4, nv);
  UP(ptor1, nt); UP(ptor2, nt); UP(ptor3, nt); UP(ptor4, nt); UP(V1, nt); UP(V2, nt); UP(V3, nt); UP(pcot1, nt); UP(pcot2, nt);
  UP(phb1, nh); UP(phb2, nh); UP(phb3, nh); UP(r0hb, nh);
  UP(inxn2, ns * ns); UP(inxn3, ns * ns * ns); UP(inxn3hb, ns * ns * ns); UP(inxn4, ns * ns * ns * ns);
  UP(TBL_Eclmb_QEq, (size_t)ff->ntable * nb);
#undef UP
  // interleave the vdW and Coulomb tables: one 32-byte record per (inxn, itb) node
  std::vector<double4> tnb((size_t)ff->ntable * nb);
  for (size_t x = 0; x < nb; x++)
    for (size_t i = 0; i < (size_t)ff->ntable; i++) {
      size_t k = 2 * (i + (size_t)ff->ntable * x);
      tnb[x * ff->ntable + i] = make_double4(ff->TBL_Evdw[k], ff->TBL_Evdw[k + 1], ff->TBL_Eclmb[k], ff->TBL_Eclmb[k + 1]);
    }
  RXG_TRY(upload(c, tnb.data(), tnb.size(), &d.TBL_nb));
  std::vector<double2> tq2((size_t)ff->ntable * nb);
  for (size_t x = 0; x < nb; x++)
    for (size_t i = 0; i < (size_t)ff->ntable; i++) {
      size_t k = i + (size_t)ff->ntable * x;
      tq2[k] = make_double2(ff->TBL_Eclmb_QEq[k], i + 1 < (size_t)ff->ntable ? ff->TBL_Eclmb_QEq[k + 1] : 0.0);
    }
  RXG_TRY(upload(c, tq2.data(), tq2.size(), &d.TBL_qeq2));
  d.ntype_pqeq = 0;
  d.pq_same = 0;
  d.isPolarizable = d.inxnpqeq = nullptr; d.Zpqeq = d.Kspqeq = nullptr; d.TBL_pcc = d.TBL_psc = d.TBL_pss = nullptr;
  if (c->cfg.isPQEq) {
    if (ff->ntype_pqeq < 1 || !ff->isPolarizable || !ff->Zpqeq || !ff->Kspqeq || !ff->inxnpqeq || !ff->TBL_Eclmb_pcc ||
        !ff->TBL_Eclmb_psc || !ff->TBL_Eclmb_pss) {
      c->err = "rxg_set_forcefield: isPQEq is set but the PQEq parameters of rxg_ff are missing";
      return RXG_ERR_ARG;
    }
    const size_t np = ff->ntype_pqeq, np2 = np * np, ntab = ff->ntable;
    d.ntype_pqeq = (int)np;
    RXG_TRY(upload(c, ff->isPolarizable, np, &d.isPolarizable)); RXG_TRY(upload(c, ff->inxnpqeq, np2, &d.inxnpqeq));
    RXG_TRY(upload(c, ff->Zpqeq, np, &d.Zpqeq)); RXG_TRY(upload(c, ff->Kspqeq, np, &d.Kspqeq));
    // TBL(ntype_pqeq2, NTABLE, 0:1) column-major -> {E(itb), E(itb+1), dE(itb), dE(itb+1)} per (inxn, itb)
    auto pack = [&](const double *T, const double4 **dst) -> int {
      std::vector<double4> t4(np2 * ntab);
      for (size_t x = 0; x < np2; x++)
        for (size_t i = 0; i < ntab; i++) {
          const size_t e0 = x + np2 * i, e1 = x + np2 * (i + 1), d0 = x + np2 * (i + ntab), d1 = x + np2 * (i + 1 + ntab);
          const bool last = i + 1 >= ntab;
          t4[x * ntab + i] = make_double4(T[e0], last ? 0.0 : T[e1], T[d0], last ? 0.0 : T[d1]);
        }
      return upload(c, t4.data(), t4.size(), dst);
    };
    RXG_TRY(pack(ff->TBL_Eclmb_pcc, &d.TBL_pcc)); RXG_TRY(pack(ff->TBL_Eclmb_psc, &d.TBL_psc)); RXG_TRY(pack(ff->TBL_Eclmb_pss, &d.TBL_pss));
    const size_t tb = sizeof(double) * np2 * ntab * 2;
    const char *ns_env = getenv("RXG_PQEQ_NO_TABLE_CACHE");
    d.pq_same = !(ns_env && ns_env[0] == '1') && memcmp(ff->TBL_Eclmb_pcc, ff->TBL_Eclmb_psc, tb) == 0 && memcmp(ff->TBL_Eclmb_psc, ff->TBL_Eclmb_pss, tb) == 0;
  }
  if (!c->d_ff) RXG_CUDA(cudaMalloc((void **)&c->d_ff, sizeof(DevFF)));
  RXG_CUDA(cudaMemcpy(c->d_ff, &d, sizeof(DevFF), cudaMemcpyHostToDevice));
  c->have_ff = true;
  return RXG_OK;
}

int rxg_set_box(rxg_handle h, const rxg_box *box) {
  Ctx *c = (Ctx *)h;
  if (!c || !box) return RXG_ERR_ARG;
  RXG_CUDA(cudaSetDevice(c->dev));
  if (c->have_box) { c->err = "rxg_set_box may be called once per handle"; return RXG_ERR_STATE; }
  c->box = *box;
  c->box.nbmesh = nullptr;
  RXG_TRY(setup_grid(c, c->gb, box->cc, box->lcsize, RXG_MAXLAYERS));
  RXG_TRY(setup_grid(c, c->gnb, box->nbcc, box->nblcsize, RXG_MAXLAYERS_NB));
  // stencil -> z-runs, preserving the reference's mesh order (src/init.F90:563-592: i, j outer, k inner)
  std::vector<int> runs;
  for (int m = 0; m < box->nbnmesh; m++) {
    int dx = box->nbmesh[3 * m], dy = box->nbmesh[3 * m + 1], dz = box->nbmesh[3 * m + 2];
    size_t nr = runs.size() / 4;
    if (nr && runs[4 * (nr - 1)] == dx && runs[4 * (nr - 1) + 1] == dy && runs[4 * (nr - 1) + 3] + 1 == dz)
      runs[4 * (nr - 1) + 3] = dz;
    else { runs.push_back(dx); runs.push_back(dy); runs.push_back(dz); runs.push_back(dz); }
  }
  c->nruns = (int)(runs.size() / 4);
  c->h_runs = runs;
  c->stencil_reach = 0;   // cells the stencil reaches along any axis: rows of cells this far inside the domain take no ghost column
  for (int m = 0; m < box->nbnmesh; m++)
    for (int a = 0; a < 3; a++) c->stencil_reach = std::max(c->stencil_reach, std::abs(box->nbmesh[3 * m + a]));
  RXG_CUDA(cudaMalloc((void **)&c->d_runs, sizeof(int) * (runs.size() + 4)));
  RXG_CUDA(cudaMemcpy(c->d_runs, runs.data(), sizeof(int) * runs.size(), cudaMemcpyHostToDevice));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  c->have_box = true;
  return RXG_OK;
}

int rxg_comm_unique_id(void *out128) {
  if (!out128) return RXG_ERR_ARG;
  ncclUniqueId id;
  std::string err;
  if (!nccl_api().load(err)) return RXG_ERR_NCCL;
  if (nccl_api().GetUniqueId(&id) != ncclSuccess) return RXG_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return RXG_OK;
}

int rxg_comm_init(rxg_handle h, int rank, int nranks, const void *id) {
  Ctx *c = (Ctx *)h;
  if (!c) return RXG_ERR_ARG;
  if (nranks == 1) return RXG_OK;
  if (!c->have_box) { c->err = "rxg_comm_init: rxg_set_box must be called first (the neighbour table decides which peer windows to open)"; return RXG_ERR_STATE; }
  if (!id) { c->err = "rxg_comm_init: null ncclUniqueId"; return RXG_ERR_ARG; }
  RXG_CUDA(cudaSetDevice(c->dev));
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  if (!nccl_api().load(c->err)) return RXG_ERR_NCCL;
  ncclResult_t r = nccl_api().CommInitRank(&c->comm, nranks, uid, rank);
  if (r != ncclSuccess) { c->err = std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(r); return RXG_ERR_NCCL; }
  // ---- peer windows for the per-iteration ghost refreshes (halo_refresh_peer).  Needs every neighbour on this node with
  // peer access; otherwise (or with RXG_PEER_HALO=0) the NCCL send/recv path stays.
  const char *ph = getenv("RXG_PEER_HALO");
  int want = !(ph && ph[0] == '0');
  c->peer.assign(nranks, nullptr);
  // The window layout must be the same on every rank (a rank addresses its neighbours' windows with its own offsets), but
  // NBUFFER is a per-rank setting: size the buffers from the largest one.  3 fields x NBUFFER atoms per buffer: no refresh can
  // exceed a buffer, so no rank can fail alone.
  int nbmax = c->NB;
  RXG_CUDA(cudaMemcpy(c->d_flag + 19, &nbmax, sizeof(int), cudaMemcpyHostToDevice));
  r = nccl_api().AllReduce(c->d_flag + 19, c->d_flag + 19, 1, ncclInt, ncclMax, c->comm, c->st);
  if (r != ncclSuccess) { c->err = std::string("ncclAllReduce: ") + nccl_api().GetErrorString(r); return RXG_ERR_NCCL; }
  RXG_CUDA(cudaStreamSynchronize(c->st));
  RXG_CUDA(cudaMemcpy(&nbmax, c->d_flag + 19, sizeof(int), cudaMemcpyDeviceToHost));
  c->pw_cap = (size_t)3 * (size_t)nbmax;
  const size_t wbytes = sizeof(double) * (PW_HDR + 12 * c->pw_cap + PW_AR_DOUBLES);
  int *d_ok = c->d_flag + 2;
  cudaIpcMemHandle_t *d_h = nullptr, *d_all = nullptr;
  std::vector<cudaIpcMemHandle_t> all(nranks);
  int ok = want;
  if (ok && cudaMalloc((void **)&c->pw, wbytes) != cudaSuccess) { cudaGetLastError(); c->pw = nullptr; ok = 0; }
  RXG_CUDA(cudaMalloc((void **)&d_h, sizeof(cudaIpcMemHandle_t)));
  RXG_CUDA(cudaMalloc((void **)&d_all, sizeof(cudaIpcMemHandle_t) * nranks));
  RXG_CUDA(cudaMalloc((void **)&c->d_pushcnt, 2 * sizeof(int)));
  RXG_CUDA(cudaMemset(c->d_pushcnt, 0, 2 * sizeof(int)));
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    RXG_CUDA(cudaMemset(c->pw, 0, sizeof(double) * PW_HDR));
    RXG_CUDA(cudaMemset(c->pw + PW_HDR + 12 * c->pw_cap, 0, sizeof(double) * PW_AR_DOUBLES));
    if (cudaIpcGetMemHandle(&mine, c->pw) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  RXG_CUDA(cudaMemcpy(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  RXG_CUDA(cudaDeviceSynchronize());
  r = nccl_api().AllGather(d_h, d_all, sizeof(cudaIpcMemHandle_t), ncclChar, c->comm, c->st);
  if (r != ncclSuccess) { c->err = std::string("ncclAllGather: ") + nccl_api().GetErrorString(r); return RXG_ERR_NCCL; }
  RXG_CUDA(cudaStreamSynchronize(c->st));
  RXG_CUDA(cudaMemcpy(all.data(), d_all, sizeof(cudaIpcMemHandle_t) * nranks, cudaMemcpyDeviceToHost));
  if (ok) {
    c->peer[rank] = c->pw;
    for (int t = 0; t < 6 && ok; t++) {
      const int nb = c->box.target_node[t];
      if (nb < 0 || nb >= nranks) { ok = 0; break; }
      if (c->peer[nb]) continue;
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[nb], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      c->peer[nb] = (double *)p;
    }
  }
  // the all-reduce through the windows needs every rank's window, not only the neighbours'
  const char *pa_env = getenv("RXG_PEER_ALLREDUCE");
  int all_ok = ok && nranks <= PW_MAXR && !(pa_env && pa_env[0] == '0');
  for (int q = 0; q < nranks && all_ok; q++) {
    if (c->peer[q]) continue;
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); all_ok = 0; break; }
    c->peer[q] = (double *)p;
  }
  // every rank must take the same path: all-reduce(min) of the local outcome
  int oks[2] = {ok, all_ok};
  RXG_CUDA(cudaMemcpy(d_ok, oks, 2 * sizeof(int), cudaMemcpyHostToDevice));
  r = nccl_api().AllReduce(d_ok, d_ok, 2, ncclInt, ncclMin, c->comm, c->st);
  if (r != ncclSuccess) { c->err = std::string("ncclAllReduce: ") + nccl_api().GetErrorString(r); return RXG_ERR_NCCL; }
  RXG_CUDA(cudaStreamSynchronize(c->st));
  RXG_CUDA(cudaMemcpy(oks, d_ok, 2 * sizeof(int), cudaMemcpyDeviceToHost));
  RXG_CUDA(cudaMemset(d_ok, 0, 2 * sizeof(int)));
  c->peer_ok = oks[0] != 0;
  c->peer_all = oks[0] != 0 && oks[1] != 0;
  cudaFree(d_h); cudaFree(d_all);
  return RXG_OK;
}

int rxg_comm_peer_halo(rxg_handle h) { return h ? (((Ctx *)h)->peer_ok ? 1 : 0) + (((Ctx *)h)->peer_all ? 2 : 0) : 0; }

int rxg_destroy(rxg_handle h) {
  Ctx *c = (Ctx *)h;
  if (!c) return RXG_OK;
  if (c->st) {
    cudaSetDevice(c->dev);
    cudaStreamSynchronize(c->st);
    if (c->st2) cudaStreamSynchronize(c->st2);
    if (c->comm && c->d_flag) {
      // Teardown is collective: a neighbour's kernels may still be storing into this rank's peer window.  An all-reduce in
      // stream order completes only after every rank has drained its own stream up to its rxg_destroy.
      nccl_api().AllReduce(c->d_flag + 19, c->d_flag + 19, 1, ncclInt, ncclSum, c->comm, c->st);
      cudaStreamSynchronize(c->st);
    }
    for (void *p : c->allocs) cudaFree(p);
    for (void *p : c->ff_allocs) cudaFree(p);
    for (void *p : {(void *)c->col, (void *)c->val, (void *)c->col16, (void *)c->win_desc, (void *)c->ucol, (void *)c->umask, (void *)c->d_blk, (void *)c->d_blk64, (void *)c->d_runs, (void *)c->d_ff})
      if (p) cudaFree(p);
    for (int k = 0; k < 2; k++) { if (c->sbuf[k]) cudaFree(c->sbuf[k]); if (c->rbuf[k]) cudaFree(c->rbuf[k]); }
    for (size_t r = 0; r < c->peer.size(); r++)
      if (c->peer[r] && c->peer[r] != c->pw) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->pw) cudaFree(c->pw);
    if (c->d_pushcnt) cudaFree(c->d_pushcnt);
    if (c->comm) nccl_api().CommDestroy(c->comm);
    if (c->wl) cudaFree(c->wl);
    for (void *p : {(void *)c->nbrlist, (void *)c->nbrindx, (void *)c->bown, (void *)c->BO[0], (void *)c->BO[1], (void *)c->BO[2], (void *)c->BO[3], (void *)c->dln[0],
                    (void *)c->dln[1], (void *)c->dln[2], (void *)c->cB[0], (void *)c->cB[1], (void *)c->cB[2], (void *)c->dBOp, (void *)c->A0,
                    (void *)c->A1, (void *)c->A2, (void *)c->A3, (void *)c->cdslot})
      if (p) cudaFree(p);
    if (c->h_acc) cudaFreeHost(c->h_acc);
    if (c->h_int) cudaFreeHost(c->h_int);
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    for (cudaEvent_t e : c->ph_ev) cudaEventDestroy(e);
    if (c->st2) { cudaStreamSynchronize(c->st2); cudaStreamDestroy(c->st2); }
    if (c->ev_bfork) cudaEventDestroy(c->ev_bfork);
    if (c->ev_bjoin) cudaEventDestroy(c->ev_bjoin);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->evs) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {c->evm0, c->evm1}) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(c->st);
  }
  delete c;
  return RXG_OK;
}

const char *rxg_last_error(rxg_handle h) { return h ? ((Ctx *)h)->err.c_str() : "null handle"; }

int rxg_hint(rxg_handle h, int flags) {
  Ctx *c = (Ctx *)h;
  if (!c) return RXG_ERR_ARG;
  c->hint = flags;
  return RXG_OK;
}

static int check_ready(Ctx *c, int natoms) {
  if (!c) return RXG_ERR_ARG;
  if (!c->have_ff || !c->have_box) { c->err = "rxg_set_forcefield / rxg_set_box not called"; return RXG_ERR_STATE; }
  if (natoms < 0 || natoms > c->NB) { c->err = "natoms outside [0, nbuffer]"; return RXG_ERR_ARG; }
  cudaSetDevice(c->dev);
  return RXG_OK;
}

int rxg_qeq(rxg_handle h, const int *natoms, const double *atype, double *pos, double *q, double *qsfp, double *qsfv,
            int *nstep_qeq) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, natoms ? *natoms : -1));
  const int n = *natoms;
  const int hint = (n == c->natoms) ? c->hint : 0;   // promises of rxg_hint hold for an unchanged atom count only
  c->hint = 0;
  c->natoms = n;
  if (!(hint & RXG_HINT_ATOMS_ON_DEVICE)) {
    RXG_TRY(h2d_planes(c, c->atype, atype, 1, n));
    RXG_TRY(h2d_planes(c, c->pos, pos, 3, n));
  }
  if (!(hint & RXG_HINT_Q_ON_DEVICE)) RXG_TRY(h2d_planes(c, c->q, q, 1, n));
  if (c->cfg.isQEq == 2) { RXG_TRY(h2d_planes(c, c->qsfp, qsfp, 1, n)); }
  {
    Timer t(c, 0);
    c->lists_shared = false;
    RXG_TRY(qeq_device(c, c->fuse_api));   // RXG_FUSE_API=1: build halo + list so that the next rxg_force may reuse them
  }
  RXG_TRY(d2h_planes(c, q, c->q, 1, n));   // resident charges (ghost entries of the host array are rewritten by every COPYATOMS)
  c->pos_deferred = (hint & RXG_HINT_DEFER_POS) != 0;
  if (!c->pos_deferred) RXG_TRY(d2h_planes(c, pos, c->pos, 3, n));
  if (c->cfg.isQEq == 1 && qsfp && qsfv) { RXG_TRY(d2h_planes(c, qsfp, c->qsfp, 1, n)); RXG_TRY(d2h_planes(c, qsfv, c->qsfv, 1, n)); }
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (nstep_qeq) *nstep_qeq = c->nstep_qeq;
  return RXG_OK;
}

int rxg_pqeq(rxg_handle h, const int *natoms, const double *atype, double *pos, double *q, double *spos, double *qsfp,
             double *qsfv, int *nstep_qeq) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, natoms ? *natoms : -1));
  if (!c->cfg.isPQEq || !spos) { c->err = "rxg_pqeq: the handle was created with isPQEq = 0, or spos is null"; return RXG_ERR_ARG; }
  RXG_TRY(h2d_planes(c, c->spos, spos, 3, *natoms));
  RXG_TRY(rxg_qeq(h, natoms, atype, pos, q, qsfp, qsfv, nstep_qeq));   // qeq_device runs the PQEq variant for this handle
  RXG_TRY(d2h_planes(c, spos, c->spos, 3, *natoms));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  return RXG_OK;
}

int rxg_spos_upload(rxg_handle h, int natoms, const double *spos) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, natoms));
  if (!c->cfg.isPQEq || !spos) { c->err = "rxg_spos_upload: isPQEq = 0 or null pointer"; return RXG_ERR_ARG; }
  RXG_TRY(h2d_planes(c, c->spos, spos, 3, natoms));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  return RXG_OK;
}
int rxg_spos_download(rxg_handle h, int natoms, double *spos) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, natoms));
  if (!c->cfg.isPQEq || !spos) { c->err = "rxg_spos_download: isPQEq = 0 or null pointer"; return RXG_ERR_ARG; }
  RXG_TRY(d2h_planes(c, spos, c->spos, 3, natoms));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  return RXG_OK;
}
long long rxg_pqeq_skips(rxg_handle h) { return h ? ((Ctx *)h)->pqeq_skips : 0; }

int rxg_force(rxg_handle h, const int *natoms, const double *atype, double *pos, double *f, const double *q, double *PE,
              double *astr) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, natoms ? *natoms : -1));
  const int n = *natoms;
  const int hint = (n == c->natoms) ? c->hint : 0;
  c->hint = 0;
  const bool atoms_here = (hint & RXG_HINT_ATOMS_ON_DEVICE) != 0;
  // RXG_FUSE_API=1: if this call follows rxg_qeq with bit-identical atoms (the host passes back what rxg_qeq returned),
  // the halo and the 10 A list of that QEq are reused exactly as rxg_md_run does; otherwise the literal path runs.
  // Without the host's promise (rxg_hint) the identity is verified on the device, at the price of uploading the atoms.
  bool reuse = false;
  if (atoms_here) {
    reuse = c->fuse_api && c->lists_shared && n > 0;
    if (reuse) c->timers_ms[22] += 1;
  } else if (c->fuse_api && c->lists_shared && n == c->natoms && n > 0) {
    double *stage = c->tmp;   // [4][NB] scratch
    for (int p = 0; p < 3; p++)
      RXG_CUDA(cudaMemcpyAsync(stage + (size_t)p * c->NB, pos + (size_t)p * c->NB, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
    RXG_CUDA(cudaMemcpyAsync(stage + 3 * (size_t)c->NB, atype, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
    c->timers_ms[20] += 32.0 * n;
    RXG_CUDA(cudaMemsetAsync(c->d_flag + 15, 0, sizeof(int), c->st));
    LAUNCH(c, k_same_atoms, cdiv(n, 256), 256, 0, n, c->NB, stage, c->pos, c->atype, c->d_flag + 15);
    RXG_CUDA(cudaMemcpyAsync(c->h_int + 15, c->d_flag + 15, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    RXG_CUDA(cudaStreamSynchronize(c->st));
    reuse = c->h_int[15] == 0;
    if (reuse) c->timers_ms[22] += 1;   // rxg_force calls that reused the halo and list of the preceding rxg_qeq
  }
  c->natoms = n;
  if (!reuse && !atoms_here) {
    RXG_TRY(h2d_planes(c, c->atype, atype, 1, n));
    RXG_TRY(h2d_planes(c, c->pos, pos, 3, n));
  }
  if (!(hint & RXG_HINT_Q_ON_DEVICE)) RXG_TRY(h2d_planes(c, c->q, q, 1, n));
  {
    Timer t(c, 1);
    RXG_TRY(force_staged(c, reuse));
  }
  c->lists_shared = false;
  RXG_TRY(d2h_planes(c, f, c->f, 3, n));
  RXG_TRY(d2h_planes(c, pos, c->pos, 3, n));   // also delivers the copy a hinted rxg_move / rxg_qeq deferred
  c->pos_deferred = false;
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (PE) for (int k = 0; k < 14; k++) PE[k] = c->PE[k];
  if (astr) for (int k = 0; k < 6; k++) astr[k] += c->astr[k];
  return RXG_OK;
}

int rxg_move(rxg_handle h, int *natoms, double *atype, double *pos, double *v, double *q, double *qs, double *qt, double *qsfp,
             double *qsfv) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, natoms ? *natoms : -1));
  const int n = *natoms;
  const int hint = c->hint;
  c->hint = 0;
  c->natoms = n;
  RXG_TRY(h2d_planes(c, c->atype, atype, 1, n));
  RXG_TRY(h2d_planes(c, c->pos, pos, 3, n));
  // Everything else travels only if an atom actually leaves this rank's domain or arrives (on one rank: wraps around the
  // box).  Most steps nothing does: then the reference's MODE_MOVE only perturbs pos by its normalise/de-normalise round
  // trip, and 21 of the 28 per-atom planes of PCIe traffic are saved.
  bool full = false;
  // RXG_HINT_CHARGES_STAY: the device's q / qs / qt are current and the host does not read them before the next QEq returns
  const bool charges_stay = (hint & RXG_HINT_CHARGES_STAY) && n == c->natoms_prev_move;
  c->lazy_upload = [&]() -> int {
    RXG_TRY(h2d_planes(c, c->v, v, 3, n));
    if (!charges_stay) RXG_TRY(h2d_planes(c, c->q, q, 1, n));
    RXG_TRY(h2d_planes(c, c->qsfp, qsfp, 1, n));
    RXG_TRY(h2d_planes(c, c->qsfv, qsfv, 1, n));
    // qs/qt travel with the atom in the reference (src/comm.F90:164-171); the device keeps them packed
    if (qs && qt && n > 0 && !charges_stay) {   // two planes into scratch, packed on the device (tmp is free until the compaction, which runs after the packs)
      double *sa = c->tmp + 13 * (size_t)c->NB, *sb = c->tmp + 14 * (size_t)c->NB;
      RXG_CUDA(cudaMemcpyAsync(sa, qs, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
      RXG_CUDA(cudaMemcpyAsync(sb, qt, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
      LAUNCH(c, k_planes_to_pairs, cdiv(n, 256), 256, 0, n, sa, sb, c->qst);
      c->timers_ms[20] += 16.0 * n;
    }
    full = true;
    return RXG_OK;
  };
  c->lists_shared = false;
  int rc;
  {
    Timer t(c, 2);
    phase_mark(c, 4);
    rc = halo_move(c);
    phase_mark(c, 0);
  }
  c->lazy_upload = nullptr;
  RXG_TRY(rc);
  const int m = c->natoms;
  // RXG_HINT_DEFER_POS: when nothing migrated, the only change is the ulp-level round trip of pos, which the next call
  // (hinted ATOMS_ON_DEVICE) consumes on the device and a later call hands back
  c->pos_deferred = (hint & RXG_HINT_DEFER_POS) && !full;
  if (!c->pos_deferred) RXG_TRY(d2h_planes(c, pos, c->pos, 3, m));
  if (full) {
    RXG_TRY(d2h_planes(c, atype, c->atype, 1, m));
    RXG_TRY(d2h_planes(c, v, c->v, 3, m));
    if (!charges_stay) RXG_TRY(d2h_planes(c, q, c->q, 1, m));
    RXG_TRY(d2h_planes(c, qsfp, c->qsfp, 1, m));
    RXG_TRY(d2h_planes(c, qsfv, c->qsfv, 1, m));
    if (qs && qt && m > 0 && !charges_stay) {
      double *sa = c->tmp, *sb = c->tmp + (size_t)c->NB;
      LAUNCH(c, k_pairs_to_planes, cdiv(m, 256), 256, 0, m, c->qst, sa, sb);
      RXG_CUDA(cudaMemcpyAsync(qs, sa, sizeof(double) * m, cudaMemcpyDeviceToHost, c->st));
      RXG_CUDA(cudaMemcpyAsync(qt, sb, sizeof(double) * m, cudaMemcpyDeviceToHost, c->st));
      c->timers_ms[21] += 16.0 * m;
    }
  }
  RXG_CUDA(cudaStreamSynchronize(c->st));
  *natoms = m;
  c->natoms_prev_move = m;
  return RXG_OK;
}

int rxg_fetch_bonds(rxg_handle h, int *nbrlist, double *BO0) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, 0));
  // reference layout: nbrlist(NBUFFER,0:MAXNEIGHBS) / BO(0,NBUFFER,MAXNEIGHBS), atom index fastest (src/init.F90:163,175)
  const int n = c->cp[6], MAXN = c->MAXN, NB = c->NB;
  std::vector<int> cnt(n), ptr(n + 1), lst((size_t)c->nbonds);
  std::vector<double> bo((size_t)c->nbonds);
  RXG_CUDA(cudaMemcpy(cnt.data(), c->nbrcnt, sizeof(int) * n, cudaMemcpyDeviceToHost));
  RXG_CUDA(cudaMemcpy(ptr.data(), c->bptr, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost));
  RXG_CUDA(cudaMemcpy(lst.data(), c->nbrlist, sizeof(int) * c->nbonds, cudaMemcpyDeviceToHost));
  RXG_CUDA(cudaMemcpy(bo.data(), c->BO[0], sizeof(double) * c->nbonds, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) {
    if (nbrlist) nbrlist[i] = cnt[i];
    for (int s = 0; s < cnt[i] && s < MAXN; s++) {
      if (nbrlist) nbrlist[(size_t)(s + 1) * NB + i] = lst[(size_t)ptr[i] + s] + 1;   // 1-based atom indices
      if (BO0) BO0[(size_t)s * NB + i] = bo[(size_t)ptr[i] + s];
    }
  }
  return RXG_OK;
}

// it_timer(1:30) of the reference (src/module.F90:215-217), in SECONDS (the reference keeps system_clock ticks and divides by
// the clock rate when it prints, src/main.F90:148-180); slot 24 is the QEq iteration count.
int rxg_it_timer(rxg_handle h, double *it_timer_sec) {
  Ctx *c = (Ctx *)h;
  if (!c || !it_timer_sec) return RXG_ERR_ARG;
  cudaSetDevice(c->dev);
  phase_harvest(c);
  for (int k = 1; k <= 30; k++) it_timer_sec[k - 1] = c->ph_sec[k];
  return RXG_OK;
}

int rxg_timers(rxg_handle h, double *t) {
  Ctx *c = (Ctx *)h;
  if (!c || !t) return RXG_ERR_ARG;
  c->timers_ms[25] = (double)c->win_launches; c->timers_ms[26] = (double)c->rows_launches;
  c->timers_ms[27] = (double)c->win_g_built; c->timers_ms[28] = (double)c->win_max;
  for (int k = 0; k < 30; k++) t[k] = c->timers_ms[k];
  return RXG_OK;
}

long long rxg_launch_count(rxg_handle h) { return h ? ((Ctx *)h)->launches : 0; }

// ---- device-resident stepping -----------------------------------------------------------------------------
int rxg_state_upload(rxg_handle h, int natoms, const double *atype, const double *pos, const double *v, const double *q,
                     const double *qsfp, const double *qsfv) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, natoms));
  c->natoms = natoms;
  for (int k = 0; k < 7; k++) c->cp[k] = natoms;
  RXG_TRY(h2d_planes(c, c->atype, atype, 1, natoms));
  RXG_TRY(h2d_planes(c, c->pos, pos, 3, natoms));
  if (v) RXG_TRY(h2d_planes(c, c->v, v, 3, natoms)); else RXG_CUDA(cudaMemsetAsync(c->v, 0, sizeof(double) * 3 * c->NB, c->st));
  if (q) RXG_TRY(h2d_planes(c, c->q, q, 1, natoms)); else RXG_CUDA(cudaMemsetAsync(c->q, 0, sizeof(double) * c->NB, c->st));
  if (qsfp) RXG_TRY(h2d_planes(c, c->qsfp, qsfp, 1, natoms)); else RXG_CUDA(cudaMemsetAsync(c->qsfp, 0, sizeof(double) * c->NB, c->st));
  if (qsfv) RXG_TRY(h2d_planes(c, c->qsfv, qsfv, 1, natoms)); else RXG_CUDA(cudaMemsetAsync(c->qsfv, 0, sizeof(double) * c->NB, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  return RXG_OK;
}

int rxg_md_prime(rxg_handle h) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, c ? c->natoms : -1));
  RXG_TRY(qeq_device(c));
  RXG_TRY(force_device(c));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  return RXG_OK;
}

int rxg_md_run(rxg_handle h, int nsteps, double dt, int qstep, double Lex_w2, int step0) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, c ? c->natoms : -1));
  if (qstep < 1) qstep = 1;
  cudaEventRecord(c->evm0, c->st);
  auto wall = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (int nstep = step0; nstep < step0 + nsteps; nstep++) {
    int n = c->natoms;
    // vkick(1) ; qsfv,qsfp ; pos += dt v      (src/main.F90:64-72)
    const bool ef = c->cfg.isEfield != 0;
    LAUNCH(c, k_md_first_half, cdiv(n, 256), 256, 0, n, c->NB, dt, Lex_w2, c->itype, c->d_ff, c->pos, c->v, c->f, c->q, c->qsfp, c->qsfv, !ef);
    if (ef) {   // LinearMomentum between the kick and the drift (src/main.F90:70-72)
      RXG_CUDA(cudaMemsetAsync(c->d_acc + 52, 0, sizeof(double) * 4, c->st));
      LAUNCH(c, k_momentum, cdiv(n, 256), 256, 0, n, c->NB, c->itype, c->d_ff, c->v, c->d_acc + 52);
      RXG_TRY(allreduce_acc(c, 52, 4));
      LAUNCH(c, k_sub_vcm_drift, cdiv(n, 256), 256, 0, n, c->NB, dt, c->v, c->pos, c->d_acc + 52, true);
    }
    double t0 = wall();
    phase_mark(c, 4);
    RXG_TRY(halo_move(c));                                   // :75
    phase_mark(c, 0);
    RXG_CUDA(cudaStreamSynchronize(c->st));
    double t1 = wall();
    c->lists_shared = false;
    if (nstep % qstep == 0) RXG_TRY(qeq_device(c, c->fuse)); // :77-83
    RXG_CUDA(cudaStreamSynchronize(c->st));
    double t2 = wall();
    RXG_TRY(force_staged(c, c->lists_shared));               // :84
    c->lists_shared = false;
    double t3 = wall();
    c->timers_ms[6] += t1 - t0; c->timers_ms[4] += t2 - t1; c->timers_ms[5] += t3 - t2;
    n = c->natoms;
    // kinetic stress, vkick(1), qsfv                        (src/main.F90:86-98)
    LAUNCH(c, k_md_second_half, cdiv(n, 256), 256, 0, n, c->NB, dt, Lex_w2, c->itype, c->d_ff, c->v, c->f, c->q, c->qsfp, c->qsfv, c->d_acc + 40);
  }
  cudaEventRecord(c->evm1, c->st);
  RXG_CUDA(cudaStreamSynchronize(c->st));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->evm0, c->evm1);
  c->timers_ms[3] += ms;
  c->timers_ms[7] += nsteps;
  return RXG_OK;
}

int rxg_state_download(rxg_handle h, int *natoms, double *atype, double *pos, double *v, double *f, double *q, double *qsfp,
                       double *qsfv) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, c ? c->natoms : -1));
  const int n = c->natoms;
  if (natoms) *natoms = n;
  if (atype) RXG_TRY(d2h_planes(c, atype, c->atype, 1, n));
  if (pos) RXG_TRY(d2h_planes(c, pos, c->pos, 3, n));
  if (v) RXG_TRY(d2h_planes(c, v, c->v, 3, n));
  if (f) RXG_TRY(d2h_planes(c, f, c->f, 3, n));
  if (q) RXG_TRY(d2h_planes(c, q, c->q, 1, n));
  if (qsfp) RXG_TRY(d2h_planes(c, qsfp, c->qsfp, 1, n));
  if (qsfv) RXG_TRY(d2h_planes(c, qsfv, c->qsfv, 1, n));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  return RXG_OK;
}

int rxg_md_observe(rxg_handle h, double *PE, double *KE, double *qsum, int *nstep_qeq, double *astr) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, c ? c->natoms : -1));
  const int n = c->natoms;
  RXG_CUDA(cudaMemsetAsync(c->d_acc + 48, 0, sizeof(double) * 2, c->st));
  LAUNCH(c, k_observe, cdiv(n, 256), 256, 0, n, c->NB, c->itype, c->d_ff, c->v, c->q, c->d_acc + 48);
  RXG_CUDA(cudaMemcpyAsync(c->h_acc + 40, c->d_acc + 40, sizeof(double) * 10, cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (PE) {
    c->PE[0] = 0;
    for (int k = 1; k < 14; k++) c->PE[0] += c->PE[k];   // PRINTE, src/main.F90:232
    for (int k = 0; k < 14; k++) PE[k] = c->PE[k];
  }
  if (KE) *KE = c->h_acc[48];
  if (qsum) *qsum = c->h_acc[49];
  if (nstep_qeq) *nstep_qeq = c->nstep_qeq;
  if (astr) for (int k = 0; k < 6; k++) astr[k] = c->astr[k] + c->h_acc[40 + k];
  return RXG_OK;
}

// ---- thermostat hooks: what the host's mdmode 4/5/7/8 code needs from resident velocities (src/main.F90:49-62,684-770)
int rxg_md_velocity_stats(rxg_handle h, double *stats) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, c ? c->natoms : -1));
  if (!stats) return RXG_ERR_ARG;
  const int n = c->natoms, ns = c->ff.nso < VSTAT_MAXT ? c->ff.nso : VSTAT_MAXT;
  double *d = c->tmp;   // scratch
  RXG_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * 6 * VSTAT_MAXT, c->st));
  if (n > 0) LAUNCH(c, k_types, cdiv(n, 256), 256, 0, c->atype, n, c->itype, c->gid);   // valid right after rxg_state_upload too
  if (n > 0) LAUNCH(c, k_velocity_stats, cdiv(n, 256), 256, 0, n, c->NB, c->itype, c->d_ff, c->v, d);
  RXG_CUDA(cudaMemcpyAsync(stats, d, sizeof(double) * 6 * ns, cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  return RXG_OK;
}
int rxg_md_velocity_affine(rxg_handle h, const double *scale, const double *shift) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, c ? c->natoms : -1));
  if (!scale || !shift) return RXG_ERR_ARG;
  const int n = c->natoms;
  double *d = c->tmp;
  RXG_CUDA(cudaMemcpyAsync(d, scale, sizeof(double) * c->ff.nso, cudaMemcpyHostToDevice, c->st));
  if (n > 0) LAUNCH(c, k_types, cdiv(n, 256), 256, 0, c->atype, n, c->itype, c->gid);
  if (n > 0) LAUNCH(c, k_velocity_affine, cdiv(n, 256), 256, 0, n, c->NB, c->itype, d, shift[0], shift[1], shift[2], c->v);
  RXG_CUDA(cudaStreamSynchronize(c->st));
  return RXG_OK;
}

// ---- kernel-level check and timing of the CG's sparse product (tests/test_gpu_spmv.py, tools/spmv_bench.py) -------
// x2: {x1, x2} per atom (residents and ghosts of the last QEq, interleaved, by ATOM index); out4: {a, b, ghost a, ghost b}
// per resident by atom index, a = sum_j H_ij x1_j etc. -- exactly what the production launch (spmv_launch, honouring the
// RXG_SPMV* switches) hands to k_cg_dots.  reps > 1 repeats the launch and returns the average CUDA-event time.
int rxg_debug_spmv(rxg_handle h, const double *x2, double *out4, int reps, double *ms_avg) {
  Ctx *c = (Ctx *)h;
  RXG_TRY(check_ready(c, c ? c->natoms : -1));
  if (!c->list_is_qeq || c->nnz <= 0) { c->err = "rxg_debug_spmv: no QEq matrix on the device (call rxg_qeq first)"; return RXG_ERR_STATE; }
  const int n = c->natoms, nt = c->cp[6];
  double2 *stage = (double2 *)(c->tmp + 8 * (size_t)c->NB);   // rowsum occupies tmp[0 .. 4*NB)
  if (x2) {
    RXG_CUDA(cudaMemcpyAsync(stage, x2, sizeof(double2) * nt, cudaMemcpyHostToDevice, c->st));
    LAUNCH(c, k_to_slots, cdiv(nt, 256), 256, 0, nt, c->gnb.order, stage, c->xs);
  }
  RXG_CUDA(cudaMemsetAsync(c->d_acc + ACC_DONE, 0, sizeof(double), c->st));
  RXG_CUDA(cudaMemsetAsync(c->tmp, 0, sizeof(double4) * nt, c->st));
  if (reps < 1) reps = 1;
  RXG_TRY(spmv_launch(c));   // warm-up / the checked launch
  cudaEventRecord(c->ev0, c->st);
  for (int r = 1; r < reps; r++) RXG_TRY(spmv_launch(c));
  cudaEventRecord(c->ev1, c->st);
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (ms_avg) {
    float ms = 0;
    if (reps > 1) cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    *ms_avg = reps > 1 ? ms / (reps - 1) : 0.0;
  }
  if (out4) {
    std::vector<double4> rs(nt);
    std::vector<int> ord(nt);
    RXG_CUDA(cudaMemcpy(rs.data(), c->tmp, sizeof(double4) * nt, cudaMemcpyDeviceToHost));
    RXG_CUDA(cudaMemcpy(ord.data(), c->gnb.order, sizeof(int) * nt, cudaMemcpyDeviceToHost));
    for (int s = 0; s < nt; s++) {
      const int i = ord[s];
      if (i < n) { out4[4 * (size_t)i] = rs[s].x; out4[4 * (size_t)i + 1] = rs[s].y; out4[4 * (size_t)i + 2] = rs[s].z; out4[4 * (size_t)i + 3] = rs[s].w; }
    }
  }
  return RXG_OK;
}

// ---- introspection for the parity tests ----------------------------------------------------------------------
int rxg_debug_fetch(rxg_handle h, const char *name, void *out, long long cap, long long *count) {
  Ctx *c = (Ctx *)h;
  if (!c || !name) return RXG_ERR_ARG;
  cudaSetDevice(c->dev);
  cudaStreamSynchronize(c->st);
  const std::string s(name);
  const long long n6 = c->cp[6] > c->natoms ? c->cp[6] : c->natoms, nat = c->natoms, NS = n6 * c->MAXN;
  const void *src = nullptr;
  long long bytes = 0, cnt = 0;
  std::vector<double> tmpd;
  auto dev = [&](const void *p, long long n, size_t esz) { src = p; cnt = n; bytes = n * (long long)esz; };
  if (s == "copyptr") {
    cnt = 7;
    if (count) *count = cnt;
    if (out) { if (cap < 28) return RXG_ERR_ARG; memcpy(out, c->cp, 28); }
    return RXG_OK;
  }
  if (s == "nnz") {
    if (count) *count = 1;
    if (out) { if (cap < 8) return RXG_ERR_ARG; memcpy(out, &c->nnz, 8); }
    return RXG_OK;
  }
  if (s == "win_meta") {   // {list carries the window stream, cells per group, stencil runs, groups, resident cells x/y/z, ghost layers}
    const DevGrid &g = c->gnb;
    const int G = std::max(1, c->win_g_built);
    const int meta[8] = {c->win_built ? 1 : 0, G, c->nruns, g.nc[0] * g.nc[1] * cdiv(g.nc[2], G), g.nc[0], g.nc[1], g.nc[2], g.L};
    if (count) *count = 8;
    if (out) { if (cap < 32) return RXG_ERR_ARG; memcpy(out, meta, 32); }
    return RXG_OK;
  }
  // 3-vector planes are returned compact [3][n6]
  if (s == "pos" || s == "f" || s == "v" || (s == "spos" && c->spos)) {
    const double *p = s == "pos" ? c->pos : (s == "f" ? c->f : (s == "v" ? c->v : c->spos));
    cnt = 3 * n6;
    if (count) *count = cnt;
    if (out) {
      if (cap < cnt * 8) return RXG_ERR_ARG;
      for (int a = 0; a < 3; a++)
        RXG_CUDA(cudaMemcpy((double *)out + a * n6, p + (size_t)a * c->NB, sizeof(double) * n6, cudaMemcpyDeviceToHost));
    }
    return RXG_OK;
  }
  if (s == "atype") dev(c->atype, n6, 8);
  else if (s == "q") dev(c->q, n6, 8);
  else if (s == "qst") dev(c->qst, 2 * n6, 8);
  else if (s == "gst") dev(c->gst, 2 * nat, 8);
  else if (s == "hsq") dev(c->hsq, 4 * n6, 8);
  else if (s == "frcindx") dev(c->frcindx, n6, 4);
  else if (s == "itype") dev(c->itype, n6, 4);
  else if (s == "gid") dev(c->gid, n6, 4);
  else if (s == "cell_bonded") dev(c->gb.cell_of, n6, 4);
  else if (s == "cell_nb") dev(c->gnb.cell_of, n6, 4);
  else if (s == "nbrcnt") dev(c->nbrcnt, n6, 4);
  else if (s == "nbrlist") dev(c->nbrlist, NS, 4);
  else if (s == "nbrindx") dev(c->nbrindx, NS, 4);
  else if (s == "rowbeg") dev(c->rowbeg, nat, 8);
  else if (s == "rowend") dev(c->rowend, nat, 8);
  else if (s == "col") dev(c->col, c->nnz, 4);   // converted to atom indices below
  else if (s == "col_slot") dev(c->col, c->nnz, 4);   // raw: slot | ghost bit
  else if (s == "col16" && c->col16 && c->win_built) dev(c->col16, c->nnz, 2);
  else if (s == "rowlen" && c->win_built) dev(c->rowlen, n6, 4);
  else if (s == "win_desc" && c->win_desc && c->win_built)
    dev(c->win_desc, 2LL * c->gnb.nc[0] * c->gnb.nc[1] * cdiv(c->gnb.nc[2], std::max(1, c->win_g_built)) * (c->nruns + 1), 4);
  else if (s == "cellstart_nb") dev(c->gnb.start, c->gnb.ncell + 1, 4);
  else if (s == "val") dev(c->val, c->nnz, 8);
  else if (s == "ucol") dev(c->ucol, c->nunion, 4);
  else if (s == "umask") dev(c->umask, c->nunion, 1);
  else if (s == "uoff") dev(c->uoff, n6 + 1, 8);
  else if (s == "rowoff") dev(c->rowoff, n6 + 1, 8);
  else if (s == "order_nb") dev(c->gnb.order, n6, 4);
  else if (s == "acc") dev(c->d_acc, 128, 8);
  else if (s == "BO0") dev(c->BO[0], NS, 8);
  else if (s == "BO1") dev(c->BO[1], NS, 8);
  else if (s == "BO2") dev(c->BO[2], NS, 8);
  else if (s == "BO3") dev(c->BO[3], NS, 8);
  else if (s == "dln_BOp1") dev(c->dln[0], NS, 8);
  else if (s == "dln_BOp2") dev(c->dln[1], NS, 8);
  else if (s == "dln_BOp3") dev(c->dln[2], NS, 8);
  else if (s == "dBOp") dev(c->dBOp, NS, 8);
  else if (s == "A0") dev(c->A0, NS, 8);
  else if (s == "A1") dev(c->A1, NS, 8);
  else if (s == "A2") dev(c->A2, NS, 8);
  else if (s == "A3") dev(c->A3, NS, 8);
  else if (s == "delta") dev(c->delta, n6, 8);
  else if (s == "deltap1") dev(c->deltap1, n6, 8);
  else if (s == "deltap2") dev(c->deltap2, n6, 8);
  else if (s == "nlp") dev(c->nlp, n6, 8);
  else if (s == "dDlp") dev(c->dDlp, n6, 8);
  else if (s == "deltalp") dev(c->deltalp, n6, 8);
  else if (s == "cdbnd") dev(c->cdbnd, n6, 8);
  else if (s == "ccbnd") dev(c->ccbnd, n6, 8);
  else if (s == "prow" && c->prow) dev(c->prow, 4 * n6, 8);   // by slot: {fpqeq, sum H Z, column sum, Z}
  else { c->err = "rxg_debug_fetch: unknown name " + s; return RXG_ERR_ARG; }
  static const char *slot_names[] = {"nbrlist", "nbrindx", "BO0", "BO1", "BO2", "BO3", "dln_BOp1", "dln_BOp2", "dln_BOp3", "dBOp", "A0", "A1", "A2", "A3"};
  bool is_slot = false;
  for (const char *nm : slot_names) is_slot = is_slot || s == nm;
  if (is_slot) {   // compact bond slots -> the reference's padded view [n6][MAXN]
    const size_t esz = (s == "nbrlist" || s == "nbrindx") ? 4 : 8;
    if (count) *count = NS;
    if (!out) return RXG_OK;
    if (cap < (long long)(NS * esz)) return RXG_ERR_ARG;
    std::vector<int> hc(n6), hp(n6 + 1);
    RXG_CUDA(cudaMemcpy(hc.data(), c->nbrcnt, sizeof(int) * n6, cudaMemcpyDeviceToHost));
    RXG_CUDA(cudaMemcpy(hp.data(), c->bptr, sizeof(int) * (n6 + 1), cudaMemcpyDeviceToHost));
    std::vector<char> hv((size_t)c->nbonds * esz + 8);
    if (c->nbonds) RXG_CUDA(cudaMemcpy(hv.data(), src, (size_t)c->nbonds * esz, cudaMemcpyDeviceToHost));
    memset(out, 0, NS * esz);
    for (long long i = 0; i < n6; i++)
      memcpy((char *)out + (size_t)i * c->MAXN * esz, hv.data() + (size_t)hp[i] * esz, (size_t)hc[i] * esz);
    return RXG_OK;
  }
  if (count) *count = cnt;
  if (out) {
    if (cap < bytes) return RXG_ERR_ARG;
    if (bytes) RXG_CUDA(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    if (s == "col" && cnt > 0) {   // slot | ghost-bit  ->  atom index, as the reference's nbplist holds it
      std::vector<int> ord(n6);
      RXG_CUDA(cudaMemcpy(ord.data(), c->gnb.order, sizeof(int) * n6, cudaMemcpyDeviceToHost));
      int *o = (int *)out;
      for (long long k = 0; k < cnt; k++) o[k] = ord[o[k] & 0x7fffffff];
    }
  }
  return RXG_OK;
}

}   // extern "C"
