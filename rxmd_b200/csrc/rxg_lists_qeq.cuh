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

// rows of the padded scratch list -> compact bond storage (slot of (i,s) = bptr[i] + s)
__global__ void k_compact_bonds(int ntot, int MAXN, const int *__restrict__ nbrcnt, const int *__restrict__ bptr,
                                const int *__restrict__ pad, int *__restrict__ lst, int *__restrict__ own) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int i = t / MAXN, s = t % MAXN;
  if (i >= ntot || s >= nbrcnt[i]) return;
  lst[(size_t)bptr[i] + s] = pad[(size_t)i * MAXN + s];
  own[(size_t)bptr[i] + s] = i;
}
// reverse index: nbrindx(i,i1) = j1 with nbrlist(j,j1) == i, src/main.F90:383-398
__global__ void k_nbrindx(int ntot, int MAXN, const int *__restrict__ nbrcnt, const int *__restrict__ bptr,
                          const int *__restrict__ nbrlist, int *__restrict__ nbrindx, int *__restrict__ bad) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int i = t / MAXN, i1 = t % MAXN;
  if (i >= ntot || i1 >= nbrcnt[i]) return;
  int j = nbrlist[(size_t)bptr[i] + i1];
  int found = -1;
  int nj = nbrcnt[j];
  const int *rowj = nbrlist + bptr[j];
  for (int j1 = 0; j1 < nj; j1++)
    if (rowj[j1] == i) found = j1;
  nbrindx[(size_t)bptr[i] + i1] = found;
  if (found < 0) atomicExch(bad, 1);
}

// ---------------------------------------------------------------------------------------------------
// A3 (+D1): 10 A pair list over the stencil runs.  One warp per RESIDENT CELL: every stencil run is loaded once and
// tested against all atoms of the cell (the cell's atoms sit in shared memory and are broadcast to the lanes).
//   QEQ=false: GetNonbondingPairList, fp64 dr2 <= rctap2 (src/main.F90:456-458)
//   QEQ=true : qeq_initialize, real(4) dr2 < rctap2 and hessian = lerp of TBL_Eclmb_QEq in r^2 (src/qeq.F90:222-240)
// FILL=false counts, FILL=true writes col (and val).
// Layout: rows lie in HBM in cell order (row of slot s starts at rowoff[s]); consumers address them by atom through
// rowbeg[i] / rowend[i].  A column entry is the neighbour's SLOT in the cell-sorted sequence (so that gathers from
// slot-ordered vectors are nearly contiguous), with bit 31 set when the neighbour is a ghost (Est weighting, Q3).
// Inside a row the entries keep the reference's order: stencil cells in mesh order, descending index inside a cell.
constexpr int PL_WARPS = 8, PL_MAXRUNS = 128;
// Row alignment in entries (run-time argument `ralign` of k_pairlist): 4 by default (32 B of val, 16 B of col: the
// bulk-copy granularity); 16 with the optional 16-bit column stream, whose 16-entry blocks must not straddle two rows.
constexpr int COL_GHOST = (int)0x80000000, COL_MASK = 0x7fffffff;
// MODE 0: FORCE list, 1: QEq list, 2: both at once (FORCE predicate; hessian = 0 where only the QEq predicate fails)
template <int MODE, bool FILL>
__global__ void __launch_bounds__(PL_WARPS * 32) k_pairlist(DevGrid g, const DevFF *__restrict__ ffp, const int *__restrict__ runs,
                                                            int nruns, int natoms, int ncell_res, int *__restrict__ slotcnt,
                                                            const long long *__restrict__ rowoff, long long *__restrict__ rowbeg,
                                                            long long *__restrict__ rowend, int *__restrict__ col,
                                                            double *__restrict__ val, int maxrow, int *__restrict__ ovf,
                                                            unsigned long long *__restrict__ nnz_real, int ralign) {
  __shared__ int sh_s[PL_WARPS][PL_MAXRUNS];       // first slot of each stencil run
  __shared__ int sh_p[PL_WARPS][PL_MAXRUNS + 1];   // exclusive prefix of the run lengths: position of each run in the flat candidate sequence
  __shared__ double4 sh_a[PL_WARPS][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rc = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (rc >= ncell_res) return;
  const DevFF &ff = *ffp;
  // resident cells only own rows (src/main.F90:438-440, src/qeq.F90:202-204)
  const int c3 = rc % g.nc[2], c2 = (rc / g.nc[2]) % g.nc[1], c1 = rc / (g.nc[2] * g.nc[1]);
  const int cid = ((c1 + g.L) * g.dim[1] + (c2 + g.L)) * g.dim[2] + (c3 + g.L);
  const int a0 = g.start[cid], a1 = g.start[cid + 1];
  if (a1 == a0) return;
  const float rctap2f = (float)ff.rctap2;
  for (int ab = a0; ab < a1; ab += 32) {
    const int nb = min(32, a1 - ab);
    const int myslot = ab + lane;
    double4 me = make_double4(0, 0, 0, 0);
    int mi = natoms;
    if (lane < nb) { me = g.sorted[myslot]; mi = rec_index(me.w); }
    sh_a[wid][lane] = me;
    const bool mine = mi < natoms;           // a ghost inside a resident cell owns no row (cannot happen after MOVE)
    long long mybase = (FILL && lane < nb) ? rowoff[myslot] : 0;
    int mycnt = 0;
    __syncwarp();
    for (int rb = 0; rb < nruns; rb += PL_MAXRUNS) {
      const int nr = min(PL_MAXRUNS, nruns - rb);
      // ---- bounds of this batch of stencil runs, lane parallel (independent loads), and their prefix sums: the runs are
      // walked as ONE flat candidate sequence, 32 candidates per step, so that short runs do not leave lanes idle
      int carry = 0;
      for (int r0 = 0; r0 < nr; r0 += 32) {
        const int r = r0 + lane;
        int s = 0, len = 0;
        if (r < nr) {
          const int4 rr = *reinterpret_cast<const int4 *>(runs + 4 * (rb + r));
          int b1 = c1 + rr.x, b2 = c2 + rr.y, z0 = c3 + rr.z, z1 = c3 + rr.w;
          if (b1 >= -g.L && b1 < g.nc[0] + g.L && b2 >= -g.L && b2 < g.nc[1] + g.L) {
            if (z0 < -g.L) z0 = -g.L;
            if (z1 >= g.nc[2] + g.L) z1 = g.nc[2] + g.L - 1;
            if (z1 >= z0) {
              int cbase = ((b1 + g.L) * g.dim[1] + (b2 + g.L)) * g.dim[2] + g.L;
              s = g.start[cbase + z0];
              len = g.start[cbase + z1 + 1] - s;
            }
          }
        }
        int inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int y = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += y;
        }
        if (r < nr) { sh_s[wid][r] = s; sh_p[wid][r] = carry + inc - len; }
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) sh_p[wid][nr] = carry;
      __syncwarp();
      const int total = carry;
      int r = 0;   // run that holds this lane's candidate; only ever moves forward
      for (int f0 = 0; f0 < total; f0 += 32) {
        const int fpos = f0 + lane;
        const bool have = fpos < total;
        if (have)
          while (fpos >= sh_p[wid][r + 1]) r++;
        const int cslot = have ? sh_s[wid][r] + (fpos - sh_p[wid][r]) : 0;
        double4 o = make_double4(0, 0, 0, 0);
        if (have) o = g.sorted[cslot];
        const int jt = rec_type(o.w);
        const int cval = cslot | ((have && rec_index(o.w) >= natoms) ? COL_GHOST : 0);
        for (int a = 0; a < nb; a++) {
          const double4 at = sh_a[wid][a];
          const double dr2 = dist2_rn(sub_rn(at.x, o.x), sub_rn(at.y, o.y), sub_rn(at.z, o.z));
          const bool acc = have && (cslot != ab + a) && (MODE == 1 ? ((float)dr2 < rctap2f) : (dr2 <= ff.rctap2));
          const unsigned mask = __ballot_sync(0xffffffffu, acc);
          if (FILL) {
            const long long wb = __shfl_sync(0xffffffffu, mybase + mycnt, a);
            if (acc) {
              const long long w = wb + __popc(mask & ((1u << lane) - 1u));
              col[w] = cval;
              if (MODE >= 1) {
                // the hessian lerp is evaluated by k_hessian over the compacted rows (full lanes); here only the
                // fp32-rounded r^2 and the bond type are parked in the 8 bytes of the value slot
                int inxn = ff.inxn2[(rec_type(at.w) - 1) + ff.nso * (jt - 1)];
                if (!(MODE == 1 || (float)dr2 < rctap2f)) inxn = 0;
                val[w] = __hiloint2double(inxn, __float_as_int((float)dr2));
              }
            }
          }
          if (lane == a) mycnt += __popc(mask);
        }
      }
      __syncwarp();
    }
    if (!FILL) {   // exact entry count (without row padding): the algorithmic-bytes figure of the roofline uses it;
                   // longest row: picks the SpMV launch shape
      const int real = __reduce_add_sync(0xffffffffu, (lane < nb && mine) ? mycnt : 0);
      const int longest = __reduce_max_sync(0xffffffffu, (lane < nb && mine) ? mycnt : 0);
      if (lane == 0 && real) { atomicAdd(nnz_real, (unsigned long long)real); atomicMax(ovf + 16, longest); }
    }
    if (lane < nb && mine) {
      if (!FILL) {
        slotcnt[myslot] = (mycnt + ralign - 1) & ~(ralign - 1);
        if (mycnt > maxrow) atomicMax(ovf, mycnt);
      } else {
        rowbeg[mi] = mybase;
        rowend[mi] = mybase + mycnt;
        // row padding: hessian 0, column = the row's last real column (written by some lane of this warp before the
        // __syncwarp above), so that the padding does not stretch the 16-bit offsets of its block (k_col16)
        const int padcol = mycnt > 0 ? __ldcg(col + mybase + mycnt - 1) : myslot;
        for (int p = mycnt; p < ((mycnt + ralign - 1) & ~(ralign - 1)); p++) {
          col[mybase + p] = padcol;
          if (MODE >= 1) val[mybase + p] = 0.0;
        }
      }
    }
    __syncwarp();
  }
}


// D1: hessian(j1,i) = (1-drtb)*TBL_Eclmb_QEq(itb,inxn) + drtb*TBL_Eclmb_QEq(itb+1,inxn) with real(4) dr2 (src/qeq.F90:234-240).
// One warp per row, lanes stride over the compacted entries; reads {float r^2, inxn} parked by k_pairlist.
__global__ void __launch_bounds__(256) k_hessian(long long nnz, const DevFF *__restrict__ ffp, double *__restrict__ val) {
  // flat over the padded entry range: row padding carries inxn = 0 (written by k_pairlist), which yields 0
  const DevFF &ff = *ffp;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    const double packed = val[k];
    const int inxn = __double2hiint(packed);
    const double d2 = (double)__int_as_float(__double2loint(packed));   // real(4) dr2 promoted back
    const int itb = (int)mul_rn(d2, ff.UDRi);
    const double drtb = mul_rn(sub_rn(d2, mul_rn((double)itb, ff.UDR)), ff.UDRi);
    double h = 0.0;
    if (inxn > 0 && itb >= 1 && itb < ff.ntable) {
      const double2 T = ff.TBL_qeq2[(size_t)(inxn - 1) * ff.ntable + (itb - 1)];   // {T(itb), T(itb+1)}: one 16-byte gather
      h = add_rn(mul_rn(sub_rn(1.0, drtb), T.x), mul_rn(drtb, T.y));
    }
    val[k] = h;
  }
}

// 16-bit column stream for the CG SpMV: per 16-entry block of the (padded) entry sequence one 32-bit base = the smallest
// slot of the block, per entry the 15-bit offset from it, bit 15 = ghost column.  Rows start on 16-entry boundaries, so
// the entries of a block belong to one row and come from one or two neighbouring stencil runs: offsets stay far below
// 32768; *ovf is raised otherwise and the 32-bit stream is used.  Canonical CSR is 12 B per entry; this
// stream moves 10.25 B.  OPT-IN (RXG_COL16=1): measured on B200 at 979 776 atoms it cuts the SpMV's DRAM bytes by 14 %
// (4.69 -> 4.07 GB) but its time by 1 % only (1.033 -> 1.019 ms) -- the kernel is bound by L1 data-pipe wavefronts
// (74 % of peak: the 16-byte gathers of x take 7.5 wavefronts per warp request), not by HBM -- and the conversion pass
// costs 0.4 ms per step, so the default stays the 32-bit stream.
__global__ void __launch_bounds__(256) k_col16(long long nnz, const int *__restrict__ col, unsigned short *__restrict__ col16,
                                               int *__restrict__ cbase, int *__restrict__ ovf) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {   // nnz is a multiple of 16
    const int v = col[k];
    const int slot = v & COL_MASK;
    int m = slot;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o, 16));
    const int off = slot - m;
    if (off > 0x7fff && *(volatile int *)ovf == 0) atomicExch(ovf, 1);
    col16[k] = (unsigned short)((off & 0x7fff) | (v < 0 ? 0x8000 : 0));
    if ((k & 15) == 0) cbase[k >> 4] = m;
  }
}

int ensure_bond_capacity(Ctx *c, long long need);   // rxg_api.cu

inline int build_nbrlist(Ctx *c) {
  const int n = c->cp[6];
  RXG_CUDA(cudaMemsetAsync(c->d_flag, 0, 2 * sizeof(int), c->st));
  RXG_CUDA(cudaMemsetAsync(c->nbrcnt, 0, sizeof(int) * n, c->st));
  LAUNCH(c, k_nbrlist, cdiv(n, 128), 128, 0, c->gb, c->d_ff, n, c->cfg.nmincell, c->MAXN, c->nbrcnt, c->nbrpad, c->d_flag);
  RXG_TRY(ensure_blk(c, n));
  RXG_TRY(device_scan<int>(c, c->nbrcnt, n, c->bptr, c->d_blk, c->d_flag + 2));
  RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (c->h_int[0] > c->MAXN) {
    c->err = "ERROR: overflow of max # in neighbor list, " + std::to_string(c->h_int[0]);
    return RXG_ERR_MAXNEIGHBS;
  }
  c->nbonds = c->h_int[2];
  RXG_TRY(ensure_bond_capacity(c, c->nbonds));
  LAUNCH(c, k_compact_bonds, cdiv((long long)n * c->MAXN, 256), 256, 0, n, c->MAXN, c->nbrcnt, c->bptr, c->nbrpad, c->nbrlist, c->bown);
  LAUNCH(c, k_nbrindx, cdiv((long long)n * c->MAXN, 256), 256, 0, n, c->MAXN, c->nbrcnt, c->bptr, c->nbrlist, c->nbrindx, c->d_flag + 1);
  RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (c->h_int[1]) {
    c->err = "ERROR: inconsistency between nbrlist and nbrindx found";
    return RXG_ERR_STATE;
  }
  return RXG_OK;
}

// `hessian`: run k_hessian over the parked (r^2, type) pairs (QEq); PQEq fills `val` itself (k_pqeq_rows)
template <int MODE>
int build_pairlist(Ctx *c, bool hessian = true) {
  const int n = c->natoms, nt = c->cp[6];
  RXG_CUDA(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->st));
  RXG_CUDA(cudaMemsetAsync(c->d_flag + 16, 0, sizeof(int), c->st));
  RXG_CUDA(cudaMemsetAsync(c->d_acc + 33, 0, sizeof(double), c->st));
  RXG_CUDA(cudaMemsetAsync(c->rowcnt, 0, sizeof(int) * (size_t)(nt + 1), c->st));
  RXG_CUDA(cudaMemsetAsync(c->rowbeg, 0, sizeof(long long) * (size_t)n, c->st));
  RXG_CUDA(cudaMemsetAsync(c->rowend, 0, sizeof(long long) * (size_t)n, c->st));
  const int ncell_res = c->gnb.nc[0] * c->gnb.nc[1] * c->gnb.nc[2];
  const int grid = cdiv((long long)ncell_res * 32, PL_WARPS * 32);
  const bool want16 = MODE >= 1 && !c->strict && c->use_col16;
  const int ralign = want16 ? 16 : 4;
  LAUNCH(c, (k_pairlist<MODE, false>), grid, PL_WARPS * 32, 0, c->gnb, c->d_ff, c->d_runs, c->nruns, n, ncell_res, c->rowcnt,
         c->rowoff, c->rowbeg, c->rowend, c->col, c->val, c->cfg.maxneighbs10, c->d_flag, (unsigned long long *)(c->d_acc + 33), ralign);
  RXG_TRY(ensure_blk(c, nt));
  RXG_TRY(device_scan<long long>(c, c->rowcnt, nt, c->rowoff, c->d_blk64, (long long *)(c->d_acc + 32)));
  RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaMemcpyAsync(c->h_int + 16, c->d_flag + 16, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaMemcpyAsync(c->h_acc + 32, c->d_acc + 32, 2 * sizeof(long long), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (c->h_int[0] > c->cfg.maxneighbs10) {
    c->err = "ERROR: nbplist greater then MAXNEIGHBS10, value " + std::to_string(c->h_int[0]);
    return RXG_ERR_MAXNEIGHBS10;
  }
  long long nnz = *(long long *)(c->h_acc + 32);
  if (nnz > c->nnz_cap) {
    if (c->col) cudaFree(c->col);
    if (c->val) cudaFree(c->val);
    if (c->col16) cudaFree(c->col16);
    if (c->cbase) cudaFree(c->cbase);
    c->nnz_cap = nnz + nnz / 16 + 1024;
    RXG_CUDA(cudaMalloc(&c->col, sizeof(int) * c->nnz_cap));
    RXG_CUDA(cudaMalloc(&c->val, sizeof(double) * c->nnz_cap));
    RXG_CUDA(cudaMalloc(&c->col16, sizeof(unsigned short) * c->nnz_cap));
    RXG_CUDA(cudaMalloc(&c->cbase, sizeof(int) * (c->nnz_cap / 16 + 2)));
  }
  c->nnz = nnz;
  c->maxrow = c->h_int[16];
  c->nnz_real = *(long long *)(c->h_acc + 33);
  c->list_is_qeq = MODE >= 1;
  LAUNCH(c, (k_pairlist<MODE, true>), grid, PL_WARPS * 32, 0, c->gnb, c->d_ff, c->d_runs, c->nruns, n, ncell_res, c->rowcnt,
         c->rowoff, c->rowbeg, c->rowend, c->col, c->val, c->cfg.maxneighbs10, c->d_flag, (unsigned long long *)(c->d_acc + 33), ralign);
  if (MODE >= 1 && hessian)
    LAUNCH(c, k_hessian, 148 * 16, 256, 0, nnz, c->d_ff, c->val);
  c->have_col16 = false;
  if (want16 && nnz > 0) {
    RXG_CUDA(cudaMemsetAsync(c->d_flag + 7, 0, sizeof(int), c->st));
    LAUNCH(c, k_col16, 148 * 16, 256, 0, nnz, c->col, c->col16, c->cbase, c->d_flag + 7);
    c->have_col16 = true;   // provisional: the overflow flag is read with the first CG scalars (qeq_cg_single)
  }
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


// Sum four per-lane doubles over the warp with 9 double-shuffles instead of 20 (shuffles share the L1 data pipe with
// the gathers, which is the busiest unit of the SpMV): halve the set of values each lane carries in the first two
// butterfly steps, finish with three single-value steps, then collect the four totals from lanes 0, 8, 16, 24.
__device__ __forceinline__ void warp_sum4(double &v0, double &v1, double &v2, double &v3, int lane) {
  const bool up16 = lane & 16, up8 = lane & 8;
  double s0 = up16 ? v0 : v2, s1 = up16 ? v1 : v3;          // what this lane gives away
  double k0 = up16 ? v2 : v0, k1 = up16 ? v3 : v1;          // what it keeps
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  double give = up8 ? k0 : k1, u = up8 ? k1 : k0;
  u += __shfl_xor_sync(0xffffffffu, give, 8);
  u += __shfl_xor_sync(0xffffffffu, u, 4);
  u += __shfl_xor_sync(0xffffffffu, u, 2);
  u += __shfl_xor_sync(0xffffffffu, u, 1);
  v0 = __shfl_sync(0xffffffffu, u, 0);
  v1 = __shfl_sync(0xffffffffu, u, 8);
  v2 = __shfl_sync(0xffffffffu, u, 16);
  v3 = __shfl_sync(0xffffffffu, u, 24);
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
__global__ void __launch_bounds__(256) k_gradient(const int *__restrict__ order, int ntot, int natoms, const long long *__restrict__ rowbeg, const long long *__restrict__ rowend, const int *__restrict__ col,
                                                  const double *__restrict__ val, const double2 *__restrict__ qst,
                                                  const int *__restrict__ itype, const DevFF *__restrict__ ffp,
                                                  double2 *__restrict__ gst, double *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // rows are walked in cell order (L1-friendly gathers)
  int i = natoms;
  if (slot < ntot) i = order[slot];
  double part[2] = {0.0, 0.0};
  if (i < natoms) {
    long long s = rowbeg[i], e = rowend[i];
    double gs = 0.0, gt = 0.0;
    for (long long k = s + lane; k < e; k += 32) {
      double h = __ldcs(val + k);
      int j = order[__ldcs(col + k) & COL_MASK];
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
__global__ void __launch_bounds__(256) k_hsh(const int *__restrict__ order, int ntot, int natoms, const long long *__restrict__ rowbeg, const long long *__restrict__ rowend, const int *__restrict__ col,
                                             const double *__restrict__ val, const double4 *__restrict__ hsq,
                                             const double2 *__restrict__ gst, const int *__restrict__ itype,
                                             const DevFF *__restrict__ ffp, double *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // rows are walked in cell order (L1-friendly gathers)
  int i = natoms;
  if (slot < ntot) i = order[slot];
  double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < natoms) {
    long long s = rowbeg[i], e = rowend[i];
    double4 me = hsq[i];
    double ts = 0.0, tt = 0.0, es = 0.0;
    for (long long k = s + lane; k < e; k += 32) {
      double h = __ldcs(val + k);
      int j = order[__ldcs(col + k) & COL_MASK];
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
__global__ void __launch_bounds__(256) k_spmv1(const int *__restrict__ order, int ntot, int natoms, const long long *__restrict__ rowbeg, const long long *__restrict__ rowend, const int *__restrict__ col,
                                               const double *__restrict__ val, const double2 *__restrict__ x,
                                               const double2 *__restrict__ qst, const double *__restrict__ q,
                                               double2 *__restrict__ gst, double2 *__restrict__ tst,
                                               double2 *__restrict__ ust, double2 *__restrict__ wst,
                                               const int *__restrict__ itype, const DevFF *__restrict__ ffp,
                                               double *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // rows are walked in cell order (L1-friendly gathers)
  int i = natoms;
  if (slot < ntot) i = order[slot];
  double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < natoms) {
    long long s = rowbeg[i], e = rowend[i];
    double a = 0.0, b = 0.0, ga = 0.0, gb = 0.0;
    for (long long k = s + lane; k < e; k += 32) {
      double h = __ldcs(val + k);
      int j = __ldcs(col + k);
      double2 v = x[j & COL_MASK];            // x is in slot order: neighbours of a stencil run are contiguous
      double pa = h * v.x, pb = h * v.y;
      a += pa; b += pb;
      if (j < 0) { ga += pa; gb += pb; }      // bit 31 = ghost column
    }
    warp_sum4(a, b, ga, gb, lane);
    if (lane == 0) {
      int t = itype[i] - 1;
      double eta = ffp->eta[t], chi = ffp->chi[t];
      double2 me = x[slot];
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

// ---------------------------------------------------------------------------------------------------
// TMA-staged variant of k_spmv1 (the production kernel).  Rows lie in HBM in cell order, so the rows of SP_ROWS
// consecutive slots form ONE contiguous span of (val, col).  One CTA per span: an elected thread issues two bulk
// async copies (cp.async.bulk ... mbarrier::complete_tx) that bring the span into shared memory, the CTA's warps
// then take one row each and read the matrix stream from shared memory while gathering x from L1/L2.  With several
// CTAs resident per SM the copy engine always has tens of KB in flight per SM, which is what HBM needs; the SM's
// load/store pipe is left to the gathers.  Rows start on 4-entry boundaries (16 B for col, 32 B for val).
constexpr int SP_ROWS = 4, SP_CAP = 1920;   // 8 rows x up to 480 entries: 46 KB of shared memory per CTA

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

template <bool INIT>
__global__ void __launch_bounds__(SP_ROWS * 32) k_spmv1_tma(const int *__restrict__ order, int ntot, int natoms,
                                                            const long long *__restrict__ rowoff,
                                                            const long long *__restrict__ rowbeg,
                                                            const long long *__restrict__ rowend, const int *__restrict__ col,
                                                            const double *__restrict__ val, const double2 *__restrict__ x,
                                                            const double2 *__restrict__ qst, const double *__restrict__ q,
                                                            double2 *__restrict__ gst, double2 *__restrict__ tst,
                                                            double2 *__restrict__ ust, double2 *__restrict__ wst,
                                                            const int *__restrict__ itype, const DevFF *__restrict__ ffp,
                                                            double *__restrict__ acc) {
  __shared__ __align__(128) double s_val[SP_CAP];
  __shared__ __align__(128) int s_col[SP_CAP];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int slot0 = blockIdx.x * SP_ROWS;
  const int slot1 = min(slot0 + SP_ROWS, ntot);
  const long long sb = rowoff[slot0], se = rowoff[slot1];
  const int span = (int)(se - sb);
  const bool staged = span > 0 && span <= SP_CAP;
  if (staged) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, (unsigned)span * 12u);
      bulk_g2s(s_val, val + sb, (unsigned)span * 8u, &bar);
      bulk_g2s(s_col, col + sb, (unsigned)span * 4u, &bar);
    }
  }
  const int slot = slot0 + wid;
  int i = natoms;
  if (slot < ntot) i = order[slot];
  double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  // per-row scalars are fetched while the bulk copies are in flight
  long long rs = 0, re = 0;
  double eta = 0, chi = 0, qi = 0, mu = 0;
  double2 me = make_double2(0, 0), g = make_double2(0, 0), w = make_double2(0, 0);
  if (i < natoms) {
    rs = rowbeg[i]; re = rowend[i];
    int t = itype[i] - 1;
    eta = ffp->eta[t]; chi = ffp->chi[t];
    me = x[slot];
    if (!INIT) { g = gst[i]; w = wst[i]; qi = q[i]; mu = acc[11]; }
  }
  if (staged) mbar_wait(&bar, 0);
  if (i < natoms) {
    double a = 0.0, b = 0.0, ga = 0.0, gb = 0.0;
    const int n = (int)(re - rs);
    if (staged) {
      const double *sv = s_val + (rs - sb);
      const int *sc = s_col + (rs - sb);
#pragma unroll 8
      for (int k = lane; k < n; k += 32) {
        double h = sv[k];
        int j = sc[k];
        double2 v = x[j & COL_MASK];
        double pa = h * v.x, pb = h * v.y;
        a += pa; b += pb;
        if (j < 0) { ga += pa; gb += pb; }
      }
    } else {
      for (long long k = rs + lane; k < re; k += 32) {
        double h = __ldcs(val + k);
        int j = __ldcs(col + k);
        double2 v = x[j & COL_MASK];
        double pa = h * v.x, pb = h * v.y;
        a += pa; b += pb;
        if (j < 0) { ga += pa; gb += pb; }
      }
    }
    warp_sum4(a, b, ga, gb, lane);
    if (lane == 0) {
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
        part[0] = chi * qi + 0.5 * eta * qi * qi + 0.5 * qi * (w.x - mu * w.y);
        part[1] = ts * me.x; part[2] = tt * me.y; part[3] = g.x * me.x; part[4] = g.y * me.y;
      }
    }
  }
  if (INIT) { double p2[2] = {part[0], part[1]}; block_accumulate<2>(p2, acc + 7); }
  else block_accumulate<5>(part, acc + 0);
}

// ---------------------------------------------------------------------------------------------------
// Production SpMV, split in two so that the streaming kernel carries no epilogue (measured on B200: the fused epilogue
// -- scattered loads of g, w, q by the row's lane 0, two vector stores and a 5-value CTA reduction per 4 rows -- cost
// 0.26 ms of 1.39 ms):
//   k_spmv_rows : TMA-staged matrix stream, LPR lanes per row, writes the four raw row sums {a, b, ghost a, ghost b}
//                 of H.(x1,x2) per cell-order slot;
//   k_cg_dots   : one thread per slot, coalesced; turns row sums into gradient / H.h products, Est and the dots.
template <int ROWS, int LPR, int CAPROW = 480>
__global__ void __launch_bounds__(ROWS * LPR) k_spmv_rows(const int *__restrict__ order, int ntot, int natoms,
                                                          const long long *__restrict__ rowoff, const long long *__restrict__ rowbeg,
                                                          const long long *__restrict__ rowend, const int *__restrict__ col,
                                                          const double *__restrict__ val, const double2 *__restrict__ x,
                                                          double4 *__restrict__ rowsum) {
  constexpr int CAP = ROWS * CAPROW;   // CAPROW = longest row the staged path takes (480: 10 A lists; 1216: the 12.5 A lists of PQEq)
  __shared__ __align__(128) double s_val[CAP];
  __shared__ __align__(128) int s_col[CAP];
  __shared__ __align__(8) unsigned long long bar;
  const int sub = threadIdx.x % LPR, rowid = threadIdx.x / LPR;
  const int slot0 = blockIdx.x * ROWS;
  const int slot1 = min(slot0 + ROWS, ntot);
  const long long sb = rowoff[slot0], se = rowoff[slot1];
  const int span = (int)(se - sb);
  const bool staged = span > 0 && span <= CAP;
  if (staged) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, (unsigned)span * 12u);
      bulk_g2s(s_val, val + sb, (unsigned)span * 8u, &bar);
      bulk_g2s(s_col, col + sb, (unsigned)span * 4u, &bar);
    }
  }
  const int slot = slot0 + rowid;
  int i = natoms;
  if (slot < ntot) i = order[slot];
  long long rs = 0, re = 0;
  if (i < natoms) { rs = rowbeg[i]; re = rowend[i]; }
  if (staged) mbar_wait(&bar, 0);
  double a = 0.0, b = 0.0, ga = 0.0, gb = 0.0;
  if (i < natoms) {
    if (staged) {
      const int n = (int)(re - rs);
      const double *sv = s_val + (rs - sb);
      const int *sc = s_col + (rs - sb);
#pragma unroll 8
      for (int k = sub; k < n; k += LPR) {
        double h = sv[k];
        int j = sc[k];
        double2 v = x[j & COL_MASK];            // x is in slot order: neighbours of a stencil run are contiguous
        double pa = h * v.x, pb = h * v.y;
        a += pa; b += pb;
        if (j < 0) { ga += pa; gb += pb; }      // bit 31 = ghost column
      }
    } else {
      for (long long k = rs + sub; k < re; k += LPR) {
        double h = __ldcs(val + k);
        int j = __ldcs(col + k);
        double2 v = x[j & COL_MASK];
        double pa = h * v.x, pb = h * v.y;
        a += pa; b += pb;
        if (j < 0) { ga += pa; gb += pb; }
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o);
    ga += __shfl_xor_sync(0xffffffffu, ga, o); gb += __shfl_xor_sync(0xffffffffu, gb, o);
  }
  if (sub == 0 && i < natoms) rowsum[slot] = make_double4(a, b, ga, gb);
}

// k_spmv_rows with the 16-bit column stream of k_col16: 10.125 B per entry from HBM instead of 12.
template <int ROWS, int LPR, bool STAGE_VAL = true>
__global__ void __launch_bounds__(ROWS * LPR) k_spmv_rows16(const int *__restrict__ order, int ntot, int natoms,
                                                            const long long *__restrict__ rowoff, const long long *__restrict__ rowbeg,
                                                            const long long *__restrict__ rowend, const unsigned short *__restrict__ col16,
                                                            const int *__restrict__ cbase, const int *__restrict__ col,
                                                            const double *__restrict__ val, const double2 *__restrict__ x,
                                                            double4 *__restrict__ rowsum) {
  constexpr int CAP = ROWS * 480;
  __shared__ __align__(128) double s_val[STAGE_VAL ? CAP : 16];
  __shared__ __align__(128) unsigned short s_col[CAP];
  __shared__ int s_base[CAP / 16 + 2];
  __shared__ __align__(8) unsigned long long bar;
  const int sub = threadIdx.x % LPR, rowid = threadIdx.x / LPR;
  const int slot0 = blockIdx.x * ROWS;
  const int slot1 = min(slot0 + ROWS, ntot);
  const long long sb = rowoff[slot0], se = rowoff[slot1];
  const int span = (int)(se - sb);
  const bool staged = span > 0 && span <= CAP;
  if (staged) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, (unsigned)span * (STAGE_VAL ? 10u : 2u));
      if (STAGE_VAL) bulk_g2s(s_val, val + sb, (unsigned)span * 8u, &bar);
      bulk_g2s(s_col, col16 + sb, (unsigned)span * 2u, &bar);
    }
    const long long b0 = sb >> 4;
    const int nb = span >> 4;
    for (int t = threadIdx.x; t < nb; t += ROWS * LPR) s_base[t] = cbase[b0 + t];
  }
  const int slot = slot0 + rowid;
  int i = natoms;
  if (slot < ntot) i = order[slot];
  long long rs = 0, re = 0;
  if (i < natoms) { rs = rowbeg[i]; re = rowend[i]; }
  if (staged) { mbar_wait(&bar, 0); __syncthreads(); }
  double a = 0.0, b = 0.0, ga = 0.0, gb = 0.0;
  if (i < natoms) {
    if (staged) {
      const int n = (int)(re - rs);
      const int p0 = (int)(rs - sb);
      const double *sv = STAGE_VAL ? s_val + p0 : val + rs;
      const unsigned short *sc = s_col + p0;
      const int *sbs = s_base + (p0 >> 4);
#pragma unroll 8
      for (int k = sub; k < n; k += LPR) {
        const double h = STAGE_VAL ? sv[k] : __ldcs(sv + k);
        const unsigned c16 = sc[k];
        const double2 v = x[sbs[k >> 4] + (int)(c16 & 0x7fffu)];
        const double pa = h * v.x, pb = h * v.y;
        a += pa; b += pb;
        if (c16 & 0x8000u) { ga += pa; gb += pb; }
      }
    } else {
      for (long long k = rs + sub; k < re; k += LPR) {
        double h = __ldcs(val + k);
        int j = __ldcs(col + k);
        double2 v = x[j & COL_MASK];
        double pa = h * v.x, pb = h * v.y;
        a += pa; b += pb;
        if (j < 0) { ga += pa; gb += pb; }
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o);
    ga += __shfl_xor_sync(0xffffffffu, ga, o); gb += __shfl_xor_sync(0xffffffffu, gb, o);
  }
  if (sub == 0 && i < natoms) rowsum[slot] = make_double4(a, b, ga, gb);
}

template <bool INIT>
__global__ void __launch_bounds__(256) k_cg_dots(const int *__restrict__ order, int ntot, int natoms, const double4 *__restrict__ rowsum,
                                                 const double2 *__restrict__ x, const double *__restrict__ q,
                                                 double2 *__restrict__ gst, double2 *__restrict__ tst, double2 *__restrict__ ust,
                                                 double2 *__restrict__ wst, const int *__restrict__ itype,
                                                 const DevFF *__restrict__ ffp, double *__restrict__ acc) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  int i = natoms;
  if (slot < ntot) i = order[slot];
  if (i < natoms) {
    const double4 r = rowsum[slot];
    const double2 me = x[slot];
    const int t = itype[i] - 1;
    const double eta = ffp->eta[t], chi = ffp->chi[t];
    if (INIT) {
      double g1 = sub_rn(sub_rn(-chi, mul_rn(eta, me.x)), r.x);
      double g2 = sub_rn(sub_rn(-1.0, mul_rn(eta, me.y)), r.y);
      gst[i] = make_double2(g1, g2);
      wst[i] = make_double2(2.0 * r.x - r.z, 2.0 * r.y - r.w);
      part[0] = g1 * g1; part[1] = g2 * g2;
    } else {
      double ts = eta * me.x + r.x, tt = eta * me.y + r.y;
      tst[i] = make_double2(ts, tt);
      ust[i] = make_double2(2.0 * r.x - r.z, 2.0 * r.y - r.w);
      const double2 g = gst[i], w = wst[i];
      const double mu = acc[11], qi = q[i];
      part[0] = chi * qi + 0.5 * eta * qi * qi + 0.5 * qi * (w.x - mu * w.y);
      part[1] = ts * me.x; part[2] = tt * me.y; part[3] = g.x * me.x; part[4] = g.y * me.y;
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
                             double2 *__restrict__ hst, double2 *__restrict__ xs, const int *__restrict__ slot_of,
                             double *__restrict__ q, double *__restrict__ acc) {
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
  xs[slot_of[i]] = h;
}
__global__ void k_h_from_g2(int natoms, const double2 *__restrict__ gst, double2 *__restrict__ hst, double2 *__restrict__ xs,
                            const int *__restrict__ slot_of) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < natoms) { double2 g = gst[i]; hst[i] = g; xs[slot_of[i]] = g; }
}
// xs[slot] = v[order[slot]] for residents and ghosts
__global__ void k_to_slots(int ntot, const int *__restrict__ order, const double2 *__restrict__ v, double2 *__restrict__ xs) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < ntot) xs[s] = v[order[s]];
}

// ---------------------------------------------------------------------------------------------------
// STRICT-ORDER validation path (RXG_STRICT_ORDER=1): the same CG with every sum taken in the reference's serial
// order and without FMA, so that the iterates are bit-identical to a serial x86-64 build of the reference.
// The reference's CG amplifies round-off (its real(4) step length keeps it in a noise-dominated regime: an FMA
// build of the same Fortran/C++ loops changes the converged charges by ~1e-5), so only this path can be compared
// at 1e-8; the production kernels above differ from it by summation order alone.  Small systems only.
__global__ void k_rows_strict_grad(const int *__restrict__ order, int natoms, const long long *__restrict__ rowbeg, const long long *__restrict__ rowend, const int *__restrict__ col,
                                   const double *__restrict__ val, const double2 *__restrict__ qst,
                                   const int *__restrict__ itype, const DevFF *__restrict__ ffp, double2 *__restrict__ gst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  double gs = 0.0, gt = 0.0;
  for (long long k = rowbeg[i]; k < rowend[i]; k++) {
    double2 x = qst[order[col[k] & COL_MASK]];
    gs = add_rn(gs, mul_rn(val[k], x.x));
    gt = add_rn(gt, mul_rn(val[k], x.y));
  }
  int t = itype[i] - 1;
  double eta = ffp->eta[t], chi = ffp->chi[t];
  double2 x = qst[i];
  gst[i] = make_double2(sub_rn(sub_rn(-chi, mul_rn(eta, x.x)), gs), sub_rn(sub_rn(-1.0, mul_rn(eta, x.y)), gt));
}
__global__ void k_rows_strict_hsh(const int *__restrict__ order, int natoms, const long long *__restrict__ rowbeg, const long long *__restrict__ rowend, const int *__restrict__ col,
                                  const double *__restrict__ val, const double4 *__restrict__ hsq,
                                  const int *__restrict__ itype, const DevFF *__restrict__ ffp, double4 *__restrict__ rowbuf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  int t = itype[i] - 1;
  double eta = ffp->eta[t], chi = ffp->chi[t];
  double4 me = hsq[i];
  double ts = mul_rn(eta, me.x), tt = mul_rn(eta, me.y);
  double es = add_rn(mul_rn(chi, me.z), mul_rn(mul_rn(mul_rn(0.5, eta), me.z), me.z));
  for (long long k = rowbeg[i]; k < rowend[i]; k++) {
    int j = order[col[k] & COL_MASK];
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
