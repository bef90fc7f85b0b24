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
        double4 o = ldg256(g.sorted + k);
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
// One work item of the CG's sparse product (k_spmv_items): up to `rg` consecutive rows of a block of <= 8 rows of one cell.
// Its values are one contiguous span of `val` (rows lie in HBM in cell order), its columns the block's union stream.
struct __align__(16) SpItem {
  long long voff;   // first value of the item's first row
  long long uoff;   // first entry of the block's union stream
  int un;           // entries of the union stream (padded to 16)
  int slot0;        // cell-order slot of the item's first row
  int nrp;          // rows of the item (1..rg)
  int vlen;         // values in the span (rows padded to 4)
  int rshift;       // position of the item's first row in the block's 8-bit row sets
  int rs[4];        // start of each row relative to voff
  int pad[3];
};
static_assert(sizeof(SpItem) == 64, "SpItem is one 64-byte record");
// MODE 0: FORCE list, 1: QEq list, 2: both at once (FORCE predicate; hessian = 0 where only the QEq predicate fails)
// CAPPED (fill only): there was no count pass; rows were laid out with capacities taken from the previous step's counts
// (k_row_caps), so every write is guarded by its row's capacity, an overflow raises ovf[20], and this pass also produces what
// the count pass would have (entry total, longest row).  Every fill of a QEq list records the rows' counts by global atom id
// (cnt_tab) for the next step.
struct CntTab { int2 *tab; unsigned mask; const int *gid; };
// Window-relative column stream of k_spmv_win (below).  A GROUP is up to G consecutive resident cells of one z-column of the
// grid; its WINDOW is the concatenation, in stencil-run order, of the slot ranges that the runs of ANY cell of the group cover
// (run r of the group = cells [g0 + z0_r, g0 + gl - 1 + z1_r] of column (c1 + dx_r, c2 + dy_r), clipped to the grid).  The fill
// pass writes, next to the 32-bit slot column, the entry's position in its group's window (15 bits) with the ghost flag in bit
// 15, and the exact length of every row by slot; `winmax` collects the largest window.
struct WinOut { unsigned short *c16; int *rowlen; const int2 *desc; int G; int park; };   // park: the 16-bit column rides in the parked value word, k_hessian stores it (coalesced)
// bounds of stencil run `rr` for the cells [za, zb] of column (c1, c2): first slot and number of slots (0 if outside the grid)
__device__ __forceinline__ void run_span(const DevGrid &g, int4 rr, int c1, int c2, int za, int zb, int &s, int &len) {
  s = 0; len = 0;
  const int b1 = c1 + rr.x, b2 = c2 + rr.y;
  int z0 = za + rr.z, z1 = zb + rr.w;
  if (b1 >= -g.L && b1 < g.nc[0] + g.L && b2 >= -g.L && b2 < g.nc[1] + g.L) {
    if (z0 < -g.L) z0 = -g.L;
    if (z1 >= g.nc[2] + g.L) z1 = g.nc[2] + g.L - 1;
    if (z1 >= z0) {
      const int cbase = ((b1 + g.L) * g.dim[1] + (b2 + g.L)) * g.dim[2] + g.L;
      s = g.start[cbase + z0];
      len = g.start[cbase + z1 + 1] - s;
    }
  }
}
// Window descriptors, written once per list build (k_win_desc): for group `grp` and stencil run r the pair
//   { first slot of the run's span, (position in the window << 12) | number of slots },
// and in entry [nruns] the window's total {0, total << 12}.  A span of more than 4095 slots marks the group as unfit (total = 2^19).
__global__ void k_win_desc(DevGrid g, const int *__restrict__ runs, int nruns, int G, int ngroups, int2 *__restrict__ desc, int *__restrict__ winmax) {
  const int lane = threadIdx.x & 31;
  const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (grp >= ngroups) return;
  const int ngz = (g.nc[2] + G - 1) / G;
  const int gz = grp % ngz, c2 = (grp / ngz) % g.nc[1], c1 = grp / (ngz * g.nc[1]);
  const int g0 = gz * G, g1 = min(g0 + G, g.nc[2]) - 1;
  int2 *d = desc + (size_t)grp * (nruns + 1);
  int wcarry = 0;
  bool unfit = false;
  for (int r0 = 0; r0 < nruns; r0 += 32) {
    const int r = r0 + lane;
    int ws = 0, wlen = 0;
    if (r < nruns) run_span(g, *reinterpret_cast<const int4 *>(runs + 4 * r), c1, c2, g0, g1, ws, wlen);
    int winc = wlen;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += y;
    }
    const int pos = wcarry + winc - wlen;
    unfit = unfit || __any_sync(0xffffffffu, wlen > 4095);
    if (r < nruns) d[r] = make_int2(ws, (min(pos, 0x7ffff) << 12) | min(wlen, 4095));
    wcarry += __shfl_sync(0xffffffffu, winc, 31);
  }
  if (lane == 0) {
    d[nruns] = make_int2(0, ((unfit || wcarry > 0x7ffff) ? 0x7ffff : wcarry) << 12);
    if (wcarry > __ldcg(winmax)) atomicMax(winmax, wcarry);   // the largest window sizes the launches of k_spmv_win
  }
}

__device__ __forceinline__ unsigned cnt_hash(int gid, unsigned mask) { return ((unsigned)gid * 2654435761u) & mask; }
#ifndef RXG_PL_MINB
#define RXG_PL_MINB 5   // resident CTAs per SM the fill pass is compiled for (48 registers; 64 without the cap)
#endif
template <int MODE, bool FILL, bool UNION, bool HFUSE = false, bool CAPPED = false>
__global__ void __launch_bounds__(PL_WARPS * 32, FILL ? RXG_PL_MINB : 1) k_pairlist(DevGrid g, const DevFF *__restrict__ ffp, const int *__restrict__ runs,
                                                            int nruns, int natoms, int ncell_res, int *__restrict__ slotcnt,
                                                            const long long *__restrict__ rowoff, long long *__restrict__ rowbeg,
                                                            long long *__restrict__ rowend, int *__restrict__ col,
                                                            double *__restrict__ val, int maxrow, int *__restrict__ ovf,
                                                            unsigned long long *__restrict__ nnz_real, int ralign,
                                                            int *__restrict__ ucnt, const long long *__restrict__ uoff,
                                                            int *__restrict__ ucol, unsigned char *__restrict__ umask,
                                                            SpItem *__restrict__ items, int *__restrict__ nitems, int rg, CntTab ct, WinOut wo) {
  __shared__ int sh_s[PL_WARPS][PL_MAXRUNS];       // first slot of each stencil run
  __shared__ int sh_w[PL_WARPS][PL_MAXRUNS];       // (fill, wo.c16) window position of the run's first slot: add the candidate's slot
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
  const bool win = FILL && wo.c16 != nullptr;
  // this cell's group (k_spmv_win) and its window descriptors (k_win_desc): run r starts at window position (d.y >> 12) with slot d.x
  const int2 *wdesc = win ? wo.desc + (size_t)((c1 * g.nc[1] + c2) * ((g.nc[2] + wo.G - 1) / wo.G) + c3 / wo.G) * (nruns + 1) : nullptr;
  for (int ab = a0; ab < a1; ab += 32) {
    const int nb = min(32, a1 - ab);
    const int myslot = ab + lane;
    double4 me = make_double4(0, 0, 0, 0);
    int mi = natoms;
    if (lane < nb) { me = g.sorted[myslot]; mi = rec_index(me.w); }
    sh_a[wid][lane] = me;
    const bool mine = mi < natoms;           // a ghost inside a resident cell owns no row (cannot happen after MOVE)
    long long mybase = (FILL && lane < nb) ? rowoff[myslot] : 0;
    const int mycap = (CAPPED && lane < nb) ? (int)(rowoff[myslot + 1] - mybase) : 0;   // this row's capacity (k_row_caps)
    // write positions travel through the warp as 32-bit offsets from the batch's first row (one shuffle per test instead of two)
    const long long base0 = FILL ? __shfl_sync(0xffffffffu, mybase, 0) : 0;
    const int myrel = (int)(mybase - base0);
    int mycnt = 0;
    // union stream of the CG SpMV (k_spmv_items): per block of up to 8 consecutive rows of this cell, the candidates accepted
    // by at least one of them, in candidate order, with the 8-bit set of accepting rows.  ub* = running entry count of the
    // (up to four) blocks of this batch of 32 rows; identical on every lane.
    int ub0 = 0, ub1 = 0, ub2 = 0, ub3 = 0;
    long long uw0 = 0, uw1 = 0, uw2 = 0, uw3 = 0;
    if (UNION && FILL) {
      uw0 = uoff[ab];
      if (nb > 8) uw1 = uoff[ab + 8];
      if (nb > 16) uw2 = uoff[ab + 16];
      if (nb > 24) uw3 = uoff[ab + 24];
    }
    int lastcol = ab;
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
          run_span(g, rr, c1, c2, c3, c3, s, len);
          if (win) { const int2 d = __ldg(wdesc + rb + r); sh_w[wid][r] = (d.y >> 12) - d.x; }
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
        if (have) o = ldg256(g.sorted + cslot);
        const int jt = rec_type(o.w);
        const int cval = cslot | ((have && rec_index(o.w) >= natoms) ? COL_GHOST : 0);
        // position in the group's window + ghost flag (garbage when the window exceeds 15 bits: k_spmv_win then reads `col`)
        const unsigned short c16v = (win && have) ? (unsigned short)(((sh_w[wid][r] + cslot) & 0x7fff) | (cval < 0 ? 0x8000 : 0)) : 0;
        unsigned um = 0;   // rows of this batch whose list takes this lane's candidate
        for (int a = 0; a < nb; a++) {
          const double4 at = sh_a[wid][a];
          const double dr2 = dist2_rn(sub_rn(at.x, o.x), sub_rn(at.y, o.y), sub_rn(at.z, o.z));
          const bool acc = have && (cslot != ab + a) && (MODE == 1 ? ((float)dr2 < rctap2f) : (dr2 <= ff.rctap2));
          const unsigned mask = __ballot_sync(0xffffffffu, acc);
          if (FILL) {
            // one shuffle carries the row's write offset (20 bits) and, CAPPED, what is left of its capacity (12 bits, clamped:
            // a step adds at most 32 entries)
            const unsigned pk = __shfl_sync(0xffffffffu, (unsigned)(myrel + mycnt) | (CAPPED ? ((unsigned)min(max(mycap - mycnt, 0), 4095) << 20) : 0u), a);
            const int before = __popc(mask & ((1u << lane) - 1u));
            const long long w = base0 + (int)(pk & 0xfffffu) + before;
            if (acc && (!CAPPED || before < (int)(pk >> 20))) {
              col[w] = cval;
              if (win && !wo.park) wo.c16[w] = c16v;
              if (MODE >= 1) {
                // the hessian lerp is evaluated by k_hessian over the compacted rows (full lanes); here only the
                // fp32-rounded r^2 and the bond type are parked in the 8 bytes of the value slot
                int inxn = ff.inxn2[(rec_type(at.w) - 1) + ff.nso * (jt - 1)];
                if (!(MODE == 1 || (float)dr2 < rctap2f)) inxn = 0;
                if (HFUSE) {   // experiment (RXG_HESS_FUSE=1): the lerp of k_hessian evaluated here, by the accepted lanes only
                  const double d2 = (double)(float)dr2;
                  const int itb = (int)mul_rn(d2, ff.UDRi);
                  const double drtb = mul_rn(sub_rn(d2, mul_rn((double)itb, ff.UDR)), ff.UDRi);
                  double h = 0.0;
                  if (inxn > 0 && itb >= 1 && itb < ff.ntable) {
                    const double2 T = ff.TBL_qeq2[(size_t)(inxn - 1) * ff.ntable + (itb - 1)];
                    h = add_rn(mul_rn(sub_rn(1.0, drtb), T.x), mul_rn(drtb, T.y));
                  }
                  val[w] = h;
                } else
                  val[w] = __hiloint2double(max(inxn, 0) | ((win && wo.park) ? ((int)c16v << 16) : 0), __float_as_int((float)dr2));
              }
            }
          }
          if (lane == a) mycnt += __popc(mask);
          // (MODE 2: the row also holds the pairs that pass FORCE's fp64 test only; their hessian is 0 and they keep their
          // place in the value stream, so the union must list them too)
          if (UNION && acc) um |= 1u << a;
        }
        if (UNION) {
#define RXG_UBLOCK(B, UB, UW)                                                          \
          if (nb > 8 * B) {                                                            \
            const unsigned m8 = (um >> (8 * B)) & 0xffu;                               \
            const unsigned bal = __ballot_sync(0xffffffffu, m8 != 0);                  \
            if (FILL && m8) {                                                          \
              const long long w = UW + UB + __popc(bal & ((1u << lane) - 1u));         \
              ucol[w] = cval;                                                          \
              umask[w] = (unsigned char)m8;                                            \
            }                                                                          \
            UB += __popc(bal);                                                         \
          }
          RXG_UBLOCK(0, ub0, uw0) RXG_UBLOCK(1, ub1, uw1) RXG_UBLOCK(2, ub2, uw2) RXG_UBLOCK(3, ub3, uw3)
#undef RXG_UBLOCK
          if (FILL) {   // a valid column for the padding of the union stream
            const unsigned any = __ballot_sync(0xffffffffu, um != 0);
            if (any) lastcol = __shfl_sync(0xffffffffu, cval, 31 - __clz(any)) & COL_MASK;
          }
        }
      }
      __syncwarp();
    }
    if (UNION) {   // close the union blocks: counts padded to 16 entries (64 B of ucol, 16 B of umask: bulk-copy granularity)
      const int ubs[4] = {ub0, ub1, ub2, ub3};
      const long long uws[4] = {uw0, uw1, uw2, uw3};
#pragma unroll
      for (int B = 0; B < 4; B++) {
        if (nb > 8 * B) {
          const int padded = (ubs[B] + 15) & ~15;
          if (!FILL) { if (lane == 0) ucnt[ab + 8 * B] = padded; }
          else if (lane < padded - ubs[B]) { ucol[uws[B] + ubs[B] + lane] = lastcol; umask[uws[B] + ubs[B] + lane] = 0; }
        }
      }
      if (FILL) {
        // work items of the SpMV: lane t describes pass (t % ppb) of block (t / ppb) of this batch of <= 32 rows
        const int ppb = 8 / rg, B = lane / ppb, r0 = (lane % ppb) * rg;
        const int row0 = 8 * B + r0;                       // first row of the item within the batch
        const bool valid = row0 < nb && B < 4;
        const int nrB = min(8, nb - 8 * B);                // rows of the block
        const int nrp = valid ? min(rg, nrB - r0) : 0;
        const int plen = (lane < nb && mine) ? ((mycnt + ralign - 1) & ~(ralign - 1)) : 0;
        SpItem I;
        I.voff = __shfl_sync(0xffffffffu, mybase, valid ? row0 : 0);
        I.uoff = B == 0 ? uw0 : (B == 1 ? uw1 : (B == 2 ? uw2 : uw3));
        I.un = ((B == 0 ? ub0 : (B == 1 ? ub1 : (B == 2 ? ub2 : ub3))) + 15) & ~15;
        I.slot0 = ab + row0; I.nrp = nrp; I.rshift = r0;
        I.pad[0] = I.pad[1] = I.pad[2] = 0;
        int vlen = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int src = (valid && j < nrp) ? row0 + j : 0;
          const long long bj = __shfl_sync(0xffffffffu, mybase, src);
          const int lj = __shfl_sync(0xffffffffu, plen, src);
          I.rs[j] = (valid && j < nrp) ? (int)(bj - I.voff) : 0;
          if (valid && j < nrp) vlen = (int)(bj - I.voff) + lj;
        }
        I.vlen = vlen;
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        int base = 0;
        if (lane == 0) base = atomicAdd(nitems, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (valid) {
          const int4 *src = reinterpret_cast<const int4 *>(&I);
          int4 *dst = reinterpret_cast<int4 *>(items + base + __popc(bal & ((1u << lane) - 1u)));
          dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        }
      }
    }
    if (win && lane < nb) wo.rowlen[myslot] = mine ? (CAPPED ? min(mycnt, mycap) : mycnt) : -1;
    if (!FILL || CAPPED) {   // exact entry count (without row padding): the algorithmic-bytes figure of the roofline uses it;
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
        int kept = mycnt;
        if (CAPPED) {
          if (mycnt > maxrow) atomicMax(ovf + 26, mycnt);          // the MAXNEIGHBS10 trap, checked by the host with the overflow flag
                                                                   // (its own slot: ovf[0] is reused by the cell and bonded-list builds that may follow)
          if (mycnt > mycap) {   // the row outgrew last step's count + slack: the host rebuilds (ovf[22..25]: one such row, for diagnostics)
            if (atomicExch(ovf + 20, 1) == 0) { ovf[22] = ct.gid[mi]; ovf[23] = mycnt; ovf[24] = mycap; ovf[25] = mi; }
            kept = mycap;
          }
        }
        rowbeg[mi] = mybase;
        rowend[mi] = mybase + kept;
        // row padding: hessian 0, column = the row's last real column (written by some lane of this warp before the
        // __syncwarp above).  CAPPED: the whole unused capacity is padding (k_hessian runs flat over all entries).
        const int padcol = kept > 0 ? __ldcg(col + mybase + kept - 1) : myslot;
        const int pend = CAPPED ? mycap : ((mycnt + ralign - 1) & ~(ralign - 1));
        for (int p = kept; p < pend; p++) {
          col[mybase + p] = padcol;
          if (MODE >= 1) val[mybase + p] = 0.0;
        }
        if (MODE >= 1 && ct.tab) {   // this step's count, by global id, is next step's capacity
          const int gi = ct.gid[mi];
          ct.tab[cnt_hash(gi, ct.mask)] = make_int2(gi, mycnt);
        }
      }
    }
    __syncwarp();
  }
}


// D1: hessian(j1,i) = (1-drtb)*TBL_Eclmb_QEq(itb,inxn) + drtb*TBL_Eclmb_QEq(itb+1,inxn) with real(4) dr2 (src/qeq.F90:234-240).
// One warp per row, lanes stride over the compacted entries; reads {float r^2, inxn} parked by k_pairlist.
__global__ void __launch_bounds__(256) k_hessian(long long nnz, const DevFF *__restrict__ ffp, double *__restrict__ val, unsigned short *__restrict__ c16) {
  // flat over the padded entry range: row padding carries inxn = 0 (written by k_pairlist), which yields 0
  const DevFF &ff = *ffp;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    const double packed = val[k];
    const int hi = __double2hiint(packed);
    const int inxn = hi & 0xffff;
    if (c16) c16[k] = (unsigned short)((unsigned)hi >> 16);   // the window-relative column parked beside the bond type (k_spmv_win)
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

// Row capacities without a count pass: last step's count of the same atom (looked up by global id) plus a slack, rounded to
// the row alignment; atoms the table does not know (new on this rank, or a hash collision) get the default capacity.
__global__ void k_row_caps(DevGrid g, int ntot, int natoms, CntTab ct, int slack, int defcap, int ralign, int *__restrict__ slotcnt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ntot) return;
  const int i = g.order[s];
  int cap = 0;
  if (i < natoms) {
    const int cid = g.cell_of[i];
    const int c3 = cid % g.dim[2] - g.L, c2 = (cid / g.dim[2]) % g.dim[1] - g.L, c1 = cid / (g.dim[2] * g.dim[1]) - g.L;
    if (c1 >= 0 && c1 < g.nc[0] && c2 >= 0 && c2 < g.nc[1] && c3 >= 0 && c3 < g.nc[2]) {   // rows exist in resident cells only
      const int gi = ct.gid[i];
      const int2 e = ct.tab[cnt_hash(gi, ct.mask)];
      cap = e.x == gi ? e.y + slack : defcap;
      cap = (cap + ralign - 1) & ~(ralign - 1);
    }
  }
  slotcnt[s] = cap;
}

int ensure_bond_capacity(Ctx *c, long long need);   // rxg_api.cu
int win_pick_group(Ctx *c);                          // rxg_api.cu

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
int build_pairlist(Ctx *c, bool hessian = true, bool allow_capped = false) {
  const int n = c->natoms, nt = c->cp[6];
  if (c->cfg.maxneighbs10 > 32000) { c->err = "MAXNEIGHBS10 above 32000 is not supported (20-bit row offsets in the list's fill pass)"; return RXG_ERR_ARG; }
  RXG_CUDA(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->st));
  RXG_CUDA(cudaMemsetAsync(c->d_flag + 16, 0, sizeof(int), c->st));
  RXG_CUDA(cudaMemsetAsync(c->d_acc + 33, 0, sizeof(double), c->st));
  RXG_CUDA(cudaMemsetAsync(c->rowcnt, 0, sizeof(int) * (size_t)(nt + 1), c->st));
  const bool un_on = MODE >= 1 && c->spmv_kind == 0;   // the union stream is built only for the kernel that walks it
  if (un_on) RXG_CUDA(cudaMemsetAsync(c->ucnt, 0, sizeof(int) * (size_t)(nt + 1), c->st));
  RXG_CUDA(cudaMemsetAsync(c->rowbeg, 0, sizeof(long long) * (size_t)n, c->st));
  RXG_CUDA(cudaMemsetAsync(c->rowend, 0, sizeof(long long) * (size_t)n, c->st));
  const int ncell_res = c->gnb.nc[0] * c->gnb.nc[1] * c->gnb.nc[2];
  const int grid = cdiv((long long)ncell_res * 32, PL_WARPS * 32);
  const int ralign = (MODE >= 1 && c->spmv_kind == 2 && !c->strict) ? c->win_ralign : 4;   // (RXG_WIN_RALIGN=16 puts the rows of k_spmv_win on 128-byte boundaries of the value stream: measured, no gain)
  CntTab ct;
  ct.tab = MODE >= 1 ? c->cnt_tab : nullptr; ct.mask = c->cnt_mask; ct.gid = c->gid;
  // no count pass when last step's counts are on record (QEq lists of the production path only)
  const bool capped = MODE >= 1 && allow_capped && c->caps_on && c->caps_valid && !c->strict && !un_on && c->cnt_tab;
  c->list_capped = capped;
  if (capped) c->timers_ms[23] += 1;   // list builds without a count pass
  RXG_CUDA(cudaMemsetAsync(c->d_flag + 20, 0, 2 * sizeof(int), c->st));
  RXG_CUDA(cudaMemsetAsync(c->d_flag + 26, 0, sizeof(int), c->st));
  // window SpMV: the fill pass also writes the 16-bit window-relative columns and the row lengths by slot
  const bool win_on = MODE >= 1 && c->spmv_kind == 2 && !c->strict;
  WinOut wo;
  wo.c16 = nullptr; wo.rowlen = c->rowlen; wo.desc = nullptr; wo.G = 1; wo.park = 0;
  if (win_on) {
    if (c->win_g <= 0) c->win_g = win_pick_group(c);
    wo.G = c->win_g;
    wo.c16 = c->col16;   // (allocated with col / val below, before the fill pass)
    // window descriptors of every group (k_spmv_win reads them instead of redoing the layout in every CTA of every product)
    const int ngroups = c->gnb.nc[0] * c->gnb.nc[1] * cdiv(c->gnb.nc[2], wo.G);
    const size_t need = (size_t)ngroups * (size_t)(c->nruns + 1);
    if (need > c->win_desc_cap) {
      if (c->win_desc) cudaFree(c->win_desc);
      c->win_desc_cap = need + need / 8;
      RXG_CUDA(cudaMalloc(&c->win_desc, sizeof(int2) * c->win_desc_cap));
    }
    LAUNCH(c, k_win_desc, cdiv((long long)ngroups * 32, 256), 256, 0, c->gnb, c->d_runs, c->nruns, wo.G, ngroups, c->win_desc, c->d_flag + 21);
    wo.desc = c->win_desc;
  }
  c->win_built = false;
#define RXG_PL_ARGS c->gnb, c->d_ff, c->d_runs, c->nruns, n, ncell_res, c->rowcnt, c->rowoff, c->rowbeg, c->rowend, c->col, c->val, c->cfg.maxneighbs10,   \
                    c->d_flag, (unsigned long long *)(c->d_acc + 33), ralign, c->ucnt, c->uoff, c->ucol, c->umask, c->items, c->d_flag + 17
  if (capped) LAUNCH(c, k_row_caps, cdiv(nt, 256), 256, 0, c->gnb, nt, n, ct, c->caps_slack, ((c->maxrow + 8 + ralign - 1) & ~(ralign - 1)), ralign, c->rowcnt);
  else if (un_on) LAUNCH(c, (k_pairlist<MODE, false, (MODE >= 1)>), grid, PL_WARPS * 32, 0, RXG_PL_ARGS, 4, ct, wo);
  else LAUNCH(c, (k_pairlist<MODE, false, false>), grid, PL_WARPS * 32, 0, RXG_PL_ARGS, 4, ct, wo);
  RXG_TRY(ensure_blk(c, nt));
  RXG_TRY(device_scan<long long>(c, c->rowcnt, nt, c->rowoff, c->d_blk64, (long long *)(c->d_acc + 32)));
  if (un_on) RXG_TRY(device_scan<long long>(c, c->ucnt, nt, c->uoff, c->d_blk64, (long long *)(c->d_acc + 34)));
  RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaMemcpyAsync(c->h_int + 16, c->d_flag + 16, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  if (win_on && !capped) RXG_CUDA(cudaMemcpyAsync(c->h_int + 21, c->d_flag + 21, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaMemcpyAsync(c->h_acc + 32, c->d_acc + 32, 3 * sizeof(long long), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (!capped && c->h_int[0] > c->cfg.maxneighbs10) {
    c->err = "ERROR: nbplist greater then MAXNEIGHBS10, value " + std::to_string(c->h_int[0]);
    return RXG_ERR_MAXNEIGHBS10;
  }
  long long nnz = *(long long *)(c->h_acc + 32);
  if (nnz > c->nnz_cap) {
    if (c->col) cudaFree(c->col);
    if (c->val) cudaFree(c->val);
    if (c->col16) { cudaFree(c->col16); c->col16 = nullptr; }
    c->nnz_cap = nnz + nnz / 16 + 1024;
    RXG_CUDA(cudaMalloc(&c->col, sizeof(int) * c->nnz_cap));
    RXG_CUDA(cudaMalloc(&c->val, sizeof(double) * c->nnz_cap));
  }
  if (win_on) {
    if (!c->col16) RXG_CUDA(cudaMalloc(&c->col16, sizeof(unsigned short) * c->nnz_cap));
    wo.c16 = c->col16;
    if (!capped) c->win_max = c->h_int[21];   // (capped: check_capped_flags reads it with the list's other flags)
  }
  const long long nun = un_on ? *(long long *)(c->h_acc + 34) : 0;
  if (nun > c->un_cap) {
    if (c->ucol) cudaFree(c->ucol);
    if (c->umask) cudaFree(c->umask);
    c->un_cap = nun + nun / 16 + 1024;
    RXG_CUDA(cudaMalloc(&c->ucol, sizeof(int) * c->un_cap));
    RXG_CUDA(cudaMalloc(&c->umask, (size_t)c->un_cap));
  }
  c->nnz = nnz;
  c->nunion = nun;
  if (!capped) {   // (capped: the fill pass produces them; list_stats_after_fill reads them at the CG's first synchronisation)
    c->maxrow = c->h_int[16];
    c->nnz_real = *(long long *)(c->h_acc + 33);
  }
  c->list_is_qeq = MODE >= 1;
  // rows per SpMV work item: four while four of the longest rows fit a stage of k_spmv_items, else two (12.5 A lists of PQEq)
  c->spmv_rg = c->maxrow <= 480 ? 4 : 2;
  RXG_CUDA(cudaMemsetAsync(c->d_flag + 17, 0, sizeof(int), c->st));
  const bool hfuse = MODE >= 1 && hessian && c->hess_fuse && !un_on && !capped;
  wo.park = (win_on && hessian && !hfuse) ? 1 : 0;
  if (capped) LAUNCH(c, (k_pairlist<MODE, true, false, false, (MODE >= 1)>), grid, PL_WARPS * 32, 0, RXG_PL_ARGS, c->spmv_rg, ct, wo);
  else if (un_on) LAUNCH(c, (k_pairlist<MODE, true, (MODE >= 1)>), grid, PL_WARPS * 32, 0, RXG_PL_ARGS, c->spmv_rg, ct, wo);
  else if (hfuse) LAUNCH(c, (k_pairlist<MODE, true, false, (MODE >= 1)>), grid, PL_WARPS * 32, 0, RXG_PL_ARGS, c->spmv_rg, ct, wo);
  else LAUNCH(c, (k_pairlist<MODE, true, false>), grid, PL_WARPS * 32, 0, RXG_PL_ARGS, c->spmv_rg, ct, wo);
  if (win_on) { c->win_built = true; c->win_g_built = wo.G; }
#undef RXG_PL_ARGS
  if (MODE >= 1 && c->cnt_tab) c->caps_valid = true;
  if (MODE >= 1 && hessian && !hfuse)
    LAUNCH(c, k_hessian, 148 * 16, 256, 0, nnz, c->d_ff, c->val, wo.park ? c->col16 : (unsigned short *)nullptr);
  c->nitems = un_on ? -1 : 0;   // read back with the CG's first synchronisation (spmv_launch)
  return RXG_OK;
}

// ---------------------------------------------------------------------------------------------------
// QEq CG.  d_acc slots: 0 Est, 1 hshs, 2 hsht, 3 g.h (s), 4 g.h (t), 5 sum qs, 6 sum qt, 7 g.g (s) new, 8 g.g (t) new,
// 9 g.g (s) old, 10 g.g (t) old, 11 mu, 12-13 PQEq ghost-column sums of the iteration, 14-17 PQEq (rxg_pqeq.cuh),
// 20-24 the CG's control block (k_cg_ctrl): Est of the previous iteration, stopped flag, completed iterations, lmin_s, lmin_t.
constexpr int ACC_GEST2 = 20, ACC_DONE = 21, ACC_NITER = 22, ACC_LMIN = 23;
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
// ---------------------------------------------------------------------------------------------------
// Bulk-copy (TMA) helpers.  Rows lie in HBM in cell order, so the rows of consecutive slots form ONE contiguous span of
// (val, col); rows start on 4-entry boundaries (16 B for col, 32 B for val).

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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n"   // suspend-time hint: the wait may park the warp
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
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
                                                          double4 *__restrict__ rowsum, const double *__restrict__ acc, int stage,
                                                          const int *__restrict__ grp) {
  constexpr int CAP = ROWS * CAPROW;   // CAPROW = longest row the staged path takes (480: 10 A lists; 1216: the 12.5 A lists of PQEq)
  const double cg_done = acc[ACC_DONE];   // tested below, so that this load flies together with the row-offset loads
  __shared__ __align__(128) double s_val[CAP];
  __shared__ __align__(128) int s_col[CAP];
  __shared__ __align__(8) unsigned long long bar;
  const int sub = threadIdx.x % LPR, rowid = threadIdx.x / LPR;
  const int slot0 = (grp ? grp[blockIdx.x] : blockIdx.x) * ROWS;   // grp: the interior or the boundary row groups only (spmv_launch)
  const int slot1 = min(slot0 + ROWS, ntot);
  const long long sb = rowoff[slot0], se = rowoff[slot1];
  if (cg_done != 0.0) return;   // the CG has stopped (k_cg_ctrl): iterations enqueued ahead of the host's check do nothing
  const int span = (int)(se - sb);
  const bool staged = stage && span > 0 && span <= CAP;   // stage == 0: every CTA reads straight from HBM (the path long rows take)
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

// ---------------------------------------------------------------------------------------------------
// Window SpMV (k_spmv_win).  k_spmv_rows is bound by the L1 data pipe, not by HBM: every stored entry costs a 16-byte gather of
// x through the L1 tag stage (7.5 wavefronts per warp request, half of them L2 round trips), on top of the staged stream.
// Here a CTA owns a GROUP of up to G consecutive resident cells of one z-column.  All rows of the group draw their columns
// from the same few stencil runs, so the CTA first copies the group's WINDOW of x -- one cp.async.bulk (TMA, UBLKCP) per
// stencil run, ~70 copies of ~0.7 KB, 40-75 KB in all -- into shared memory, and the rows then gather x from shared memory
// (29-cycle LDS.128, no tag stage, no L2 round trip) through the 16-bit window-relative column stream that k_pairlist's fill
// pass writes (struct WinOut): bits 0-14 position in the window, bit 15 ghost flag.  The matrix stream is 8 B of value +
// 2 B of column per entry, read once, coalesced, straight into registers (evict-first), eight entries in flight per lane;
// a warp takes whole rows, so a row sum needs one warp reduction and no atomics.  A group whose window does not fit the
// shared-memory budget (or 15 bits) falls back to the 32-bit columns and global gathers for that CTA only.
// Row table: start (rowoff) and exact length (rowlen, written by the fill pass) by cell-order slot, so that nothing on the
// path depends on the atom-order indirection.
// the rows of one group, streamed by one warp: rows wid, wid + NW, ... .  GH: the group may take ghost columns (bit 15 of a column)
template <int NW, int U, bool GH>
__device__ __forceinline__ void win_rows(const double2 *__restrict__ s_x, int a0, int nrows, int lane, int wid, const long long *__restrict__ rowoff,
                                         const int *__restrict__ rowlen, const unsigned short *__restrict__ col16,
                                         const double *__restrict__ val, double4 *__restrict__ rowsum, unsigned long long *bar) {
  int t = wid;
  if (t >= nrows) { mbar_wait(bar, 0); return; }
  long long rs = __ldg(rowoff + a0 + t);
  int n = __ldg(rowlen + a0 + t);
  // the next row's start and length are requested one row ahead, so their round trip hides behind the current row
  long long rs_next = 0;
  int n_next = 0;
  if (t + NW < nrows) { rs_next = __ldg(rowoff + a0 + t + NW); n_next = __ldg(rowlen + a0 + t + NW); }
  double h[U];
  unsigned short c[U];
  int k0 = 0;
  auto load = [&]() {
    const double *pv = val + rs + k0 + lane;
    const unsigned short *pc = col16 + rs + k0 + lane;
    const int rem = n - k0 - lane;
#pragma unroll
    for (int u = 0; u < U; u++) {
      h[u] = 0.0; c[u] = 0;
      if (32 * u < rem) { h[u] = __ldcs(pv + 32 * u); c[u] = __ldcs(pc + 32 * u); }
    }
  };
  load();                // the first batch is requested BEFORE the wait for the window: the two round trips overlap
  mbar_wait(bar, 0);
  double a = 0.0, b = 0.0, ga = 0.0, gb = 0.0;
  const char *sxb = reinterpret_cast<const char *>(s_x);
  for (;;) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned cu = c[u];
      const double2 v = *reinterpret_cast<const double2 *>(sxb + (GH ? ((cu << 4) & 0x7fff0u) : (cu << 4)));
      a = fma(h[u], v.x, a);
      b = fma(h[u], v.y, b);
      if (GH && (cu & 0x8000u)) { ga = fma(h[u], v.x, ga); gb = fma(h[u], v.y, gb); }
    }
    k0 += 32 * U;
    if (k0 >= n) {
      warp_sum4(a, b, ga, gb, lane);
      if (lane == 0 && n >= 0) rowsum[a0 + t] = make_double4(a, b, ga, gb);   // n < 0: a ghost in a resident cell owns no row
      a = b = ga = gb = 0.0;
      t += NW;
      if (t >= nrows) break;
      rs = rs_next; n = n_next; k0 = 0;
      if (t + NW < nrows) { rs_next = __ldg(rowoff + a0 + t + NW); n_next = __ldg(rowlen + a0 + t + NW); }
    }
    load();
  }
}

// Short rows (sparse systems: ~120 entries per row): a full warp per row spends more instructions on the row's reduction and
// bookkeeping than on its entries.  Here a warp takes TWO rows at a time, 16 lanes each, in lockstep (the number of batches is
// the longer row's), and one 9-shuffle reduction serves both rows.
template <int NW, int U, bool GH>
__device__ __forceinline__ void win_rows_half(const double2 *__restrict__ s_x, int a0, int nrows, int lane, int wid, const long long *__restrict__ rowoff,
                                              const int *__restrict__ rowlen, const unsigned short *__restrict__ col16,
                                              const double *__restrict__ val, double4 *__restrict__ rowsum, unsigned long long *bar) {
  const int sub = lane & 15, half = lane >> 4;
  int t = 2 * wid + half;                       // this half-warp's row; the warp advances by 2 NW rows per pass
  if (2 * wid >= nrows) { mbar_wait(bar, 0); return; }
  long long rs = 0;
  int n = -1;
  if (t < nrows) { rs = __ldg(rowoff + a0 + t); n = __ldg(rowlen + a0 + t); }
  long long rs_next = 0;
  int n_next = -1;
  if (t + 2 * NW < nrows) { rs_next = __ldg(rowoff + a0 + t + 2 * NW); n_next = __ldg(rowlen + a0 + t + 2 * NW); }
  double h[U];
  unsigned short c[U];
  int k0 = 0;
  auto load = [&]() {
    const double *pv = val + rs + k0 + sub;
    const unsigned short *pc = col16 + rs + k0 + sub;
    const int rem = n - k0 - sub;
#pragma unroll
    for (int u = 0; u < U; u++) {
      h[u] = 0.0; c[u] = 0;
      if (16 * u < rem) { h[u] = __ldcs(pv + 16 * u); c[u] = __ldcs(pc + 16 * u); }
    }
  };
  load();
  mbar_wait(bar, 0);
  const char *sxb = reinterpret_cast<const char *>(s_x);
  for (;;) {   // one pass = two rows
    const int nmax = max(n, __shfl_xor_sync(0xffffffffu, n, 16));
    double a = 0.0, b = 0.0, ga = 0.0, gb = 0.0;
    for (;;) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned cu = c[u];
        const double2 v = *reinterpret_cast<const double2 *>(sxb + (GH ? ((cu << 4) & 0x7fff0u) : (cu << 4)));
        a = fma(h[u], v.x, a);
        b = fma(h[u], v.y, b);
        if (GH && (cu & 0x8000u)) { ga = fma(h[u], v.x, ga); gb = fma(h[u], v.y, gb); }
      }
      k0 += 16 * U;
      if (k0 >= nmax) break;   // warp-uniform
      load();
    }
    // four sums per half-warp: halve what a lane carries twice, then two plain steps; totals on lanes 0 / 4 / 8 / 12 of the half
    {
      const bool up8 = lane & 8, up4 = lane & 4;
      const double s0 = up8 ? a : ga, s1 = up8 ? b : gb;
      double k0v = up8 ? ga : a, k1v = up8 ? gb : b;
      k0v += __shfl_xor_sync(0xffffffffu, s0, 8);
      k1v += __shfl_xor_sync(0xffffffffu, s1, 8);
      const double give = up4 ? k0v : k1v;
      double u = up4 ? k1v : k0v;
      u += __shfl_xor_sync(0xffffffffu, give, 4);
      u += __shfl_xor_sync(0xffffffffu, u, 2);
      u += __shfl_xor_sync(0xffffffffu, u, 1);
      const int base = lane & 16;
      a = __shfl_sync(0xffffffffu, u, base);
      b = __shfl_sync(0xffffffffu, u, base + 4);
      ga = __shfl_sync(0xffffffffu, u, base + 8);
      gb = __shfl_sync(0xffffffffu, u, base + 12);
    }
    if (sub == 0 && n >= 0) rowsum[a0 + t] = make_double4(a, b, ga, gb);   // n < 0: no row (past the group, or a ghost's slot)
    t += 2 * NW;
    if (t - half >= nrows) break;   // warp-uniform: the pass's first row is past the group
    rs = rs_next; n = n_next; k0 = 0;
    rs_next = 0; n_next = -1;
    if (t + 2 * NW < nrows) { rs_next = __ldg(rowoff + a0 + t + 2 * NW); n_next = __ldg(rowlen + a0 + t + 2 * NW); }
    load();
  }
}

template <int NW, int U, int MINB, int LPR = 32>
__global__ void __launch_bounds__(NW * 32, MINB) k_spmv_win(DevGrid g, int nruns, int G, int reach, const int2 *__restrict__ desc,
                                                            const long long *__restrict__ rowoff, const int *__restrict__ rowlen,
                                                            const int *__restrict__ col, const unsigned short *__restrict__ col16,
                                                            const double *__restrict__ val, const double2 *__restrict__ x,
                                                            double4 *__restrict__ rowsum, const double *__restrict__ acc, int wcap) {
  extern __shared__ __align__(128) unsigned char win_smem[];
  double2 *s_x = reinterpret_cast<double2 *>(win_smem);
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, NW);   // one arrival (with its share of the expected bytes) per warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double cg_done = acc[ACC_DONE];
  const int ngz = (g.nc[2] + G - 1) / G;
  const int gz = blockIdx.x % ngz, c2 = (blockIdx.x / ngz) % g.nc[1], c1 = blockIdx.x / (ngz * g.nc[1]);
  const int g0 = gz * G, g1 = min(g0 + G, g.nc[2]) - 1;
  const int cid0 = ((c1 + g.L) * g.dim[1] + (c2 + g.L)) * g.dim[2] + (g0 + g.L);
  const int a0 = g.start[cid0], a1 = g.start[cid0 + (g1 - g0) + 1];
  // Two load chains start here side by side: window descriptors -> bulk copies of x (this block), and group bounds -> row
  // start/length -> first batch of the matrix stream (win_rows); neither waits for the other until the first gather.  Warp w
  // issues the copies of the runs r = w (mod NW): a bulk copy takes uniform operands, so a warp issues its copies one lane at
  // a time.
  const int2 *d = desc + (size_t)blockIdx.x * (nruns + 1);
  // (this lane's first descriptor is requested together with the window's total, not after the decision that depends on it)
  const int r_first = wid + NW * lane;
  int2 e_first = make_int2(0, 0);
  if (r_first < nruns) e_first = __ldg(d + r_first);
  const int tot = __ldg(&d[nruns].y) >> 12;
  const bool staged = tot <= wcap && tot <= 32768;
  {
    unsigned issued = 0;
    if (staged && cg_done == 0.0 && a1 > a0) {   // (a group without rows -- vacuum -- copies nothing)
      for (int r = r_first; r < nruns; r += NW * 32) {
        const int2 e = r == r_first ? e_first : __ldg(d + r);
        const int wlen = e.y & 0xfff, pos = e.y >> 12;
        if (wlen > 0) {
          bulk_g2s(s_x + pos, x + e.x, (unsigned)wlen * 16u, &bar);
          issued += (unsigned)wlen * 16u;
        }
      }
    }
    issued = __reduce_add_sync(0xffffffffu, issued);
    // (a copy may complete before its warp's arrive: the phase cannot, it needs all NW arrivals; the tx-count is signed)
    if (lane == 0) mbar_expect_tx(&bar, issued);
  }
  const int nrows = (cg_done != 0.0) ? 0 : a1 - a0;   // the CG has stopped (k_cg_ctrl): nothing to do (no copy was issued either)
  if (staged) {
    // a group at least `reach` cells inside the resident grid takes no ghost column: its rows skip the ghost sums altogether
    const bool inner = c1 >= reach && c1 < g.nc[0] - reach && c2 >= reach && c2 < g.nc[1] - reach && g0 >= reach && g1 < g.nc[2] - reach;
    if (LPR == 16) {
      if (inner) win_rows_half<NW, U, false>(s_x, a0, nrows, lane, wid, rowoff, rowlen, col16, val, rowsum, &bar);
      else win_rows_half<NW, U, true>(s_x, a0, nrows, lane, wid, rowoff, rowlen, col16, val, rowsum, &bar);
    } else {
      if (inner) win_rows<NW, U, false>(s_x, a0, nrows, lane, wid, rowoff, rowlen, col16, val, rowsum, &bar);
      else win_rows<NW, U, true>(s_x, a0, nrows, lane, wid, rowoff, rowlen, col16, val, rowsum, &bar);
    }
  } else {
    mbar_wait(&bar, 0);
    // the window does not fit: 32-bit columns, x gathered from global memory (this CTA only)
    for (int t = wid; t < nrows; t += NW) {
      const long long rs = rowoff[a0 + t];
      const int n = rowlen[a0 + t];
      if (n < 0) continue;
      const double *pv = val + rs;
      const int *pc = col + rs;
      double a = 0.0, b = 0.0, ga = 0.0, gb = 0.0;
      for (int k = lane; k < n; k += 32) {
        const double hh = __ldcs(pv + k);
        const int cc = __ldcs(pc + k);
        const double2 v = x[cc & COL_MASK];
        const double pa = hh * v.x, pb = hh * v.y;
        a += pa; b += pb;
        if (cc < 0) { ga += pa; gb += pb; }
      }
      warp_sum4(a, b, ga, gb, lane);
      if (lane == 0) rowsum[a0 + t] = make_double4(a, b, ga, gb);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Cell-blocked SpMV (k_spmv_items; experiment, RXG_SPMV=items: correct, tested, slower -- DESIGN.md 4.3).  The rows of one cell take nearly the same columns (the union of the lists of the ~3 atoms
// of a 3 A cell is 0.44x the sum of their lengths at RDX density), and k_spmv_rows pays one 16-byte gather of x per stored
// entry: its L1 data pipe, not HBM, is what saturates.  Here a work item (SpItem, written by k_pairlist's fill pass) is up
// to RG consecutive rows of a block of <= 8 rows of one cell, with the block's UNION stream: per union entry a column (4 B)
// and the 8-bit set of rows that hold it (1 B).  A lane gathers x[col] once and feeds it to every row of the set; the row's
// value sits at the row's running position in its own compacted value stream (ballot + popc), so the fp64 values are stored
// once, exactly as CSR stores them: 8 B of value + 5 B x 0.44 of column/mask = 10.2 B per stored entry instead of 12 B, and
// 0.44 gathers instead of 1.
// Execution: ONE persistent CTA per SM -- a producer warp and SI_CONS consumer warps around a shared-memory ring (~200 KB)
// that is allocated item by item in exactly the bytes each item needs.  The producer's elected lane reads item records
// (three items ahead, in registers), reserves ring space, and issues three bulk async copies per item (cp.async.bulk, UBLKCP
// in SASS: values, columns, row sets) that complete on the slot's `full` mbarrier; space is reclaimed in item order as the
// consumers signal `empty`.  Consumer warp w owns items w, w+SI_CONS, ... of its CTA and walks each alone: no cross-warp
// reduction, no CTA barrier.  Per 128 union entries it issues four gathers of x, and while they are in flight multiplies out
// the previous 128 from shared memory (branch-free: a row that lacks the column multiplies by a zero value).  HBM always has
// the ring's free part (tens of KB per SM) in flight, whatever the gather latency of the walks.
// An item that does not fit the ring is walked straight from global memory.
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int SI_CONS = 12;    // consumer warps per CTA
constexpr int SI_SLOTS = 32;   // items in flight per CTA (descriptors); power of two
struct SiSlot { int off; int staged; };   // ring offset of the item's data; 0 = read from global memory

template <int RG, bool STAGED>
__device__ __forceinline__ void spmv_item_walk(const SpItem &I, const double *__restrict__ vv, const int *__restrict__ uc,
                                               const unsigned char *__restrict__ um, const double2 *__restrict__ x,
                                               double4 *__restrict__ rowsum, int lane) {
  const unsigned lt = (1u << lane) - 1u;
  int pos[RG];
  double sa[RG], sb[RG], sga[RG], sgb[RG];
#pragma unroll
  for (int j = 0; j < RG; j++) { pos[j] = I.rs[j]; sa[j] = 0.0; sb[j] = 0.0; sga[j] = 0.0; sgb[j] = 0.0; }
  const int un = I.un, rshift = I.rshift, nrp = I.nrp;
  constexpr int G = 4;   // steps (of 32 union entries) per group: G gathers of x in flight per lane
  int c[G], cn[G];
  unsigned mm[G], mn[G];
  double2 xv[G], xn[G];
  auto load_group = [&](int u0, int (&cc)[G], unsigned (&mk)[G], double2 (&xx)[G]) {
#pragma unroll
    for (int t = 0; t < G; t++) {
      const int u = u0 + 32 * t + lane;
      cc[t] = 0; mk[t] = 0u;
      if (u < un) {
        cc[t] = STAGED ? uc[u] : __ldcs(uc + u);
        mk[t] = ((unsigned)(STAGED ? um[u] : __ldcs(um + u)) >> rshift) & ((1u << RG) - 1u);
      }
    }
#pragma unroll
    for (int t = 0; t < G; t++) {
      xx[t] = make_double2(0.0, 0.0);
      if (mk[t]) xx[t] = x[cc[t] & COL_MASK];
    }
  };
  load_group(0, c, mm, xv);
  for (int u0 = 0; u0 < un; u0 += 32 * G) {
    if (u0 + 32 * G < un) load_group(u0 + 32 * G, cn, mn, xn);   // next group's gathers fly while this one is multiplied out
#pragma unroll
    for (int t = 0; t < G; t++) {
      const double gx = c[t] < 0 ? xv[t].x : 0.0, gy = c[t] < 0 ? xv[t].y : 0.0;   // bit 31 = ghost column (Est weighting)
#pragma unroll
      for (int j = 0; j < RG; j++) {
        if (j < nrp) {   // warp-uniform
          const bool bit = (mm[t] >> j) & 1u;
          const unsigned bal = __ballot_sync(0xffffffffu, bit);
          double h = 0.0;
          if (bit) { const int k = pos[j] + __popc(bal & lt); h = STAGED ? vv[k] : __ldcs(vv + k); }
          sa[j] = fma(h, xv[t].x, sa[j]); sb[j] = fma(h, xv[t].y, sb[j]);
          sga[j] = fma(h, gx, sga[j]); sgb[j] = fma(h, gy, sgb[j]);
          pos[j] += __popc(bal);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < G; t++) { c[t] = cn[t]; mm[t] = mn[t]; xv[t] = xn[t]; }
  }
#pragma unroll
  for (int j = 0; j < RG; j++) {
    if (j < nrp) {
      warp_sum4(sa[j], sb[j], sga[j], sgb[j], lane);
      if (lane == 0) rowsum[I.slot0 + j] = make_double4(sa[j], sb[j], sga[j], sgb[j]);
    }
  }
}

template <int RG>
__global__ void __launch_bounds__((SI_CONS + 1) * 32, 1) k_spmv_items(const SpItem *__restrict__ items, int nitems,
                                                                      const int *__restrict__ ucol, const unsigned char *__restrict__ umask,
                                                                      const double *__restrict__ val, const double2 *__restrict__ x,
                                                                      double4 *__restrict__ rowsum, const double *__restrict__ acc, int ring_bytes,
                                                                      int stage_on) {
  static_assert(RG <= 4, "at most four rows per item");
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) unsigned long long full[SI_SLOTS], empty[SI_SLOTS];
  __shared__ SiSlot slots[SI_SLOTS];
  __shared__ __align__(16) SpItem recs[SI_SLOTS];
  if (acc[ACC_DONE] != 0.0) return;   // the CG has stopped (k_cg_ctrl)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SI_SLOTS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int g = gridDim.x;
  if (w == SI_CONS) {
    // ---- producer: reserve ring space in item order, copy, publish; reclaim space in item order
    if (lane == 0) {
      int4 qa[4], qb[4], qc[4];
      auto fetch = [&](int4 (&q)[4], int it) {
        if (it < nitems) {
          const int4 *src = reinterpret_cast<const int4 *>(items + it);
          q[0] = __ldg(src); q[1] = __ldg(src + 1); q[2] = __ldg(src + 2); q[3] = __ldg(src + 3);
        }
      };
      fetch(qa, blockIdx.x); fetch(qb, blockIdx.x + g); fetch(qc, blockIdx.x + 2 * g);
      int head = 0;          // next free byte of the ring
      int tail = 0;          // oldest item whose space is still held (item index k of this CTA)
      int k = 0;
      for (int it = blockIdx.x; it < nitems; it += g, k++) {
        const int s = k & (SI_SLOTS - 1);
        const long long voff = ((long long)(unsigned)qa[0].x) | ((long long)qa[0].y << 32);
        const long long uoff = ((long long)(unsigned)qa[0].z) | ((long long)qa[0].w << 32);
        const int un = qa[1].x, vlen = qa[1].w;
        const int bv = (vlen * 8 + 127) & ~127, bc = (un * 4 + 127) & ~127, bm = (un + 127) & ~127;
        const int need = bv + bc + bm;
        const bool staged = stage_on && vlen > 0 && un > 0 && need <= ring_bytes / 2;
        // Wait for the slot's previous user (item k - SI_SLOTS) and for `need` contiguous bytes.  Live data is the circular
        // interval [start of the oldest unreleased item, head); an unstaged item holds zero bytes at the head of its time.
        const int nb = staged ? need : 0;
        int off = head;
        for (;;) {
          bool ok = k - tail < SI_SLOTS;
          if (ok && nb > 0) {
            if (tail == k) { head = 0; off = 0; }                     // nothing live: restart at the ring's origin
            else {
              const int tail_off = slots[tail & (SI_SLOTS - 1)].off;
              if (head >= tail_off) {                                  // live data does not wrap
                if (head + nb <= ring_bytes) off = head;
                else if (nb < tail_off) off = 0;                       // leave the ring's end unused and wrap
                else ok = false;
              } else {                                                 // live data wraps: free = [head, tail_off)
                if (head + nb < tail_off) off = head; else ok = false;
              }
            }
          }
          if (ok) break;
          mbar_wait(&empty[tail & (SI_SLOTS - 1)], (tail / SI_SLOTS) & 1);   // reclaim the oldest item's space
          tail++;
        }
        head = off + nb;
        slots[s].off = off; slots[s].staged = staged ? 1 : 0;
        int4 *dst = reinterpret_cast<int4 *>(&recs[s]);
        dst[0] = qa[0]; dst[1] = qa[1]; dst[2] = qa[2]; dst[3] = qa[3];
        if (staged) {
          mbar_expect_tx(&full[s], (unsigned)vlen * 8u + (unsigned)un * 5u);   // (arrives and sets the byte count)
          bulk_g2s(ring + off, val + voff, (unsigned)vlen * 8u, &full[s]);
          bulk_g2s(ring + off + bv, ucol + uoff, (unsigned)un * 4u, &full[s]);
          bulk_g2s(ring + off + bv + bc, umask + uoff, (unsigned)un, &full[s]);
        } else {
          mbar_arrive(&full[s]);   // nothing to wait for: the consumer reads this item from global memory
        }
#pragma unroll
        for (int q = 0; q < 4; q++) { qa[q] = qb[q]; qb[q] = qc[q]; }
        fetch(qc, it + 3 * g);
      }
    }
  } else {
    for (int k = w, it = blockIdx.x + w * g; it < nitems; k += SI_CONS, it += SI_CONS * g) {
      const int s = k & (SI_SLOTS - 1);
      mbar_wait(&full[s], (k / SI_SLOTS) & 1);
      const SpItem I = recs[s];
      const int off = slots[s].off;
      if (slots[s].staged) {
        const int bv = (I.vlen * 8 + 127) & ~127, bc = (I.un * 4 + 127) & ~127;
        spmv_item_walk<RG, true>(I, reinterpret_cast<const double *>(ring + off), reinterpret_cast<const int *>(ring + off + bv),
                                 ring + off + bv + bc, x, rowsum, lane);
      } else {
        spmv_item_walk<RG, false>(I, val + I.voff, ucol + I.uoff, umask + I.uoff, x, rowsum, lane);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Interior / boundary split of the CG's sparse product (multi-rank).  A row whose cell lies at least `lay` cells (the
// stencil's reach) inside the resident grid along every axis has no ghost column, so its product does not depend on the
// ghost refresh of (hs,ht) and can run WHILE the neighbours' values travel; the other rows follow once they have arrived.
// k_group_class: class of each group of `rows` consecutive slots (what one CTA of k_spmv_rows takes): 1 = every slot in an
// interior cell (or a ghost slot, which owns no row), 0 = boundary.  k_group_lists compacts both classes.
__global__ void k_group_class(DevGrid g, int ntot, int rows, int lay, int *__restrict__ cls) {
  const int grp = blockIdx.x * blockDim.x + threadIdx.x;
  if ((long long)grp * rows >= ntot) return;
  int interior = 1;
  for (int s = grp * rows; s < min(ntot, (grp + 1) * rows); s++) {
    const int cid = g.cell_of[g.order[s]];
    const int c3 = cid % g.dim[2] - g.L, c2 = (cid / g.dim[2]) % g.dim[1] - g.L, c1 = cid / (g.dim[2] * g.dim[1]) - g.L;
    const bool res = c1 >= 0 && c1 < g.nc[0] && c2 >= 0 && c2 < g.nc[1] && c3 >= 0 && c3 < g.nc[2];
    if (!res) continue;   // ghost slot: no row
    if (c1 < lay || c1 >= g.nc[0] - lay || c2 < lay || c2 >= g.nc[1] - lay || c3 < lay || c3 >= g.nc[2] - lay) interior = 0;
  }
  cls[grp] = interior;
}
__global__ void k_group_lists(int ngrp, const int *__restrict__ cls, const int *__restrict__ off, int *__restrict__ lst_int,
                              int *__restrict__ lst_bnd) {
  const int grp = blockIdx.x * blockDim.x + threadIdx.x;
  if (grp >= ngrp) return;
  if (cls[grp]) lst_int[off[grp]] = grp;       // off = exclusive scan of cls
  else lst_bnd[grp - off[grp]] = grp;
}

template <bool INIT>
__global__ void __launch_bounds__(256) k_cg_dots(const int *__restrict__ order, int ntot, int natoms, const double4 *__restrict__ rowsum,
                                                 const double2 *__restrict__ x, const double *__restrict__ q,
                                                 double2 *__restrict__ gst, double2 *__restrict__ tst, double2 *__restrict__ ust,
                                                 double2 *__restrict__ wst, const int *__restrict__ itype,
                                                 const DevFF *__restrict__ ffp, double *__restrict__ acc) {
  if (!INIT && acc[ACC_DONE] != 0.0) return;
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

// The CG's control step on the device (one thread, after the all-reduce of the five dots): the stop rule of
// src/qeq.F90:114-115 on Est, the real(4) step lengths of :133, and the roll of the g.g sums -- what the host
// did between two synchronisations per iteration in round 1.  Once the stop flag is set, the iteration's remaining kernels
// and every later iteration already enqueued return at their first instruction, so the host only looks at the flag every
// few iterations.  pq: PQEq's ghost-column recurrences (qs_ghost += lmin_s hs_ghost => acc[14] += lmin_s * acc[12]).
__global__ void k_cg_ctrl(double *__restrict__ acc, double tol, int pq) {
  if (acc[ACC_DONE] != 0.0) return;
  const double GEst1 = acc[0], GEst2 = acc[ACC_GEST2];
  bool stop = 0.5 * (fabs(GEst2) + fabs(GEst1)) < tol;                                        // src/qeq.F90:114
  if (!stop && fabs(GEst2) > 0.0 && fabs(__ddiv_rn(GEst1, GEst2) - 1.0) < tol) stop = true;   // src/qeq.F90:115
  if (stop) { acc[ACC_DONE] = 1.0; return; }
  acc[ACC_GEST2] = GEst1;
  const float lmin_s = (float)__ddiv_rn(acc[3], acc[1]);   // real(4) :: lmin, src/qeq.F90:23,133
  const float lmin_t = (float)__ddiv_rn(acc[4], acc[2]);
  acc[ACC_LMIN] = (double)lmin_s; acc[ACC_LMIN + 1] = (double)lmin_t;
  if (pq) { acc[14] += (double)lmin_s * acc[12]; acc[15] += (double)lmin_t * acc[13]; }
  acc[9] = acc[7]; acc[10] = acc[8];
  acc[5] = 0.0; acc[6] = 0.0; acc[7] = 0.0; acc[8] = 0.0;
  acc[0] = 0.0; acc[1] = 0.0; acc[2] = 0.0; acc[3] = 0.0; acc[4] = 0.0; acc[12] = 0.0; acc[13] = 0.0;   // the next iteration's dots start from zero
  acc[ACC_NITER] += 1.0;
}
// qs,qt step (src/qeq.F90:136-137) + gradient / Est-bookkeeping recurrences + partial sums (sum qs, sum qt, g.g)
__global__ void __launch_bounds__(256) k_cg_update1(int natoms, const double2 *__restrict__ hst,
                                                    const double2 *__restrict__ tst, const double2 *__restrict__ ust,
                                                    double2 *__restrict__ qst, double2 *__restrict__ gst,
                                                    double2 *__restrict__ wst, double *__restrict__ acc) {
  if (acc[ACC_DONE] != 0.0) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double part[4] = {0.0, 0.0, 0.0, 0.0};
  if (i < natoms) {
    const double ls = acc[ACC_LMIN], lt = acc[ACC_LMIN + 1];   // real(4) lmin promoted (k_cg_ctrl), SURVEY Q3
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
  if (acc[ACC_DONE] != 0.0) return;
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
__global__ void k_flag_to_acc(const int *__restrict__ flag, double *__restrict__ acc) { *acc = *flag ? 1.0 : 0.0; }
// hs = gs, ht = gt (src/qeq.F90:90-91); also arms the CG's control block (GEst2 = 1e99, :94)
__global__ void k_h_from_g2(int natoms, const double2 *__restrict__ gst, double2 *__restrict__ hst, double2 *__restrict__ xs,
                            const int *__restrict__ slot_of, double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    acc[ACC_GEST2] = 1e99; acc[ACC_DONE] = 0.0; acc[ACC_NITER] = 0.0;
    acc[0] = 0.0; acc[1] = 0.0; acc[2] = 0.0; acc[3] = 0.0; acc[4] = 0.0; acc[5] = 0.0; acc[6] = 0.0; acc[12] = 0.0; acc[13] = 0.0;
  }
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
                                   const int *__restrict__ itype, const DevFF *__restrict__ ffp, double2 *__restrict__ gst,
                                   const double *__restrict__ fpq) {
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
  double g1 = sub_rn(sub_rn(-chi, mul_rn(eta, x.x)), gs);
  if (fpq) g1 = sub_rn(g1, fpq[i]);   // PQEq: - fpqeq(i), src/pqeq.F90:463
  gst[i] = make_double2(g1, sub_rn(sub_rn(-1.0, mul_rn(eta, x.y)), gt));
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
