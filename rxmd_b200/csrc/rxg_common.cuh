// rxg_common.cuh -- shared declarations of the B200 hot-path library (device context, helpers).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>   // types only: the library is dlopen()ed at rxg_comm_init (see NcclApi)
#include <dlfcn.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>
#include "../../include/rxmd_b200.h"

#define RXG_MAXLAYERS 5       // reference src/module.F90:44
#define RXG_MAXLAYERS_NB 10   // reference src/module.F90:45

namespace rxg {

// NCCL entry points resolved at run time.  Linking libnccl at load time would pin whichever libnccl.so.2 the loader
// finds first; a host process that loads its own NCCL later (PyTorch bundles a newer one) would then break.  With
// dlopen at rxg_comm_init the library joins whatever NCCL the process already uses, or the system one otherwise.
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string &err) {
    if (handle) return true;
    const char *env = getenv("RXG_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n) continue;
      handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (handle) break;
    }
    if (!handle) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define RXG_SYM(field, name)                                                   \
  field = reinterpret_cast<decltype(field)>(dlsym(handle, name));              \
  if (!field) { err = std::string("libnccl lacks ") + name; return false; }
    RXG_SYM(GetUniqueId, "ncclGetUniqueId") RXG_SYM(CommInitRank, "ncclCommInitRank") RXG_SYM(CommDestroy, "ncclCommDestroy")
    RXG_SYM(Send, "ncclSend") RXG_SYM(Recv, "ncclRecv") RXG_SYM(AllReduce, "ncclAllReduce") RXG_SYM(AllGather, "ncclAllGather") RXG_SYM(GroupStart, "ncclGroupStart")
    RXG_SYM(GroupEnd, "ncclGroupEnd") RXG_SYM(GetErrorString, "ncclGetErrorString")
#undef RXG_SYM
    return true;
  }
};
inline NcclApi &nccl_api() { static NcclApi a; return a; }

// ---- round-to-nearest fp64 arithmetic that ptxas may not contract into FMA.  Used wherever a result
// feeds a comparison or an index (cell ids, cut-off tests, table weights) so that it is bit-identical to
// an x86-64 build of the reference without FMA (hard part 3).
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
// gfortran sum(a(1:3)*b(1:3)) == ((a1*b1 + a2*b2) + a3*b3)
__device__ __forceinline__ double dot3_rn(double a0, double a1, double a2, double b0, double b1, double b2) {
  return add_rn(add_rn(mul_rn(a0, b0), mul_rn(a1, b1)), mul_rn(a2, b2));
}
__device__ __forceinline__ double dist2_rn(double dx, double dy, double dz) {
  return add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
}
__device__ __forceinline__ int nint_d(double x) { return (int)llround(x); }
// One 32-byte record in ONE load instruction (LDG.E.256 on sm_100a) through the read-only path.  nvcc splits a double4 load
// into two 16-byte requests; the gather-bound kernels (table nodes, candidate records, {x,y,z,q} packs) pay the L1 tag stage
// per request, so halving the requests is what counts.  The record must be 32-byte aligned and not written by the kernel.
__device__ __forceinline__ double4 ldg256(const double4 *p) {
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

// Force-field tables resident in HBM (device pointers); filled by rxg_set_forcefield from rxg_ff.
// Indexing is 0-based here: type t = ity-1, bond type x = inxn-1; inxn* tables keep 1-based VALUES (0 = none).
struct DevFF {
  int nso, nboty, nvaty, ntoty, nhbty, ntable;
  double vpar1, vpar2, cutoff_vpar30, rctap, rctap2, UDR, UDRi;
  const double *Val, *Valval, *Valangle, *Vale, *mass, *plp1, *plp2, *nlpopt;
  const double *povun2, *povun3, *povun4, *povun5, *povun6, *povun7, *povun8, *pval3, *pval5, *chi, *eta;
  const double *cBOp1, *cBOp3, *cBOp5, *pbo2h, *pbo4h, *pbo6h, *pbo2, *pbo4, *pbo6, *swtch;
  const double *rc2, *pboc1, *pboc3, *pboc4, *pboc5, *ovc, *v13cor, *Desig, *Depi, *Depipi, *pbe1, *pbe2, *povun1;
  const double *theta00, *pval1, *pval2, *pval4, *pval6, *pval7, *pval8, *pval9, *pval10;
  const double *ppen1, *ppen2, *ppen3, *ppen4, *pcoa1, *pcoa2, *pcoa3, *pcoa4;
  const double *ptor1, *ptor2, *ptor3, *ptor4, *V1, *V2, *V3, *pcot1, *pcot2;
  const double *phb1, *phb2, *phb3, *r0hb;
  const int *inxn2, *inxn3, *inxn3hb, *inxn4;
  const double *TBL_Eclmb_QEq;   // (NTABLE, nboty) column-major, as given
  const double2 *TBL_qeq2;       // [(inxn-1)*NTABLE + (itb-1)] = {T(itb,inxn), T(itb+1,inxn)}
  const double4 *TBL_nb;         // [(inxn-1)*NTABLE + (itb-1)] = {Evdw, CEvdw, Eclmb, CEclmb}: one 32-byte sector per node
  // PQEq (module pqeq_vars, reference src/module.F90:285-304); ntype_pqeq == 0 when isPQEq is off
  int ntype_pqeq;
  const int *isPolarizable, *inxnpqeq;   // [ntype_pqeq], (ntype_pqeq,ntype_pqeq) column-major, 1-based values
  const double *Zpqeq, *Kspqeq;
  // TBL_Eclmb_pcc/psc/pss re-packed: [(inxn-1)*NTABLE + (itb-1)] = {E(itb), E(itb+1), dE(itb), dE(itb+1)}: one 32-byte gather per lerp
  const double4 *TBL_pcc, *TBL_psc, *TBL_pss;
  int pq_same;   // the three tables hold identical numbers (every element has Rc == Rs): one record serves all terms of a pair
};

// One linked-cell grid (replaces header/llist/nacell, reference src/main.F90:277-318) as a counting sort.
struct DevGrid {
  int nc[3], L, dim[3];
  int ncell;            // dim[0]*dim[1]*dim[2], z fastest
  double cs[3];         // normalised cell size
  int *cell_of;         // [NB] linear cell id of each atom (-1: atype==0, skipped like the reference)
  int *start;           // [ncell+1] exclusive prefix of the per-cell counts
  int *fill;            // [ncell] scratch
  int *slot_of;         // [NB] inverse of `order`: slot of each atom in the cell-sorted sequence
  int *order;           // [NB] atom indices sorted by cell; inside a cell DESCENDING index (= the reference's
                        //      head-insertion order, so rows come out in the reference's own order)
  double4 *sorted;      // [NB] {x,y,z, bits(index | type<<32)} in `order` order
};

struct Ctx {
  rxg_config cfg;
  rxg_box box;
  bool have_ff = false, have_box = false;
  int dev = 0;
  cudaStream_t st = nullptr;
  std::string err;
  long long launches = 0;
  // ---- per-atom arrays, capacity NB ------------------------------------------------------------------
  int NB = 0, MAXN = 0;
  int natoms = 0;
  int cp[7] = {0, 0, 0, 0, 0, 0, 0};   // copyptr(0:6), reference src/module.F90:234
  double *pos = nullptr, *v = nullptr, *f = nullptr;            // [3*NB], x|y|z planes
  double *fsl = nullptr;    // [3*NB] by SLOT: partner forces of the non-bonded and H-bond kernels, folded into f by k_fsl_to_f
  double *atype = nullptr, *q = nullptr, *qsfp = nullptr, *qsfv = nullptr;
  double2 *qst = nullptr;   // {qs, qt}      (reference qs(:), qt(:))
  double4 *hsq = nullptr;   // {hs, ht, q, -} gather pack of get_hsh; .z mirrors q for residents+ghosts
  double2 *gst = nullptr;   // {gs, gt}
  double2 *hst = nullptr;   // {hs, ht} of the single-pass CG (residents + ghosts)
  double2 *tst = nullptr;   // {eta*hs + H.hs, same for t} per resident row
  double2 *ust = nullptr;   // resident-weighted H.h sums (Est bookkeeping)
  double2 *wst = nullptr;   // resident-weighted H.qs, H.qt
  int *itype = nullptr, *gid = nullptr, *frcindx = nullptr;
  // ---- PQEq (isPQEq): shell displacements and the per-call products of qeq_initialize (reference src/pqeq.F90) ----------
  double *spos = nullptr;   // [3*NB] spos(NBUFFER,3), planes
  double4 *sps = nullptr;   // [NB] by SLOT: {sx, sy, sz, Zpqeq(type)} of residents and ghosts
  double4 *prow = nullptr;  // [NB] by SLOT (resident rows): {fpqeq, sum_j H_ij Z_j, column sum of the shell-core coupling over resident rows, Z_i}
  double *pcs = nullptr;    // [NB] by SLOT: column sums of the shell-core coupling taken by atomics (ghost columns)
  double *qsl = nullptr;    // [NB] by SLOT: q (final charges, for the shell relaxation and ENbond_PQEq)
  long long pqeq_skips = 0;
  // ---- COPYATOMS bookkeeping -----------------------------------------------------------------------------
  int *gsrc = nullptr;      // [NB] vprocs 1 1 1 only: the resident each ghost is a periodic image of (halo_self)
  bool halo_self = false;
  int *sel = nullptr;       // concatenated selection lists of the six stages of the last MODE_COPY
  int sel_cap = 0, selptr[7] = {0, 0, 0, 0, 0, 0, 0};
  int ns[7] = {0, 0, 0, 0, 0, 0, 0}, nr[7] = {0, 0, 0, 0, 0, 0, 0};   // atoms sent / received per stage (1..6)
  double *sbuf[2] = {nullptr, nullptr}, *rbuf[2] = {nullptr, nullptr};
  size_t xbuf_cap = 0;
  long long moved = 0, nccl_msgs = 0;
  std::function<int()> lazy_upload;   // rxg_move: per-atom state beyond atype/pos is uploaded only once an atom actually migrates
  ncclComm_t comm = nullptr;
  // ---- peer-memory halo refresh (multi-rank, one node): every rank owns a window that its neighbours write into directly
  // over NVLink (cudaIpc), see halo_refresh_peer.  Window = 256 B of flags + 6 stages x 2 parities x pw_cap doubles.
  double *pw = nullptr;
  size_t pw_cap = 0;                 // doubles per buffer
  std::vector<double *> peer;        // [nranks] window of each neighbour rank (my own for myself), nullptr otherwise
  bool peer_ok = false;
  int pseq = 0;                      // refresh sequence number, identical on every rank
  int arseq = 0;                     // all-reduce sequence number
  bool peer_all = false;             // every rank's window is open here: CG scalars are reduced through the windows
  int *d_pushcnt = nullptr;          // [2] block-completion counters of the push kernels
  // list build without a count pass: last step's row counts by global atom id (direct-mapped table), see k_row_caps
  int2 *cnt_tab = nullptr;
  unsigned cnt_mask = 0;
  bool caps_on = true, caps_valid = false, list_capped = false;
  bool caps_cooldown = true, caps_slack_env = false;
  int caps_slack = 8, caps_skip = 0, caps_fails = 0;   // caps_skip: list builds that keep the count pass after an overflow
  long long caps_overflows = 0;
  double *q_save = nullptr;   // [NB] charges at QEq entry, restored if a capped list overflows and the call starts over
  bool hess_fuse = false;   // RXG_HESS_FUSE=1: experiment, the hessian lerp inside the list's fill pass instead of k_hessian
  bool eval_occ = true;     // RXG_EVAL_OCC=0 switches back to the angle / torsion / H-bond evaluators compiled without a register cap
  // interior / boundary split of the SpMV (multi-rank): the ghost refresh runs on st2 beside the interior rows
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // FORCE's charge-independent part beside the QEq CG of the same step (force_device stages, launch_bonded_side):
  // 0 nothing done, 1 bonded cells + bonded list built, 2 bonded terms launched on st2 (joined at the end of the CG)
  cudaEvent_t ev_bfork = nullptr, ev_bjoin = nullptr;
  int bonded_stage = 0;
  bool bonded_env = false;  // RXG_BONDED_OVERLAP=1 (experiment): FORCE's charge-independent part beside the CG; default: FORCE after QEq, in the reference's order
  bool overlap = false, overlap_env = true;
  int *grp_cls = nullptr, *grp_off = nullptr, *grp_int = nullptr, *grp_bnd = nullptr;   // [NB/2+2] each
  int ngrp = 0, ngrp_int = 0, grp_rows = 0, stencil_reach = 0;
  int natoms_prev_move = -1;   // residents after the last rxg_move / QEq (RXG_HINT_CHARGES_STAY is honoured only if the count still matches)
  int hint = 0;             // rxg_hint: the host's promises about the arrays of the next entry-point call
  bool pos_deferred = false;   // a hinted call skipped the copy-back of pos; the next un-hinted copy-back delivers it
  bool fuse = true, fuse_api = false, lists_shared = false;   // md_run: QEq builds halo + 10 A list once for QEq and FORCE of the same step
  int qeq_mode = 0;         // 0 single-pass CG (default), 1 two-pass (literal kernels), strict => literal serial order
  double *tmp = nullptr;    // [12*NB] scratch for MOVE compaction
  double2 *xs = nullptr;    // [NB] CG gather vector in CELL-SORTED order (slot space): {hs,ht} (or {qs,qt} at start)
  double4 *pqa = nullptr;   // [NB] {x,y,z,q} by atom (bonded kernels)
  double4 *pqs = nullptr;   // [NB] {x,y,z,q} by slot (non-bonded gathers)
  int4 *tgs = nullptr;      // [NB] {itype, gid, atom index, -} by slot
  int2 *gts = nullptr;      // [NB] {gid, itype} by slot (k_enbond_half)
  bool enbond_queue = true; // RXG_ENBOND_QUEUE=0: k_enbond<true> (no survivor compaction)
  // ---- cells -------------------------------------------------------------------------------------------
  DevGrid gb, gnb;
  int *d_runs = nullptr;    // stencil runs {dx,dy,dzlo,dzhi}
  int nruns = 0;
  // ---- lists -------------------------------------------------------------------------------------------
  int *nbrcnt = nullptr, *nbrpad = nullptr;       // [NB] counts, [NB*MAXN] padded scratch rows written by k_nbrlist
  int *bptr = nullptr;                            // [NB+1] first bond slot of each atom (exclusive scan of nbrcnt)
  int *nbrlist = nullptr, *nbrindx = nullptr;     // [bond_cap] compact: neighbour index / slot of the reverse bond
  int *bown = nullptr;                            // [bond_cap] atom that owns each slot
  long long bond_cap = 0, nbonds = 0;
  long long *rowoff = nullptr;   // [NB+1] 10 A list: row offsets by CELL-ORDER slot (rows lie in HBM in cell order)
  long long *rowbeg = nullptr, *rowend = nullptr;   // [NB] the same rows addressed by atom index
  int *rowcnt = nullptr;
  int *col = nullptr;            // [nnz_cap]
  double *val = nullptr;         // [nnz_cap] hessian (QEq list only)
  // union stream of the cell-blocked CG SpMV (k_spmv_items): per block of <= 8 consecutive rows of a cell, the columns taken by
  // at least one of the rows (ucol) with the set of rows that take each (umask); blocks start at uoff[first slot of the block]
  int *ucnt = nullptr;               // [NB+2] padded entry count of the block that starts at a slot (0 elsewhere)
  long long *uoff = nullptr;         // [NB+2] exclusive scan of ucnt
  int *ucol = nullptr;               // [un_cap]
  unsigned char *umask = nullptr;    // [un_cap]
  long long un_cap = 0, nunion = 0;
  struct SpItem *items = nullptr;    // [NB] work items of k_spmv_items, written by the list's fill pass
  int nitems = 0, spmv_rg = 4;       // rows per item (4, or 2 for lists with rows longer than 480 entries)
  int spmv_kind = 0, spmv_shape = 0, spmv_stage = 1, spmv_ring = 0, spmv_ring_env = 0, spmv_grid[2] = {0, 0};   // RXG_SPMV / _SHAPE / _STAGE / _RING (rxg_api.cu)
  // window SpMV (k_spmv_win, spmv_kind 2): 16-bit window-relative columns, exact row lengths by slot, group size in cells
  unsigned short *col16 = nullptr;   // [nnz_cap]
  int *rowlen = nullptr;             // [NB+2]
  int2 *win_desc = nullptr;          // [groups][nruns + 1] window descriptors (k_win_desc)
  size_t win_desc_cap = 0;
  int win_ralign = 4, win_lpr = 32, win_g = 0, win_g_env = 0, win_wcap_env = 0, win_nw = 8, win_u = 0, win_max = 0, win_smem_target = 0, win_smem_set[6] = {0, 0, 0, 0, 0, 0};
  bool win_built = false;            // the current list carries col16 / rowlen for group size win_g_built
  int win_g_built = 0;
  std::vector<int> h_runs;           // host copy of the stencil runs (group-size estimate)
  long long win_launches = 0, rows_launches = 0;   // sparse products taken by k_spmv_win / by k_spmv_rows
  long long nnz_cap = 0, nnz = 0, nnz_real = 0;   // nnz counts the row padding, nnz_real does not
  bool list_is_qeq = false;
  int maxrow = 0;                // longest row of the current list (entries)
  // ---- bond-order products, [bond_cap] (compact bond slots) unless noted -----------------------------------------------------------
  double *BO[4] = {nullptr, nullptr, nullptr, nullptr}, *dln[3] = {nullptr, nullptr, nullptr}, *dBOp = nullptr;
  double *A0 = nullptr, *A1 = nullptr, *A2 = nullptr, *A3 = nullptr;
  double *cB[3] = {nullptr, nullptr, nullptr};   // accumulated bond coefficients (cf1,cf2,cf3 of ForceBbo) per directed slot
  double *cdslot = nullptr;                        // cdbnd contributions addressed to the partner of a slot
  double *delta = nullptr, *deltap1 = nullptr, *deltap2 = nullptr, *nlp = nullptr, *dDlp = nullptr, *deltalp = nullptr;
  double *ccbnd = nullptr, *cdbnd = nullptr;
  double *s3 = nullptr;      // [3*NB] per-centre sums of E3b (CE3body_d(1), CEval(6), CEval(5))
  double2 *sbo = nullptr;    // [NB] {prod_SBO, sum_SBO1} per centre
  int2 *wl = nullptr;        // angle / torsion work lists
  long long wl_cap3 = 0, wl_cap4 = 0, wl_caph = 0, n_angles = 0, n_torsions = 0, n_hbonds = 0;
  // ---- scalars ---------------------------------------------------------------------------------------------
  double *d_acc = nullptr;   // [128] reduction targets: 0-39 QEq / PQEq CG and list statistics, 40-55 integrator and observables, 80-103 FORCE (PE, stress)
  int *d_flag = nullptr;     // [8]  error / overflow flags
  int *d_blk = nullptr;      // scan scratch
  long long *d_blk64 = nullptr;
  int nblk_cap = 0;
  int *d_cnt = nullptr, *h_cnt = nullptr;   // [4 + 4*64] count / capacity / error table of exchange_counts (device, pinned host)
  double *h_acc = nullptr;   // pinned mirror of d_acc
  int *h_int = nullptr;      // pinned
  double PE[14] = {0};
  double astr[6] = {0};
  int nstep_qeq = 0;
  DevFF ff;                 // host copy of the device pointer table
  DevFF *d_ff = nullptr;
  std::vector<void *> ff_allocs, allocs;
  bool strict = false;      // RXG_STRICT_ORDER=1: serial-order, FMA-free CG for bit-level validation (small systems)
  double timers_ms[30] = {0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm0 = nullptr, evm1 = nullptr, evk[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t evs[32] = {};   // SpMV timing, one pair per iteration of a CG batch
  int cg_batch = 4;           // CG iterations enqueued between two looks at the stop flag (RXG_CG_BATCH, 1..16)
  bool grad_pending = false;
  // phase clock (rxg_it_timer)
  bool ph_on = true;
  std::vector<cudaEvent_t> ph_ev;
  std::vector<int> ph_slot;
  size_t ph_used = 0;
  double ph_sec[31] = {0};
  // ---- staging (pinned) ------------------------------------------------------------------------------------
};

// Phase clock behind rxg_it_timer: one CUDA event per phase boundary on the library's stream, resolved lazily (no
// synchronisation on the hot path).  Slots are the reference's it_timer indices (src/module.F90:215-217, table printed at
// src/main.F90:148-180): 1 QEq, 3 LINKEDLIST, 4 COPYATOMS, 5 NEIGHBORLIST, 6 BOCALC, 7 ENbond, 8 Ebond, 9 Elnpr, 10 Ehb,
// 11 E3b, 12 E4b, 13 ForceBondedTerms, 15 GetNonbondingPairList, 16 qeq_initialize, 18 get_hsh, 19 get_gradient.
constexpr int PH_QEQ = 64;
inline void phase_harvest(Ctx *c) {
  if (c->ph_used < 2) { c->ph_used = 0; return; }
  cudaEventSynchronize(c->ph_ev[c->ph_used - 1]);
  for (size_t i = 0; i + 1 < c->ph_used; i++) {
    if (c->ph_slot[i] <= 0) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ph_ev[i], c->ph_ev[i + 1]) != cudaSuccess) continue;
    c->ph_sec[c->ph_slot[i] & 63] += 1e-3 * ms;
    if (c->ph_slot[i] & PH_QEQ) c->ph_sec[1] += 1e-3 * ms;   // phases inside QEq also count towards it_timer(1)
  }
  c->ph_used = 0;
}
// start of phase `slot` (0 = end of a timed region); the previous phase ends here
inline void phase_mark(Ctx *c, int slot) {
  if (!c->ph_on) return;
  if (c->ph_used >= 4096) phase_harvest(c);
  if (c->ph_used >= c->ph_ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { c->ph_on = false; return; }
    c->ph_ev.push_back(e);
    c->ph_slot.push_back(0);
  }
  cudaEventRecord(c->ph_ev[c->ph_used], c->st);
  c->ph_slot[c->ph_used] = slot;
  c->ph_used++;
}

#define RXG_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      c->err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__); \
      return RXG_ERR_CUDA;                                                                               \
    }                                                                                                    \
  } while (0)

#define RXG_TRY(call)            \
  do {                           \
    int rc_ = (call);            \
    if (rc_ != RXG_OK) return rc_; \
  } while (0)

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

#define LAUNCH(c, kern, grid, block, smem, ...)                 \
  do {                                                          \
    kern<<<(grid), (block), (smem), (c)->st>>>(__VA_ARGS__);    \
    (c)->launches++;                                            \
  } while (0)

}   // namespace rxg
