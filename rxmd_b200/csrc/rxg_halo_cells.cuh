// rxg_halo_cells.cuh -- COPYATOMS as kernels (kernel group E) and linked-cell binning (A1).
// Reference: src/comm.F90 (COPYATOMS), src/main.F90:277-318 (LINKEDLIST), :596-681 (coordinate transforms).
#pragma once
#include "rxg_common.cuh"

namespace rxg {

constexpr int SCAN_BLK = 1024;

// ---------------------------------------------------------------------------------------------------
// block-level exclusive scan helper (blockDim.x == SCAN_BLK)
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T *total) {
  __shared__ T warp_sums[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  T x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    T s = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    warp_sums[lane] = s;
  }
  __syncthreads();
  T base = wid ? warp_sums[wid - 1] : (T)0;
  if (total) *total = warp_sums[31];
  T r = base + x - v;
  __syncthreads();
  return r;
}

// generic 3-phase exclusive scan over int counts -> T offsets (T = int or long long)
template <typename T>
__global__ void k_scan_phase1(const int *__restrict__ in, long long n, T *__restrict__ blk) {
  long long i = (long long)blockIdx.x * SCAN_BLK + threadIdx.x;
  T v = (i < n) ? (T)in[i] : (T)0;
  T tot;
  block_excl_scan<T>(v, &tot);
  if (threadIdx.x == 0) blk[blockIdx.x] = tot;
}
template <typename T>
__global__ void k_scan_phase2(T *__restrict__ blk, int nblk, T *__restrict__ total_out) {
  // single block; serial over chunks of SCAN_BLK
  __shared__ T carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += SCAN_BLK) {
    int i = base + threadIdx.x;
    T v = (i < nblk) ? blk[i] : (T)0;
    T tot;
    T ex = block_excl_scan<T>(v, &tot);
    if (i < nblk) blk[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}
template <typename T>
__global__ void k_scan_phase3(const int *__restrict__ in, long long n, const T *__restrict__ blk, T *__restrict__ out) {
  long long i = (long long)blockIdx.x * SCAN_BLK + threadIdx.x;
  T v = (i < n) ? (T)in[i] : (T)0;
  T ex = block_excl_scan<T>(v, nullptr);
  if (i < n) out[i] = ex + blk[blockIdx.x];
  if (i == n - 1) out[n] = ex + blk[blockIdx.x] + v;   // closing entry
}

// out[0..n] = exclusive prefix of in[0..n), out[n] = total.  total also left in *d_total (device) if given.
template <typename T>
int device_scan(Ctx *c, const int *in, long long n, T *out, T *blk_scratch, T *d_total) {
  if (n <= 0) {
    RXG_CUDA(cudaMemsetAsync(out, 0, sizeof(T), c->st));
    if (d_total) RXG_CUDA(cudaMemsetAsync(d_total, 0, sizeof(T), c->st));
    return RXG_OK;
  }
  int nblk = cdiv(n, SCAN_BLK);
  LAUNCH(c, k_scan_phase1<T>, nblk, SCAN_BLK, 0, in, n, blk_scratch);
  LAUNCH(c, k_scan_phase2<T>, 1, SCAN_BLK, 0, blk_scratch, nblk, d_total);
  LAUNCH(c, k_scan_phase3<T>, nblk, SCAN_BLK, 0, in, n, blk_scratch, out);
  return RXG_OK;
}

// ---------------------------------------------------------------------------------------------------
// coordinate transforms on resident+ghost positions (in place), src/main.F90:613-681
struct BoxDev {
  double H[9], Hi[9], OBOX[3], LBOX[3];
};

__global__ void k_to_norm(double *__restrict__ pos, int NB, int n, BoxDev b) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = pos[i], y = pos[NB + i], z = pos[2 * NB + i];
  pos[i] = sub_rn(dot3_rn(b.Hi[0], b.Hi[3], b.Hi[6], x, y, z), b.OBOX[0]);
  pos[NB + i] = sub_rn(dot3_rn(b.Hi[1], b.Hi[4], b.Hi[7], x, y, z), b.OBOX[1]);
  pos[2 * NB + i] = sub_rn(dot3_rn(b.Hi[2], b.Hi[5], b.Hi[8], x, y, z), b.OBOX[2]);
}
__global__ void k_to_real(double *__restrict__ pos, int NB, int n, BoxDev b) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = add_rn(pos[i], b.OBOX[0]), y = add_rn(pos[NB + i], b.OBOX[1]), z = add_rn(pos[2 * NB + i], b.OBOX[2]);
  pos[i] = dot3_rn(b.H[0], b.H[3], b.H[6], x, y, z);
  pos[NB + i] = dot3_rn(b.H[1], b.H[4], b.H[7], x, y, z);
  pos[2 * NB + i] = dot3_rn(b.H[2], b.H[5], b.H[8], x, y, z);
}
// the reference's QCOPY1/QCOPY2 calls convert every position to normalised coordinates and back without moving
// anything else (src/comm.F90:222-227,260-264); keeping the round trip keeps positions bit-identical
__global__ void k_roundtrip(double *__restrict__ pos, int NB, int n, BoxDev b, int times) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = pos[i], y = pos[NB + i], z = pos[2 * NB + i];
  for (int t = 0; t < times; t++) {
    double sx = sub_rn(dot3_rn(b.Hi[0], b.Hi[3], b.Hi[6], x, y, z), b.OBOX[0]);
    double sy = sub_rn(dot3_rn(b.Hi[1], b.Hi[4], b.Hi[7], x, y, z), b.OBOX[1]);
    double sz = sub_rn(dot3_rn(b.Hi[2], b.Hi[5], b.Hi[8], x, y, z), b.OBOX[2]);
    sx = add_rn(sx, b.OBOX[0]); sy = add_rn(sy, b.OBOX[1]); sz = add_rn(sz, b.OBOX[2]);
    x = dot3_rn(b.H[0], b.H[3], b.H[6], sx, sy, sz);
    y = dot3_rn(b.H[1], b.H[4], b.H[7], sx, sy, sz);
    z = dot3_rn(b.H[2], b.H[5], b.H[8], sx, sy, sz);
  }
  pos[i] = x; pos[NB + i] = y; pos[2 * NB + i] = z;
}

inline BoxDev make_boxdev(const rxg_box &b) {
  BoxDev d;
  for (int i = 0; i < 9; i++) { d.H[i] = b.HH[i]; d.Hi[i] = b.HHi[i]; }
  for (int i = 0; i < 3; i++) { d.OBOX[i] = b.OBOX[i]; d.LBOX[i] = b.LBOX[i]; }
  return d;
}

// ---------------------------------------------------------------------------------------------------
// halo selection: inBuffer(), src/comm.F90:551-576.  upper: lbox-dr < rr ; lower: rr <= dr
__device__ __forceinline__ bool in_buffer(bool upper, double lbox, double dr, double rr) {
  return upper ? (sub_rn(lbox, dr) < rr) : (rr <= dr);
}

__global__ void k_sel_count(const double *__restrict__ coord, const double *__restrict__ atype, int n, bool upper,
                            double lbox, double dr, int skip_dead, int *__restrict__ blk) {
  int i = blockIdx.x * SCAN_BLK + threadIdx.x;
  int fl = 0;
  if (i < n) fl = in_buffer(upper, lbox, dr, coord[i]) && !(skip_dead && atype[i] < 0.0);
  int tot;
  block_excl_scan<int>(fl, &tot);
  if (threadIdx.x == 0) blk[blockIdx.x] = tot;
}

// ---------------------------------------------------------------------------------------------------
// COPYATOMS as pack -> exchange -> unpack.  Messages are SoA inside the buffer: field f of atom k at [f*cnt + k].
// The sender keeps the list of local indices it selected in each of the six stages (`sel`); every later refresh
// (MODE_QCOPY*) and the force copy-back (MODE_CPBK) reuse those lists, so only MODE_COPY / MODE_MOVE ever test
// positions (the reference re-tests them on every call, src/comm.F90:273-288, with the same outcome).
constexpr int NE_COPY = 9;    // x y z atype q qs qt hs ht           (reference ne=10 also carries frcindx)
constexpr int NE_MOVE = 12;   // x y z vx vy vz atype q qs qt qsfp qsfv  (+ sx sy sz with PQEq, src/comm.F90:153,165-167)

// store_atoms for MODE_COPY (src/comm.F90:406-447): stable selection, coordinate shift by -/+LBOX (xshift :531)
__global__ void k_pack_copy(const double *__restrict__ pos, int NB, const double *__restrict__ atype,
                            const double *__restrict__ q, const double2 *__restrict__ qst, const double4 *__restrict__ hsq,
                            int n, int axis, bool upper, double lbox, double dr, double sft, const int *__restrict__ blkoff,
                            int cnt, int *__restrict__ sel, double *__restrict__ buf) {
  int i = blockIdx.x * SCAN_BLK + threadIdx.x;
  int fl = 0;
  if (i < n) fl = in_buffer(upper, lbox, dr, pos[(size_t)axis * NB + i]);
  int ex = block_excl_scan<int>(fl, nullptr);
  if (!fl) return;
  int k = blkoff[blockIdx.x] + ex;
  if (k >= cnt) return;
  sel[k] = i;
  double p[3] = {pos[i], pos[NB + i], pos[2 * (size_t)NB + i]};
  p[axis] = add_rn(p[axis], sft);
  double2 s = qst[i];
  double4 h = hsq[i];
  size_t c = cnt;
  buf[k] = p[0]; buf[c + k] = p[1]; buf[2 * c + k] = p[2];
  buf[3 * c + k] = atype[i]; buf[4 * c + k] = q[i];
  buf[5 * c + k] = s.x; buf[6 * c + k] = s.y; buf[7 * c + k] = h.x; buf[8 * c + k] = h.y;
}
// append_atoms for MODE_COPY (src/comm.F90:497-522)
// `sel` (self-exchange only, else null): the sender's selection list IS this rank's, so ghost m is an image of local atom
// sel[k]; gsrc[m] = the RESIDENT it descends from (an image of an image resolves through the earlier stage's entry)
__global__ void k_unpack_copy(double *__restrict__ pos, int NB, double *__restrict__ atype, double *__restrict__ q,
                              double2 *__restrict__ qst, double4 *__restrict__ hsq, int cnt, int dst0,
                              const double *__restrict__ buf, const int *__restrict__ sel, int natoms, int *__restrict__ gsrc) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  size_t c = cnt;
  int m = dst0 + k;
  if (sel) { int i = sel[k]; gsrc[m] = i < natoms ? i : gsrc[i]; }
  pos[m] = buf[k]; pos[NB + m] = buf[c + k]; pos[2 * (size_t)NB + m] = buf[2 * c + k];
  atype[m] = buf[3 * c + k];
  double qq = buf[4 * c + k];
  q[m] = qq;
  qst[m] = make_double2(buf[5 * c + k], buf[6 * c + k]);
  hsq[m] = make_double4(buf[7 * c + k], buf[8 * c + k], qq, 0.0);
}

// store_atoms for MODE_MOVE: atoms that left through the stage's face; the original is marked dead (atype=-1)
__global__ void k_pack_move(const double *__restrict__ pos, const double *__restrict__ v, int NB, double *__restrict__ atype,
                            const double *__restrict__ q, const double2 *__restrict__ qst, const double *__restrict__ qsfp,
                            const double *__restrict__ qsfv, const double *__restrict__ spos, int n, int axis, bool upper,
                            double lbox, double sft, const int *__restrict__ blkoff, int cnt, double *__restrict__ buf) {
  int i = blockIdx.x * SCAN_BLK + threadIdx.x;
  int fl = 0;
  if (i < n) fl = in_buffer(upper, lbox, 0.0, pos[(size_t)axis * NB + i]) && !(atype[i] < 0.0);
  int ex = block_excl_scan<int>(fl, nullptr);
  if (!fl) return;
  int k = blkoff[blockIdx.x] + ex;
  if (k >= cnt) return;
  double p[3] = {pos[i], pos[NB + i], pos[2 * (size_t)NB + i]};
  p[axis] = add_rn(p[axis], sft);
  double2 s = qst[i];
  size_t c = cnt;
  buf[k] = p[0]; buf[c + k] = p[1]; buf[2 * c + k] = p[2];
  buf[3 * c + k] = v[i]; buf[4 * c + k] = v[NB + i]; buf[5 * c + k] = v[2 * (size_t)NB + i];
  buf[6 * c + k] = atype[i]; buf[7 * c + k] = q[i]; buf[8 * c + k] = s.x; buf[9 * c + k] = s.y;
  buf[10 * c + k] = qsfp[i]; buf[11 * c + k] = qsfv[i];
  if (spos) { buf[12 * c + k] = spos[i]; buf[13 * c + k] = spos[(size_t)NB + i]; buf[14 * c + k] = spos[2 * (size_t)NB + i]; }
  atype[i] = -1.0;
}
__global__ void k_unpack_move(double *__restrict__ pos, double *__restrict__ v, int NB, double *__restrict__ atype,
                              double *__restrict__ q, double2 *__restrict__ qst, double *__restrict__ qsfp,
                              double *__restrict__ qsfv, double *__restrict__ spos, int cnt, int dst0,
                              const double *__restrict__ buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  size_t c = cnt;
  int m = dst0 + k;
  if (spos) { spos[m] = buf[12 * c + k]; spos[(size_t)NB + m] = buf[13 * c + k]; spos[2 * (size_t)NB + m] = buf[14 * c + k]; }
  pos[m] = buf[k]; pos[NB + m] = buf[c + k]; pos[2 * (size_t)NB + m] = buf[2 * c + k];
  v[m] = buf[3 * c + k]; v[NB + m] = buf[4 * c + k]; v[2 * (size_t)NB + m] = buf[5 * c + k];
  atype[m] = buf[6 * c + k]; q[m] = buf[7 * c + k];
  qst[m] = make_double2(buf[8 * c + k], buf[9 * c + k]);
  qsfp[m] = buf[10 * c + k]; qsfv[m] = buf[11 * c + k];
}

// finalize(MODE_MOVE): stable removal of dead atoms, src/comm.F90:238-256
__global__ void k_alive_flag(const double *__restrict__ atype, int n, int *__restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = nint_d(atype[i]) > 0;
}
__global__ void k_move_compact(const int *__restrict__ flag, const int *__restrict__ dst, int n, int NB,
                               const double *__restrict__ pos, const double *__restrict__ v,
                               const double *__restrict__ atype, const double *__restrict__ q,
                               const double2 *__restrict__ qst, const double *__restrict__ qsfp,
                               const double *__restrict__ qsfv, const double *__restrict__ spos, double *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  int m = dst[i];
  for (int a = 0; a < 3; a++) { out[(size_t)a * NB + m] = pos[(size_t)a * NB + i]; out[(size_t)(3 + a) * NB + m] = v[(size_t)a * NB + i]; }
  out[(size_t)6 * NB + m] = atype[i];
  out[(size_t)7 * NB + m] = q[i];
  out[(size_t)8 * NB + m] = qst[i].x;
  out[(size_t)9 * NB + m] = qst[i].y;
  out[(size_t)10 * NB + m] = qsfp[i];
  out[(size_t)11 * NB + m] = qsfv[i];
  if (spos)
    for (int a = 0; a < 3; a++) out[(size_t)(12 + a) * NB + m] = spos[(size_t)a * NB + i];
}
__global__ void k_move_restore(const double *__restrict__ in, int n, int NB, double *__restrict__ pos, double *__restrict__ v,
                               double *__restrict__ atype, double *__restrict__ q, double2 *__restrict__ qst,
                               double *__restrict__ qsfp, double *__restrict__ qsfv, double *__restrict__ spos) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  if (spos)
    for (int a = 0; a < 3; a++) spos[(size_t)a * NB + m] = in[(size_t)(12 + a) * NB + m];
  for (int a = 0; a < 3; a++) { pos[(size_t)a * NB + m] = in[(size_t)a * NB + m]; v[(size_t)a * NB + m] = in[(size_t)(3 + a) * NB + m]; }
  atype[m] = in[(size_t)6 * NB + m];
  q[m] = in[(size_t)7 * NB + m];
  qst[m] = make_double2(in[(size_t)8 * NB + m], in[(size_t)9 * NB + m]);
  qsfp[m] = in[(size_t)10 * NB + m];
  qsfv[m] = in[(size_t)11 * NB + m];
}

// value refreshes through the stored selection lists.  which: 1 = (qs,qt) [MODE_QCOPY1], 2 = (hs,ht,q) [MODE_QCOPY2],
// 3 = (hs,ht) only (single-pass CG), 4 = q only (FORCE reusing the QEq halo), 5 = spos (PQEq: the reference packs spos
// into MODE_COPY itself, src/comm.F90:122,129-131; here it follows through the same selection lists)
__global__ void k_pack_vals(int which, const int *__restrict__ sel, int cnt, const double2 *__restrict__ qst,
                            const double4 *__restrict__ hsq, const double2 *__restrict__ hst, const double *__restrict__ q,
                            const double *__restrict__ spos, int NB, double *__restrict__ buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  int i = sel[k];
  size_t c = cnt;
  if (which == 4) { buf[k] = q[i]; return; }
  if (which == 5) { buf[k] = spos[i]; buf[c + k] = spos[(size_t)NB + i]; buf[2 * c + k] = spos[2 * (size_t)NB + i]; return; }
  if (which == 1) { double2 s = qst[i]; buf[k] = s.x; buf[c + k] = s.y; }
  else if (which == 2) { double4 h = hsq[i]; buf[k] = h.x; buf[c + k] = h.y; buf[2 * c + k] = h.z; }
  else { double2 h = hst[i]; buf[k] = h.x; buf[c + k] = h.y; }
}
__global__ void k_unpack_vals(int which, int cnt, int dst0, const double *__restrict__ buf, double2 *__restrict__ qst,
                              double4 *__restrict__ hsq, double2 *__restrict__ hst, double2 *__restrict__ xs,
                              const int *__restrict__ slot_of, double *__restrict__ q, double *__restrict__ spos, int NB) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  size_t c = cnt;
  int m = dst0 + k;
  if (which == 4) { q[m] = buf[k]; return; }
  if (which == 5) { spos[m] = buf[k]; spos[(size_t)NB + m] = buf[c + k]; spos[2 * (size_t)NB + m] = buf[2 * c + k]; return; }
  if (which == 1) qst[m] = make_double2(buf[k], buf[c + k]);
  else if (which == 2) { double qq = buf[2 * c + k]; hsq[m] = make_double4(buf[k], buf[c + k], qq, 0.0); q[m] = qq; }
  else { double2 v = make_double2(buf[k], buf[c + k]); hst[m] = v; xs[slot_of[m]] = v; }
}
// the same refreshes when every ghost is a periodic image of a resident of this rank: one gather through gsrc
__global__ void k_refresh_self(int which, int natoms, int ntot, const int *__restrict__ gsrc, double2 *__restrict__ qst,
                               double4 *__restrict__ hsq, double2 *__restrict__ hst, double2 *__restrict__ xs,
                               const int *__restrict__ slot_of, double *__restrict__ q, double *__restrict__ spos, int NB) {
  int m = natoms + blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= ntot) return;
  const int i = gsrc[m];
  if (which == 4) { q[m] = q[i]; return; }
  if (which == 5) { spos[m] = spos[i]; spos[(size_t)NB + m] = spos[(size_t)NB + i]; spos[2 * (size_t)NB + m] = spos[2 * (size_t)NB + i]; return; }
  if (which == 1) qst[m] = qst[i];
  else if (which == 2) { double4 h = hsq[i]; hsq[m] = make_double4(h.x, h.y, h.z, 0.0); q[m] = h.z; }
  else { double2 v = hst[i]; hst[m] = v; xs[slot_of[m]] = v; }
}
// ---------------------------------------------------------------------------------------------------
// Peer-memory variant of the value refreshes (MODE_QCOPY1/2 and friends) for ranks that share a node: instead of
// pack -> ncclSend/ncclRecv -> unpack, the sender's kernel gathers the selected values and STORES them straight into the
// receiver's window over NVLink, then publishes a sequence number; the receiver's kernel waits for that number and scatters
// the values to its ghosts.  One push + one pull kernel per direction, no library call on the critical path of a CG
// iteration.  Stage order (x, then y, then z, so that edge and corner images are forwarded) is kept by stream order: a
// rank's y-push runs after its x-pull.
// Buffer reuse: message number s of a stage goes to parity s&1.  A rank can only send number s+2 after it pulled number
// s+1 from the same neighbour, which that neighbour pushed after its own pulls of number s (same stream) -- so the buffer
// being overwritten has been consumed.
constexpr int PW_HDR = 32;   // doubles (256 B) reserved for the 12 flags
// all-reduce area that follows the 12 halo buffers: [2 parities][PW_MAXR ranks][PW_ARW doubles] then [2][PW_MAXR] int flags
constexpr int PW_MAXR = 64, PW_ARW = 8;
constexpr size_t PW_AR_DOUBLES = 2 * PW_MAXR * PW_ARW + PW_MAXR;   // values + flags (2*PW_MAXR ints)
constexpr long long PEER_SPIN_LIMIT = 4000000000LL;   // ~2 s at 1.9 GHz, then the pull gives up and raises an error flag

struct PeerDir {   // the two directions of one axis (blockIdx.y)
  const int *sel[2];
  int cnt[2], dst0[2], nblk[2];
  double *data[2];   // push: the target's buffer; pull: my buffer
  int *flag[2];
};
__global__ void k_peer_push(int which, PeerDir d, const double2 *__restrict__ qst,
                            const double4 *__restrict__ hsq, const double2 *__restrict__ hst, const double *__restrict__ q,
                            const double *__restrict__ spos, int NB, int seq, int *__restrict__ counters) {
  const int dir = blockIdx.y;
  if ((int)blockIdx.x >= d.nblk[dir]) return;
  const int *__restrict__ sel = d.sel[dir];
  const int cnt = d.cnt[dir];
  double *__restrict__ rdata = d.data[dir];
  int *rflag = d.flag[dir];
  int *counter = counters + dir;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < cnt) {
    const int i = sel[k];
    const size_t c = cnt;
    if (which == 4) rdata[k] = q[i];
    else if (which == 5) { rdata[k] = spos[i]; rdata[c + k] = spos[(size_t)NB + i]; rdata[2 * c + k] = spos[2 * (size_t)NB + i]; }
    else if (which == 1) { double2 s = qst[i]; rdata[k] = s.x; rdata[c + k] = s.y; }
    else if (which == 2) { double4 h = hsq[i]; rdata[k] = h.x; rdata[c + k] = h.y; rdata[2 * c + k] = h.z; }
    else { double2 h = hst[i]; rdata[k] = h.x; rdata[c + k] = h.y; }
  }
  __threadfence_system();   // my stores are visible system-wide before anything I (or my block) signal afterwards
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(counter, 1);
    if (done == d.nblk[dir] - 1) {   // last block of this direction: every block's data is out
      *counter = 0;
      __threadfence_system();
      asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(rflag), "r"(seq) : "memory");
    }
  }
}
__global__ void k_peer_pull(int which, PeerDir d, int seq,
                            double2 *__restrict__ qst, double4 *__restrict__ hsq, double2 *__restrict__ hst,
                            double2 *__restrict__ xs, const int *__restrict__ slot_of, double *__restrict__ q,
                            double *__restrict__ spos, int NB, int *__restrict__ err) {
  const int dir = blockIdx.y;
  if ((int)blockIdx.x >= d.nblk[dir]) return;
  const int cnt = d.cnt[dir], dst0 = d.dst0[dir];
  const double *__restrict__ ldata = d.data[dir];
  const int *lflag = d.flag[dir];
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    int v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(lflag) : "memory");
      if (v != seq && clock64() - t0 > PEER_SPIN_LIMIT) { atomicExch(err, 1); break; }
    } while (v != seq);
  }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const size_t c = cnt;
  const int m = dst0 + k;
  // the window was written by another GPU: read it past L1 (which is not coherent with remote stores)
  if (which == 4) { q[m] = __ldcg(ldata + k); return; }
  if (which == 5) { spos[m] = __ldcg(ldata + k); spos[(size_t)NB + m] = __ldcg(ldata + c + k); spos[2 * (size_t)NB + m] = __ldcg(ldata + 2 * c + k); return; }
  if (which == 1) qst[m] = make_double2(__ldcg(ldata + k), __ldcg(ldata + c + k));
  else if (which == 2) { double qq = __ldcg(ldata + 2 * c + k); hsq[m] = make_double4(__ldcg(ldata + k), __ldcg(ldata + c + k), qq, 0.0); q[m] = qq; }
  else { double2 v = make_double2(__ldcg(ldata + k), __ldcg(ldata + c + k)); hst[m] = v; xs[slot_of[m]] = v; }
}

// an axis whose neighbour is this rank itself (vprocs = 1 along it): ghost cp[stage]+j is the image of local atom sel[j]
__global__ void k_refresh_axis_local(int which, PeerDir d, double2 *__restrict__ qst, double4 *__restrict__ hsq,
                                     double2 *__restrict__ hst, double2 *__restrict__ xs, const int *__restrict__ slot_of,
                                     double *__restrict__ q, double *__restrict__ spos, int NB) {
  const int dir = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d.cnt[dir]) return;
  const int i = d.sel[dir][k], m = d.dst0[dir] + k;
  if (which == 4) { q[m] = q[i]; return; }
  if (which == 5) { spos[m] = spos[i]; spos[(size_t)NB + m] = spos[(size_t)NB + i]; spos[2 * (size_t)NB + m] = spos[2 * (size_t)NB + i]; return; }
  if (which == 1) qst[m] = qst[i];
  else if (which == 2) { double4 h = hsq[i]; hsq[m] = make_double4(h.x, h.y, h.z, 0.0); q[m] = h.z; }
  else { double2 v = hst[i]; hst[m] = v; xs[slot_of[m]] = v; }
}

inline int halo_refresh_peer_axis(Ctx *c, int which, int axis, int seq) {
  const int nf = (which == 2 || which == 5) ? 3 : (which == 4 ? 1 : 2);
  const int par = seq & 1;
  int *err = c->d_flag + 3;   // read back with the CG scalars of the iteration (qeq_cg_single)
  const int d0 = 2 * axis + 1;
  const int ns[2] = {c->ns[d0], c->ns[d0 + 1]}, nr[2] = {c->nr[d0], c->nr[d0 + 1]};
  const int tgt[2] = {c->box.target_node[2 * axis], c->box.target_node[2 * axis + 1]};
  if ((size_t)nf * (size_t)std::max(std::max(ns[0], ns[1]), std::max(nr[0], nr[1])) > c->pw_cap) {
    c->err = "peer halo window too small for this halo (raise nbuffer)";
    return RXG_ERR_NBUFFER;
  }
  PeerDir ps, pl;
  for (int k = 0; k < 2; k++) {   // my selection of stage d0+k lands in the target's buffer of the same stage
    const int slot = (d0 - 1 + k) * 2 + par;
    double *base = c->peer[tgt[k]];
    ps.sel[k] = c->sel + c->selptr[d0 - 1 + k]; ps.cnt[k] = ns[k]; ps.dst0[k] = 0; ps.nblk[k] = cdiv(std::max(ns[k], 1), 256);
    ps.data[k] = base + PW_HDR + (size_t)slot * c->pw_cap; ps.flag[k] = (int *)base + slot;
    pl.sel[k] = nullptr; pl.cnt[k] = nr[k]; pl.dst0[k] = c->cp[d0 - 1 + k]; pl.nblk[k] = cdiv(std::max(nr[k], 1), 256);
    pl.data[k] = c->pw + PW_HDR + (size_t)slot * c->pw_cap; pl.flag[k] = (int *)c->pw + slot;
  }
  LAUNCH(c, k_peer_push, dim3(std::max(ps.nblk[0], ps.nblk[1]), 2), 256, 0, which, ps, c->qst, c->hsq, c->hst, c->q, c->spos, c->NB, seq, c->d_pushcnt);
  LAUNCH(c, k_peer_pull, dim3(std::max(pl.nblk[0], pl.nblk[1]), 2), 256, 0, which, pl, seq, c->qst, c->hsq, c->hst, c->xs, c->gnb.slot_of, c->q,
         c->spos, c->NB, err);
  return RXG_OK;
}

// MPI_ALLREDUCE(SUM) of a few doubles through the windows: every rank stores its `count` values into slot [me] of every
// rank's all-reduce area, publishes the sequence number, waits for the other ranks' numbers in its own area and adds the
// contributions in RANK ORDER -- every rank gets the bit-identical sum (ncclAllReduce does not promise an order).
// One block of PW_MAXR threads; thread r talks to rank r.
struct PeerAll { double *win[PW_MAXR]; };
__global__ void __launch_bounds__(PW_MAXR) k_peer_allreduce(PeerAll pa, int nranks, int me, size_t aroff, int seq, double *__restrict__ acc,
                                                           int count, int *__restrict__ err) {
  const int r = threadIdx.x, par = seq & 1;
  if (r < nranks) {
    double *dst = pa.win[r] + aroff + ((size_t)par * PW_MAXR + me) * PW_ARW;
    for (int k = 0; k < count; k++) dst[k] = acc[k];
    __threadfence_system();
    int *flag = (int *)(pa.win[r] + aroff + 2 * PW_MAXR * PW_ARW) + par * PW_MAXR + me;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
    const int *mine = (const int *)(pa.win[me] + aroff + 2 * PW_MAXR * PW_ARW) + par * PW_MAXR + r;
    const long long t0 = clock64();
    int v;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v != seq && clock64() - t0 > PEER_SPIN_LIMIT) { atomicExch(err, 1); break; }
    } while (v != seq);
    __threadfence();
  }
  __syncthreads();
  if (r < count) {
    const double *src = pa.win[me] + aroff + (size_t)par * PW_MAXR * PW_ARW;
    double sum = 0.0;
    for (int q = 0; q < nranks; q++) sum = add_rn(sum, __ldcg(src + (size_t)q * PW_ARW + r));
    acc[r] = sum;
  }
}

// MODE_CPBK: ghost forces of one stage travel back to the rank that owns the source atoms and are added there
// (src/comm.F90:385-396, 474-482); the owner addresses them through its own selection list
__global__ void k_pack_force(const double *__restrict__ f, int NB, int lo, int cnt, double *__restrict__ buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  size_t c = cnt;
  buf[k] = f[lo + k]; buf[c + k] = f[(size_t)NB + lo + k]; buf[2 * c + k] = f[2 * (size_t)NB + lo + k];
}
__global__ void k_unpack_force(double *__restrict__ f, int NB, const int *__restrict__ sel, int cnt, const double *__restrict__ buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  size_t c = cnt;
  int s = sel[k];
  atomicAdd(&f[s], buf[k]);
  atomicAdd(&f[(size_t)NB + s], buf[c + k]);
  atomicAdd(&f[2 * (size_t)NB + s], buf[2 * c + k]);
}

// qs(:), qt(:) of the host <-> the packed {qs,qt} of the device (rxg_move): two planes of scratch, no host loop
__global__ void k_planes_to_pairs(int n, const double *__restrict__ a, const double *__restrict__ b, double2 *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_double2(a[i], b[i]);
}
__global__ void k_pairs_to_planes(int n, const double2 *__restrict__ in, double *__restrict__ a, double *__restrict__ b) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { double2 v = in[i]; a[i] = v.x; b[i] = v.y; }
}

// are the host's residents bit-identical to the device's?  (flag != 0 if not)
__global__ void k_same_atoms(int n, int NB, const double *__restrict__ stage, const double *__restrict__ pos,
                             const double *__restrict__ atype, int *__restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool same = stage[i] == pos[i] && stage[(size_t)NB + i] == pos[(size_t)NB + i] && stage[2 * (size_t)NB + i] == pos[2 * (size_t)NB + i] &&
              stage[3 * (size_t)NB + i] == atype[i];
  if (!same) atomicExch(flag, 1);
}

// itype = nint(atype), gtype = l2g(atype): src/pot.F90:37-42, src/main.F90:582-593
__global__ void k_types(const double *__restrict__ atype, int n, int *__restrict__ itype, int *__restrict__ gid) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = atype[i];
  int ity = nint_d(a);
  itype[i] = ity;
  gid[i] = nint_d(mul_rn(sub_rn(a, (double)ity), 1e13));
}

// ---------------------------------------------------------------------------------------------------
static const int h_cptridx[7] = {0, 0, 0, 2, 2, 4, 4};

inline int ensure_blk(Ctx *c, long long n) {
  int need = cdiv(n, SCAN_BLK) + 8;
  if (need > c->nblk_cap) {
    if (c->d_blk) cudaFree(c->d_blk);
    if (c->d_blk64) cudaFree(c->d_blk64);
    c->nblk_cap = need * 2;
    RXG_CUDA(cudaMalloc(&c->d_blk, sizeof(int) * 2 * c->nblk_cap));   // two directions of an axis side by side
    RXG_CUDA(cudaMalloc(&c->d_blk64, sizeof(long long) * c->nblk_cap));
  }
  return RXG_OK;
}

inline int ensure_xbuf(Ctx *c, size_t doubles_each) {
  if (doubles_each > c->xbuf_cap) {
    for (int k = 0; k < 2; k++) {
      if (c->sbuf[k]) cudaFree(c->sbuf[k]);
      if (c->rbuf[k]) cudaFree(c->rbuf[k]);
    }
    c->xbuf_cap = doubles_each + doubles_each / 4 + 1024;
    for (int k = 0; k < 2; k++) {
      RXG_CUDA(cudaMalloc(&c->sbuf[k], sizeof(double) * c->xbuf_cap));
      RXG_CUDA(cudaMalloc(&c->rbuf[k], sizeof(double) * c->xbuf_cap));
    }
  }
  return RXG_OK;
}

#define RXG_NCCL(call)                                                                                      \
  do {                                                                                                      \
    ncclResult_t r_ = (call);                                                                               \
    if (r_ != ncclSuccess) {                                                                                \
      c->err = std::string("NCCL error: ") + nccl_api().GetErrorString(r_) + " at " + __FILE__ + ":" + std::to_string(__LINE__); \
      return RXG_ERR_NCCL;                                                                                  \
    }                                                                                                       \
  } while (0)

// send_recv (src/comm.F90:291-364) for the two stages of one axis at once.  Stage `up` (dflag 2a+1) sends to the
// +neighbour and receives from the -neighbour; stage `dn` (dflag 2a+2) the other way round.  cs/cr are counts in
// doubles.  Neighbour == myself is the reference's self-copy branch (:305-315); here it is a pointer alias.
inline int exchange_axis(Ctx *c, int axis, const size_t cs[2], const size_t cr[2], double *recv[2], bool reverse) {
  const int tp = c->box.target_node[2 * axis], tm = c->box.target_node[2 * axis + 1];
  const int me = c->box.myid;
  // forward: buffer 0 (upper selection) goes to +neighbour, buffer 1 to -neighbour.  reverse (CPBK): buffer 0 holds
  // forces of stage-`up` ghosts, which came from the -neighbour, so it goes back there.
  const int dst0 = reverse ? tm : tp, dst1 = reverse ? tp : tm;
  const int src0 = reverse ? tp : tm, src1 = reverse ? tm : tp;
  if (dst0 == me && dst1 == me) {
    recv[0] = c->sbuf[0]; recv[1] = c->sbuf[1];
    return RXG_OK;
  }
  if (!c->comm) { c->err = "rxg_comm_init was not called for a multi-rank decomposition"; return RXG_ERR_NCCL; }
  recv[0] = c->rbuf[0]; recv[1] = c->rbuf[1];
  RXG_NCCL(nccl_api().GroupStart());
  if (cs[0]) RXG_NCCL(nccl_api().Send(c->sbuf[0], cs[0], ncclDouble, dst0, c->comm, c->st));
  if (cr[0]) RXG_NCCL(nccl_api().Recv(c->rbuf[0], cr[0], ncclDouble, src0, c->comm, c->st));
  if (cs[1]) RXG_NCCL(nccl_api().Send(c->sbuf[1], cs[1], ncclDouble, dst1, c->comm, c->st));
  if (cr[1]) RXG_NCCL(nccl_api().Recv(c->rbuf[1], cr[1], ncclDouble, src1, c->comm, c->st));
  RXG_NCCL(nccl_api().GroupEnd());
  c->nccl_msgs += 4;
  return RXG_OK;
}

// rank of the +/- neighbour of rank r along an axis (src/init.F90:79-97: vID x fastest, periodic)
inline int neighbour_rank(const rxg_box &b, int r, int axis, int dir) {
  int v[3] = {r % b.vprocs[0], (r / b.vprocs[0]) % b.vprocs[1], r / (b.vprocs[0] * b.vprocs[1])};
  v[axis] = (v[axis] + dir + b.vprocs[axis]) % b.vprocs[axis];
  return v[0] + v[1] * b.vprocs[0] + v[2] * b.vprocs[0] * b.vprocs[1];
}

// The receive counts of a MODE_COPY / MODE_MOVE axis phase (the reference probes the message size, :335-336) AND the
// collective decision on its capacity traps.  Every rank contributes {ns_up, ns_dn, free atom slots, local error code} to
// ONE all-gather; from the gathered table every rank evaluates every rank's "na+nr > NBUFFER" condition (src/comm.F90:467-472)
// and sees every rank's local error, so all ranks return the same code BEFORE any send, recv or peer push is issued -- a rank
// that stopped alone would leave its neighbours waiting in ncclRecv or in k_peer_pull's spin.
// `cur` = atoms this rank holds before the axis' arrivals, `local_err` = RXG_OK or the trap this rank already hit.
inline int exchange_counts(Ctx *c, int axis, const int ns[2], int nr[2], int cur, int local_err) {
  const int tp = c->box.target_node[2 * axis], tm = c->box.target_node[2 * axis + 1];
  const int me = c->box.myid;
  if (!c->comm) {
    if (tp != me || tm != me) { c->err = "rxg_comm_init was not called for a multi-rank decomposition"; return RXG_ERR_NCCL; }
    nr[0] = ns[0]; nr[1] = ns[1];
    if (local_err != RXG_OK) return local_err;
    if (cur + nr[0] + nr[1] > c->NB) {   // src/comm.F90:467-472
      c->err = "ERROR: over capacity in append_atoms; na+nr > NBUFFER " + std::to_string(cur + nr[0] + nr[1]) + " > " + std::to_string(c->NB);
      return RXG_ERR_NBUFFER;
    }
    return RXG_OK;
  }
  const int nranks = c->box.nprocs;
  if (nranks > 64) { c->err = "exchange_counts: more than 64 ranks"; return RXG_ERR_ARG; }
  c->h_cnt[0] = ns[0]; c->h_cnt[1] = ns[1]; c->h_cnt[2] = c->NB - cur; c->h_cnt[3] = local_err;
  RXG_CUDA(cudaMemcpyAsync(c->d_cnt, c->h_cnt, 4 * sizeof(int), cudaMemcpyHostToDevice, c->st));
  RXG_NCCL(nccl_api().AllGather(c->d_cnt, c->d_cnt + 4, 4, ncclInt, c->comm, c->st));
  RXG_CUDA(cudaMemcpyAsync(c->h_cnt + 4, c->d_cnt + 4, 4 * sizeof(int) * nranks, cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  c->nccl_msgs++;
  const int *T = c->h_cnt + 4;   // [rank][4]
  // stage `up` (buffer 0) travels to the + neighbour, so what I receive in slot 0 is the - neighbour's upper selection
  nr[0] = T[4 * tm + 0]; nr[1] = T[4 * tp + 1];
  int verdict = RXG_OK, who = -1;
  for (int r = 0; r < nranks && verdict == RXG_OK; r++) {
    if (T[4 * r + 3] != RXG_OK) { verdict = T[4 * r + 3]; who = r; break; }
    const int rm = neighbour_rank(c->box, r, axis, -1), rp = neighbour_rank(c->box, r, axis, +1);
    if (T[4 * rm + 0] + T[4 * rp + 1] > T[4 * r + 2]) { verdict = RXG_ERR_NBUFFER; who = r; }
  }
  if (verdict != RXG_OK) {
    if (who != me || c->err.empty() || local_err == RXG_OK)
      c->err = (verdict == RXG_ERR_NBUFFER ? std::string("ERROR: over capacity in append_atoms; na+nr > NBUFFER") : std::string("ERROR: capacity trap")) +
               " on rank " + std::to_string(who) + " (all ranks stop together)";
    return verdict;
  }
  return RXG_OK;
}

// selection + counts of both directions of one axis (one host sync)
inline int select_axis(Ctx *c, int axis, int n, const double dr[3], int skip_dead, int ns[2]) {
  ns[0] = ns[1] = 0;
  if (n <= 0) return RXG_OK;
  RXG_TRY(ensure_blk(c, n));
  const int nblk = cdiv(n, SCAN_BLK);
  for (int k = 0; k < 2; k++) {
    int *blk = c->d_blk + (size_t)k * c->nblk_cap;
    LAUNCH(c, k_sel_count, nblk, SCAN_BLK, 0, c->pos + (size_t)axis * c->NB, c->atype, n, k == 0, c->box.LBOX[axis], dr[axis], skip_dead, blk);
    LAUNCH(c, k_scan_phase2<int>, 1, SCAN_BLK, 0, blk, nblk, c->d_flag + 4 + k);
  }
  RXG_CUDA(cudaMemcpyAsync(c->h_int + 4, c->d_flag + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  ns[0] = c->h_int[4]; ns[1] = c->h_int[5];
  return RXG_OK;
}

// COPYATOMS(MODE_COPY, dr), src/comm.F90:2-100
inline int halo_copy(Ctx *c, const double dr[3]) {
  const int NB = c->NB;
  BoxDev b = make_boxdev(c->box);
  c->cp[0] = c->natoms;
  c->selptr[0] = 0;
  c->halo_self = true;   // every neighbour is this rank itself (vprocs 1 1 1): ghosts are periodic images of residents
  for (int t = 0; t < 6; t++) c->halo_self = c->halo_self && c->box.target_node[t] == c->box.myid;
  if (c->natoms > 0) LAUNCH(c, k_to_norm, cdiv(c->natoms, 256), 256, 0, c->pos, NB, c->natoms, b);
  for (int axis = 0; axis < 3; axis++) {
    const int d0 = 2 * axis + 1, d1 = d0 + 1;
    const int n = c->cp[h_cptridx[d0]];
    int ns[2], nr[2];
    RXG_TRY(select_axis(c, axis, n, dr, 0, ns));
    int local_err = RXG_OK;
    if (c->selptr[d0 - 1] + ns[0] + ns[1] > c->sel_cap) { c->err = "ERROR: over capacity in store_atoms (selection lists)"; local_err = RXG_ERR_NBUFFER; }
    c->selptr[d0] = c->selptr[d0 - 1] + ns[0];
    c->selptr[d1] = c->selptr[d0] + ns[1];
    RXG_TRY(exchange_counts(c, axis, ns, nr, c->cp[d0 - 1], local_err));   // every rank leaves here with the same verdict
    size_t need = (size_t)NE_COPY * (size_t)std::max(std::max(ns[0], ns[1]), std::max(nr[0], nr[1]));
    RXG_TRY(ensure_xbuf(c, need));
    const int nblk = cdiv(n > 0 ? n : 1, SCAN_BLK);
    for (int k = 0; k < 2; k++)
      if (ns[k] > 0)
        LAUNCH(c, k_pack_copy, nblk, SCAN_BLK, 0, c->pos, NB, c->atype, c->q, c->qst, c->hsq, n, axis, k == 0, c->box.LBOX[axis],
               dr[axis], k == 0 ? -c->box.LBOX[axis] : c->box.LBOX[axis], c->d_blk + (size_t)k * c->nblk_cap, ns[k],
               c->sel + c->selptr[d0 - 1 + k], c->sbuf[k]);
    size_t cs[2] = {(size_t)NE_COPY * ns[0], (size_t)NE_COPY * ns[1]}, cr[2] = {(size_t)NE_COPY * nr[0], (size_t)NE_COPY * nr[1]};
    double *rb[2];
    RXG_TRY(exchange_axis(c, axis, cs, cr, rb, false));
    c->cp[d0] = c->cp[d0 - 1] + nr[0];
    c->cp[d1] = c->cp[d0] + nr[1];
    for (int k = 0; k < 2; k++) {
      c->ns[d0 + k] = ns[k]; c->nr[d0 + k] = nr[k];
      if (nr[k] > 0)
        LAUNCH(c, k_unpack_copy, cdiv(nr[k], 256), 256, 0, c->pos, NB, c->atype, c->q, c->qst, c->hsq, nr[k], c->cp[d0 - 1 + k], rb[k],
               c->halo_self ? c->sel + c->selptr[d0 - 1 + k] : (const int *)nullptr, c->natoms, c->gsrc);
    }
  }
  if (c->cp[6] > 0) LAUNCH(c, k_to_real, cdiv(c->cp[6], 256), 256, 0, c->pos, NB, c->cp[6], b);
  return RXG_OK;
}

// COPYATOMS(MODE_QCOPY1|2) (+ which=3: hs,ht only).  `roundtrips` position round trips are applied at the end
// (the reference does one per call, src/comm.F90:222-227,260-264)
inline int halo_refresh(Ctx *c, int which, int roundtrips) {
  const int nf = (which == 2 || which == 5) ? 3 : (which == 4 ? 1 : 2);
  const int nghost = c->cp[6] - c->natoms;
  if (c->halo_self && nghost > 0)
    LAUNCH(c, k_refresh_self, cdiv(nghost, 256), 256, 0, which, c->natoms, c->cp[6], c->gsrc, c->qst, c->hsq, c->hst, c->xs, c->gnb.slot_of,
           c->q, c->spos, c->NB);
  const int seq = ++c->pseq;   // advances identically on every rank (same sequence of refresh calls)
  for (int axis = 0; axis < 3 && !c->halo_self; axis++) {
    const int d0 = 2 * axis + 1;
    const bool self_axis = c->box.target_node[2 * axis] == c->box.myid && c->box.target_node[2 * axis + 1] == c->box.myid;
    if (self_axis) {   // stage `up` sends my upper selection to myself: it arrives as my stage-d0 ghosts (and likewise down)
      PeerDir d;
      for (int k = 0; k < 2; k++) { d.sel[k] = c->sel + c->selptr[d0 - 1 + k]; d.cnt[k] = c->ns[d0 + k]; d.dst0[k] = c->cp[d0 - 1 + k]; }
      const int mx = std::max(d.cnt[0], d.cnt[1]);
      if (mx > 0) LAUNCH(c, k_refresh_axis_local, dim3(cdiv(mx, 256), 2), 256, 0, which, d, c->qst, c->hsq, c->hst, c->xs, c->gnb.slot_of, c->q, c->spos, c->NB);
      continue;
    }
    if (c->peer_ok) {   // neighbours' kernels store into my window over NVLink
      RXG_TRY(halo_refresh_peer_axis(c, which, axis, seq));
      continue;
    }
    const int ns[2] = {c->ns[d0], c->ns[d0 + 1]}, nr[2] = {c->nr[d0], c->nr[d0 + 1]};
    if (ns[0] + ns[1] + nr[0] + nr[1] == 0) continue;
    RXG_TRY(ensure_xbuf(c, (size_t)nf * (size_t)std::max(std::max(ns[0], ns[1]), std::max(nr[0], nr[1]))));
    for (int k = 0; k < 2; k++)
      if (ns[k] > 0) LAUNCH(c, k_pack_vals, cdiv(ns[k], 256), 256, 0, which, c->sel + c->selptr[d0 - 1 + k], ns[k], c->qst, c->hsq, c->hst, c->q, c->spos, c->NB, c->sbuf[k]);
    size_t cs[2] = {(size_t)nf * ns[0], (size_t)nf * ns[1]}, cr[2] = {(size_t)nf * nr[0], (size_t)nf * nr[1]};
    double *rb[2];
    RXG_TRY(exchange_axis(c, axis, cs, cr, rb, false));
    for (int k = 0; k < 2; k++)
      if (nr[k] > 0) LAUNCH(c, k_unpack_vals, cdiv(nr[k], 256), 256, 0, which, nr[k], c->cp[d0 - 1 + k], rb[k], c->qst, c->hsq, c->hst, c->xs, c->gnb.slot_of, c->q, c->spos, c->NB);
  }
  if (roundtrips > 0 && c->cp[6] > 0)
    LAUNCH(c, k_roundtrip, cdiv(c->cp[6], 256), 256, 0, c->pos, c->NB, c->cp[6], make_boxdev(c->box), roundtrips);
  return RXG_OK;
}
inline int halo_qcopy(Ctx *c, int which) { return halo_refresh(c, which, 1); }

// COPYATOMS(MODE_CPBK): stages in reverse order z, y, x (src/comm.F90:72-78)
inline int halo_cpbk(Ctx *c) {
  for (int axis = 2; axis >= 0; axis--) {
    const int d0 = 2 * axis + 1;
    // what I received as ghosts goes back (nr), what I had sent comes home (ns)
    const int nback[2] = {c->nr[d0], c->nr[d0 + 1]}, nhome[2] = {c->ns[d0], c->ns[d0 + 1]};
    if (nback[0] + nback[1] + nhome[0] + nhome[1] == 0) continue;
    RXG_TRY(ensure_xbuf(c, (size_t)3 * (size_t)std::max(std::max(nback[0], nback[1]), std::max(nhome[0], nhome[1]))));
    for (int k = 0; k < 2; k++)
      if (nback[k] > 0) LAUNCH(c, k_pack_force, cdiv(nback[k], 256), 256, 0, c->f, c->NB, c->cp[d0 - 1 + k], nback[k], c->sbuf[k]);
    size_t cs[2] = {(size_t)3 * nback[0], (size_t)3 * nback[1]}, cr[2] = {(size_t)3 * nhome[0], (size_t)3 * nhome[1]};
    double *rb[2];
    RXG_TRY(exchange_axis(c, axis, cs, cr, rb, true));
    for (int k = 0; k < 2; k++)
      if (nhome[k] > 0) LAUNCH(c, k_unpack_force, cdiv(nhome[k], 256), 256, 0, c->f, c->NB, c->sel + c->selptr[d0 - 1 + k], nhome[k], rb[k]);
  }
  return RXG_OK;
}

// COPYATOMS(MODE_MOVE, dr=0): atom migration (src/comm.F90:151-171, 238-256)
inline int halo_move(Ctx *c) {
  const int NB = c->NB;
  BoxDev b = make_boxdev(c->box);
  const double zero[3] = {0.0, 0.0, 0.0};
  double *sp = c->cfg.isPQEq ? c->spos : nullptr;
  const int NE_MOVE = c->cfg.isPQEq ? 15 : rxg::NE_MOVE;
  c->cp[0] = c->natoms;
  if (c->natoms > 0) LAUNCH(c, k_to_norm, cdiv(c->natoms, 256), 256, 0, c->pos, NB, c->natoms, b);
  for (int axis = 0; axis < 3; axis++) {
    const int d0 = 2 * axis + 1, d1 = d0 + 1;
    const int n = c->cp[h_cptridx[d0]];
    int ns[2], nr[2];
    RXG_TRY(select_axis(c, axis, n, zero, 1, ns));
    RXG_TRY(exchange_counts(c, axis, ns, nr, c->cp[d0 - 1], RXG_OK));
    if (ns[0] + ns[1] + nr[0] + nr[1] > 0 && c->lazy_upload) {   // first atom that leaves or arrives: now the rest of the state is needed
      RXG_TRY(c->lazy_upload());
      c->lazy_upload = nullptr;
    }
    RXG_TRY(ensure_xbuf(c, (size_t)NE_MOVE * (size_t)std::max(std::max(ns[0], ns[1]), std::max(nr[0], nr[1]))));
    const int nblk = cdiv(n > 0 ? n : 1, SCAN_BLK);
    // the two selections of one axis are disjoint (an atom cannot leave through both faces), so marking the first
    // direction's atoms dead before packing the second does not change the second's selection or its scan
    for (int k = 0; k < 2; k++)
      if (ns[k] > 0)
        LAUNCH(c, k_pack_move, nblk, SCAN_BLK, 0, c->pos, c->v, NB, c->atype, c->q, c->qst, c->qsfp, c->qsfv, sp, n, axis, k == 0,
               c->box.LBOX[axis], k == 0 ? -c->box.LBOX[axis] : c->box.LBOX[axis], c->d_blk + (size_t)k * c->nblk_cap, ns[k], c->sbuf[k]);
    size_t cs[2] = {(size_t)NE_MOVE * ns[0], (size_t)NE_MOVE * ns[1]}, cr[2] = {(size_t)NE_MOVE * nr[0], (size_t)NE_MOVE * nr[1]};
    double *rb[2];
    RXG_TRY(exchange_axis(c, axis, cs, cr, rb, false));
    c->cp[d0] = c->cp[d0 - 1] + nr[0];
    c->cp[d1] = c->cp[d0] + nr[1];
    for (int k = 0; k < 2; k++)
      if (nr[k] > 0)
        LAUNCH(c, k_unpack_move, cdiv(nr[k], 256), 256, 0, c->pos, c->v, NB, c->atype, c->q, c->qst, c->qsfp, c->qsfv, sp, nr[k], c->cp[d0 - 1 + k], rb[k]);
    c->moved += ns[0] + ns[1] + nr[0] + nr[1];
  }
  const int n6 = c->cp[6];
  // compaction is needed whenever something was sent away or arrived (on one rank: whenever something wrapped)
  if (c->moved > 0) {
    int *flag = c->gb.cell_of;   // scratch (rebuilt by the next binning)
    RXG_TRY(ensure_blk(c, n6));
    LAUNCH(c, k_alive_flag, cdiv(n6, 256), 256, 0, c->atype, n6, flag);
    int nblk = cdiv(n6, SCAN_BLK);
    LAUNCH(c, k_scan_phase1<int>, nblk, SCAN_BLK, 0, flag, (long long)n6, c->d_blk);
    LAUNCH(c, k_scan_phase2<int>, 1, SCAN_BLK, 0, c->d_blk, nblk, c->d_flag + 5);
    LAUNCH(c, k_scan_phase3<int>, nblk, SCAN_BLK, 0, flag, (long long)n6, c->d_blk, c->rowcnt /* >= NB+1 ints */);
    LAUNCH(c, k_move_compact, cdiv(n6, 256), 256, 0, flag, c->rowcnt, n6, NB, c->pos, c->v, c->atype, c->q, c->qst, c->qsfp, c->qsfv, sp, c->tmp);
    RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag + 5, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    RXG_CUDA(cudaStreamSynchronize(c->st));
    int ni = c->h_int[0];
    if (ni > 0) LAUNCH(c, k_move_restore, cdiv(ni, 256), 256, 0, c->tmp, ni, NB, c->pos, c->v, c->atype, c->q, c->qst, c->qsfp, c->qsfv, sp);
    c->natoms = ni;
    c->moved = 0;
  }
  if (c->natoms > 0) LAUNCH(c, k_to_real, cdiv(c->natoms, 256), 256, 0, c->pos, NB, c->natoms, b);
  for (int d = 0; d <= 6; d++) c->cp[d] = c->natoms;
  for (int d = 0; d <= 6; d++) { c->ns[d] = 0; c->nr[d] = 0; c->selptr[d] = 0; }
  return RXG_OK;
}

// ---------------------------------------------------------------------------------------------------
// LINKEDLIST as a counting sort (A1).  Cell of an atom: l = floor((HHi.r - OBOX)/cellDims), src/main.F90:299-308
__global__ void k_cell_ids(const double *__restrict__ pos, const double *__restrict__ atype, int NB, int n, BoxDev b,
                           DevGrid g, int *__restrict__ err) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (nint_d(atype[i]) == 0) { g.cell_of[i] = -1; return; }
  double x = pos[i], y = pos[NB + i], z = pos[2 * NB + i];
  double rn[3];
  rn[0] = sub_rn(dot3_rn(b.Hi[0], b.Hi[3], b.Hi[6], x, y, z), b.OBOX[0]);
  rn[1] = sub_rn(dot3_rn(b.Hi[1], b.Hi[4], b.Hi[7], x, y, z), b.OBOX[1]);
  rn[2] = sub_rn(dot3_rn(b.Hi[2], b.Hi[5], b.Hi[8], x, y, z), b.OBOX[2]);
  int l[3];
  bool ok = true;
  for (int a = 0; a < 3; a++) {
    l[a] = (int)floor(__ddiv_rn(rn[a], g.cs[a]));
    if (l[a] < -g.L || l[a] >= g.nc[a] + g.L) ok = false;
  }
  if (!ok) { g.cell_of[i] = -1; atomicExch(err, 1); return; }
  int cid = ((l[0] + g.L) * g.dim[1] + (l[1] + g.L)) * g.dim[2] + (l[2] + g.L);
  g.cell_of[i] = cid;
  atomicAdd(&g.fill[cid], 1);
}
__global__ void k_cell_scatter(int n, DevGrid g) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cid = g.cell_of[i];
  if (cid < 0) return;
  int k = atomicAdd(&g.fill[cid], 1);
  g.order[g.start[cid] + k] = i;
}
// per cell: sort the member indices DESCENDING (the reference's head insertion yields exactly this order), then
// emit the packed candidate records
__global__ void k_cell_finish(const double *__restrict__ pos, const int *__restrict__ itype, int NB, DevGrid g) {
  int cid = blockIdx.x * blockDim.x + threadIdx.x;
  if (cid >= g.ncell) return;
  int s = g.start[cid], e = g.start[cid + 1];
  for (int a = s + 1; a < e; a++) {
    int key = g.order[a];
    int bidx = a - 1;
    while (bidx >= s && g.order[bidx] < key) { g.order[bidx + 1] = g.order[bidx]; bidx--; }
    g.order[bidx + 1] = key;
  }
  for (int a = s; a < e; a++) {
    int i = g.order[a];
    g.slot_of[i] = a;
    long long packed = (long long)(unsigned)i | ((long long)itype[i] << 32);
    const double x = pos[i], y = pos[NB + i], z = pos[2 * NB + i];
    g.sorted[a] = make_double4(x, y, z, __longlong_as_double(packed));
  }
}

inline int bin_grid(Ctx *c, DevGrid &g) {
  const int n = c->cp[6];
  BoxDev b = make_boxdev(c->box);
  RXG_CUDA(cudaMemsetAsync(g.fill, 0, sizeof(int) * g.ncell, c->st));
  RXG_CUDA(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->st));
  LAUNCH(c, k_cell_ids, cdiv(n, 256), 256, 0, c->pos, c->atype, c->NB, n, b, g, c->d_flag);
  RXG_TRY(ensure_blk(c, g.ncell));
  RXG_TRY(device_scan<int>(c, g.fill, g.ncell, g.start, c->d_blk, (int *)nullptr));
  RXG_CUDA(cudaMemsetAsync(g.fill, 0, sizeof(int) * g.ncell, c->st));
  LAUNCH(c, k_cell_scatter, cdiv(n, 256), 256, 0, n, g);
  LAUNCH(c, k_cell_finish, cdiv(g.ncell, 128), 128, 0, c->pos, c->itype, c->NB, g);
  RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (c->h_int[0]) {
    c->err = "LINKEDLIST: atom outside the layered cell grid";
    return RXG_ERR_STATE;
  }
  return RXG_OK;
}

}   // namespace rxg
