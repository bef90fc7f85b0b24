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

// MODE_COPY stage (store_atoms + self send_recv + append_atoms fused; src/comm.F90:367-528): ghost m is written
// at copyptr(dflag-1)+rank in selection order (stable), shifted by -/+LBOX along the stage's axis.
__global__ void k_copy_append(double *__restrict__ pos, int NB, double *__restrict__ atype, double *__restrict__ q,
                              double2 *__restrict__ qst, double4 *__restrict__ hsq, int *__restrict__ frcindx, int n,
                              int axis, bool upper, double lbox, double dr, double sft, const int *__restrict__ blkoff,
                              int dst0, int cap) {
  int i = blockIdx.x * SCAN_BLK + threadIdx.x;
  int fl = 0;
  if (i < n) fl = in_buffer(upper, lbox, dr, pos[axis * NB + i]);
  int ex = block_excl_scan<int>(fl, nullptr);
  if (!fl) return;
  int m = dst0 + blkoff[blockIdx.x] + ex;
  if (m >= cap) return;   // capacity is checked on the host before the launch
  double p[3] = {pos[i], pos[NB + i], pos[2 * NB + i]};
  p[axis] = add_rn(p[axis], sft);
  pos[m] = p[0]; pos[NB + m] = p[1]; pos[2 * NB + m] = p[2];
  atype[m] = atype[i];
  q[m] = q[i];
  qst[m] = qst[i];
  hsq[m] = hsq[i];
  frcindx[m] = i;
}

// MODE_MOVE stage: atoms that left through the stage's face are re-appended with the shifted coordinate and the
// original is marked dead (atype = -1), src/comm.F90:406-447.
__global__ void k_move_append(double *__restrict__ pos, double *__restrict__ v, int NB, double *__restrict__ atype,
                              double *__restrict__ q, double2 *__restrict__ qst, double *__restrict__ qsfp,
                              double *__restrict__ qsfv, int n, int axis, bool upper, double lbox, double sft,
                              const int *__restrict__ blkoff, int dst0, int cap) {
  int i = blockIdx.x * SCAN_BLK + threadIdx.x;
  int fl = 0;
  if (i < n) fl = in_buffer(upper, lbox, 0.0, pos[axis * NB + i]) && !(atype[i] < 0.0);
  int ex = block_excl_scan<int>(fl, nullptr);
  if (!fl) return;
  int m = dst0 + blkoff[blockIdx.x] + ex;
  if (m >= cap) return;
  double p[3] = {pos[i], pos[NB + i], pos[2 * NB + i]};
  p[axis] = add_rn(p[axis], sft);
  pos[m] = p[0]; pos[NB + m] = p[1]; pos[2 * NB + m] = p[2];
  v[m] = v[i]; v[NB + m] = v[NB + i]; v[2 * NB + m] = v[2 * NB + i];
  atype[m] = atype[i];
  q[m] = q[i];
  qst[m] = qst[i];
  qsfp[m] = qsfp[i];
  qsfv[m] = qsfv[i];
  atype[i] = -1.0;
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
                               const double *__restrict__ qsfv, double *__restrict__ out) {
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
}
__global__ void k_move_restore(const double *__restrict__ in, int n, int NB, double *__restrict__ pos, double *__restrict__ v,
                               double *__restrict__ atype, double *__restrict__ q, double2 *__restrict__ qst,
                               double *__restrict__ qsfp, double *__restrict__ qsfv) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  for (int a = 0; a < 3; a++) { pos[(size_t)a * NB + m] = in[(size_t)a * NB + m]; v[(size_t)a * NB + m] = in[(size_t)(3 + a) * NB + m]; }
  atype[m] = in[(size_t)6 * NB + m];
  q[m] = in[(size_t)7 * NB + m];
  qst[m] = make_double2(in[(size_t)8 * NB + m], in[(size_t)9 * NB + m]);
  qsfp[m] = in[(size_t)10 * NB + m];
  qsfv[m] = in[(size_t)11 * NB + m];
}

// MODE_QCOPY1 / MODE_QCOPY2 on one axis phase: ghosts [lo,hi) take the fresh values of their source atom
// (src/comm.F90:183-207; the source index is what MODE_COPY recorded in frcindx)
__global__ void k_qcopy1(double2 *__restrict__ qst, const int *__restrict__ frcindx, int lo, int hi) {
  int m = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (m < hi) qst[m] = qst[frcindx[m]];
}
__global__ void k_qcopy2(double4 *__restrict__ hsq, double *__restrict__ q, const int *__restrict__ frcindx, int lo, int hi) {
  int m = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (m < hi) {
    double4 t = hsq[frcindx[m]];
    hsq[m] = t;
    q[m] = t.z;
  }
}
// MODE_CPBK on one axis phase: ghost forces are added back onto their source (src/comm.F90:385-396,474-482)
__global__ void k_cpbk(double *__restrict__ f, int NB, const int *__restrict__ frcindx, int lo, int hi) {
  int m = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (m < hi) {
    int s = frcindx[m];
    atomicAdd(&f[s], f[m]);
    atomicAdd(&f[NB + s], f[NB + m]);
    atomicAdd(&f[2 * NB + s], f[2 * NB + m]);
  }
}

__global__ void k_iota(int *__restrict__ a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
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
    RXG_CUDA(cudaMalloc(&c->d_blk, sizeof(int) * c->nblk_cap));
    RXG_CUDA(cudaMalloc(&c->d_blk64, sizeof(long long) * c->nblk_cap));
  }
  return RXG_OK;
}

// COPYATOMS(MODE_COPY, dr): src/comm.F90:2-100 for a rank whose six neighbours are itself (periodic self images)
inline int halo_copy(Ctx *c, const double dr[3]) {
  const int NB = c->NB;
  BoxDev b = make_boxdev(c->box);
  c->cp[0] = c->natoms;
  if (c->natoms > 0) {
    LAUNCH(c, k_to_norm, cdiv(c->natoms, 256), 256, 0, c->pos, NB, c->natoms, b);
    LAUNCH(c, k_iota, cdiv(c->natoms, 256), 256, 0, c->frcindx, c->natoms);
  }
  for (int d = 1; d <= 6; d++) {
    const int axis = (d - 1) / 2;
    const bool upper = (d % 2) == 1;
    const int n = c->cp[h_cptridx[d]];
    const double sft = upper ? -c->box.LBOX[axis] : c->box.LBOX[axis];
    int ns = 0;
    if (n > 0) {
      RXG_TRY(ensure_blk(c, n));
      int nblk = cdiv(n, SCAN_BLK);
      LAUNCH(c, k_sel_count, nblk, SCAN_BLK, 0, c->pos + (size_t)axis * NB, c->atype, n, upper, c->box.LBOX[axis], dr[axis], 0, c->d_blk);
      LAUNCH(c, k_scan_phase2<int>, 1, SCAN_BLK, 0, c->d_blk, nblk, c->d_flag + 4);
      RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag + 4, sizeof(int), cudaMemcpyDeviceToHost, c->st));
      RXG_CUDA(cudaStreamSynchronize(c->st));
      ns = c->h_int[0];
      if (c->cp[d - 1] + ns > NB) {
        c->err = "ERROR: over capacity in append_atoms (NBUFFER)";
        return RXG_ERR_NBUFFER;
      }
      if (ns > 0)
        LAUNCH(c, k_copy_append, nblk, SCAN_BLK, 0, c->pos, NB, c->atype, c->q, c->qst, c->hsq, c->frcindx, n, axis, upper,
               c->box.LBOX[axis], dr[axis], sft, c->d_blk, c->cp[d - 1], NB);
    }
    c->cp[d] = c->cp[d - 1] + ns;
  }
  if (c->cp[6] > 0) LAUNCH(c, k_to_real, cdiv(c->cp[6], 256), 256, 0, c->pos, NB, c->cp[6], b);
  return RXG_OK;
}

// COPYATOMS(MODE_QCOPY1|2): value refresh of the ghosts created by the last MODE_COPY (+ the position round trip)
inline int halo_qcopy(Ctx *c, int which) {
  BoxDev b = make_boxdev(c->box);
  for (int ph = 0; ph < 3; ph++) {
    int lo = c->cp[2 * ph], hi = c->cp[2 * ph + 2];
    if (hi > lo) {
      if (which == 1) LAUNCH(c, k_qcopy1, cdiv(hi - lo, 256), 256, 0, c->qst, c->frcindx, lo, hi);
      else LAUNCH(c, k_qcopy2, cdiv(hi - lo, 256), 256, 0, c->hsq, c->q, c->frcindx, lo, hi);
    }
  }
  if (c->cp[6] > 0) LAUNCH(c, k_roundtrip, cdiv(c->cp[6], 256), 256, 0, c->pos, c->NB, c->cp[6], b, 1);
  return RXG_OK;
}

// COPYATOMS(MODE_CPBK): reverse order z, y, x
inline int halo_cpbk(Ctx *c) {
  for (int ph = 2; ph >= 0; ph--) {
    int lo = c->cp[2 * ph], hi = c->cp[2 * ph + 2];
    if (hi > lo) LAUNCH(c, k_cpbk, cdiv(hi - lo, 256), 256, 0, c->f, c->NB, c->frcindx, lo, hi);
  }
  return RXG_OK;
}

// COPYATOMS(MODE_MOVE, dr=0): migration; on one rank atoms that left the box re-enter through the opposite face
inline int halo_move(Ctx *c) {
  const int NB = c->NB;
  BoxDev b = make_boxdev(c->box);
  c->cp[0] = c->natoms;
  if (c->natoms > 0) LAUNCH(c, k_to_norm, cdiv(c->natoms, 256), 256, 0, c->pos, NB, c->natoms, b);
  for (int d = 1; d <= 6; d++) {
    const int axis = (d - 1) / 2;
    const bool upper = (d % 2) == 1;
    const int n = c->cp[h_cptridx[d]];
    const double sft = upper ? -c->box.LBOX[axis] : c->box.LBOX[axis];
    int ns = 0;
    if (n > 0) {
      RXG_TRY(ensure_blk(c, n));
      int nblk = cdiv(n, SCAN_BLK);
      LAUNCH(c, k_sel_count, nblk, SCAN_BLK, 0, c->pos + (size_t)axis * NB, c->atype, n, upper, c->box.LBOX[axis], 0.0, 1, c->d_blk);
      LAUNCH(c, k_scan_phase2<int>, 1, SCAN_BLK, 0, c->d_blk, nblk, c->d_flag + 4);
      RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag + 4, sizeof(int), cudaMemcpyDeviceToHost, c->st));
      RXG_CUDA(cudaStreamSynchronize(c->st));
      ns = c->h_int[0];
      if (c->cp[d - 1] + ns > NB) {
        c->err = "ERROR: over capacity in append_atoms (NBUFFER)";
        return RXG_ERR_NBUFFER;
      }
      if (ns > 0)
        LAUNCH(c, k_move_append, nblk, SCAN_BLK, 0, c->pos, c->v, NB, c->atype, c->q, c->qst, c->qsfp, c->qsfv, n, axis, upper,
               c->box.LBOX[axis], sft, c->d_blk, c->cp[d - 1], NB);
    }
    c->cp[d] = c->cp[d - 1] + ns;
  }
  const int n6 = c->cp[6];
  if (n6 > c->natoms) {   // something moved: compact (stable)
    int *flag = c->gb.cell_of;   // scratch (rebuilt by the next binning)
    int *dst = c->gb.order;
    RXG_TRY(ensure_blk(c, n6));
    LAUNCH(c, k_alive_flag, cdiv(n6, 256), 256, 0, c->atype, n6, flag);
    // dst needs n6+1 entries; order[] has NB >= n6 entries; keep the total in d_flag[5]
    int nblk = cdiv(n6, SCAN_BLK);
    LAUNCH(c, k_scan_phase1<int>, nblk, SCAN_BLK, 0, flag, (long long)n6, c->d_blk);
    LAUNCH(c, k_scan_phase2<int>, 1, SCAN_BLK, 0, c->d_blk, nblk, c->d_flag + 5);
    LAUNCH(c, k_scan_phase3<int>, nblk, SCAN_BLK, 0, flag, (long long)n6, c->d_blk, c->rowcnt /*>= NB+1 ints*/);
    (void)dst;
    LAUNCH(c, k_move_compact, cdiv(n6, 256), 256, 0, flag, c->rowcnt, n6, NB, c->pos, c->v, c->atype, c->q, c->qst, c->qsfp, c->qsfv, c->tmp);
    RXG_CUDA(cudaMemcpyAsync(c->h_int, c->d_flag + 5, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    RXG_CUDA(cudaStreamSynchronize(c->st));
    int ni = c->h_int[0];
    LAUNCH(c, k_move_restore, cdiv(ni, 256), 256, 0, c->tmp, ni, NB, c->pos, c->v, c->atype, c->q, c->qst, c->qsfp, c->qsfv);
    c->natoms = ni;
  }
  if (c->natoms > 0) LAUNCH(c, k_to_real, cdiv(c->natoms, 256), 256, 0, c->pos, NB, c->natoms, b);
  for (int d = 0; d <= 6; d++) c->cp[d] = c->natoms;
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
    long long packed = (long long)(unsigned)i | ((long long)itype[i] << 32);
    g.sorted[a] = make_double4(pos[i], pos[NB + i], pos[2 * NB + i], __longlong_as_double(packed));
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
