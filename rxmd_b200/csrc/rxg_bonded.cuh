// rxg_bonded.cuh -- bond orders (kernel group B), energy/force terms (C1-C6), bonded-force finalisation (C7),
// FORCE orchestration and the device-resident integrator kernels.
// Reference: src/bo.F90 (BOCALC), src/pot.F90 (FORCE and all terms), src/main.F90:64-98,192-207.
//
// Restructuring relative to the reference (results equal up to fp64 summation order):
//  * every ForceB/ForceBbo call is linear in per-bond data, so energy kernels only ACCUMULATE the coefficient
//    triple (cf1,cf2,cf3 of src/pot.F90:1331) per directed bond slot, and cdbnd per atom; one gather kernel per
//    atom then produces the bond forces and ccbnd without atomics (k_final1), a second applies ccbnd (k_final2).
//  * the serial, order-dependent ForceBondedTerms loop (src/pot.F90:125-140) becomes the
//    predicate `j < i` on reference local indices inside k_final1.
//  * ENbond evaluates each resident's full 10 A row (both directions of a pair) so that no force is scattered;
//    energies are still counted once, on the side the reference counts them (gid(j) < gid(i)).
#pragma once
#include "rxg_lists_qeq.cuh"

namespace rxg {

constexpr double MAXANGLE = 0.999999999999, MINANGLE = -0.999999999999, NSMALL = 1e-10;   // src/module.F90:85-87
constexpr double PI_RX = 3.14159265358979;                                                  // src/module.F90:90
constexpr double MINBO0 = 1e-4, CUTOF2_ESUB = 1e-4;                                         // src/module.F90:61-62
constexpr double CECHRGE = 23.02;                                                           // src/module.F90:683
constexpr double RCHB2 = 100.0;                                                             // src/module.F90:677-678

// d_acc layout for FORCE: 16+k = PE(k) k=1..13 ; 32.. reserved (nnz) ; 34..39 astr ; 40..45 kinetic astr ; 48 KE ; 49 sum q
constexpr int ACC_PE = 80, ACC_ASTR = 98;   // FORCE's own block of d_acc[128]: its charge-independent part may run beside the QEq CG (slots 0-39)

struct Bonds {   // per directed bond slot, compact: slot of (atom i, s-th neighbour) = ptr[i] + s
  int MAXN;
  const int *ptr, *own;   // first slot of an atom; atom that owns a slot
  const int *cnt, *lst, *idx;
  double *BO0, *BO1, *BO2, *BO3, *dln1, *dln2, *dln3, *dBOp, *A0, *A1, *A2, *A3;
  double *cB0, *cB1, *cB2, *cdslot;
};

// ---------------------------------------------------------------------------------------------------
// B1: BOPRIM, src/bo.F90:28-118.  One thread per atom; each directed slot evaluates the (bitwise symmetric)
// pair expression itself instead of mirroring through nbrindx; deltap(i,1) is the atom's own row sum.
__global__ void __launch_bounds__(128) k_boprim(int ntot, const double *__restrict__ pos, int NB, const int *__restrict__ itype,
                                                const DevFF *__restrict__ ffp, Bonds B, double *__restrict__ deltap1,
                                                double *__restrict__ deltap2, double *__restrict__ cdbnd,
                                                double *__restrict__ ccbnd, double *__restrict__ s3) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot) return;
  const DevFF &ff = *ffp;
  const int ity = itype[i];
  cdbnd[i] = 0.0;
  ccbnd[i] = 0.0;
  s3[i] = 0.0; s3[NB + i] = 0.0; s3[2 * (size_t)NB + i] = 0.0;
  if (ity <= 0) { deltap1[i] = 0.0; deltap2[i] = 0.0; return; }
  const double xi = pos[i], yi = pos[NB + i], zi = pos[2 * NB + i];
  double dp = -ff.Val[ity - 1];
  const int n = B.cnt[i];
  for (int s = 0; s < n; s++) {
    size_t a = (size_t)B.ptr[i] + s;
    int j = B.lst[a];
    int x = ff.inxn2[(ity - 1) + ff.nso * (itype[j] - 1)] - 1;
    double dx = sub_rn(xi, pos[j]), dy = sub_rn(yi, pos[NB + j]), dz = sub_rn(zi, pos[2 * NB + j]);
    double dr2 = dist2_rn(dx, dy, dz);
    double b0 = 0, b1 = 0, b2 = 0, b3 = 0, l1 = 0, l2 = 0, l3 = 0, db = 0;
    if (x >= 0 && dr2 <= ff.rc2[x]) {
      // dr2**p as exp(p*log(dr2)) with one shared logarithm (src/bo.F90:67-69); relative error ~1e-15
      const double lg = log(dr2);
      double a1 = ff.cBOp1[x] * exp(ff.pbo2h[x] * lg);
      double a2 = ff.cBOp3[x] * exp(ff.pbo4h[x] * lg);
      double a3 = ff.cBOp5[x] * exp(ff.pbo6h[x] * lg);
      b1 = ff.swtch[3 * x] * exp(a1);
      b2 = ff.swtch[3 * x + 1] * exp(a2);
      b3 = ff.swtch[3 * x + 2] * exp(a3);
      b1 = (1.0 + ff.cutoff_vpar30) * b1;
      if ((b1 + b2) + b3 > ff.cutoff_vpar30) {
        l1 = ff.swtch[3 * x] * ff.pbo2[x] * a1 / dr2;
        l2 = ff.swtch[3 * x + 1] * ff.pbo4[x] * a2 / dr2;
        l3 = ff.swtch[3 * x + 2] * ff.pbo6[x] * a3 / dr2;
        db = (b1 * l1 + b2 * l2) + b3 * l3;
        b1 = b1 - ff.cutoff_vpar30;
        b0 = (b1 + b2) + b3;
        dp += b0;
      } else {
        b1 = b2 = b3 = 0.0;
      }
    }
    B.BO0[a] = b0; B.BO1[a] = b1; B.BO2[a] = b2; B.BO3[a] = b3;
    B.dln1[a] = l1; B.dln2[a] = l2; B.dln3[a] = l3; B.dBOp[a] = db;
    B.cB0[a] = 0.0; B.cB1[a] = 0.0; B.cB2[a] = 0.0; B.cdslot[a] = 0.0;
  }
  deltap1[i] = dp;
  deltap2[i] = dp + ff.Val[ity - 1] - ff.Valval[ity - 1];   // src/bo.F90:149-152
}

// B2: BOFULL, src/bo.F90:121-298; each slot computes its own side (A2/A3 are side specific, the rest symmetric)
__global__ void __launch_bounds__(128) k_bofull(int ntot, const int *__restrict__ itype, const DevFF *__restrict__ ffp, Bonds B,
                                                const double *__restrict__ deltap1, const double *__restrict__ deltap2,
                                                double *__restrict__ delta) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot) return;
  const DevFF &ff = *ffp;
  const int ity = itype[i];
  if (ity <= 0) { delta[i] = 0.0; return; }
  const double Vi = ff.Val[ity - 1];
  const double dpi = deltap1[i], dp2i = deltap2[i];
  const double e1i = exp(-ff.vpar1 * dpi), e2i = exp(-ff.vpar2 * dpi);
  double dsum = 0.0;
  const int n = B.cnt[i];
  for (int s = 0; s < n; s++) {
    size_t a = (size_t)B.ptr[i] + s;
    int j = B.lst[a];
    int jty = itype[j];
    int x = ff.inxn2[(ity - 1) + ff.nso * (jty - 1)] - 1;
    if (x < 0) continue;
    const double Vj = ff.Val[jty - 1];
    const double dpj = deltap1[j], dp2j = deltap2[j];
    const double e1j = exp(-ff.vpar1 * dpj), e2j = exp(-ff.vpar2 * dpj);
    double fn2 = e1i + e1j;
    double fn3 = (-1.0 / ff.vpar2) * log(0.5 * (e2i + e2j));
    double fn23 = fn2 + fn3;
    double BOp0 = B.BO0[a], bp2 = B.BO2[a], bp3 = B.BO3[a];
    double fn1 = 0.5 * ((Vi + fn2) / (Vi + fn23) + (Vj + fn2) / (Vj + fn23));
    const bool no_ovc = ff.ovc[x] < 1e-3, no_v13 = ff.v13cor[x] < 1e-3;
    if (no_ovc) fn1 = 1.0;
    double BOpsqr = BOp0 * BOp0;
    double p3 = ff.pboc3[x], p4 = ff.pboc4[x], p5 = ff.pboc5[x];
    double fn4 = 1.0 / (1.0 + exp(-p3 * (p4 * BOpsqr - dp2i) + p5));
    double fn5 = 1.0 / (1.0 + exp(-p3 * (p4 * BOpsqr - dp2j) + p5));
    if (no_v13) { fn4 = 1.0; fn5 = 1.0; }
    double fn45 = fn4 * fn5, fn145 = fn1 * fn45, fn1145 = fn1 * fn145;
    double B0 = BOp0 * fn145, B2 = bp2 * fn1145, B3 = bp3 * fn1145;
    if (B0 < 1e-10) B0 = 0.0;
    if (B2 < 1e-10) B2 = 0.0;
    if (B3 < 1e-10) B3 = 0.0;
    double B1 = B0 - B2 - B3;
    double u1ij = Vi + fn23, u1ji = Vj + fn23;
    double u1ij_inv2 = 1.0 / (u1ij * u1ij), u1ji_inv2 = 1.0 / (u1ji * u1ji);
    double Cf1A = 0.5 * fn3 * (u1ij_inv2 + u1ji_inv2);
    double Cf1B = -0.5 * ((u1ij - fn3) * u1ij_inv2 + (u1ji - fn3) * u1ji_inv2);
    double e22 = e2i + e2j;
    double Cf1ij = (-Cf1A * ff.pboc1[x] * e1i) + (Cf1B * e2i) / e22;
    double p34 = p3 * p4;
    double u45ij = p5 + p3 * dp2i - p34 * BOpsqr;
    double u45ji = p5 + p3 * dp2j - p34 * BOpsqr;
    double x45ij = exp(u45ij), x45ji = exp(u45ji);
    double ex1 = 1.0 / (1.0 + x45ij), ex2 = 1.0 / (1.0 + x45ji);
    double ex12 = ex1 * ex2;
    double Cf45ij = -x45ij * ex12 * ex1, Cf45ji = -x45ji * ex12 * ex2;
    if (no_ovc) Cf1ij = 0.0;
    if (no_v13) { Cf45ij = 0.0; Cf45ji = 0.0; }
    double fn45_inv = 1.0 / fn45;
    double Cf1ij_div1 = Cf1ij / fn1;
    B.BO0[a] = B0; B.BO1[a] = B1; B.BO2[a] = B2; B.BO3[a] = B3;
    B.A0[a] = fn145;
    B.A1[a] = -2.0 * p34 * BOp0 * (Cf45ij + Cf45ji) * fn45_inv;
    double a2 = Cf1ij_div1 + (p3 * Cf45ij * fn45_inv);
    B.A2[a] = a2;
    B.A3[a] = a2 + Cf1ij_div1;
    dsum += B0;
  }
  delta[i] = -Vi + dsum;   // src/bo.F90:292-295
}

// ---------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_add(double (&v)[NV], double *__restrict__ acc) {
  block_accumulate<NV>(v, acc);
}

__device__ __forceinline__ void atomic_add3(double *__restrict__ f, int NB, int i, double x, double y, double z) {
  atomicAdd(&f[i], x);
  atomicAdd(&f[NB + i], y);
  atomicAdd(&f[2 * NB + i], z);
}

// gather packs: {x,y,z,q} by atom (bonded kernels) and, in cell-sorted slot order, {x,y,z,q} + {itype,gid,atom}
// (non-bonded kernels: the neighbours of a stencil run are contiguous slots)
__global__ void k_pack_pq(int ntot, const double *__restrict__ pos, int NB, const double *__restrict__ q,
                          const int *__restrict__ itype, const int *__restrict__ gid, const int *__restrict__ slot_of,
                          double4 *__restrict__ pqa, double4 *__restrict__ pqs, int4 *__restrict__ tgs, int2 *__restrict__ gts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot) return;
  double4 p = make_double4(pos[i], pos[NB + i], pos[2 * (size_t)NB + i], q[i]);
  pqa[i] = p;
  int s = slot_of[i];
  pqs[s] = p;
  tgs[s] = make_int4(itype[i], gid[i], i, 0);
  gts[s] = make_int2(gid[i], itype[i]);   // the 8 bytes the half-list test of k_enbond_half gathers per list entry
}

// C6: ENbond, src/pot.F90:676-781.  One warp per resident row of the 10 A list.
// HALF=true is the literal form (gid(j)<gid(i), forces scattered to j); HALF=false evaluates the full row.
template <bool HALF>
__global__ void __launch_bounds__(256) k_enbond(int ntot, int natoms, const long long *__restrict__ rowbeg,
                                                const long long *__restrict__ rowend, const int *__restrict__ col,
                                                const double4 *__restrict__ pqs, const int4 *__restrict__ tgs,
                                                const DevFF *__restrict__ ffp, double *__restrict__ f, double *__restrict__ fsl, int NB,
                                                double *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // rows are walked in cell order
  double part[3] = {0.0, 0.0, 0.0};
  double vir[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  int4 ti = make_int4(0, 0, natoms, 0);
  if (slot < ntot) ti = tgs[slot];
  const int i = ti.z;
  if (i < natoms) {
    const DevFF &ff = *ffp;
    const double4 pi = pqs[slot];
    double fx = 0, fy = 0, fz = 0;
    long long s = rowbeg[i], e = rowend[i];
    for (long long k = s + lane; k < e; k += 32) {
      int js = __ldcs(col + k) & COL_MASK;
      int4 tj = tgs[js];
      bool lower = tj.y < ti.y;
      if (HALF ? !lower : (tj.y == ti.y)) continue;
      double4 pj = ldg256(pqs + js);
      double dx = sub_rn(pi.x, pj.x), dy = sub_rn(pi.y, pj.y), dz = sub_rn(pi.z, pj.z);
      double dr2 = dist2_rn(dx, dy, dz);
      if (!(dr2 <= ff.rctap2)) continue;
      int inxn = ff.inxn2[(ti.x - 1) + ff.nso * (tj.x - 1)];
      int itb = (int)mul_rn(dr2, ff.UDRi);
      if (inxn <= 0 || itb < 1 || itb + 1 > ff.ntable) continue;   // out of bounds in the reference
      double drtb = mul_rn(sub_rn(dr2, mul_rn((double)itb, ff.UDR)), ff.UDRi);
      double drtb1 = 1.0 - drtb;
      const double4 *T = ff.TBL_nb + (size_t)(inxn - 1) * ff.ntable + (itb - 1);
      double4 T0 = ldg256(T), T1 = ldg256(T + 1);
      double qij = pi.w * pj.w;
      double PEvdw = drtb1 * T0.x + drtb * T1.x;
      double CEvdw = drtb1 * T0.y + drtb * T1.y;
      double PEclmb = (drtb1 * T0.z + drtb * T1.z) * qij;
      double CEclmb = (drtb1 * T0.w + drtb * T1.w) * qij;
      if (lower) { part[0] += PEvdw; part[1] += PEclmb; }
      double cc = CEvdw + CEclmb;
      fx -= cc * dx; fy -= cc * dy; fz -= cc * dz;
      if (!HALF) {   // pair virial, half from each side: sum_atoms pos_a f_b of src/pot.F90:65-72 restricted to this pair
        double h = -0.5 * cc;
        vir[0] += h * dx * dx; vir[1] += h * dy * dy; vir[2] += h * dz * dz;
        vir[3] += h * dy * dz; vir[4] += h * dz * dx; vir[5] += h * dx * dy;
      }
      // the partner's share goes to the SLOT-ordered accumulator: the partners of consecutive lanes are consecutive slots
      // (a stencil run), so the fp64 reductions of a warp fall into a few sectors; indexed by atom they were 32 scattered ones
      if (HALF) atomic_add3(fsl, NB, js, cc * dx, cc * dy, cc * dz);
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) {
      atomic_add3(f, NB, i, fx, fy, fz);
      part[2] = CECHRGE * (ff.chi[ti.x - 1] * pi.w + 0.5 * ff.eta[ti.x - 1] * pi.w * pi.w);   // src/pot.F90:708
    }
  }
  block_add<3>(part, acc + ACC_PE + 11);
  if (!HALF) block_add<6>(vir, acc + ACC_ASTR);
}

// The literal half-list form of ENbond (gid(j) < gid(i), partner forces scattered), as the production path runs it.
// k_enbond<true> tests `lower` per list entry and lets the ~50 % of lanes that fail idle through the expensive part (a 32-byte
// position gather, two 32-byte table-node gathers, ~40 fp64 operations, three reductions): a latency chain of four dependent
// gathers per 32 entries for 16 useful pairs.  Here the test reads 8 bytes per entry ({gid, type} by slot, nearly contiguous),
// survivors queue up in shared memory (ring of 64 per warp, partner slot | type << 26) and the expensive part runs on full
// warps of them.  Same pairs, same arithmetic per pair; only the order in which a row's pair forces are summed changes.
__global__ void __launch_bounds__(256) k_enbond_half(int ntot, int natoms, const long long *__restrict__ rowbeg,
                                                     const long long *__restrict__ rowend, const int *__restrict__ col,
                                                     const double4 *__restrict__ pqs, const int4 *__restrict__ tgs,
                                                     const int2 *__restrict__ gts, const DevFF *__restrict__ ffp,
                                                     double *__restrict__ f, double *__restrict__ fsl, int NB, double *__restrict__ acc) {
  __shared__ int sh_q[8][64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // rows are walked in cell order
  double part[3] = {0.0, 0.0, 0.0};
  int4 ti = make_int4(0, 0, natoms, 0);
  if (slot < ntot) ti = tgs[slot];
  const int i = ti.z;
  if (i < natoms) {
    const DevFF &ff = *ffp;
    const double4 pi = pqs[slot];
    double fx = 0, fy = 0, fz = 0;
    int qh = 0, qn = 0;   // ring state, identical on every lane
    auto eval = [&](bool full) {
      if (full || lane < qn) {
        const unsigned e = (unsigned)sh_q[wid][(qh + lane) & 63];
        const int js = (int)(e & 0x3ffffffu), tjx = (int)(e >> 26);
        const double4 pj = ldg256(pqs + js);
        const double dx = sub_rn(pi.x, pj.x), dy = sub_rn(pi.y, pj.y), dz = sub_rn(pi.z, pj.z);
        const double dr2 = dist2_rn(dx, dy, dz);
        const int inxn = ff.inxn2[(ti.x - 1) + ff.nso * (tjx - 1)];
        const int itb = (int)mul_rn(dr2, ff.UDRi);
        // (out of bounds in the reference)
        if (dr2 <= ff.rctap2 && inxn > 0 && itb >= 1 && itb + 1 <= ff.ntable) {
          const double drtb = mul_rn(sub_rn(dr2, mul_rn((double)itb, ff.UDR)), ff.UDRi);
          const double drtb1 = 1.0 - drtb;
          const double4 *T = ff.TBL_nb + (size_t)(inxn - 1) * ff.ntable + (itb - 1);
          const double4 T0 = ldg256(T), T1 = ldg256(T + 1);
          const double qij = pi.w * pj.w;
          const double PEvdw = drtb1 * T0.x + drtb * T1.x;
          const double CEvdw = drtb1 * T0.y + drtb * T1.y;
          const double PEclmb = (drtb1 * T0.z + drtb * T1.z) * qij;
          const double CEclmb = (drtb1 * T0.w + drtb * T1.w) * qij;
          part[0] += PEvdw; part[1] += PEclmb;
          const double cc = CEvdw + CEclmb;
          fx -= cc * dx; fy -= cc * dy; fz -= cc * dz;
          // the partner's share goes to the SLOT-ordered accumulator (k_fsl_to_f folds it into f)
          atomic_add3(fsl, NB, js, cc * dx, cc * dy, cc * dz);
        }
      }
      __syncwarp();
    };
    const long long s = rowbeg[i], e = rowend[i];
    for (long long k0 = s; k0 < e; k0 += 32) {
      const long long k = k0 + lane;
      const bool in = k < e;
      const int js = in ? (__ldcs(col + k) & COL_MASK) : 0;
      int2 gt = make_int2(0, 0);
      if (in) gt = gts[js];
      const bool lower = in && gt.x < ti.y;
      const unsigned m = __ballot_sync(0xffffffffu, lower);
      if (lower) sh_q[wid][(qh + qn + __popc(m & ((1u << lane) - 1u))) & 63] = js | (gt.y << 26);
      qn += __popc(m);
      __syncwarp();
      if (qn >= 32) { eval(true); qh = (qh + 32) & 63; qn -= 32; }
    }
    if (qn > 0) eval(false);
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) {
      atomic_add3(f, NB, i, fx, fy, fz);
      part[2] = CECHRGE * (ff.chi[ti.x - 1] * pi.w + 0.5 * ff.eta[ti.x - 1] * pi.w * pi.w);   // src/pot.F90:708
    }
  }
  block_add<3>(part, acc + ACC_PE + 11);
}

// Elnpr preparation loop, src/pot.F90:183-209 (all atoms)
__global__ void k_elnpr_prep(int ntot, const int *__restrict__ itype, const DevFF *__restrict__ ffp,
                             const double *__restrict__ delta, double *__restrict__ nlp, double *__restrict__ dDlp,
                             double *__restrict__ deltalp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot) return;
  int ity = itype[i];
  if (ity <= 0) return;
  const DevFF &ff = *ffp;
  int t = ity - 1;
  double deltaE = -ff.Vale[t] + ff.Val[t] + delta[i];
  double dEh = deltaE * 0.5;
  int idEh = (int)dEh;
  double u = 2.0 + deltaE - 2 * idEh;
  double explp1 = exp(-ff.plp1[t] * (u * u));
  dDlp[i] = 2.0 * ff.plp1[t] * explp1 * u;
  double nl = explp1 - (double)idEh;
  nlp[i] = nl;
  deltalp[i] = (ff.mass[t] > 21.0) ? 0.0 : ff.nlpopt[t] - nl;
}

// C1 + C2: Ebond (src/pot.F90:926-977) and the Elnpr main loop (src/pot.F90:211-306); one thread per resident,
// all coefficient updates land in the atom's own slots.
__global__ void __launch_bounds__(128) k_ebond_elnpr(int natoms, const int *__restrict__ itype, const int *__restrict__ gid,
                                                     const DevFF *__restrict__ ffp, Bonds B, const double *__restrict__ delta,
                                                     const double *__restrict__ dDlp, const double *__restrict__ deltalp,
                                                     double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double part[4] = {0.0, 0.0, 0.0, 0.0};   // PE(1..4)
  if (i < natoms && itype[i] > 0) {
    const DevFF &ff = *ffp;
    const int ity = itype[i], t = ity - 1, iid = gid[i];
    const int n = B.cnt[i];
    double sum_ovun1 = 0.0, sum_ovun2 = 0.0;
    for (int s = 0; s < n; s++) {
      size_t a = (size_t)B.ptr[i] + s;
      int j = B.lst[a];
      int x = ff.inxn2[t + ff.nso * (itype[j] - 1)] - 1;
      if (x < 0) continue;
      double b0 = B.BO0[a], b1 = B.BO1[a], b2 = B.BO2[a], b3 = B.BO3[a];
      sum_ovun1 += ff.povun1[x] * ff.Desig[x] * b0;
      sum_ovun2 += (delta[j] - deltalp[j]) * (b2 + b3);
      if (gid[j] < iid) {   // Ebond
        double bp = pow(b1, ff.pbe2[x]);
        double ex = exp(ff.pbe1[x] * (1.0 - bp));
        part[0] += -ff.Desig[x] * b1 * ex - ff.Depi[x] * b2 - ff.Depipi[x] * b3;
        double CEbo = -ff.Desig[x] * ex * (1.0 - ff.pbe1[x] * ff.pbe2[x] * bp);
        B.cB0[a] += CEbo;
        B.cB1[a] += -ff.Depi[x] - CEbo;
        B.cB2[a] += -ff.Depipi[x] - CEbo;
      }
    }
    const double dlp = deltalp[i], dl = delta[i], dD = dDlp[i];
    double expvd2 = exp(-75.0 * dlp);
    double dElp = ff.plp2[t] * ((1.0 + expvd2) + 75.0 * dlp * expvd2) / ((1.0 + expvd2) * (1.0 + expvd2));
    double expovun1 = ff.povun3[t] * exp(ff.povun4[t] * sum_ovun2);
    double deltalpcorr = dl - dlp / (1.0 + expovun1);
    double expovun2 = exp(ff.povun2[t] * deltalpcorr);
    double DlpV_i = 1.0 / (deltalpcorr + ff.Val[t] + 1e-8);
    double expovun2n = 1.0 / expovun2;
    double expovun6 = exp(ff.povun6[t] * deltalpcorr);
    double expovun8 = ff.povun7[t] * exp(ff.povun8[t] * sum_ovun2);
    double div1 = 1.0 / (1.0 + expovun1), div2 = 1.0 / (1.0 + expovun2), div2n = 1.0 / (1.0 + expovun2n),
           div8 = 1.0 / (1.0 + expovun8);
    double PElp = ff.plp2[t] * dlp / (1.0 + expvd2);
    double PEover = sum_ovun1 * DlpV_i * deltalpcorr * div2;
    double PEunder = -ff.povun5[t] * (1.0 - expovun6) * div2n * div8;
    part[1] = PElp; part[2] = PEover; part[3] = PEunder;
    double CElp1 = dElp * dD;
    double CEover1 = deltalpcorr * DlpV_i * div2;
    double CEover2 = sum_ovun1 * DlpV_i * div2 * (1.0 - deltalpcorr * DlpV_i - ff.povun2[t] * deltalpcorr * div2n);
    double CEover3 = CEover2 * (1.0 - dD * div1);
    double CEover4 = CEover2 * dlp * ff.povun4[t] * expovun1 * (div1 * div1);
    double CEunder1 = (ff.povun5[t] * ff.povun6[t] * expovun6 * div8 + PEunder * ff.povun2[t] * expovun2n) * div2n;
    double CEunder2 = -PEunder * ff.povun8[t] * expovun8 * div8;
    double CEunder3 = CEunder1 * (1.0 - dD * div1);
    double CEunder4 = CEunder1 * dlp * ff.povun4[t] * expovun1 * (div1 * div1) + CEunder2;
    for (int s = 0; s < n; s++) {
      size_t a = (size_t)B.ptr[i] + s;
      int j = B.lst[a];
      int x = ff.inxn2[t + ff.nso * (itype[j] - 1)] - 1;
      if (x < 0) continue;
      double bpp = B.BO2[a] + B.BO3[a];
      double dj = delta[j] - deltalp[j], oj = 1.0 - dDlp[j];
      double CEover5 = CEover1 * ff.povun1[x] * ff.Desig[x];
      double CElp_b = CElp1 + CEover3 + CEover5 + CEunder3;
      double CElp_bpp = CEover4 * dj + CEunder4 * dj;
      B.cB0[a] += CElp_b;          // coeff = (b, b+bpp, b+bpp) -> cf = (b, bpp, bpp)
      B.cB1[a] += CElp_bpp;
      B.cB2[a] += CElp_bpp;
      B.cdslot[a] += CEover4 * oj * bpp + CEunder4 * oj * bpp;   // cdbnd(j) += CElp_d
    }
  }
  block_add<4>(part, acc + ACC_PE + 1);
}

// ---------------------------------------------------------------------------------------------------
// direction forces of an angle i-j-k (ForceA3, src/pot.F90:1462-1521): returns fij (on i) and fjk (=-force on k)
__device__ __forceinline__ void a3_forces(double coeff, const double *da0, double n0, const double *da1, double n1,
                                          double *fij, double *fjk) {
  double C00 = n0 * n0, C01 = (da0[0] * da1[0] + da0[1] * da1[1]) + da0[2] * da1[2], C11 = n1 * n1;
  double coCC = coeff * (1.0 / (n0 * n1));
  double Ci1 = -(C01 / C00), Ck2 = C01 / C11;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    fij[c] = coCC * (Ci1 * da0[c] + da1[c]);
    fjk[c] = -coCC * (-da0[c] + Ck2 * da1[c]);
  }
}


// ---------------------------------------------------------------------------------------------------
// Angles and torsions are evaluated in two steps: an ENUMERATION kernel applies the reference's cut-off tests
// (cheap, very divergent: ~8 of 66 candidate pairs and ~13 of ~50 candidate quadruples survive per atom) and
// appends the survivors to a work list in HBM; an EVALUATION kernel then runs one thread per surviving
// angle / torsion, so the expensive fp64 part (15 exp, pow, acos ...) executes with full warps.
__device__ __forceinline__ int warp_append(bool valid, int *__restrict__ counter, int lane) {
  unsigned m = __ballot_sync(0xffffffffu, valid);
  int base = 0;
  if (lane == 0 && m) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  return valid ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

// reserve `n` consecutive work-list entries for this thread with ONE atomic per warp; returns the thread's first index
__device__ __forceinline__ int warp_reserve(int n, int *__restrict__ counter) {
  const int lane = threadIdx.x & 31;
  int inc = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  int total = __shfl_sync(0xffffffffu, inc, 31);
  int base = 0;
  if (lane == 0 && total > 0) base = atomicAdd(counter, total);
  base = __shfl_sync(0xffffffffu, base, 0);
  return base + inc - n;
}

// per-centre sums of E3b (src/pot.F90:360-366), one thread per resident
__global__ void k_e3b_sums(int natoms, const int *__restrict__ itype, Bonds B, double2 *__restrict__ sbo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= natoms || itype[j] <= 0) return;
  const int n = B.cnt[j];
  const size_t row = (size_t)B.ptr[j];
  double sum_BO8 = 0.0, sum_SBO1 = 0.0;
  for (int s = 0; s < n; s++) {
    double b0 = B.BO0[row + s];
    double b2 = b0 * b0, b4 = b2 * b2;
    sum_BO8 -= b4 * b4;
    sum_SBO1 += B.BO2[row + s] + B.BO3[row + s];
  }
  sbo[j] = make_double2(exp(sum_BO8), sum_SBO1);
}

// C3a: E3b enumeration (src/pot.F90:356-386 tests).  One thread per directed bond slot (j,i1): partners k1 > i1.
__global__ void __launch_bounds__(256) k_e3b_enum(int nslots, int natoms, const int *__restrict__ itype,
                                                  const DevFF *__restrict__ ffp, Bonds B, int2 *__restrict__ wl, int cap,
                                                  int *__restrict__ counter) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const DevFF &ff = *ffp;
  int j = natoms, i1 = 0, n = 0, tj = 0, ity = 0;
  size_t row = 0;
  double BOij0 = 0.0;
  bool live = false;
  if (a < nslots) {
    j = B.own[a];
    if (j < natoms && itype[j] > 0) {
      row = (size_t)B.ptr[j];
      i1 = a - (int)row;
      n = B.cnt[j];
      BOij0 = B.BO0[a];
      live = (BOij0 - CUTOF2_ESUB > 0.0) && i1 + 1 < n;
      if (live) { tj = itype[j] - 1; ity = itype[B.lst[a]]; }
    }
  }
  auto valid = [&](int k1) {
    double BOjk0 = B.BO0[row + k1];
    if (!((BOjk0 - CUTOF2_ESUB > 0.0) && (BOij0 * BOjk0 > CUTOF2_ESUB))) return false;
    int kty = itype[B.lst[row + k1]];
    return ff.inxn3[(ity - 1) + ff.nso * (tj + ff.nso * (kty - 1))] != 0;
  };
  int cnt = 0;
  if (live)
    for (int k1 = i1 + 1; k1 < n; k1++) cnt += valid(k1);
  int w = warp_reserve(cnt, counter);
  if (cnt > 0)
    for (int k1 = i1 + 1; k1 < n; k1++)
      if (valid(k1)) {
        if (w < cap) wl[w] = make_int2(j, i1 | (k1 << 8));
        w++;
      }
}

// C3b: E3b evaluation (src/pot.F90:388-541), one thread per angle.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_e3b_eval(int nwork, const int2 *__restrict__ wl, const double4 *__restrict__ pq, int NB,
                                                  const int *__restrict__ itype, const DevFF *__restrict__ ffp, Bonds B,
                                                  const double *__restrict__ delta, const double *__restrict__ nlp,
                                                  const double *__restrict__ dDlp, const double2 *__restrict__ sbo,
                                                  double *__restrict__ s3, double *__restrict__ f, double *__restrict__ acc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double part[3] = {0.0, 0.0, 0.0};   // PE(5..7)
  if (t < nwork) {
    const DevFF &ff = *ffp;
    const int2 w = wl[t];
    const int j = w.x, i1 = w.y & 0xff, k1 = (w.y >> 8) & 0xff;
    const size_t row = (size_t)B.ptr[j];
    const int jty = itype[j], tj = jty - 1;
    const int i = B.lst[row + i1], k = B.lst[row + k1];
    const int ity = itype[i], kty = itype[k];
    const int x = ff.inxn3[(ity - 1) + ff.nso * (tj + ff.nso * (kty - 1))] - 1;
    const double BOij = B.BO0[row + i1] - CUTOF2_ESUB, BOjk = B.BO0[row + k1] - CUTOF2_ESUB;
    const double2 sb = sbo[j];
    const double prod_SBO = sb.x, sum_SBO1 = sb.y;
    const double dj = delta[j];
    const double delta_ang = dj + ff.Val[tj] - ff.Valangle[tj];
    const double nlpj = nlp[j], dDlpj = dDlp[j];
    const double4 pj = pq[j], pi = pq[i], pk = pq[k];
    double rij[3] = {pi.x - pj.x, pi.y - pj.y, pi.z - pj.z};
    double rjk[3] = {pj.x - pk.x, pj.y - pk.y, pj.z - pk.z};
    double nij = sqrt((rij[0] * rij[0] + rij[1] * rij[1]) + rij[2] * rij[2]);
    double njk = sqrt((rjk[0] * rjk[0] + rjk[1] * rjk[1]) + rjk[2] * rjk[2]);
    double cos_ijk = -((rij[0] * rjk[0] + rij[1] * rjk[1]) + rij[2] * rjk[2]) / (nij * njk);
    if (cos_ijk > MAXANGLE) cos_ijk = MAXANGLE;
    if (cos_ijk < MINANGLE) cos_ijk = MINANGLE;
    double theta_ijk = acos(cos_ijk);
    double sin_ijk = sin(theta_ijk);
    // --- valence angle
    double pv1 = ff.pval1[x], pv2 = ff.pval2[x], pv3 = ff.pval3[tj], pv4 = ff.pval4[x], pv5 = ff.pval5[tj];
    double exp3ij = exp(-pv3 * pow(BOij, pv4)), exp3jk = exp(-pv3 * pow(BOjk, pv4));
    double fn7ij = 1.0 - exp3ij, fn7jk = 1.0 - exp3jk;
    double pv6 = ff.pval6[x], pv7 = ff.pval7[x];
    double exp6 = exp(pv6 * delta_ang), exp7 = exp(-pv7 * delta_ang);
    double trm8 = 1.0 + exp6 + exp7;
    double fn8j = pv5 - (pv5 - 1.0) * (2.0 + exp6) / trm8;
    double pv8 = ff.pval8[x], pv9 = ff.pval9[x], pv10 = ff.pval10[x];
    double SBO = sum_SBO1 + (1.0 - prod_SBO) * (-delta_ang - pv8 * nlpj);
    double SBO2 = 0.0, CSBO2 = 0.0;
    if (SBO > 0) { SBO2 = pow(SBO, pv9); CSBO2 = pv9 * pow(SBO, pv9 - 1.0); }
    if (SBO > 1) { SBO2 = 2.0 - pow(2.0 - SBO, pv9); CSBO2 = pv9 * pow(2.0 - SBO, pv9 - 1.0); }
    if (SBO > 2) { SBO2 = 2.0; CSBO2 = 0.0; }
    double th00 = ff.theta00[x];
    double e10 = exp(-pv10 * (2.0 - SBO2));
    double theta0 = PI_RX - th00 * (1.0 - e10);
    double theta_diff = theta0 - theta_ijk;
    double exp2 = exp(-pv2 * theta_diff * theta_diff);
    double PEval = fn7ij * fn7jk * fn8j * (pv1 - pv1 * exp2);
    double Cf7ij = pv3 * pv4 * pow(BOij, pv4 - 1.0) * exp3ij;
    double Cf7jk = pv3 * pv4 * pow(BOjk, pv4 - 1.0) * exp3jk;
    double Cf8j = (1.0 - pv5) / (trm8 * trm8) * (pv6 * exp6 * trm8 - (2.0 + exp6) * (pv6 * exp6 - pv7 * exp7));
    double Ctheta0 = pv10 * th00 * e10;
    double dSBO1 = -8.0 * prod_SBO * (delta_ang + pv8 * nlpj);
    double dSBO2 = (prod_SBO - 1.0) * (1.0 - pv8 * dDlpj);
    double CEval1 = Cf7ij * fn7jk * fn8j * pv1 * (1.0 - exp2);
    double CEval2 = fn7ij * Cf7jk * fn8j * pv1 * (1.0 - exp2);
    double CEval3 = fn7ij * fn7jk * Cf8j * pv1 * (1.0 - exp2);
    double CEval4 = 2.0 * pv1 * pv2 * fn7ij * fn7jk * fn8j * exp2 * theta_diff;
    double CEval5 = CEval4 * Ctheta0 * CSBO2;
    double CEval6 = CEval5 * dSBO1;
    double CEval7 = CEval5 * dSBO2;
    double CEval8 = CEval4 / sin_ijk;
    // --- penalty
    double pp2 = ff.ppen2[x], pp3 = ff.ppen3[x], pp4 = ff.ppen4[x];
    double exp_pen3 = exp(-pp3 * dj), exp_pen4 = exp(pp4 * dj);
    double trm_pen34 = 1.0 + exp_pen3 + exp_pen4;
    double fn9 = (2.0 + exp_pen3) / trm_pen34;
    double PEpen = ff.ppen1[x] * fn9 * exp(-pp2 * (BOij - 2.0) * (BOij - 2.0)) * exp(-pp2 * (BOjk - 2.0) * (BOjk - 2.0));
    double Cf9j = (-pp3 * exp_pen3 * trm_pen34 - (2.0 + exp_pen3) * (-pp3 * exp_pen3 + pp4 * exp_pen4)) / (trm_pen34 * trm_pen34);
    double CEpen1 = Cf9j / fn9 * PEpen;
    double CEpen2 = -2.0 * pp2 * (BOij - 2.0) * PEpen;
    double CEpen3 = -2.0 * pp2 * (BOjk - 2.0) * PEpen;
    // --- 3-body conjugation
    double sum_BOi = delta[i] + ff.Val[ity - 1], sum_BOk = delta[k] + ff.Val[kty - 1];
    double delta_val = dj + ff.Val[tj] - ff.Valval[tj];
    double pc2 = ff.pcoa2[x], pc3 = ff.pcoa3[x], pc4 = ff.pcoa4[x];
    double exp_coa2 = exp(pc2 * delta_val);
    double ui = -BOij + sum_BOi, uk = -BOjk + sum_BOk;
    double PEcoa = ff.pcoa1[x] / (1.0 + exp_coa2) * exp(-pc3 * (ui * ui)) * exp(-pc3 * (uk * uk)) *
                   exp(-pc4 * ((BOij - 1.5) * (BOij - 1.5))) * exp(-pc4 * ((BOjk - 1.5) * (BOjk - 1.5)));
    double CEcoa1 = -2.0 * pc4 * (BOij - 1.5) * PEcoa;
    double CEcoa2 = -2.0 * pc4 * (BOjk - 1.5) * PEcoa;
    double CEcoa3 = -pc2 * exp_coa2 / (1.0 + exp_coa2) * PEcoa;
    double CEcoa4 = -2.0 * pc3 * ui * PEcoa;
    double CEcoa5 = -2.0 * pc3 * uk * PEcoa;
    part[0] = PEval; part[1] = PEpen; part[2] = PEcoa;
    // ForceB on BO_ij and BO_jk: both bonds are slots of the centre
    atomicAdd(&B.cB0[row + i1], CEpen2 + CEcoa1 - CEcoa4 + CEval1);
    atomicAdd(&B.cB0[row + k1], CEpen3 + CEcoa2 - CEcoa5 + CEval2);
    // cdbnd(i) += CE3body_d(2) ; cdbnd(k) += CE3body_d(3)  -> addressed to the partners of those slots
    atomicAdd(&B.cdslot[row + i1], CEcoa4);
    atomicAdd(&B.cdslot[row + k1], CEcoa5);
    // the reference's loop over ALL neighbours of j per angle (src/pot.F90:526-532) is linear in three per-centre
    // sums: sum CE3body_d(1), sum CEval(6), sum CEval(5); k_final1 applies them to every bond of j
    atomicAdd(&s3[j], CEpen1 + CEcoa3 + CEval3 + CEval7);
    atomicAdd(&s3[NB + j], CEval6);
    atomicAdd(&s3[2 * (size_t)NB + j], CEval5);
    double fij[3], fjk[3];
    a3_forces(CEval8, rij, nij, rjk, njk, fij, fjk);
    atomic_add3(f, NB, i, fij[0], fij[1], fij[2]);
    atomic_add3(f, NB, j, -fij[0] + fjk[0], -fij[1] + fjk[1], -fij[2] + fjk[2]);
    atomic_add3(f, NB, k, -fjk[0], -fjk[1], -fjk[2]);
  }
  block_add<3>(part, acc + ACC_PE + 5);
}

// C5a: Ehb enumeration: donor-H bonds (i, slot s) with jty == 2 and BO > MINBO0 (src/pot.F90:590-595).
// hydrogen is hard-coded as atom type 2
__global__ void k_ehb_enum(int natoms, const int *__restrict__ itype, const DevFF *__restrict__ ffp, Bonds B,
                           int2 *__restrict__ wl, int cap, int *__restrict__ counter) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  const DevFF &ff = *ffp;
  const int ity = itype[i];
  if (ity <= 0 || ff.nso < 2) return;
  bool donor = false;   // any acceptor type with inxn3hb(ity, 2, kty) != 0 ?
  for (int kt = 0; kt < ff.nso; kt++) donor |= ff.inxn3hb[(ity - 1) + ff.nso * (1 + ff.nso * kt)] != 0;
  if (!donor) return;
  const int n = B.cnt[i];
  const size_t row = (size_t)B.ptr[i];
  for (int s = 0; s < n; s++) {
    if (itype[B.lst[row + s]] == 2 && B.BO0[row + s] > MINBO0) {
      int w = atomicAdd(counter, 1);
      if (w < cap) wl[w] = make_int2(i, s);
    }
  }
}
// C5b: Ehb evaluation, src/pot.F90:597-661.  One warp per donor-H bond; the lanes scan i's 10 A row.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_ehb_eval(int nwork, const int2 *__restrict__ wl, const int *__restrict__ slot_of,
                                                  const double4 *__restrict__ pqs, const int4 *__restrict__ tgs, int NB,
                                                  const DevFF *__restrict__ ffp, Bonds B, const long long *__restrict__ rowbeg,
                                                  const long long *__restrict__ rowend, const int *__restrict__ col,
                                                  double *__restrict__ f, double *__restrict__ fsl, double *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double part[1] = {0.0};
  if (t < nwork) {
    const DevFF &ff = *ffp;
    const int2 w = wl[t];
    const int i = w.x, s = w.y;
    const size_t row = (size_t)B.ptr[i];
    const int j = B.lst[row + s];
    const double bo = B.BO0[row + s];
    const int si = slot_of[i], sj = slot_of[j];
    const int ity = tgs[si].x, jty = 2;
    const double4 pi = pqs[si], pj = pqs[sj];
    double rij[3] = {pi.x - pj.x, pi.y - pj.y, pi.z - pj.z};
    double nij = sqrt((rij[0] * rij[0] + rij[1] * rij[1]) + rij[2] * rij[2]);
    double cb = 0.0, fi[3] = {0, 0, 0}, fj[3] = {0, 0, 0};
    for (long long kk = rowbeg[i] + lane; kk < rowend[i]; kk += 32) {
      int ks = col[kk] & COL_MASK;
      int4 tk = tgs[ks];
      int inxnhb = ff.inxn3hb[(ity - 1) + ff.nso * ((jty - 1) + ff.nso * (tk.x - 1))];
      if (!((j != tk.z) && (i != tk.z) && (inxnhb != 0))) continue;
      const double4 pk = ldg256(pqs + ks);
      double rik2 = dist2_rn(sub_rn(pi.x, pk.x), sub_rn(pi.y, pk.y), sub_rn(pi.z, pk.z));
      if (!(rik2 < RCHB2)) continue;
      int x = inxnhb - 1;
      double rjk[3] = {pj.x - pk.x, pj.y - pk.y, pj.z - pk.z};
      double njk = sqrt((rjk[0] * rjk[0] + rjk[1] * rjk[1]) + rjk[2] * rjk[2]);
      double cos_ijk = -((rij[0] * rjk[0] + rij[1] * rjk[1]) + rij[2] * rjk[2]) / (nij * njk);
      if (cos_ijk > MAXANGLE) cos_ijk = MAXANGLE;
      if (cos_ijk < MINANGLE) cos_ijk = MINANGLE;
      double cos_xhz1 = 1.0 - cos_ijk;
      double s2 = 0.5 * cos_xhz1;          // sin^2(theta/2) = (1 - cos theta)/2 : the reference's acos/sin pair is not needed
      double sin_xhz4 = s2 * s2;
      double r0 = ff.r0hb[x], p1 = ff.phb1[x], p2 = ff.phb2[x], p3 = ff.phb3[x];
      double exp_hb2 = exp(-p2 * bo);
      double exp_hb3 = exp(-p3 * (r0 / njk + njk / r0 - 2.0));
      double PEhb = p1 * (1.0 - exp_hb2) * exp_hb3 * sin_xhz4;
      part[0] += PEhb;
      cb += p1 * p2 * exp_hb2 * exp_hb3 * sin_xhz4;
      double CEhb2 = -0.5 * p1 * (1.0 - exp_hb2) * exp_hb3 * cos_xhz1;
      double CEhb3 = -PEhb * p3 * (-r0 / (njk * njk) + 1.0 / r0) * (1.0 / njk);
      double fij[3], fjk[3];
      a3_forces(CEhb2, rij, nij, rjk, njk, fij, fjk);
      double ff3[3] = {CEhb3 * rjk[0], CEhb3 * rjk[1], CEhb3 * rjk[2]};
#pragma unroll
      for (int c = 0; c < 3; c++) {
        fi[c] += fij[c];
        fj[c] += -fij[c] + fjk[c] - ff3[c];
      }
      atomic_add3(fsl, NB, ks, -fjk[0] + ff3[0], -fjk[1] + ff3[1], -fjk[2] + ff3[2]);   // slot-ordered accumulator, see k_enbond
    }
    cb = warp_sum(cb);
#pragma unroll
    for (int c = 0; c < 3; c++) { fi[c] = warp_sum(fi[c]); fj[c] = warp_sum(fj[c]); }
    if (lane == 0 && cb != 0.0) {
      atomicAdd(&B.cB0[row + s], cb);   // ForceB(i,j1,...,CEhb(1))
      atomic_add3(f, NB, i, fi[0], fi[1], fi[2]);
      atomic_add3(f, NB, j, fj[0], fj[1], fj[2]);
    }
  }
  block_add<1>(part, acc + ACC_PE + 10);
}

// normalised cross product with the reference's floor on the norm (cross_product, src/pot.F90:1524-1543)
__device__ __forceinline__ double cross_n(const double *a, double na, const double *b, double nb, double *c) {
  double a0 = a[0] / na, a1 = a[1] / na, a2 = a[2] / na;
  double b0 = b[0] / nb, b1 = b[1] / nb, b2 = b[2] / nb;
  c[0] = a1 * b2 - a2 * b1;
  c[1] = a2 * b0 - a0 * b2;
  c[2] = a0 * b1 - a1 * b0;
  double n = sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
  return n < NSMALL ? NSMALL : n;
}

// C4a: E4b enumeration (tests of src/pot.F90:1022-1081).  One thread per directed bond slot (j,k1) = central bond j-k.
__global__ void __launch_bounds__(256) k_e4b_enum(int nslots, int natoms, const int *__restrict__ itype, const int *__restrict__ gid,
                                                  const DevFF *__restrict__ ffp, Bonds B, int2 *__restrict__ wl, int cap,
                                                  int *__restrict__ counter) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const DevFF &ff = *ffp;
  int j = natoms, k = 0, k1 = 0, nj = 0, nk = 0, jty = 0, kty = 0;
  size_t rowj = 0, rowk = 0;
  double BOjk0 = 0.0;
  bool live = false;
  if (a < nslots) {
    j = B.own[a];
    if (j < natoms && itype[j] > 0) {
      BOjk0 = B.BO0[a];
      k = B.lst[a];
      live = (BOjk0 > CUTOF2_ESUB) && (gid[j] < gid[k]);
      if (live) {
        rowj = (size_t)B.ptr[j]; rowk = (size_t)B.ptr[k];
        k1 = a - (int)rowj;
        nj = B.cnt[j]; nk = B.cnt[k];
        jty = itype[j]; kty = itype[k];
      }
    }
  }
  auto valid = [&](int i1, int l1) {
    double BOij0 = B.BO0[rowj + i1], BOkl0 = B.BO0[rowk + l1];
    int i = B.lst[rowj + i1], l = B.lst[rowk + l1];
    if (!((BOij0 > CUTOF2_ESUB) && ((BOij0 * BOjk0) > CUTOF2_ESUB) && (i != k) && (BOkl0 > CUTOF2_ESUB) &&
          (BOjk0 * BOkl0 > CUTOF2_ESUB) && (i != l) && (j != l) && ((BOij0 * (BOjk0 * BOjk0) * BOkl0) > MINBO0)))
      return false;
    return ff.inxn4[(itype[i] - 1) + ff.nso * ((jty - 1) + ff.nso * ((kty - 1) + ff.nso * (itype[l] - 1)))] != 0;
  };
  int cnt = 0;
  if (live)
    for (int i1 = 0; i1 < nj; i1++) {
      if (!(B.BO0[rowj + i1] > CUTOF2_ESUB) || i1 == k1) continue;
      for (int l1 = 0; l1 < nk; l1++) cnt += valid(i1, l1);
    }
  int w = warp_reserve(cnt, counter);
  if (cnt > 0)
    for (int i1 = 0; i1 < nj; i1++) {
      if (!(B.BO0[rowj + i1] > CUTOF2_ESUB) || i1 == k1) continue;
      for (int l1 = 0; l1 < nk; l1++)
        if (valid(i1, l1)) {
          if (w < cap) wl[w] = make_int2(j, k1 | (i1 << 8) | (l1 << 16));
          w++;
        }
    }
}

// C4b: E4b evaluation (src/pot.F90:1083-1205), one thread per torsion i-j-k-l.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_e4b_eval(int nwork, const int2 *__restrict__ wl, const double4 *__restrict__ pq, int NB,
                                                  const int *__restrict__ itype, const DevFF *__restrict__ ffp, Bonds B,
                                                  const double *__restrict__ delta, double *__restrict__ cdbnd,
                                                  double *__restrict__ f, double *__restrict__ acc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double part[2] = {0.0, 0.0};   // PE(8..9)
  if (t < nwork) {
    const DevFF &ff = *ffp;
    const int2 w = wl[t];
    const int j = w.x, k1 = w.y & 0xff, i1 = (w.y >> 8) & 0xff, l1 = (w.y >> 16) & 0xff;
    const size_t rowj = (size_t)B.ptr[j];
    const int k = B.lst[rowj + k1];
    const size_t rowk = (size_t)B.ptr[k];
    const int i = B.lst[rowj + i1], l = B.lst[rowk + l1];
    const int ity = itype[i], jty = itype[j], kty = itype[k], lty = itype[l];
    const int x = ff.inxn4[(ity - 1) + ff.nso * ((jty - 1) + ff.nso * ((kty - 1) + ff.nso * (lty - 1)))] - 1;
    const double BOij = B.BO0[rowj + i1] - CUTOF2_ESUB, BOjk = B.BO0[rowj + k1] - CUTOF2_ESUB, BOkl = B.BO0[rowk + l1] - CUTOF2_ESUB;
    const double BOpi_jk = B.BO2[rowj + k1];
    const double delta_ang_jk = (delta[j] + ff.Val[jty - 1] - ff.Valangle[jty - 1]) + (delta[k] + ff.Val[kty - 1] - ff.Valangle[kty - 1]);
    const double4 pi = pq[i], pj = pq[j], pk = pq[k], pl = pq[l];
    double rij[3] = {pi.x - pj.x, pi.y - pj.y, pi.z - pj.z};
    double rjk[3] = {pj.x - pk.x, pj.y - pk.y, pj.z - pk.z};
    double rkl[3] = {pk.x - pl.x, pk.y - pl.y, pk.z - pl.z};
    double nij = sqrt((rij[0] * rij[0] + rij[1] * rij[1]) + rij[2] * rij[2]);
    double njk = sqrt((rjk[0] * rjk[0] + rjk[1] * rjk[1]) + rjk[2] * rjk[2]);
    double nkl = sqrt((rkl[0] * rkl[0] + rkl[1] * rkl[1]) + rkl[2] * rkl[2]);
    double cos_ijk = -((rij[0] * rjk[0] + rij[1] * rjk[1]) + rij[2] * rjk[2]) / (nij * njk);
    if (cos_ijk > MAXANGLE) cos_ijk = MAXANGLE;
    if (cos_ijk < MINANGLE) cos_ijk = MINANGLE;
    double theta_ijk = acos(cos_ijk);
    double sin_ijk = sin(theta_ijk);
    double tan_ijk_i = 1.0 / tan(theta_ijk);
    double crs_ijk[3];
    double ncr1 = cross_n(rij, nij, rjk, njk, crs_ijk);
    double pt1 = ff.ptor1[x], pt2 = ff.ptor2[x], pt3 = ff.ptor3[x], pt4 = ff.ptor4[x];
    double V1 = ff.V1[x], V2 = ff.V2[x], V3 = ff.V3[x], pc1 = ff.pcot1[x], pc2 = ff.pcot2[x];
    double et1 = exp(-pt2 * BOij), et2 = exp(-pt2 * BOjk), et3 = exp(-pt2 * BOkl);
    double exp_tor3 = exp(-pt3 * delta_ang_jk), exp_tor4 = exp(pt4 * delta_ang_jk);
    double exp_tor34_i = 1.0 / (1.0 + exp_tor3 + exp_tor4);
    double fn10 = (1.0 - et1) * (1.0 - et2) * (1.0 - et3);
    double fn11 = (2.0 + exp_tor3) * exp_tor34_i;
    double fn12 = exp(-pc2 * ((BOij - 1.5) * (BOij - 1.5) + (BOjk - 1.5) * (BOjk - 1.5) + (BOkl - 1.5) * (BOkl - 1.5)));
    double btb2 = 2.0 - BOpi_jk - fn11;
    double exp_tor1 = exp(pt1 * (btb2 * btb2));
    double cos_jkl = -((rjk[0] * rkl[0] + rjk[1] * rkl[1]) + rjk[2] * rkl[2]) / (njk * nkl);
    if (cos_jkl > MAXANGLE) cos_jkl = MAXANGLE;
    if (cos_jkl < MINANGLE) cos_jkl = MINANGLE;
    double theta_jkl = acos(cos_jkl);
    double sin_jkl = sin(theta_jkl);
    double tan_jkl_i = 1.0 / tan(theta_jkl);
    double crs_jkl[3];
    double ncr2 = cross_n(rjk, njk, rkl, nkl, crs_jkl);
    double cw = ((crs_ijk[0] * crs_jkl[0] + crs_ijk[1] * crs_jkl[1]) + crs_ijk[2] * crs_jkl[2]) / (ncr1 * ncr2);
    if (cw > MAXANGLE) cw = MAXANGLE;
    if (cw < MINANGLE) cw = MINANGLE;
    double omega = acos(cw);
    double cw_sqr = cw * cw;
    double cos_2w = cos(2.0 * omega);
    double c2 = 1.0 - cos_2w;
    double c3 = 1.0 + cos(3.0 * omega);
    double Vsum = V1 * (1.0 + cw) + V2 * exp_tor1 * c2 + V3 * c3;
    double PEtors = 0.5 * fn10 * sin_ijk * sin_jkl * Vsum;
    double PEconj = pc1 * fn12 * (1.0 + (cw_sqr - 1.0) * sin_ijk * sin_jkl);
    part[0] = PEtors; part[1] = PEconj;
    double CEtors1 = 0.5 * sin_ijk * sin_jkl * Vsum;
    double CEtors2 = -pt1 * fn10 * sin_ijk * sin_jkl * V2 * exp_tor1 * btb2 * c2;
    double dfn11 = (-pt3 * exp_tor3 + (pt3 * exp_tor3 - pt4 * exp_tor4) * (2.0 + exp_tor3) * exp_tor34_i) * exp_tor34_i;
    double CEtors3 = CEtors2 * dfn11;
    double CEtors4 = CEtors1 * pt2 * et1 * (1.0 - et2) * (1.0 - et3);
    double CEtors5 = CEtors1 * pt2 * (1.0 - et1) * et2 * (1.0 - et3);
    double CEtors6 = CEtors1 * pt2 * (1.0 - et1) * (1.0 - et2) * et3;
    double cmn = -0.5 * fn10 * Vsum;
    double CEtors7 = cmn * sin_jkl * tan_ijk_i;
    double CEtors8 = cmn * sin_ijk * tan_jkl_i;
    double CEtors9 = fn10 * sin_ijk * sin_jkl * (0.5 * V1 - 2.0 * V2 * exp_tor1 * cw + 1.5 * V3 * (cos_2w + 2.0 * cw_sqr));
    double Cconj = -2.0 * pc2 * PEconj;
    double CEconj4 = -pc1 * fn12 * (cw_sqr - 1.0) * tan_ijk_i * sin_jkl;
    double CEconj5 = -pc1 * fn12 * (cw_sqr - 1.0) * sin_ijk * tan_jkl_i;
    double CEconj6 = 2.0 * pc1 * fn12 * cw * sin_ijk * sin_jkl;
    double Ca_ijk = CEconj4 + CEtors7, Ca_jkl = CEconj5 + CEtors8, Ca_ijkl = CEconj6 + CEtors9;
    atomicAdd(&cdbnd[j], CEtors3);
    atomicAdd(&cdbnd[k], CEtors3);
    atomicAdd(&B.cB0[rowj + i1], Cconj * (BOij - 1.5) + CEtors4);   // ForceB on BO_ij
    atomicAdd(&B.cB0[rowj + k1], Cconj * (BOjk - 1.5) + CEtors5);   // ForceBbo on BO_jk: coeff (b, b+t2, b) -> cf (b, t2, 0)
    atomicAdd(&B.cB1[rowj + k1], CEtors2);
    atomicAdd(&B.cB0[rowk + l1], Cconj * (BOkl - 1.5) + CEtors6);   // ForceB on BO_kl
    // --- direction forces: ForceA3(i,j,k), ForceA3(j,k,l), ForceA4(i,j,k,l)  (src/pot.F90:1369-1521)
    double f1[3], f2[3], g1[3], g2[3];
    a3_forces(Ca_ijk, rij, nij, rjk, njk, f1, f2);
    a3_forces(Ca_jkl, rjk, njk, rkl, nkl, g1, g2);
    double C00 = nij * nij, C01 = (rij[0] * rjk[0] + rij[1] * rjk[1]) + rij[2] * rjk[2],
           C02 = (rij[0] * rkl[0] + rij[1] * rkl[1]) + rij[2] * rkl[2];
    double C11 = njk * njk, C12 = (rjk[0] * rkl[0] + rjk[1] * rkl[1]) + rjk[2] * rkl[2], C22 = nkl * nkl;
    double D0 = C00 * C11 - C01 * C01, D1 = C11 * C22 - C12 * C12;
    double coDD = Ca_ijkl * (1.0 / sqrt(D0 * D1));
    double com = C01 * C12 - C02 * C11;
    double Cwi0 = C11 / D0 * com, Cwi1 = -(C12 + C01 / D0 * com), Cwi2 = C11;
    double Cwj0 = -(C12 + (C11 + C01) / D0 * com);
    double Cwj1 = -(-C12 - 2 * C02 - C22 / D1 * com - (C00 + C01) / D0 * com);
    double Cwj2 = -(C01 + C11 + C12 / D1 * com);
    double Cwl0 = -C11, Cwl1 = (C01 + C12 / D1 * com), Cwl2 = -(C11 / D1 * com);
    double Fi[3], Fj[3], Fk[3], Fl[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double hij = coDD * (Cwi0 * rij[c] + Cwi1 * rjk[c] + Cwi2 * rkl[c]);
      double hjk = coDD * ((Cwj0 + Cwi0) * rij[c] + (Cwj1 + Cwi1) * rjk[c] + (Cwj2 + Cwi2) * rkl[c]);
      double hkl = -coDD * (Cwl0 * rij[c] + Cwl1 * rjk[c] + Cwl2 * rkl[c]);
      Fi[c] = f1[c] + hij;
      Fj[c] = (-f1[c] + f2[c]) + g1[c] + (-hij + hjk);
      Fk[c] = -f2[c] + (-g1[c] + g2[c]) + (-hjk + hkl);
      Fl[c] = -g2[c] - hkl;
    }
    atomic_add3(f, NB, i, Fi[0], Fi[1], Fi[2]);
    atomic_add3(f, NB, j, Fj[0], Fj[1], Fj[2]);
    atomic_add3(f, NB, k, Fk[0], Fk[1], Fk[2]);
    atomic_add3(f, NB, l, Fl[0], Fl[1], Fl[2]);
  }
  block_add<2>(part, acc + ACC_PE + 8);
}

// ---------------------------------------------------------------------------------------------------
// C7: ForceBondedTerms (src/pot.F90:113-144) + ForceD/ForceB/ForceBbo (src/pot.F90:1230-1365) as gathers.
// pass 0: cdbnd(i) total = own accumulator + what neighbours addressed to i through their slots
__global__ void k_final0(int ntot, Bonds B, double *__restrict__ cdbnd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot) return;
  double s = cdbnd[i];
  const int n = B.cnt[i];
  for (int k = 0; k < n; k++) {
    size_t a = (size_t)B.ptr[i] + k;
    int j = B.lst[a];
    s += B.cdslot[(size_t)B.ptr[j] + B.idx[a]];
  }
  cdbnd[i] = s;
}
// pass 1: bond forces on atom i from every one of its bonds, and ccbnd(i) with the reference's order predicate
__global__ void __launch_bounds__(128) k_final1(int ntot, const double *__restrict__ pos, int NB, Bonds B,
                                                const double *__restrict__ cdbnd, const double *__restrict__ s3,
                                                double *__restrict__ ccbnd, double *__restrict__ f) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot) return;
  const int n = B.cnt[i];
  if (n == 0) { ccbnd[i] = 0.0; return; }
  const double xi = pos[i], yi = pos[NB + i], zi = pos[2 * NB + i];
  const double cdi = cdbnd[i];
  const double sdi = s3[i], s6i = s3[NB + i], s5i = s3[2 * (size_t)NB + i];   // E3b per-centre sums of atom i
  double fx = 0, fy = 0, fz = 0, cc = 0.0;
  for (int k = 0; k < n; k++) {
    size_t a = (size_t)B.ptr[i] + k;
    int j = B.lst[a];
    size_t b = (size_t)B.ptr[j] + B.idx[a];
    double cdj = cdbnd[j];
    double b0 = B.BO0[a];
    double b02 = b0 * b0, b03 = b02 * b0, b07 = b03 * b03 * b0;
    // E3b's ForceBbo over every neighbour of a centre (src/pot.F90:526-532): coeff = d1 + CEval6*BO^7 + (0,CEval5,CEval5)
    double e3c1 = (sdi + s3[j]) + (s6i + s3[NB + j]) * b07, e3c23 = s5i + s3[2 * (size_t)NB + j];
    double c1e = B.cB0[a] + B.cB0[b] + e3c1;           // energy-term coefficients of both orientations
    double c2 = B.cB1[a] + B.cB1[b] + e3c23, c3 = B.cB2[a] + B.cB2[b] + e3c23;
    double c1 = c1e + cdi + cdj;                       // + ForceD(i) and ForceD(j) (src/pot.F90:1243-1245)
    double b2 = B.BO2[a], b3 = B.BO3[a], A1 = B.A1[a], db = B.dBOp[a];
    double Cb = c1 * (B.A0[a] + b0 * A1) * db + c2 * b2 * (B.dln2[a] + A1 * db) + c3 * b3 * (B.dln3[a] + A1 * db);
    double dx = xi - pos[j], dy = yi - pos[NB + j], dz = zi - pos[2 * NB + j];
    fx -= Cb * dx; fy -= Cb * dy; fz -= Cb * dz;
    double A2 = B.A2[a];
    cc += c1e * b0 * A2 + (c2 * b2 + c3 * b3) * B.A3[a];
    cc += cdi * b0 * A2;                               // ForceD(i): Cbond(2), always consumed
    if (j < i) cc += cdj * b0 * A2;                    // ForceD(j): Cbond(3), consumed only if it ran before i (Q1)
  }
  ccbnd[i] = cc;
  atomic_add3(f, NB, i, fx, fy, fz);
}
// pass 2: f(i) -= sum_j (ccbnd(i)+ccbnd(j)) dBOp (ri-rj)     (src/pot.F90:129-135 seen from both ends of a bond)
__global__ void __launch_bounds__(128) k_final2(int ntot, const double *__restrict__ pos, int NB, Bonds B,
                                                const double *__restrict__ ccbnd, double *__restrict__ f) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot) return;
  const int n = B.cnt[i];
  if (n == 0) return;
  const double xi = pos[i], yi = pos[NB + i], zi = pos[2 * NB + i];
  const double cci = ccbnd[i];
  double fx = 0, fy = 0, fz = 0;
  for (int k = 0; k < n; k++) {
    size_t a = (size_t)B.ptr[i] + k;
    int j = B.lst[a];
    double w = (cci + ccbnd[j]) * B.dBOp[a];
    fx -= w * (xi - pos[j]); fy -= w * (yi - pos[NB + j]); fz -= w * (zi - pos[2 * NB + j]);
  }
  f[i] += fx; f[NB + i] += fy; f[2 * NB + i] += fz;
}

// virial accumulation over residents and ghosts before the copy-back, src/pot.F90:65-72
__global__ void __launch_bounds__(256) k_virial(int ntot, const double *__restrict__ pos, const double *__restrict__ f, int NB,
                                                double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double part[6] = {0, 0, 0, 0, 0, 0};
  if (i < ntot) {
    double x = pos[i], y = pos[NB + i], z = pos[2 * NB + i], fx = f[i], fy = f[NB + i], fz = f[2 * NB + i];
    part[0] = x * fx; part[1] = y * fy; part[2] = z * fz; part[3] = y * fz; part[4] = z * fx; part[5] = x * fy;
  }
  block_add<6>(part, acc + ACC_ASTR);
}

// ---------------------------------------------------------------------------------------------------
// integrator halves of the main loop, src/main.F90:64-72 and :86-98 (vkick :192-207)
__global__ void k_md_first_half(int n, int NB, double dt, double Lex_w2, const int *__restrict__ itype,
                                const DevFF *__restrict__ ffp, double *__restrict__ pos, double *__restrict__ v,
                                const double *__restrict__ f, const double *__restrict__ q, double *__restrict__ qsfp,
                                double *__restrict__ qsfv, bool drift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double dthm = dt * 0.5 / ffp->mass[itype[i] - 1];
  double sv = qsfv[i] + 0.5 * dt * Lex_w2 * (q[i] - qsfp[i]);
  qsfv[i] = sv;
  qsfp[i] = qsfp[i] + dt * sv;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    double vv = add_rn(v[(size_t)c * NB + i], mul_rn(mul_rn(1.0, dthm), f[(size_t)c * NB + i]));
    v[(size_t)c * NB + i] = vv;
    if (drift) pos[(size_t)c * NB + i] = add_rn(pos[(size_t)c * NB + i], mul_rn(dt, vv));
  }
}
// LinearMomentum, src/main.F90:773-803 (called every step when an electric field is applied, :70-71):
// acc[0] = sum m, acc[1..3] = sum m v
__global__ void __launch_bounds__(256) k_momentum(int n, int NB, const int *__restrict__ itype, const DevFF *__restrict__ ffp,
                                                  const double *__restrict__ v, double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double part[4] = {0, 0, 0, 0};
  if (i < n) {
    double m = ffp->mass[itype[i] - 1];
    part[0] = m; part[1] = m * v[i]; part[2] = m * v[NB + i]; part[3] = m * v[2 * (size_t)NB + i];
  }
  block_add<4>(part, acc);
}
__global__ void k_sub_vcm_drift(int n, int NB, double dt, double *__restrict__ v, double *__restrict__ pos,
                                const double *__restrict__ acc, bool drift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mm = acc[0];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    double vv = sub_rn(v[(size_t)c * NB + i], acc[1 + c] / mm);
    v[(size_t)c * NB + i] = vv;
    if (drift) pos[(size_t)c * NB + i] = add_rn(pos[(size_t)c * NB + i], mul_rn(dt, vv));
  }
}
// thermostat hooks for device-resident stepping (the thermostats themselves stay host logic, src/main.F90:49-62,684-770):
// per atom type t: out[6t..6t+5] += {count, sum 1/2 m v^2, sum m, sum m vx, sum m vy, sum m vz}
constexpr int VSTAT_MAXT = 20;   // MAX_ELEMENT of ScaleTemperature, src/main.F90:727
__global__ void __launch_bounds__(256) k_velocity_stats(int n, int NB, const int *__restrict__ itype, const DevFF *__restrict__ ffp,
                                                        const double *__restrict__ v, double *__restrict__ out) {
  __shared__ double sh[6 * VSTAT_MAXT];
  for (int k = threadIdx.x; k < 6 * VSTAT_MAXT; k += blockDim.x) sh[k] = 0.0;
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int t = itype[i] - 1;
    if (t >= 0 && t < VSTAT_MAXT) {
      double m = ffp->mass[t], vx = v[i], vy = v[NB + i], vz = v[2 * (size_t)NB + i];
      atomicAdd(&sh[6 * t], 1.0);
      atomicAdd(&sh[6 * t + 1], 0.5 * m * ((vx * vx + vy * vy) + vz * vz));
      atomicAdd(&sh[6 * t + 2], m);
      atomicAdd(&sh[6 * t + 3], m * vx); atomicAdd(&sh[6 * t + 4], m * vy); atomicAdd(&sh[6 * t + 5], m * vz);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 6 * VSTAT_MAXT; k += blockDim.x)
    if (sh[k] != 0.0) atomicAdd(&out[k], sh[k]);
}
// v(i) = scale[type(i)] * v(i) - shift
__global__ void k_velocity_affine(int n, int NB, const int *__restrict__ itype, const double *__restrict__ scale,
                                  double sx, double sy, double sz, double *__restrict__ v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = scale[itype[i] - 1];
  v[i] = sub_rn(mul_rn(a, v[i]), sx);
  v[NB + i] = sub_rn(mul_rn(a, v[NB + i]), sy);
  v[2 * (size_t)NB + i] = sub_rn(mul_rn(a, v[2 * (size_t)NB + i]), sz);
}
__global__ void __launch_bounds__(256) k_md_second_half(int n, int NB, double dt, double Lex_w2, const int *__restrict__ itype,
                                                        const DevFF *__restrict__ ffp, double *__restrict__ v,
                                                        const double *__restrict__ f, const double *__restrict__ q,
                                                        const double *__restrict__ qsfp, double *__restrict__ qsfv,
                                                        double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double part[6] = {0, 0, 0, 0, 0, 0};
  if (i < n) {
    double m = ffp->mass[itype[i] - 1];
    double dthm = dt * 0.5 / m;
    double vx = v[i], vy = v[NB + i], vz = v[2 * NB + i];
    part[0] = vx * vx * m; part[1] = vy * vy * m; part[2] = vz * vz * m;
    part[3] = vy * vz * m; part[4] = vz * vx * m; part[5] = vx * vy * m;
    v[i] = add_rn(vx, mul_rn(dthm, f[i]));
    v[NB + i] = add_rn(vy, mul_rn(dthm, f[NB + i]));
    v[2 * NB + i] = add_rn(vz, mul_rn(dthm, f[2 * NB + i]));
    qsfv[i] = qsfv[i] + 0.5 * dt * Lex_w2 * (q[i] - qsfp[i]);
  }
  block_add<6>(part, acc);
}
// KE = sum hmas*v^2 and sum q (PRINTE, src/main.F90:225-230)
__global__ void __launch_bounds__(256) k_observe(int n, int NB, const int *__restrict__ itype, const DevFF *__restrict__ ffp,
                                                 const double *__restrict__ v, const double *__restrict__ q,
                                                 double *__restrict__ acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double part[2] = {0, 0};
  if (i < n) {
    double hm = 0.5 * ffp->mass[itype[i] - 1];
    double vx = v[i], vy = v[NB + i], vz = v[2 * NB + i];
    part[0] = hm * ((vx * vx + vy * vy) + vz * vz);
    part[1] = q[i];
  }
  block_add<2>(part, acc);
}

// f(atom) += the slot-ordered partner forces of k_enbond / k_enbond_pqeq / k_ehb_eval (residents and ghosts)
__global__ void k_fsl_to_f(int ntot, int NB, const int *__restrict__ order, const double *__restrict__ fsl, double *__restrict__ f) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ntot) return;
  const int i = order[s];
  f[i] += fsl[s]; f[(size_t)NB + i] += fsl[(size_t)NB + s]; f[2 * (size_t)NB + i] += fsl[2 * (size_t)NB + s];
}

}   // namespace rxg
#include "rxg_pqeq.cuh"
namespace rxg {

// ---------------------------------------------------------------------------------------------------
inline Bonds make_bonds(Ctx *c) {
  Bonds B;
  B.MAXN = c->MAXN; B.ptr = c->bptr; B.own = c->bown; B.cnt = c->nbrcnt; B.lst = c->nbrlist; B.idx = c->nbrindx;
  B.BO0 = c->BO[0]; B.BO1 = c->BO[1]; B.BO2 = c->BO[2]; B.BO3 = c->BO[3];
  B.dln1 = c->dln[0]; B.dln2 = c->dln[1]; B.dln3 = c->dln[2]; B.dBOp = c->dBOp;
  B.A0 = c->A0; B.A1 = c->A1; B.A2 = c->A2; B.A3 = c->A3;
  B.cB0 = c->cB[0]; B.cB1 = c->cB[1]; B.cB2 = c->cB[2]; B.cdslot = c->cdslot;
  return B;
}

// subroutine FORCE on device-resident state, reference src/pot.F90:2-90
// `reuse`: the residents, ghosts, non-bonded cells and the 10 A list of the QEq that just ran are still valid (same step,
// FORCE-width halo, see qeq_device); only the ghost charges are refreshed.
// `stage` (device-resident stepping and RXG_FUSE_API, reuse only): the charge-independent part of FORCE runs beside the QEq CG
// of the same step (qeq_cg_single): 1 = bonded cells and bonded list (host synchronisations; before the CG starts),
// 2 = bond orders, bonded energy terms and ForceBondedTerms (on the side stream, while the CG's HBM-bound sparse products leave
// the fp64 pipes idle), 3 = ghost charges, ENbond, sums, MODE_CPBK (after the CG).  0 = everything in the reference's order.
inline int force_device(Ctx *c, bool reuse = false, int stage = 0) {
  const int NB = c->NB, n = c->natoms;
  const bool s1 = stage == 0 || stage == 1, s2 = stage == 0 || stage == 2, s3 = stage == 0 || stage == 3;
  if (s1) {
    RXG_CUDA(cudaMemsetAsync(c->f, 0, sizeof(double) * 3 * NB, c->st));
    RXG_CUDA(cudaMemsetAsync(c->fsl, 0, sizeof(double) * 3 * NB, c->st));
    RXG_CUDA(cudaMemsetAsync(c->d_acc + ACC_PE, 0, sizeof(double) * 24, c->st));
  }
  double dr[3];
  for (int a = 0; a < 3; a++) dr[a] = c->cfg.nmincell * c->box.lcsize[a];
  const bool pqeq = c->cfg.isPQEq != 0;
  if (stage == 0 || stage == 3) {
    phase_mark(c, 4);                                             // COPYATOMS
    if (!reuse) RXG_TRY(halo_copy(c, dr));                        // src/pot.F90:28
    else RXG_TRY(halo_refresh(c, 4, 1));   // ghost q; + the position round trip FORCE's own MODE_COPY would apply
    if (pqeq) RXG_TRY(halo_refresh(c, 5, 0));   // ghost spos (part of MODE_COPY in the reference, src/comm.F90:129-131)
  }
  const int nt = c->cp[6];
  if (s1) {
    if (!reuse) LAUNCH(c, k_types, cdiv(nt, 256), 256, 0, c->atype, nt, c->itype, c->gid);
    phase_mark(c, 3);                                             // LINKEDLIST
    RXG_TRY(bin_grid(c, c->gb));                                  // :30
    if (!reuse) RXG_TRY(bin_grid(c, c->gnb));                     // :31
    phase_mark(c, 5);                                             // NEIGHBORLIST
    RXG_TRY(build_nbrlist(c));                                    // :33
    phase_mark(c, 15);                                            // GetNonbondingPairList
    if (!reuse) RXG_TRY(build_pairlist<0>(c));                    // :34
    if (stage == 1) { phase_mark(c, 0); return RXG_OK; }
  }
  phase_mark(c, 6);                                             // BOCALC
  Bonds B = make_bonds(c);
  if (s2) {
    LAUNCH(c, k_boprim, cdiv(nt, 128), 128, 0, nt, c->pos, NB, c->itype, c->d_ff, B, c->deltap1, c->deltap2, c->cdbnd, c->ccbnd, c->s3);
    LAUNCH(c, k_bofull, cdiv(nt, 128), 128, 0, nt, c->itype, c->d_ff, B, c->deltap1, c->deltap2, c->delta);
  }
  // ---- energy terms (src/pot.F90:49-57)
  if (!c->wl) {   // first call: size the work lists from the resident count
    c->wl_cap3 = 16LL * NB + 1024; c->wl_cap4 = 32LL * NB + 1024; c->wl_caph = 2LL * NB + 1024;
    RXG_CUDA(cudaMalloc((void **)&c->wl, sizeof(int2) * (size_t)(c->wl_cap3 + c->wl_cap4 + c->wl_caph)));
  }
  double4 *pq = c->pqa;
  phase_mark(c, 7);                                             // ENbond
  // (stage 2 packs the positions the bonded terms read, with the charges of the previous step; stage 3 packs again once q is final)
  LAUNCH(c, k_pack_pq, cdiv(nt, 256), 256, 0, nt, c->pos, NB, c->q, c->itype, c->gid, c->gnb.slot_of, c->pqa, c->pqs, c->tgs, c->gts);
  // the full-row form needs every partner's image inside this rank's halo: true when the FORCE halo >= rctap
  bool full_ok = true;
  const double lat[3] = {c->box.lata, c->box.latb, c->box.latc};
  for (int a = 0; a < 3; a++)
    if (dr[a] * lat[a] < c->ff.rctap) full_ok = false;
  // measured on B200 (979 776-atom RDX): the literal half-list form with fp64 atomics (5.1 ms) beats the full-row form
  // (7.1 ms, bound by the L1 data pipe on table gathers), so the literal form is the default
  if (!(getenv("RXG_ENBOND_FULL") && getenv("RXG_ENBOND_FULL")[0] == '1')) full_ok = false;
  const int wgrid = cdiv((long long)n * 32, 256);
  const int ogrid = cdiv((long long)nt * 32, 256);   // warps over cell-ordered slots (ghost slots exit at once)
  if (!s3) {
  } else if (pqeq) {   // src/pot.F90:48-49
    full_ok = false;
    LAUNCH(c, k_pack_sps, cdiv(nt, 256), 256, 0, nt, c->spos, NB, c->itype, c->gnb.slot_of, c->d_ff, c->sps);
    LAUNCH(c, k_enbond_pqeq, ogrid, 256, 0, nt, n, c->rowbeg, c->rowend, c->col, c->pqs, c->tgs, c->sps, c->d_ff, c->f, c->fsl, NB, c->d_acc);
    if (c->cfg.isEfield && n > 0)   // :61
      LAUNCH(c, k_efield, cdiv(n, 256), 256, 0, n, c->q, c->itype, c->d_ff, c->cfg.eFieldDir, c->cfg.eFieldStrength, c->f, NB);
  } else if (!full_ok) {
    if (c->enbond_queue && NB < (1 << 26) && c->ff.nso < 32)
      LAUNCH(c, k_enbond_half, ogrid, 256, 0, nt, n, c->rowbeg, c->rowend, c->col, c->pqs, c->tgs, c->gts, c->d_ff, c->f, c->fsl, NB, c->d_acc);
    else LAUNCH(c, (k_enbond<true>), ogrid, 256, 0, nt, n, c->rowbeg, c->rowend, c->col, c->pqs, c->tgs, c->d_ff, c->f, c->fsl, NB, c->d_acc);
  }
  if (s2) {
  phase_mark(c, 9);                                             // Elnpr (preparation loop)
  LAUNCH(c, k_elnpr_prep, cdiv(nt, 256), 256, 0, nt, c->itype, c->d_ff, c->delta, c->nlp, c->dDlp, c->deltalp);
  phase_mark(c, 8);                                             // Ebond (one kernel with Elnpr's main loop)
  LAUNCH(c, k_ebond_elnpr, cdiv(n, 128), 128, 0, n, c->itype, c->gid, c->d_ff, B, c->delta, c->dDlp, c->deltalp, c->d_acc);
  // angles and torsions: enumerate survivors of the cut-off tests, then evaluate one per thread
  for (int attempt = 0; attempt < 2; attempt++) {
    RXG_CUDA(cudaMemsetAsync(c->d_flag + 12, 0, 3 * sizeof(int), c->st));
    int2 *wl3 = c->wl, *wl4 = c->wl + c->wl_cap3, *wlh = c->wl + c->wl_cap3 + c->wl_cap4;
    phase_mark(c, 10);                                          // Ehb
    LAUNCH(c, k_ehb_enum, cdiv(n, 128), 128, 0, n, c->itype, c->d_ff, B, wlh, (int)c->wl_caph, c->d_flag + 14);
    const int nslots = (int)c->nbonds;
    phase_mark(c, 11);                                          // E3b
    if (attempt == 0) LAUNCH(c, k_e3b_sums, cdiv(n, 128), 128, 0, n, c->itype, B, c->sbo);
    LAUNCH(c, k_e3b_enum, cdiv(nslots, 256), 256, 0, nslots, n, c->itype, c->d_ff, B, wl3, (int)c->wl_cap3, c->d_flag + 12);
    phase_mark(c, 12);                                          // E4b
    LAUNCH(c, k_e4b_enum, cdiv(nslots, 256), 256, 0, nslots, n, c->itype, c->gid, c->d_ff, B, wl4, (int)c->wl_cap4, c->d_flag + 13);
    RXG_CUDA(cudaMemcpyAsync(c->h_int + 12, c->d_flag + 12, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    RXG_CUDA(cudaStreamSynchronize(c->st));
    const long long n3 = c->h_int[12], n4 = c->h_int[13], nh = c->h_int[14];
    if (n3 <= c->wl_cap3 && n4 <= c->wl_cap4 && nh <= c->wl_caph) {
      c->n_angles = n3; c->n_torsions = n4; c->n_hbonds = nh;
      phase_mark(c, 10);
#define RXG_EVAL(kern, lo, hi, grid, block, ...)                                   \
  do {                                                                              \
    if (c->eval_occ) LAUNCH(c, (kern<hi>), grid, block, 0, __VA_ARGS__);           \
    else LAUNCH(c, (kern<lo>), grid, block, 0, __VA_ARGS__);                        \
  } while (0)
      if (nh > 0)
        RXG_EVAL(k_ehb_eval, 1, 3, cdiv(nh * 32, 256), 256, (int)nh, wlh, c->gnb.slot_of, c->pqs, c->tgs, NB, c->d_ff, B, c->rowbeg,
                 c->rowend, c->col, c->f, c->fsl, c->d_acc);
      phase_mark(c, 11);
      if (n3 > 0)
        RXG_EVAL(k_e3b_eval, 1, 5, cdiv(n3, 128), 128, (int)n3, wl3, pq, NB, c->itype, c->d_ff, B, c->delta, c->nlp, c->dDlp, c->sbo,
                 c->s3, c->f, c->d_acc);
      phase_mark(c, 12);
      if (n4 > 0)
        RXG_EVAL(k_e4b_eval, 1, 5, cdiv(n4, 128), 128, (int)n4, wl4, pq, NB, c->itype, c->d_ff, B, c->delta, c->cdbnd, c->f, c->d_acc);
#undef RXG_EVAL
      break;
    }
    if (attempt == 1) { c->err = "angle/torsion work list overflow"; return RXG_ERR_STATE; }
    if (c->wl) cudaFree(c->wl);   // grow and enumerate again (rare: first call of a denser system)
    c->wl_cap3 = std::max(c->wl_cap3, n3 + n3 / 4 + 1024);
    c->wl_cap4 = std::max(c->wl_cap4, n4 + n4 / 4 + 1024);
    c->wl_caph = std::max(c->wl_caph, nh + nh / 4 + 1024);
    RXG_CUDA(cudaMalloc((void **)&c->wl, sizeof(int2) * (size_t)(c->wl_cap3 + c->wl_cap4 + c->wl_caph)));
  }
  }   // s2
  phase_mark(c, 13);                                            // ForceBondedTerms
  if (stage == 0) LAUNCH(c, k_fsl_to_f, cdiv(nt, 256), 256, 0, nt, NB, c->gnb.order, c->fsl, c->f);
  if (s2) {
    // ---- ForceBondedTerms (src/pot.F90:63)
    LAUNCH(c, k_final0, cdiv(nt, 128), 128, 0, nt, B, c->cdbnd);
    LAUNCH(c, k_final1, cdiv(nt, 128), 128, 0, nt, c->pos, NB, B, c->cdbnd, c->s3, c->ccbnd, c->f);
    LAUNCH(c, k_final2, cdiv(nt, 128), 128, 0, nt, c->pos, NB, B, c->ccbnd, c->f);
    if (stage == 2) { phase_mark(c, 0); return RXG_OK; }
  }
  if (stage == 3) LAUNCH(c, k_fsl_to_f, cdiv(nt, 256), 256, 0, nt, NB, c->gnb.order, c->fsl, c->f);   // partner forces of Ehb (stage 2) and ENbond
  LAUNCH(c, k_virial, cdiv(nt, 256), 256, 0, nt, c->pos, c->f, NB, c->d_acc);   // :65-72
  // full-row ENbond puts both halves of a pair force on residents, so its virial is taken per pair inside the kernel
  if (full_ok) LAUNCH(c, (k_enbond<false>), ogrid, 256, 0, nt, n, c->rowbeg, c->rowend, c->col, c->pqs, c->tgs, c->d_ff, c->f, c->fsl, NB, c->d_acc);
  phase_mark(c, 4);                                             // COPYATOMS(MODE_CPBK)
  RXG_TRY(halo_cpbk(c));                                                          // :74
  phase_mark(c, 0);
  RXG_CUDA(cudaMemcpyAsync(c->h_acc + ACC_PE, c->d_acc + ACC_PE, sizeof(double) * 24, cudaMemcpyDeviceToHost, c->st));
  if (c->peer_ok) RXG_CUDA(cudaMemcpyAsync(c->h_int + 3, c->d_flag + 3, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  RXG_CUDA(cudaStreamSynchronize(c->st));
  if (c->peer_ok && c->h_int[3]) { c->err = "peer halo: a neighbour's ghost values did not arrive (timeout)"; return RXG_ERR_NCCL; }
  c->PE[0] = 0.0;
  for (int k = 1; k < 14; k++) c->PE[k] = c->h_acc[ACC_PE + k];
  for (int k = 0; k < 6; k++) c->astr[k] = c->h_acc[ACC_ASTR + k];
  return RXG_OK;
}

}   // namespace rxg
