// rxg_pqeq.cuh -- polarizable charge equilibration (PQEq) kernels: SURVEY 8a row a18.
// Reference: src/pqeq.F90 (PQEq, qeq_initialize, get_hsh, get_gradient, update_shell_positions),
//            src/pot.F90:784-923 (ENbond_PQEq), src/module.F90:359-417 (EEfield, get_coulomb_and_dcoulomb_pqeq).
//
// How the reference's PQEq differs from its QEq, and what that means here:
//   * the 10 A list becomes a 12.5 A list (rctap0_pqeq); the hessian is Cclmb0_qeq * lerp(TBL_Eclmb_pcc) evaluated with the
//     fp64 r^2 (QEq uses the real(4) r^2), and a constant vector fpqeq enters the s-gradient (src/pqeq.F90:336-343,463);
//   * get_hsh re-evaluates two more table lerps per list entry in EVERY CG iteration only to form the scalar Est
//     (:405-431).  None of those terms changes during the CG except through q, and they enter Est linearly:
//         Est = sum_i [chi q_i + eta/2 q_i^2] + 1/2 sum_i (q_i+Z_i) sum_j H_ij (q_j+Z_j) + sum_ij S_ij (q_j+Z_j) + E_ss
//     with S_ij = -Cclmb0_qeq Z_i T_sc(|shell_i - core_j|) and E_ss = 1/2 sum_ij Cclmb0_qeq Z_i Z_j T_ss(|shell_i - shell_j|).
//     k_pqeq_rows evaluates the lerps ONCE per call and keeps the column sums c_j = sum_i S_ij, so an iteration needs only
//     sum_j c_j (q_j+Z_j): O(N) work beside the single sparse product of the CG (rxg_lists_qeq.cuh).
//   * after the CG the shells relax one capped step (update_shell_positions).
// Early returns of get_coulomb_and_dcoulomb_pqeq (shell distance beyond rctap while the cores are inside): the reference
// then reads a variable it did not assign for that pair (src/pqeq.F90:219-231, 340-343).  Like the oracle, these kernels
// take the intended zero contribution and count the events (Ctx::pqeq_skips); see DESIGN.md.
#pragma once

namespace rxg {

constexpr double CCLMB0 = 332.0638, CCLMB0_QEQ = 14.4;   // src/module.F90:681-682
constexpr double EEV_KCAL = 23.060538;                   // src/module.F90:191
constexpr double MAX_SHELL_DISPLACEMENT = 1e-3;          // src/pqeq.F90:190

// get_coulomb_and_dcoulomb_pqeq (src/module.F90:386-417): E and dE such that ff = dE * rr.  false = early return.
__device__ __forceinline__ bool clmb_pqeq(const DevFF &ff, const double4 *__restrict__ T, int inxn, double rx, double ry, double rz,
                                          double &E, double &dE) {
  const double dr2 = dist2_rn(rx, ry, rz);
  E = 0.0; dE = 0.0;
  if (dr2 > ff.rctap2) return false;
  const int itb = (int)mul_rn(dr2, ff.UDRi);
  const double drtb = mul_rn(sub_rn(dr2, mul_rn((double)itb, ff.UDR)), ff.UDRi);
  const double drtb1 = sub_rn(1.0, drtb);
  if (inxn >= 1 && itb >= 1 && itb + 1 <= ff.ntable) {   // outside: the reference reads out of bounds (like SURVEY Q9)
    const double4 t = ldg256(T + (size_t)(inxn - 1) * ff.ntable + (itb - 1));
    E = add_rn(mul_rn(drtb1, t.x), mul_rn(drtb, t.y));
    dE = add_rn(mul_rn(drtb1, t.z), mul_rn(drtb, t.w));
  }
  return true;
}

// The same lerp with a one-record cache.  When the three PQEq tables are identical (DevFF::pq_same: every element has
// Rc = Rs, true for both parameter files the reference ships) the core-core, core-shell, shell-core and shell-shell terms of
// one pair read the SAME table, and because a shell sits ~1e-3 A off its core their r^2 almost always fall into the same
// table interval: the record fetched for the first term serves the others (one 32-byte gather per pair instead of three or
// four).  The arithmetic is unchanged, so the results are bit-identical with or without the cache.
struct PqCache { int itb; double4 t; };
__device__ __forceinline__ bool clmb_pqeq_c(const DevFF &ff, const double4 *__restrict__ T, int inxn, double rx, double ry, double rz,
                                            double &E, double &dE, PqCache &pc) {
  const double dr2 = dist2_rn(rx, ry, rz);
  E = 0.0; dE = 0.0;
  if (dr2 > ff.rctap2) return false;
  const int itb = (int)mul_rn(dr2, ff.UDRi);
  const double drtb = mul_rn(sub_rn(dr2, mul_rn((double)itb, ff.UDR)), ff.UDRi);
  const double drtb1 = sub_rn(1.0, drtb);
  if (inxn >= 1 && itb >= 1 && itb + 1 <= ff.ntable) {
    double4 t;
    if (ff.pq_same && itb == pc.itb) t = pc.t;
    else {
      t = ldg256(T + (size_t)(inxn - 1) * ff.ntable + (itb - 1));
      pc.itb = itb; pc.t = t;
    }
    E = add_rn(mul_rn(drtb1, t.x), mul_rn(drtb, t.y));
    dE = add_rn(mul_rn(drtb1, t.z), mul_rn(drtb, t.w));
  }
  return true;
}

// by-slot packs: sps = {spos, Zpqeq(type)}; qsl = q
__global__ void k_pack_sps(int ntot, const double *__restrict__ spos, int NB, const int *__restrict__ itype,
                           const int *__restrict__ slot_of, const DevFF *__restrict__ ffp, double4 *__restrict__ sps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntot) return;
  int t = itype[i];
  double z = (t >= 1 && t <= ffp->ntype_pqeq) ? ffp->Zpqeq[t - 1] : 0.0;
  sps[slot_of[i]] = make_double4(spos[i], spos[(size_t)NB + i], spos[2 * (size_t)NB + i], z);
}
__global__ void k_pack_qsl(int ntot, const double *__restrict__ q, const int *__restrict__ slot_of, double *__restrict__ qsl) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ntot) qsl[slot_of[i]] = q[i];
}

// qeq_initialize of src/pqeq.F90:262-365 over the compacted rows (one warp per resident row): hessian, fpqeq and the
// CG-invariant pieces of Est (header).  acc[17] += E_ss; *skips += early returns.
__global__ void __launch_bounds__(256) k_pqeq_rows(const DevGrid g, int ntot, int natoms, const long long *__restrict__ rowbeg,
                                                   const long long *__restrict__ rowend, const int *__restrict__ col,
                                                   double *__restrict__ val, const double4 *__restrict__ sps,
                                                   const DevFF *__restrict__ ffp, double4 *__restrict__ prow,
                                                   double *__restrict__ pcs, double *__restrict__ acc, int *__restrict__ skips) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double ess[1] = {0.0};
  int nskip = 0;
  double4 me = make_double4(0, 0, 0, 0);
  int i = natoms;
  if (slot < ntot) { me = g.sorted[slot]; i = rec_index(me.w); }
  if (i < natoms) {
    const DevFF &ff = *ffp;
    const int ity = rec_type(me.w), np = ff.ntype_pqeq;
    const double4 si = sps[slot];
    const bool poli = ff.isPolarizable[ity - 1] != 0;
    const double shx = add_rn(me.x, si.x), shy = add_rn(me.y, si.y), shz = add_rn(me.z, si.z);
    const float rctap2f = (float)ff.rctap2;
    double fp = 0.0, az = 0.0, cres = 0.0;
    const long long s = rowbeg[i], e = rowend[i];
    for (long long k = s + lane; k < e; k += 32) {
      const int cj = col[k];
      const int js = cj & COL_MASK;
      const double4 oj = ldg256(g.sorted + js);
      const double dx = sub_rn(me.x, oj.x), dy = sub_rn(me.y, oj.y), dz = sub_rn(me.z, oj.z);
      // a list shared with FORCE holds the fp64 '<=' pairs; qeq_initialize keeps real(4) dr2 < rctap2 (src/pqeq.F90:316)
      if (!((float)dist2_rn(dx, dy, dz) < rctap2f)) { val[k] = 0.0; continue; }
      const int jty = rec_type(oj.w);
      const double4 sj = ldg256(sps + js);
      const int ix = ff.inxnpqeq[(ity - 1) + np * (jty - 1)];
      double E, dE;
      PqCache pc; pc.itb = -1;
      clmb_pqeq_c(ff, ff.TBL_pcc, ix, dx, dy, dz, E, dE, pc);
      const double h = mul_rn(CCLMB0_QEQ, E);
      val[k] = h;
      const double hz = mul_rn(h, sj.w);
      fp += hz; az += hz;
      const bool polj = ff.isPolarizable[jty - 1] != 0;
      if (polj) {   // core_i - shell_j, :340-343; the same number is S_ji (shell_j - core_i) when row j exists
        if (!clmb_pqeq_c(ff, ff.TBL_psc, ix, sub_rn(dx, sj.x), sub_rn(dy, sj.y), sub_rn(dz, sj.z), E, dE, pc)) nskip++;
        const double t = mul_rn(mul_rn(CCLMB0_QEQ, E), sj.w);
        fp -= t;
        if (cj >= 0) cres -= t;
      }
      if (poli) {
        if (cj < 0) {   // ghost column: no row of its own on this rank, its column sum is taken here
          clmb_pqeq_c(ff, ff.TBL_psc, ix, sub_rn(shx, oj.x), sub_rn(shy, oj.y), sub_rn(shz, oj.z), E, dE, pc);
          atomicAdd(&pcs[js], -CCLMB0_QEQ * E * si.w);
        }
        if (polj) {
          clmb_pqeq_c(ff, ff.TBL_pss, ix, sub_rn(shx, add_rn(oj.x, sj.x)), sub_rn(shy, add_rn(oj.y, sj.y)),
                      sub_rn(shz, add_rn(oj.z, sj.z)), E, dE, pc);
          ess[0] += 0.5 * CCLMB0_QEQ * E * si.w * sj.w;
        }
      }
    }
    fp = warp_sum(fp); az = warp_sum(az); cres = warp_sum(cres);
    if (lane == 0) prow[slot] = make_double4(fp, az, cres, si.w);
  }
  block_accumulate<1>(ess, acc + 17);
  nskip = __reduce_add_sync(0xffffffffu, nskip);
  if (lane == 0 && nskip) atomicAdd(skips, nskip);
}

// k_cg_dots (rxg_lists_qeq.cuh) with the PQEq gradient and Est.  Ghost slots contribute sum_j c_j (...) terms:
//   INIT : acc[14] = sum_ghost c_j qs_j, acc[15] = sum_ghost c_j qt_j, acc[16] = sum_ghost c_j Z_j   (x = {qs,qt} by slot)
//   else : acc[12], acc[13] = sum_ghost c_j hs_j, c_j ht_j (x = {hs,ht}); k_cg_ctrl advances acc[14], acc[15] with lmin
template <bool INIT>
__global__ void __launch_bounds__(256) k_cg_dots_pqeq(const int *__restrict__ order, int ntot, int natoms,
                                                      const double4 *__restrict__ rowsum, const double2 *__restrict__ x,
                                                      const double *__restrict__ q, double2 *__restrict__ gst,
                                                      double2 *__restrict__ tst, double2 *__restrict__ ust,
                                                      double2 *__restrict__ wst, const int *__restrict__ itype,
                                                      const DevFF *__restrict__ ffp, const double4 *__restrict__ prow,
                                                      const double *__restrict__ pcs, const double4 *__restrict__ sps,
                                                      double *__restrict__ acc) {
  if (!INIT && acc[ACC_DONE] != 0.0) return;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  double gpart[3] = {0.0, 0.0, 0.0};
  int i = -1;
  if (slot < ntot) i = order[slot];
  if (i >= 0 && i < natoms) {
    const double4 r = rowsum[slot];
    const double2 me = x[slot];
    const int t = itype[i] - 1;
    const double eta = ffp->eta[t], chi = ffp->chi[t];
    const double4 pr = prow[slot];
    if (INIT) {
      double g1 = sub_rn(sub_rn(sub_rn(-chi, mul_rn(eta, me.x)), r.x), pr.x);   // - fpqeq(i), src/pqeq.F90:463
      double g2 = sub_rn(sub_rn(-1.0, mul_rn(eta, me.y)), r.y);
      gst[i] = make_double2(g1, g2);
      wst[i] = make_double2(r.x, r.y);   // get_hsh of PQEq weights resident and ghost columns alike (:430-433)
      part[0] = g1 * g1; part[1] = g2 * g2;
    } else {
      double ts = eta * me.x + r.x, tt = eta * me.y + r.y;
      tst[i] = make_double2(ts, tt);
      ust[i] = make_double2(r.x, r.y);
      const double2 gg = gst[i], w = wst[i];
      const double mu = acc[11], qi = q[i];
      const double qic = qi + pr.w;
      part[0] = chi * qi + 0.5 * eta * qi * qi + 0.5 * qic * ((w.x - mu * w.y) + pr.y) + (pr.z + pcs[slot]) * qic;
      part[1] = ts * me.x; part[2] = tt * me.y; part[3] = gg.x * me.x; part[4] = gg.y * me.y;
    }
  } else if (i >= natoms) {
    const double cj = pcs[slot];
    const double2 v = x[slot];
    if (INIT) { gpart[0] = cj * v.x; gpart[1] = cj * v.y; gpart[2] = cj * sps[slot].w; }
    else { gpart[0] = cj * v.x; gpart[1] = cj * v.y; }
  }
  if (INIT) {
    double p2[2] = {part[0], part[1]};
    block_accumulate<2>(p2, acc + 7);
    block_accumulate<3>(gpart, acc + 14);
  } else {
    if (slot == 0) part[0] += (acc[14] - acc[11] * acc[15]) + acc[16] + acc[17];   // ghost columns + E_ss, once per rank
    block_accumulate<5>(part, acc + 0);
    double g2[2] = {gpart[0], gpart[1]};
    block_accumulate<2>(g2, acc + 12);
  }
}

// ---------------------------------------------------------------------------------------------------
// STRICT-ORDER validation path of PQEq (RXG_STRICT_ORDER=1, small systems): the literal two-product CG of src/pqeq.F90:96-166
// with every per-row sum in list order and without FMA, like k_rows_strict_* of rxg_lists_qeq.cuh does for QEq.
// fpqeq(i) in qeq_initialize's own accumulation order (src/pqeq.F90:336-343): one thread per resident row.
__global__ void k_fpqeq_strict(const DevGrid g, int natoms, const long long *__restrict__ rowbeg, const long long *__restrict__ rowend,
                               const int *__restrict__ col, const double *__restrict__ val, const double4 *__restrict__ sps,
                               const DevFF *__restrict__ ffp, double *__restrict__ fpq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  const DevFF &ff = *ffp;
  const int slot = g.slot_of[i], np = ff.ntype_pqeq;
  const double4 me = g.sorted[slot];
  const int ity = rec_type(me.w);
  double fp = 0.0;
  for (long long k = rowbeg[i]; k < rowend[i]; k++) {
    const int js = col[k] & COL_MASK;
    const double4 oj = g.sorted[js];
    const int jty = rec_type(oj.w);
    const double4 sj = sps[js];
    fp = add_rn(fp, mul_rn(val[k], sj.w));                       // fpqeq = fpqeq + Cclmb0_qeq*pqeqc*Z_j, :336
    if (ff.isPolarizable[jty - 1]) {                             // core_i - shell_j, :340-343
      double E, dE;
      const double dx = sub_rn(me.x, oj.x), dy = sub_rn(me.y, oj.y), dz = sub_rn(me.z, oj.z);
      clmb_pqeq(ff, ff.TBL_psc, ff.inxnpqeq[(jty - 1) + np * (ity - 1)], sub_rn(dx, sj.x), sub_rn(dy, sj.y), sub_rn(dz, sj.z), E, dE);
      fp = sub_rn(fp, mul_rn(mul_rn(CCLMB0_QEQ, E), sj.w));
    }
  }
  fpq[i] = fp;
}
// get_hsh of src/pqeq.F90:368-439 per row: rowbuf[i] = {eta*hs + sum H hs, same for t, the row's share of Est, -}
__global__ void k_rows_strict_hsh_pqeq(const DevGrid g, int natoms, const long long *__restrict__ rowbeg,
                                       const long long *__restrict__ rowend, const int *__restrict__ col, const double *__restrict__ val,
                                       const double4 *__restrict__ hsq, const double4 *__restrict__ sps, const DevFF *__restrict__ ffp,
                                       double4 *__restrict__ rowbuf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  const DevFF &ff = *ffp;
  const int slot = g.slot_of[i], np = ff.ntype_pqeq;
  const double4 pi = g.sorted[slot], si = sps[slot], me = hsq[i];
  const int ity = rec_type(pi.w);
  const double eta = ff.eta[ity - 1], chi = ff.chi[ity - 1];
  const bool poli = ff.isPolarizable[ity - 1] != 0;
  const double qic = add_rn(me.z, si.w);
  const double shx = add_rn(pi.x, si.x), shy = add_rn(pi.y, si.y), shz = add_rn(pi.z, si.z);
  double ts = mul_rn(eta, me.x), tt = mul_rn(eta, me.y);
  double es = add_rn(mul_rn(chi, me.z), mul_rn(mul_rn(mul_rn(0.5, eta), me.z), me.z));
  for (long long k = rowbeg[i]; k < rowend[i]; k++) {
    const int js = col[k] & COL_MASK;
    const double4 oj = g.sorted[js], sj = sps[js];
    const int j = rec_index(oj.w), jty = rec_type(oj.w);
    const double4 x = hsq[j];
    const double qjc = add_rn(x.z, sj.w);
    const double h = val[k];
    const double Ccicj = mul_rn(mul_rn(h, qic), qjc);
    double Csicj = 0.0, Csisj = 0.0, E, dE;
    if (poli) {
      const int ix = ff.inxnpqeq[(ity - 1) + np * (jty - 1)];
      clmb_pqeq(ff, ff.TBL_psc, ix, sub_rn(shx, oj.x), sub_rn(shy, oj.y), sub_rn(shz, oj.z), E, dE);
      Csicj = mul_rn(mul_rn(mul_rn(-CCLMB0_QEQ, E), qjc), si.w);
      if (ff.isPolarizable[jty - 1]) {
        clmb_pqeq(ff, ff.TBL_pss, ix, sub_rn(shx, add_rn(oj.x, sj.x)), sub_rn(shy, add_rn(oj.y, sj.y)), sub_rn(shz, add_rn(oj.z, sj.z)), E, dE);
        Csisj = mul_rn(mul_rn(mul_rn(CCLMB0_QEQ, E), si.w), sj.w);
      }
    }
    ts = add_rn(ts, mul_rn(h, x.x));
    tt = add_rn(tt, mul_rn(h, x.y));
    es = add_rn(add_rn(es, mul_rn(0.5, add_rn(Ccicj, Csisj))), Csicj);
  }
  rowbuf[i] = make_double4(ts, tt, es, 0.0);
}

// update_shell_positions, src/pqeq.F90:187-259.  One warp per resident row; reads the by-slot copy of spos (sps), writes
// spos, so the relaxation uses the old shells of every neighbour like the reference's two loops do.
__global__ void __launch_bounds__(256) k_shell_relax(const DevGrid g, int ntot, int natoms, const long long *__restrict__ rowbeg,
                                                     const long long *__restrict__ rowend, const int *__restrict__ col,
                                                     const double4 *__restrict__ sps, const double *__restrict__ qsl,
                                                     const DevFF *__restrict__ ffp, int isEfield, int eFieldDir, double eFieldStrength,
                                                     double *__restrict__ spos, int NB, int *__restrict__ skips) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slot >= ntot) return;
  const double4 me = g.sorted[slot];
  const int i = rec_index(me.w);
  if (i >= natoms) return;
  const DevFF &ff = *ffp;
  const int ity = rec_type(me.w), np = ff.ntype_pqeq;
  if (!ff.isPolarizable[ity - 1]) return;
  const double4 si = sps[slot];
  const double Zi = si.w, Ks = ff.Kspqeq[ity - 1];
  const double shx = add_rn(me.x, si.x), shy = add_rn(me.y, si.y), shz = add_rn(me.z, si.z);
  const float rctap2f = (float)ff.rctap2;
  double fx = 0.0, fy = 0.0, fz = 0.0;
  int nskip = 0;
  const long long s = rowbeg[i], e = rowend[i];
  for (long long k = s + lane; k < e; k += 32) {
    const int js = col[k] & COL_MASK;
    const double4 oj = ldg256(g.sorted + js);
    if (!((float)dist2_rn(sub_rn(me.x, oj.x), sub_rn(me.y, oj.y), sub_rn(me.z, oj.z)) < rctap2f)) continue;   // QEq-list predicate
    const int jty = rec_type(oj.w);
    const double4 sj = ldg256(sps + js);
    const double qjc = qsl[js] + sj.w;
    const int ix = ff.inxnpqeq[(ity - 1) + np * (jty - 1)];
    double E, dE;
    PqCache pc; pc.itb = -1;
    const double dx = sub_rn(shx, oj.x), dy = sub_rn(shy, oj.y), dz = sub_rn(shz, oj.z);
    if (!clmb_pqeq_c(ff, ff.TBL_psc, ix, dx, dy, dz, E, dE, pc)) nskip++;
    // ff = -Cclmb0*sf*qjc*Z_i ; sforce -= ff     (Eq. 38)
    double cf = -CCLMB0 * dE * qjc * Zi;
    fx -= cf * dx; fy -= cf * dy; fz -= cf * dz;
    if (ff.isPolarizable[jty - 1]) {
      const double ex = sub_rn(shx, add_rn(oj.x, sj.x)), ey = sub_rn(shy, add_rn(oj.y, sj.y)), ez = sub_rn(shz, add_rn(oj.z, sj.z));
      if (!clmb_pqeq_c(ff, ff.TBL_pss, ix, ex, ey, ez, E, dE, pc)) nskip++;
      cf = CCLMB0 * dE * Zi * sj.w;
      fx -= cf * ex; fy -= cf * ey; fz -= cf * ez;
    }
  }
  fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
  nskip = __reduce_add_sync(0xffffffffu, nskip);
  if (lane == 0) {
    if (nskip) atomicAdd(skips, nskip);
    double ef[3] = {0.0, 0.0, 0.0};
    if (isEfield) ef[eFieldDir - 1] = -(Zi * eFieldStrength * EEV_KCAL);
    fx += ef[0] - Ks * si.x; fy += ef[1] - Ks * si.y; fz += ef[2] - Ks * si.z;   // Eq. 37
    double dx = fx / Ks, dy = fy / Ks, dz = fz / Ks;                                // Eq. 39
    const double ddr = sqrt(dist2_rn(dx, dy, dz));
    if (ddr > MAX_SHELL_DISPLACEMENT) {
      dx = dx / ddr * MAX_SHELL_DISPLACEMENT; dy = dy / ddr * MAX_SHELL_DISPLACEMENT; dz = dz / ddr * MAX_SHELL_DISPLACEMENT;
    }
    spos[i] = si.x + dx; spos[(size_t)NB + i] = si.y + dy; spos[2 * (size_t)NB + i] = si.z + dz;
  }
}

// ENbond_PQEq, src/pot.F90:784-923: the literal half-list form (pairs with gid(i) < gid(j) -- the opposite sense of ENbond),
// one warp per resident row, forces on the partner by fp64 atomics.
__global__ void __launch_bounds__(256) k_enbond_pqeq(int ntot, int natoms, const long long *__restrict__ rowbeg,
                                                     const long long *__restrict__ rowend, const int *__restrict__ col,
                                                     const double4 *__restrict__ pqs, const int4 *__restrict__ tgs,
                                                     const double4 *__restrict__ sps, const DevFF *__restrict__ ffp,
                                                     double *__restrict__ f, double *__restrict__ fsl, int NB, double *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double part[3] = {0.0, 0.0, 0.0};
  int4 ti = make_int4(0, 0, natoms, 0);
  if (slot < ntot) ti = tgs[slot];
  const int i = ti.z;
  if (i < natoms) {
    const DevFF &ff = *ffp;
    const int np = ff.ntype_pqeq;
    const double4 pi = pqs[slot], si = sps[slot];
    const bool poli = ff.isPolarizable[ti.x - 1] != 0;
    const double qic = pi.w + si.w;
    double fx = 0, fy = 0, fz = 0;
    const long long s = rowbeg[i], e = rowend[i];
    for (long long k = s + lane; k < e; k += 32) {
      const int js = __ldcs(col + k) & COL_MASK;
      const int4 tj = tgs[js];
      if (!(ti.y < tj.y)) continue;
      const double4 pj = ldg256(pqs + js), sj = ldg256(sps + js);
      const double dx = sub_rn(pi.x, pj.x), dy = sub_rn(pi.y, pj.y), dz = sub_rn(pi.z, pj.z);
      const double dr2 = dist2_rn(dx, dy, dz);
      const int inxn = ff.inxn2[(ti.x - 1) + ff.nso * (tj.x - 1)];
      const int itb = (int)mul_rn(dr2, ff.UDRi);
      double PEvdw = 0.0, CEvdw = 0.0;
      if (inxn > 0 && itb >= 1 && itb + 1 <= ff.ntable) {
        const double drtb = mul_rn(sub_rn(dr2, mul_rn((double)itb, ff.UDR)), ff.UDRi);
        const double drtb1 = 1.0 - drtb;
        const double4 *T = ff.TBL_nb + (size_t)(inxn - 1) * ff.ntable + (itb - 1);
        const double4 T0 = ldg256(T), T1 = ldg256(T + 1);
        PEvdw = drtb1 * T0.x + drtb * T1.x;
        CEvdw = drtb1 * T0.y + drtb * T1.y;
      }
      const bool polj = ff.isPolarizable[tj.x - 1] != 0;
      const double qjc = pj.w + sj.w;
      const int ix = ff.inxnpqeq[(ti.x - 1) + np * (tj.x - 1)];
      double E, dE;
      PqCache pc; pc.itb = -1;
      clmb_pqeq_c(ff, ff.TBL_pcc, ix, dx, dy, dz, E, dE, pc);
      double c0 = CCLMB0 * qic * qjc * dE;
      double Eclmb = CCLMB0 * E * qic * qjc;
      double gx = CEvdw * dx + c0 * dx, gy = CEvdw * dy + c0 * dy, gz = CEvdw * dz + c0 * dz;
      if (polj) {   // core_i - shell_j
        const double ex = sub_rn(dx, sj.x), ey = sub_rn(dy, sj.y), ez = sub_rn(dz, sj.z);
        clmb_pqeq_c(ff, ff.TBL_psc, ix, ex, ey, ez, E, dE, pc);
        c0 = -CCLMB0 * sj.w * qic * dE;
        gx += c0 * ex; gy += c0 * ey; gz += c0 * ez;
        Eclmb += -CCLMB0 * E * qic * sj.w;
      }
      if (poli) {   // shell_i - core_j
        const double ex = add_rn(dx, si.x), ey = add_rn(dy, si.y), ez = add_rn(dz, si.z);
        clmb_pqeq_c(ff, ff.TBL_psc, ix, ex, ey, ez, E, dE, pc);
        c0 = -CCLMB0 * si.w * qjc * dE;
        gx += c0 * ex; gy += c0 * ey; gz += c0 * ez;
        Eclmb += -CCLMB0 * E * si.w * qjc;
        if (polj) {   // shell_i - shell_j
          const double hx = sub_rn(ex, sj.x), hy = sub_rn(ey, sj.y), hz = sub_rn(ez, sj.z);
          clmb_pqeq_c(ff, ff.TBL_pss, ix, hx, hy, hz, E, dE, pc);
          c0 = CCLMB0 * si.w * sj.w * dE;
          gx += c0 * hx; gy += c0 * hy; gz += c0 * hz;
          Eclmb += CCLMB0 * E * si.w * sj.w;
        }
      }
      part[0] += PEvdw; part[1] += Eclmb;
      fx -= gx; fy -= gy; fz -= gz;
      atomic_add3(fsl, NB, js, gx, gy, gz);   // slot-ordered accumulator, see k_enbond
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) {
      atomic_add3(f, NB, i, fx, fy, fz);
      double Eshell = 0.0;
      if (poli) Eshell = 0.5 * ff.Kspqeq[ti.x - 1] * dist2_rn(si.x, si.y, si.z);
      part[2] = CECHRGE * (ff.chi[ti.x - 1] * pi.w + 0.5 * ff.eta[ti.x - 1] * (pi.w * pi.w)) + Eshell;   // :825
    }
  }
  block_add<3>(part, acc + ACC_PE + 11);
}

// EEfield, src/module.F90:359-383 (called from FORCE, src/pot.F90:61): force only, "energy to be determined" there
__global__ void k_efield(int natoms, const double *__restrict__ q, const int *__restrict__ itype, const DevFF *__restrict__ ffp,
                         int dir, double strength, double *__restrict__ f, int NB) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natoms) return;
  const double qic = q[i] + ffp->Zpqeq[itype[i] - 1];
  f[(size_t)(dir - 1) * NB + i] += -qic * strength * EEV_KCAL;
}

}   // namespace rxg
