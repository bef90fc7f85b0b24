// rxmd_oracle.cpp -- CPU oracle for the RXMD ReaxFF+QEq hot path.  TEST INFRASTRUCTURE ONLY.
//
// Loop-for-loop restatement (C++, 0-based, flat arrays) of the reference Fortran, every routine
// citing the reference file:line it follows.  Compile with -ffp-contract=off so that no FMA is
// formed (an -O3 x86-64 gfortran build of the reference has none either).
//
// PARITY PIN: the reference cannot be compiled here (no Fortran toolchain) and ships no tests; the
// only known-answer data is the README sample run (README.md:157: step-0 per-atom energies of the
// 168-atom RDX cell, 4 significant digits).  tests/test_oracle_pin.py checks the oracle against it.
// Beyond those digits: "parity unpinned" (see DESIGN.md).
//
// All simulated ranks live in one process; COPYATOMS messages are copies between ranks.
#include "rxmd_oracle.h"

#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

const double MAXANGLE = 0.999999999999, MINANGLE = -0.999999999999, NSMALL = 1e-10;   // src/module.F90:85-87
const double PI_RX = 3.14159265358979;                                                  // src/module.F90:90
const double MINBO0 = 1e-4, CUTOF2_ESUB = 1e-4;                                         // src/module.F90:61-62
const double CECHRGE = 23.02;                                                           // src/module.F90:683
const double CCLMB0 = 332.0638, CCLMB0_QEQ = 14.4;                                         // src/module.F90:681-682
const double EEV_KCAL = 23.060538;                                                      // src/module.F90:191
const double MAX_SHELL_DISPLACEMENT = 1e-3;                                             // src/pqeq.F90:190
const double RCHB2 = 100.0;                                                             // src/module.F90:677-678
const int MAXLAYERS = 5, MAXLAYERS_NB = 10;                                             // src/module.F90:44-45
enum { MODE_COPY = 1, MODE_MOVE = 2, MODE_CPBK = 3, MODE_QCOPY1 = 4, MODE_QCOPY2 = 5 };   // src/module.F90:38-39

inline int nint(double x) { return (int)std::lround(x); }
inline int l2g(double atype) {                      // src/main.F90:582-593
  int ity = nint(atype);
  return nint((atype - ity) * 1e13);
}
inline double sum3(double a, double b, double c) { return (a + b) + c; }   // gfortran sum(x(1:3))

struct Params {
  rxg_config cfg;
  int nso, nboty, nvaty, ntoty, nhbty, ntable;
  double vpar1, vpar2, cutoff_vpar30, rctap, rctap2, UDR, UDRi;
  std::vector<double> Val, Valval, Valangle, Vale, mass, plp1, plp2, nlpopt, povun2, povun3, povun4, povun5, povun6,
      povun7, povun8, pval3, pval5, chi, eta;
  std::vector<double> cBOp1, cBOp3, cBOp5, pbo2h, pbo4h, pbo6h, pbo2, pbo4, pbo6, swtch, rc2, pboc1, pboc3, pboc4,
      pboc5, ovc, v13cor, Desig, Depi, Depipi, pbe1, pbe2, povun1;
  std::vector<double> theta00, pval1, pval2, pval4, pval6, pval7, pval8, pval9, pval10, ppen1, ppen2, ppen3, ppen4,
      pcoa1, pcoa2, pcoa3, pcoa4;
  std::vector<double> ptor1, ptor2, ptor3, ptor4, V1, V2, V3, pcot1, pcot2;
  std::vector<double> phb1, phb2, phb3, r0hb;
  std::vector<int> inxn2v, inxn3v, inxn3hbv, inxn4v;
  std::vector<double> TBL_Evdw, TBL_Eclmb, TBL_Eclmb_QEq;
  // module pqeq_vars, src/module.F90:285-304 (after initialize_pqeq :488-613)
  int ntype_pqeq = 0;
  std::vector<int> isPolarizable, inxnpqeqv;
  std::vector<double> Zpqeq, Kspqeq, TBL_pcc, TBL_psc, TBL_pss;
  bool polarizable(int ity) const { return isPolarizable[ity - 1] != 0; }
  int inxnpqeq(int a, int b) const { return inxnpqeqv[(a - 1) + ntype_pqeq * (b - 1)]; }
  // TBL_Eclmb_p??(ntype_pqeq2, NTABLE, 0:1), column-major
  double tblp(const std::vector<double> &T, int inxn, int itb, int c) const {
    return T[(size_t)(inxn - 1) + (size_t)ntype_pqeq * ntype_pqeq * ((size_t)(itb - 1) + (size_t)ntable * c)];
  }
  // 1-based type ids, column-major Fortran layout (see include/rxmd_b200.h)
  int inxn2(int a, int b) const { return (a < 1 || b < 1) ? 0 : inxn2v[(a - 1) + nso * (b - 1)]; }
  int inxn3(int a, int b, int c) const { return inxn3v[(a - 1) + nso * ((b - 1) + nso * (c - 1))]; }
  int inxn3hb(int a, int b, int c) const { return inxn3hbv[(a - 1) + nso * ((b - 1) + nso * (c - 1))]; }
  int inxn4(int a, int b, int c, int d) const {
    return inxn4v[(a - 1) + nso * ((b - 1) + nso * ((c - 1) + nso * (d - 1)))];
  }
  double evdw(int c, int itb, int inxn) const { return TBL_Evdw[c + 2 * ((size_t)(itb - 1) + (size_t)ntable * (inxn - 1))]; }
  double eclmb(int c, int itb, int inxn) const { return TBL_Eclmb[c + 2 * ((size_t)(itb - 1) + (size_t)ntable * (inxn - 1))]; }
  double eqeq(int itb, int inxn) const { return TBL_Eclmb_QEq[(size_t)(itb - 1) + (size_t)ntable * (inxn - 1)]; }
};

void cpy(std::vector<double> &d, const double *s, size_t n) { d.assign(s, s + n); }
void cpyi(std::vector<int> &d, const int *s, size_t n) { d.assign(s, s + n); }

struct Grid {   // header/llist/nacell with NLAYERS ghost layers, src/main.F90:277-318
  int nc[3], L, dim[3];
  std::vector<int> header, nacell, llist;
  void setup(const int *cc, int layers, int nb) {
    L = layers;
    for (int a = 0; a < 3; a++) { nc[a] = cc[a]; dim[a] = cc[a] + 2 * L; }
    header.assign((size_t)dim[0] * dim[1] * dim[2], -1);
    nacell.assign(header.size(), 0);
    llist.assign(nb, -1);
  }
  bool inside(int c1, int c2, int c3) const {
    return c1 >= -L && c1 < nc[0] + L && c2 >= -L && c2 < nc[1] + L && c3 >= -L && c3 < nc[2] + L;
  }
  size_t idx(int c1, int c2, int c3) const { return ((size_t)(c1 + L) * dim[1] + (c2 + L)) * dim[2] + (c3 + L); }
};

struct Rank {
  rxg_box box;
  std::vector<int> nbmesh;
  int NB = 0, MAXN = 0, W10 = 0;
  int natoms = 0, copyptr[7] = {0, 0, 0, 0, 0, 0, 0};
  std::vector<double> atype, q, pos, v, f, qs, qt, gs, gt, hs, ht, qsfp, qsfv, frcindx;
  std::vector<double> spos, fpqeq;   // PQEq: shell displacements spos(NBUFFER,3) src/module.F90:286; fpqeq src/pqeq.F90:20
  long long pqeq_skips = 0;          // calls where the reference would read an undefined value (see get_clmb_pqeq)
  Grid g, nbg;
  std::vector<int> nbrcnt, nbrlist, nbrindx, nbpcnt, nbplist;
  std::vector<double> hessian;
  std::vector<double> BO[4], dln_BOp[3], dBOp, A0, A1, A2, A3, delta, deltap1, deltap2, nlp, dDlp, deltalp, ccbnd, cdbnd;
  std::vector<int> itype, gtype;
  double PE[14], astr[6];
  int nstep_qeq = 0;
  // COPYATOMS shared state (src/module.F90:26-34)
  std::vector<double> sbuffer, rbuffer;
  int ns = 0, nr = 0, na = 0, ne = 0;
  std::vector<char> commflag;
  double *X(int i) { return &pos[i]; }
};

struct World {
  Params P;
  std::vector<Rank> R;
  std::string err;
  double t_qeq = 0, t_force = 0, t_move = 0;
  bool corrected = false;
  int term_mask = 0x3f;
  // diagnostic: what a SERIAL build of the reference does where get_coulomb_and_dcoulomb_pqeq returns early -- the output
  // variable keeps the value of the last call that assigned it (see get_clmb_pqeq).  Parity uses false (zero contribution).
  bool pqeq_stale = false;
};

#define RX(r, i) r.pos[(i)]
#define RY(r, i) r.pos[(size_t)r.NB + (i)]
#define RZ(r, i) r.pos[2 * (size_t)r.NB + (i)]
#define FX(r, i) r.f[(i)]
#define FY(r, i) r.f[(size_t)r.NB + (i)]
#define FZ(r, i) r.f[2 * (size_t)r.NB + (i)]

// ---------------------------------------------------------------------------------------------
// coordinate transforms, src/main.F90:613-681
void xu2xs_inplace(Rank &r, int nmax, std::vector<double> &a) {
  const double *Hi = r.box.HHi;   // column-major: HHi(i,j) = Hi[(i-1)+3*(j-1)]
  for (int i = 0; i < nmax; i++) {
    double rr[3] = {a[i], a[(size_t)r.NB + i], a[2 * (size_t)r.NB + i]};
    for (int c = 0; c < 3; c++) {
      double s = sum3(Hi[c] * rr[0], Hi[c + 3] * rr[1], Hi[c + 6] * rr[2]);
      a[(size_t)c * r.NB + i] = s - r.box.OBOX[c];
    }
  }
}
void xs2xu_inplace(Rank &r, int nmax, std::vector<double> &a) {
  const double *H = r.box.HH;
  for (int i = 0; i < nmax; i++) {
    double rr[3] = {a[i] + r.box.OBOX[0], a[(size_t)r.NB + i] + r.box.OBOX[1], a[2 * (size_t)r.NB + i] + r.box.OBOX[2]};
    for (int c = 0; c < 3; c++) a[(size_t)c * r.NB + i] = sum3(H[c] * rr[0], H[c + 3] * rr[1], H[c + 6] * rr[2]);
  }
}

// ---------------------------------------------------------------------------------------------
// COPYATOMS, src/comm.F90:2-597.  Executed for all ranks in lock step.
struct PackSpec {
  std::vector<std::vector<double> *> p2d;   // 3-vectors, packed first
  std::vector<char> shift2d;
  std::vector<std::vector<double> *> p1d;
  int cpbk1d = -1;                          // index in p1d that carries frcindx (value = sender's index)
};

PackSpec make_spec(Rank &r, int imode, bool isPQEq) {    // src/comm.F90:104-229
  PackSpec s;
  switch (imode) {
    case MODE_COPY:
      s.p2d = {&r.pos}; s.shift2d = {1};
      if (isPQEq) { s.p2d.push_back(&r.spos); s.shift2d.push_back(0); }   // :122,129-131
      s.p1d = {&r.atype, &r.q, &r.qs, &r.qt, &r.hs, &r.ht, &r.frcindx}; s.cpbk1d = 6;
      break;
    case MODE_MOVE:
      s.p2d = {&r.pos, &r.v}; s.shift2d = {1, 0};
      if (isPQEq) { s.p2d.push_back(&r.spos); s.shift2d.push_back(0); }   // :153,165-167
      s.p1d = {&r.atype, &r.q, &r.qs, &r.qt, &r.qsfp, &r.qsfv};
      break;
    case MODE_QCOPY1: s.p1d = {&r.qs, &r.qt}; break;
    case MODE_QCOPY2: s.p1d = {&r.hs, &r.ht, &r.q}; break;
    default: break;
  }
  return s;
}

bool inBuffer(const Rank &r, int dflag, const double *dr, double rr) {   // src/comm.F90:551-576
  switch (dflag) {
    case 1: return r.box.LBOX[0] - dr[0] < rr;
    case 2: return rr <= dr[0];
    case 3: return r.box.LBOX[1] - dr[1] < rr;
    case 4: return rr <= dr[1];
    case 5: return r.box.LBOX[2] - dr[2] < rr;
    default: return rr <= dr[2];
  }
}

int COPYATOMS(World &w, int imode, const double *dr) {
  static const int dinv[7] = {0, 2, 1, 4, 3, 6, 5};
  static const int cptridx[7] = {0, 0, 0, 2, 2, 4, 4};
  static const int is_xyz[7] = {0, 1, 1, 2, 2, 3, 3};
  const int nr_ = (int)w.R.size();
  std::vector<PackSpec> spec(nr_);
  // initialize, src/comm.F90:104-229
  for (int ir = 0; ir < nr_; ir++) {
    Rank &r = w.R[ir];
    r.na = r.ns = r.nr = 0;
    r.copyptr[0] = r.natoms;
    spec[ir] = make_spec(r, imode, w.P.cfg.isPQEq != 0);
    if (imode == MODE_CPBK) {
      r.ne = 4;
    } else {
      r.ne = (int)spec[ir].p2d.size() * 3 + (int)spec[ir].p1d.size();
      if (imode == MODE_COPY)
        for (int a = 0; a < r.natoms; a++) r.frcindx[a] = a;
      int nmax = r.copyptr[6] > r.natoms ? r.copyptr[6] : r.natoms;
      xu2xs_inplace(r, nmax, r.pos);
    }
  }
  for (int dflag = 1; dflag <= 6; dflag++) {
    // --- step_preparation + store_atoms, src/comm.F90:273-288, 367-453
    for (int ir = 0; ir < nr_; ir++) {
      Rank &r = w.R[ir];
      PackSpec &S = spec[ir];
      r.ns = 0;
      r.sbuffer.clear();
      if (imode == MODE_CPBK) {
        int is = 7 - dflag;
        for (int n = r.copyptr[is - 1]; n < r.copyptr[is]; n++) {
          r.sbuffer.push_back(r.frcindx[n]);
          r.sbuffer.push_back(FX(r, n)); r.sbuffer.push_back(FY(r, n)); r.sbuffer.push_back(FZ(r, n));
          r.ns += r.ne;
        }
      } else {
        int axis = is_xyz[dflag] - 1;
        int nsel = r.copyptr[cptridx[dflag]];
        double sft = (dflag % 2 == 1) ? -r.box.LBOX[axis] : r.box.LBOX[axis];   // xshift, src/comm.F90:531-548
        for (int n = 0; n < nsel; n++) {
          if (!inBuffer(r, dflag, dr, r.pos[(size_t)axis * r.NB + n])) continue;
          for (size_t a = 0; a < S.p2d.size(); a++) {
            double t[3] = {(*S.p2d[a])[n], (*S.p2d[a])[(size_t)r.NB + n], (*S.p2d[a])[2 * (size_t)r.NB + n]};
            if (S.shift2d[a]) t[axis] = t[axis] + sft;
            r.sbuffer.push_back(t[0]); r.sbuffer.push_back(t[1]); r.sbuffer.push_back(t[2]);
          }
          for (size_t a = 0; a < S.p1d.size(); a++)
            r.sbuffer.push_back((int)a == S.cpbk1d ? (double)n : (*S.p1d[a])[n]);
          if (imode == MODE_MOVE) r.atype[n] = -1.0;
          r.ns += r.ne;
        }
      }
    }
    // --- send_recv, src/comm.F90:291-364: I send to tn1 and receive from tn2
    for (int ir = 0; ir < nr_; ir++) {
      Rank &r = w.R[ir];
      int tn2 = r.box.target_node[dinv[dflag] - 1];
      if (imode == MODE_CPBK) tn2 = r.box.target_node[(7 - dflag) - 1];
      Rank &src = w.R[tn2];
      r.rbuffer = src.sbuffer;
      r.nr = src.ns;
    }
    // --- append_atoms, src/comm.F90:456-528
    for (int ir = 0; ir < nr_; ir++) {
      Rank &r = w.R[ir];
      PackSpec &S = spec[ir];
      if (imode == MODE_CPBK) {
        for (int i = 0; i < r.nr / r.ne; i++) {
          int ine = i * r.ne;
          int m = nint(r.rbuffer[ine]);
          FX(r, m) += r.rbuffer[ine + 1]; FY(r, m) += r.rbuffer[ine + 2]; FZ(r, m) += r.rbuffer[ine + 3];
        }
      } else {
        if ((r.na + r.nr) / r.ne > r.NB || r.copyptr[dflag - 1] + r.nr / r.ne > r.NB) {
          w.err = "ERROR: over capacity in append_atoms";
          return RXG_ERR_NBUFFER;
        }
        r.copyptr[dflag] = r.copyptr[dflag - 1] + r.nr / r.ne;
        for (int i = 0; i < r.nr / r.ne; i++) {
          int off = i * r.ne;
          int m = r.copyptr[dflag - 1] + i;
          for (size_t a = 0; a < S.p2d.size(); a++) {
            (*S.p2d[a])[m] = r.rbuffer[off]; (*S.p2d[a])[(size_t)r.NB + m] = r.rbuffer[off + 1];
            (*S.p2d[a])[2 * (size_t)r.NB + m] = r.rbuffer[off + 2];
            off += 3;
          }
          for (size_t a = 0; a < S.p1d.size(); a++) (*S.p1d[a])[m] = r.rbuffer[off++];
        }
      }
      r.na += r.nr;
    }
  }
  // finalize, src/comm.F90:232-270
  for (int ir = 0; ir < nr_; ir++) {
    Rank &r = w.R[ir];
    PackSpec &S = spec[ir];
    if (imode == MODE_MOVE) {
      int ni = 0;
      for (int i = 0; i < r.copyptr[6]; i++) {
        if (nint(r.atype[i]) > 0) {
          for (auto *p : S.p2d)
            for (int c = 0; c < 3; c++) (*p)[(size_t)c * r.NB + ni] = (*p)[(size_t)c * r.NB + i];
          for (auto *p : S.p1d) (*p)[ni] = (*p)[i];
          ni++;
        }
      }
      r.natoms = ni;
    }
    if (imode != MODE_CPBK) xs2xu_inplace(r, r.copyptr[6], r.pos);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// LINKEDLIST, src/main.F90:277-318
int LINKEDLIST(World &w, Rank &r, Grid &g, const double *cellDims) {
  const double *Hi = r.box.HHi;
  std::fill(g.header.begin(), g.header.end(), -1);
  std::fill(g.nacell.begin(), g.nacell.end(), 0);
  std::fill(g.llist.begin(), g.llist.end(), -1);
  for (int n = 0; n < r.copyptr[6]; n++) {
    if (nint(r.atype[n]) == 0) continue;
    double rr[3] = {RX(r, n), RY(r, n), RZ(r, n)};
    int l[3];
    for (int c = 0; c < 3; c++) {
      double rn = sum3(Hi[c] * rr[0], Hi[c + 3] * rr[1], Hi[c + 6] * rr[2]) - r.box.OBOX[c];   // xu2xs :596-611
      l[c] = (int)std::floor(rn / cellDims[c]);
    }
    if (!g.inside(l[0], l[1], l[2])) { w.err = "LINKEDLIST: atom outside the layered cell grid"; return RXG_ERR_STATE; }
    size_t ci = g.idx(l[0], l[1], l[2]);
    g.llist[n] = g.header[ci];
    g.header[ci] = n;
    g.nacell[ci]++;
  }
  return 0;
}

// NEIGHBORLIST, src/main.F90:321-417
int NEIGHBORLIST(World &w, Rank &r, int nlayer) {
  const Params &P = w.P;
  Grid &g = r.g;
  std::fill(r.nbrcnt.begin(), r.nbrcnt.end(), 0);
  int overflow = 0;
#pragma omp parallel for collapse(3) schedule(dynamic, 4) reduction(max : overflow)
  for (int c1 = -nlayer; c1 < g.nc[0] + nlayer; c1++)
    for (int c2 = -nlayer; c2 < g.nc[1] + nlayer; c2++)
      for (int c3 = -nlayer; c3 < g.nc[2] + nlayer; c3++) {
        size_t ci = g.idx(c1, c2, c3);
        int m = g.header[ci];
        for (int m1 = 0; m1 < g.nacell[ci]; m1++) {
          int mty = nint(r.atype[m]);
          for (int c4 = -1; c4 <= 1; c4++)
            for (int c5 = -1; c5 <= 1; c5++)
              for (int c6 = -1; c6 <= 1; c6++) {
                size_t cj = g.idx(c1 + c4, c2 + c5, c3 + c6);
                int n = g.header[cj];
                for (int n1 = 0; n1 < g.nacell[cj]; n1++) {
                  if (n != m) {
                    int nty = nint(r.atype[n]);
                    int inxn = P.inxn2(mty, nty);
                    double d0 = RX(r, n) - RX(r, m), d1 = RY(r, n) - RY(r, m), d2 = RZ(r, n) - RZ(r, m);
                    double dr2 = sum3(d0 * d0, d1 * d1, d2 * d2);
                    if (inxn > 0 && dr2 < P.rc2[inxn - 1]) {   // inxn==0 guard: SURVEY Q11
                      if (r.nbrcnt[m] < r.MAXN) r.nbrlist[(size_t)m * r.MAXN + r.nbrcnt[m]] = n;
                      r.nbrcnt[m]++;
                      if (r.nbrcnt[m] > overflow) overflow = r.nbrcnt[m];
                    }
                  }
                  n = g.llist[n];
                }
              }
          m = g.llist[m];
        }
      }
  if (overflow > r.MAXN) {   // src/main.F90:401-407 (checked before use here: rows are MAXN wide)
    w.err = "ERROR: overflow of max # in neighbor list";
    return RXG_ERR_MAXNEIGHBS;
  }
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int i = 0; i < r.copyptr[6]; i++) {   // :383-398
    for (int i1 = 0; i1 < r.nbrcnt[i]; i1++) {
      int j = r.nbrlist[(size_t)i * r.MAXN + i1];
      bool found = false;
      for (int j1 = 0; j1 < r.nbrcnt[j]; j1++)
        if (i == r.nbrlist[(size_t)j * r.MAXN + j1]) { r.nbrindx[(size_t)i * r.MAXN + i1] = j1; found = true; }
      if (!found) bad++;
    }
  }
  if (bad) { w.err = "ERROR: inconsistency between nbrlist and nbrindx found"; return RXG_ERR_STATE; }
  return 0;
}

// get_coulomb_and_dcoulomb_pqeq, src/module.F90:386-417 (the code after the first `return` is dead).
// Returns false when the reference returns early (dr2 > rctap2) WITHOUT assigning Eclmb / ff.  Two call sites then read
// a variable the reference never defined for that pair (pqeqs in qeq_initialize src/pqeq.F90:340-343, sf in
// update_shell_positions :219-231): undefined behaviour in the reference.  The oracle (and the CUDA path) take the
// physically intended value, a zero contribution (the taper is 0 at the cut-off), and count the events in pqeq_skips so
// that parity tests can state that none occurred on their inputs.  Table indices outside 1..NTABLE (reference reads
// out of bounds, like SURVEY Q9) give 0.
bool get_clmb_pqeq(const Params &P, const double *rr, double &Eclmb, int inxn, const std::vector<double> &T, double *ff) {
  double dr2 = sum3(rr[0] * rr[0], rr[1] * rr[1], rr[2] * rr[2]);
  if (dr2 > P.rctap2) return false;
  int itb = (int)(dr2 * P.UDRi);
  int itb1 = itb + 1;
  double drtb = dr2 - itb * P.UDR;
  drtb = drtb * P.UDRi;
  double drtb1 = 1.0 - drtb;
  if (itb < 1 || itb1 > P.ntable) { Eclmb = 0.0; ff[0] = ff[1] = ff[2] = 0.0; return true; }
  Eclmb = drtb1 * P.tblp(T, inxn, itb, 0) + drtb * P.tblp(T, inxn, itb1, 0);
  double dEclmb = drtb1 * P.tblp(T, inxn, itb, 1) + drtb * P.tblp(T, inxn, itb1, 1);
  ff[0] = dEclmb * rr[0]; ff[1] = dEclmb * rr[1]; ff[2] = dEclmb * rr[2];
  return true;
}
#define SX(r, i) r.spos[(i)]
#define SY(r, i) r.spos[(size_t)r.NB + (i)]
#define SZ(r, i) r.spos[2 * (size_t)r.NB + (i)]

// GetNonbondingPairList (fp64, <=) src/main.F90:420-477 and qeq_initialize (fp32, <, hessian) src/qeq.F90:183-268
int PairList(World &w, Rank &r, bool qeq) {
  const Params &P = w.P;
  Grid &g = r.nbg;
  const int nmesh = r.box.nbnmesh;
  std::fill(r.nbpcnt.begin(), r.nbpcnt.end(), 0);
  int status = 0;
  const bool pqeq = qeq && P.cfg.isPQEq;
  long long skips = 0;
#pragma omp parallel for collapse(3) schedule(dynamic, 2) reduction(+ : skips)
  for (int c1 = 0; c1 < g.nc[0]; c1++)
    for (int c2 = 0; c2 < g.nc[1]; c2++)
      for (int c3 = 0; c3 < g.nc[2]; c3++) {
        size_t ci = g.idx(c1, c2, c3);
        int i = g.header[ci];
        for (int m = 0; m < g.nacell[ci]; m++, i = g.llist[i]) {
          if (i >= r.natoms) { status = 1; continue; }
          int ity = nint(r.atype[i]);
          int cnt = 0;
          double fp = 0.0;   // fpqeq(i), src/pqeq.F90:294
          for (int mn = 0; mn < nmesh; mn++) {
            int c4 = c1 + r.nbmesh[3 * mn], c5 = c2 + r.nbmesh[3 * mn + 1], c6 = c3 + r.nbmesh[3 * mn + 2];
            if (!g.inside(c4, c5, c6)) continue;   // the reference's array is sized so this never triggers
            size_t cj = g.idx(c4, c5, c6);
            int j = g.header[cj];
            for (int n = 0; n < g.nacell[cj]; n++, j = g.llist[j]) {
              if (i == j) continue;
              double d0 = RX(r, i) - RX(r, j), d1 = RY(r, i) - RY(r, j), d2 = RZ(r, i) - RZ(r, j);
              double dr2d = sum3(d0 * d0, d1 * d1, d2 * d2);
              if (!qeq) {
                if (dr2d <= P.rctap2) {
                  if (cnt < r.W10) r.nbplist[(size_t)i * r.W10 + cnt] = j;
                  cnt++;
                }
              } else {
                float dr2 = (float)dr2d;                 // real(4) :: dr2, src/qeq.F90:191
                if (dr2 < (float)P.rctap2) {             // rctap2 = 100 or 156.25, exact in fp32
                  if (cnt < r.W10 && pqeq) {             // qeq_initialize of src/pqeq.F90:262-365
                    r.nbplist[(size_t)i * r.W10 + cnt] = j;
                    int jty = nint(r.atype[j]);
                    double rr[3] = {d0, d1, d2}, pqeqc = 0.0, pqeqs = 0.0, ffd[3];
                    get_clmb_pqeq(P, rr, pqeqc, P.inxnpqeq(ity, jty), P.TBL_pcc, ffd);   // never skips: fp32 dr2 < rctap2
                    r.hessian[(size_t)i * r.W10 + cnt] = CCLMB0_QEQ * pqeqc;
                    fp = fp + CCLMB0_QEQ * pqeqc * P.Zpqeq[jty - 1];                    // Eq. 30, :336
                    if (P.polarizable(jty)) {
                      double rs[3] = {RX(r, i) - RX(r, j) - SX(r, j), RY(r, i) - RY(r, j) - SY(r, j), RZ(r, i) - RZ(r, j) - SZ(r, j)};
                      if (!get_clmb_pqeq(P, rs, pqeqs, P.inxnpqeq(jty, ity), P.TBL_psc, ffd)) skips++;
                      fp = fp - CCLMB0_QEQ * pqeqs * P.Zpqeq[jty - 1];
                    }
                  } else if (cnt < r.W10) {
                    r.nbplist[(size_t)i * r.W10 + cnt] = j;
                    int jty = nint(r.atype[j]);
                    // itb = int(dr2*UDRi): real(4)*real(8) promotes dr2 to double, src/qeq.F90:234-236
                    int itb = (int)((double)dr2 * P.UDRi);
                    double drtb = (double)dr2 - itb * P.UDR;
                    drtb = drtb * P.UDRi;
                    int inxn = P.inxn2(ity, jty);
                    double h = 0.0;
                    if (inxn > 0 && itb >= 1 && itb < P.ntable)
                      h = (1.0 - drtb) * P.eqeq(itb, inxn) + drtb * P.eqeq(itb + 1, inxn);
                    r.hessian[(size_t)i * r.W10 + cnt] = h;
                  }
                  cnt++;
                }
              }
            }
          }
          r.nbpcnt[i] = cnt;
          if (pqeq) r.fpqeq[i] = fp;
          if (cnt > r.W10) status = 2;
        }
      }
  r.pqeq_skips += skips;
  if (status == 2) { w.err = "ERROR: nbplist greater then MAXNEIGHBS10"; return RXG_ERR_MAXNEIGHBS10; }
  if (status == 1) { w.err = "PairList: ghost atom inside a resident non-bonded cell"; return RXG_ERR_STATE; }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// QEq, src/qeq.F90
void get_hsh(const Params &P, Rank &r, double &Est, double &hshs_sum, double &hsht_sum) {   // :271-318
  double e = 0, s = 0, t = 0;
#pragma omp parallel for schedule(static) reduction(+ : e, s, t)
  for (int i = 0; i < r.natoms; i++) {
    int ity = nint(r.atype[i]);
    double eta_ity = P.eta[ity - 1];
    double t_hshs = eta_ity * r.hs[i], t_hsht = eta_ity * r.ht[i];
    e = e + P.chi[ity - 1] * r.q[i] + 0.5 * eta_ity * r.q[i] * r.q[i];
    const int *nl = &r.nbplist[(size_t)i * r.W10];
    const double *hv = &r.hessian[(size_t)i * r.W10];
    for (int j1 = 0; j1 < r.nbpcnt[i]; j1++) {
      int j = nl[j1];
      t_hshs = t_hshs + hv[j1] * r.hs[j];
      t_hsht = t_hsht + hv[j1] * r.ht[j];
      double Est1 = 0.5 * hv[j1] * r.q[i] * r.q[j];
      e = e + Est1;
      if (j < r.natoms) e = e + Est1;
    }
    s = s + t_hshs * r.hs[i];
    t = t + t_hsht * r.ht[i];
  }
  Est = e; hshs_sum = s; hsht_sum = t;
}

void get_gradient(const Params &P, Rank &r, double *gg) {   // :321-363 (local part; allreduce by caller)
#pragma omp parallel for schedule(static)
  for (int i = 0; i < r.natoms; i++) {
    double gssum = 0, gtsum = 0;
    const int *nl = &r.nbplist[(size_t)i * r.W10];
    const double *hv = &r.hessian[(size_t)i * r.W10];
    for (int j1 = 0; j1 < r.nbpcnt[i]; j1++) {
      int j = nl[j1];
      gssum = gssum + hv[j1] * r.qs[j];
      gtsum = gtsum + hv[j1] * r.qt[j];
    }
    int ity = nint(r.atype[i]);
    double eta_ity = P.eta[ity - 1];
    r.gs[i] = -P.chi[ity - 1] - eta_ity * r.qs[i] - gssum;
    r.gt[i] = -1.0 - eta_ity * r.qt[i] - gtsum;
  }
  double a = 0, b = 0;
  for (int i = 0; i < r.natoms; i++) { a += r.gs[i] * r.gs[i]; b += r.gt[i] * r.gt[i]; }
  gg[0] = a; gg[1] = b;
}

int QEq(World &w) {   // src/qeq.F90:2-178
  const Params &P = w.P;
  const int nr_ = (int)w.R.size();
  int nmax;
  if (P.cfg.isQEq == 1) {
    for (auto &r : w.R) {
      for (int i = 0; i < r.natoms; i++) { r.qsfp[i] = r.q[i]; r.qsfv[i] = 0.0; }
      std::fill(r.qs.begin(), r.qs.end(), 0.0);
      std::fill(r.qt.begin(), r.qt.end(), 0.0);
      for (int i = 0; i < r.natoms; i++) r.qs[i] = r.q[i];
    }
    nmax = P.cfg.NMAXQEq;
  } else if (P.cfg.isQEq == 2) {
    for (auto &r : w.R)
      for (int i = 0; i < r.natoms; i++) {
        r.qs[i] = P.cfg.Lex_fqs * r.qsfp[i] + (1.0 - P.cfg.Lex_fqs) * r.q[i];
        r.qt[i] = 0.0;
      }
    nmax = 1;
  } else {
    return 0;
  }
  double QCopyDr[3] = {P.rctap / w.R[0].box.lata, P.rctap / w.R[0].box.latb, P.rctap / w.R[0].box.latc};
  int rc = COPYATOMS(w, MODE_COPY, QCopyDr);
  if (rc) return rc;
  for (auto &r : w.R) {
    rc = LINKEDLIST(w, r, r.nbg, r.box.nblcsize);
    if (rc) return rc;
    rc = PairList(w, r, true);
    if (rc) return rc;
  }
  COPYATOMS(w, MODE_QCOPY1, QCopyDr);
  double Gnew[2] = {0, 0}, Gold[2];
  for (auto &r : w.R) { double gg[2]; get_gradient(P, r, gg); Gnew[0] += gg[0]; Gnew[1] += gg[1]; }
  for (auto &r : w.R)
    for (int i = 0; i < r.natoms; i++) { r.hs[i] = r.gs[i]; r.ht[i] = r.gt[i]; }
  COPYATOMS(w, MODE_QCOPY2, QCopyDr);
  double GEst2 = 1e99;
  int nstep_qeq;
  for (nstep_qeq = 0; nstep_qeq < nmax; nstep_qeq++) {
    double GEst1 = 0, h_hsh[2] = {0, 0}, g_h[2] = {0, 0};
    for (int ir = 0; ir < nr_; ir++) {
      double Est, a, b;
      get_hsh(P, w.R[ir], Est, a, b);
      GEst1 += Est; h_hsh[0] += a; h_hsh[1] += b;
    }
    if (0.5 * (std::fabs(GEst2) + std::fabs(GEst1)) < P.cfg.QEq_tol) break;                     // :114
    if (std::fabs(GEst2) > 0.0 && std::fabs(GEst1 / GEst2 - 1.0) < P.cfg.QEq_tol) break;       // :115
    GEst2 = GEst1;
    for (auto &r : w.R) {
      double a = 0, b = 0;
      for (int i = 0; i < r.natoms; i++) { a += r.gs[i] * r.hs[i]; b += r.gt[i] * r.ht[i]; }
      g_h[0] += a; g_h[1] += b;
    }
    float lmin[2] = {(float)(g_h[0] / h_hsh[0]), (float)(g_h[1] / h_hsh[1])};   // real(4) :: lmin, :23,133 (Q3)
    double ssum = 0, tsum = 0;
    for (auto &r : w.R) {
      double a = 0, b = 0;
      for (int i = 0; i < r.natoms; i++) {
        r.qs[i] = r.qs[i] + (double)lmin[0] * r.hs[i];
        r.qt[i] = r.qt[i] + (double)lmin[1] * r.ht[i];
      }
      for (int i = 0; i < r.natoms; i++) { a += r.qs[i]; b += r.qt[i]; }
      ssum += a; tsum += b;
    }
    double mu = ssum / tsum;
    for (auto &r : w.R)
      for (int i = 0; i < r.natoms; i++) r.q[i] = r.qs[i] - mu * r.qt[i];
    COPYATOMS(w, MODE_QCOPY1, QCopyDr);
    Gold[0] = Gnew[0]; Gold[1] = Gnew[1];
    Gnew[0] = Gnew[1] = 0;
    for (auto &r : w.R) { double gg[2]; get_gradient(P, r, gg); Gnew[0] += gg[0]; Gnew[1] += gg[1]; }
    for (auto &r : w.R)
      for (int i = 0; i < r.natoms; i++) {
        r.hs[i] = r.gs[i] + (Gnew[0] / Gold[0]) * r.hs[i];
        r.ht[i] = r.gt[i] + (Gnew[1] / Gold[1]) * r.ht[i];
      }
    COPYATOMS(w, MODE_QCOPY2, QCopyDr);
  }
  for (auto &r : w.R) r.nstep_qeq = nstep_qeq;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// PQEq, src/pqeq.F90
void get_hsh_pqeq(const Params &P, Rank &r, double &Est, double &hshs_sum, double &hsht_sum) {   // :368-439
  double e = 0, s = 0, t = 0;
#pragma omp parallel for schedule(static) reduction(+ : e, s, t)
  for (int i = 0; i < r.natoms; i++) {
    int ity = nint(r.atype[i]);
    double eta_ity = P.eta[ity - 1];
    double t_hshs = eta_ity * r.hs[i], t_hsht = eta_ity * r.ht[i];
    double qic = r.q[i] + P.Zpqeq[ity - 1];
    double shelli[3] = {RX(r, i) + SX(r, i), RY(r, i) + SY(r, i), RZ(r, i) + SZ(r, i)};
    e = e + P.chi[ity - 1] * r.q[i] + 0.5 * eta_ity * r.q[i] * r.q[i];
    const int *nl = &r.nbplist[(size_t)i * r.W10];
    const double *hv = &r.hessian[(size_t)i * r.W10];
    for (int j1 = 0; j1 < r.nbpcnt[i]; j1++) {
      int j = nl[j1];
      int jty = nint(r.atype[j]);
      double qjc = r.q[j] + P.Zpqeq[jty - 1];
      double shellj[3] = {RX(r, j) + SX(r, j), RY(r, j) + SY(r, j), RZ(r, j) + SZ(r, j)};
      double Ccicj = 0.0, Csicj = 0.0, Csisj = 0.0, ffd[3];
      Ccicj = hv[j1] * qic * qjc;
      if (P.polarizable(ity)) {
        double dr[3] = {shelli[0] - RX(r, j), shelli[1] - RY(r, j), shelli[2] - RZ(r, j)};
        get_clmb_pqeq(P, dr, Csicj, P.inxnpqeq(ity, jty), P.TBL_psc, ffd);
        Csicj = -CCLMB0_QEQ * Csicj * qjc * P.Zpqeq[ity - 1];
        if (P.polarizable(jty)) {
          double ds[3] = {shelli[0] - shellj[0], shelli[1] - shellj[1], shelli[2] - shellj[2]};
          get_clmb_pqeq(P, ds, Csisj, P.inxnpqeq(ity, jty), P.TBL_pss, ffd);
          Csisj = CCLMB0_QEQ * Csisj * P.Zpqeq[ity - 1] * P.Zpqeq[jty - 1];
        }
      }
      t_hshs = t_hshs + hv[j1] * r.hs[j];
      t_hsht = t_hsht + hv[j1] * r.ht[j];
      double Est1 = 0.5 * (Ccicj + Csisj);
      e = e + Est1 + Csicj;   // no resident/ghost distinction here (:430-433)
    }
    s = s + t_hshs * r.hs[i];
    t = t + t_hsht * r.ht[i];
  }
  Est = e; hshs_sum = s; hsht_sum = t;
}

void get_gradient_pqeq(const Params &P, Rank &r, double *gg) {   // :442-477
#pragma omp parallel for schedule(static)
  for (int i = 0; i < r.natoms; i++) {
    double gssum = 0, gtsum = 0;
    const int *nl = &r.nbplist[(size_t)i * r.W10];
    const double *hv = &r.hessian[(size_t)i * r.W10];
    for (int j1 = 0; j1 < r.nbpcnt[i]; j1++) {
      int j = nl[j1];
      gssum = gssum + hv[j1] * r.qs[j];
      gtsum = gtsum + hv[j1] * r.qt[j];
    }
    int ity = nint(r.atype[i]);
    double eta_ity = P.eta[ity - 1];
    r.gs[i] = -P.chi[ity - 1] - eta_ity * r.qs[i] - gssum - r.fpqeq[i];
    r.gt[i] = -1.0 - eta_ity * r.qt[i] - gtsum;
  }
  double a = 0, b = 0;
  for (int i = 0; i < r.natoms; i++) { a += r.gs[i] * r.gs[i]; b += r.gt[i] * r.gt[i]; }
  gg[0] = a; gg[1] = b;
}

void update_shell_positions(const Params &P, Rank &r, bool stale) {   // :187-259
  std::vector<double> sforce(3 * (size_t)r.natoms, 0.0);
  const size_t n = r.natoms;
  long long skips = 0;
  double sf_keep[3] = {0, 0, 0};   // `sf` of the reference: one variable for both calls, never reset (stale mode only)
#pragma omp parallel for schedule(guided) reduction(+ : skips) if (!stale)
  for (int i = 0; i < r.natoms; i++) {
    int ity = nint(r.atype[i]);
    if (!P.polarizable(ity)) continue;
    double sf0 = 0, sf1 = 0, sf2 = 0;
    if (P.cfg.isEfield) {
      double e = P.Zpqeq[ity - 1] * P.cfg.eFieldStrength * EEV_KCAL;
      if (P.cfg.eFieldDir == 1) sf0 = sf0 - e; else if (P.cfg.eFieldDir == 2) sf1 = sf1 - e; else sf2 = sf2 - e;
    }
    sf0 = sf0 - P.Kspqeq[ity - 1] * SX(r, i); sf1 = sf1 - P.Kspqeq[ity - 1] * SY(r, i); sf2 = sf2 - P.Kspqeq[ity - 1] * SZ(r, i);   // Eq. 37
    double shelli[3] = {RX(r, i) + SX(r, i), RY(r, i) + SY(r, i), RZ(r, i) + SZ(r, i)};
    const int *nl = &r.nbplist[(size_t)i * r.W10];
    for (int j1 = 0; j1 < r.nbpcnt[i]; j1++) {
      int j = nl[j1];
      int jty = nint(r.atype[j]);
      double qjc = r.q[j] + P.Zpqeq[jty - 1];
      double shellj[3] = {RX(r, j) + SX(r, j), RY(r, j) + SY(r, j), RZ(r, j) + SZ(r, j)};
      double dr[3] = {shelli[0] - RX(r, j), shelli[1] - RY(r, j), shelli[2] - RZ(r, j)};
      double Esc = 0, sf[3] = {0, 0, 0};
      if (stale) { sf[0] = sf_keep[0]; sf[1] = sf_keep[1]; sf[2] = sf_keep[2]; }
      if (!get_clmb_pqeq(P, dr, Esc, P.inxnpqeq(ity, jty), P.TBL_psc, sf)) skips++;
      double ff[3] = {-CCLMB0 * sf[0] * qjc * P.Zpqeq[ity - 1], -CCLMB0 * sf[1] * qjc * P.Zpqeq[ity - 1], -CCLMB0 * sf[2] * qjc * P.Zpqeq[ity - 1]};
      sf0 = sf0 - ff[0]; sf1 = sf1 - ff[1]; sf2 = sf2 - ff[2];
      double ss[3] = {sf[0], sf[1], sf[2]};   // the same variable goes into the shell-shell call
      if (P.polarizable(jty)) {
        double ds[3] = {shelli[0] - shellj[0], shelli[1] - shellj[1], shelli[2] - shellj[2]};
        double Ess = 0;
        if (!stale) { ss[0] = ss[1] = ss[2] = 0.0; }
        if (!get_clmb_pqeq(P, ds, Ess, P.inxnpqeq(ity, jty), P.TBL_pss, ss)) skips++;
        double f2[3] = {CCLMB0 * ss[0] * P.Zpqeq[ity - 1] * P.Zpqeq[jty - 1], CCLMB0 * ss[1] * P.Zpqeq[ity - 1] * P.Zpqeq[jty - 1],
                        CCLMB0 * ss[2] * P.Zpqeq[ity - 1] * P.Zpqeq[jty - 1]};
        sf0 = sf0 - f2[0]; sf1 = sf1 - f2[1]; sf2 = sf2 - f2[2];
      }
      if (stale) { sf_keep[0] = ss[0]; sf_keep[1] = ss[1]; sf_keep[2] = ss[2]; }
    }
    sforce[i] = sf0; sforce[n + i] = sf1; sforce[2 * n + i] = sf2;
  }
  r.pqeq_skips += skips;
  for (int i = 0; i < r.natoms; i++) {   // Eq. 39, :240-256
    int ity = nint(r.atype[i]);
    if (!P.polarizable(ity)) continue;   // (the reference divides by Kspqeq = 0 for them and discards the result)
    double dr[3] = {sforce[i] / P.Kspqeq[ity - 1], sforce[n + i] / P.Kspqeq[ity - 1], sforce[2 * n + i] / P.Kspqeq[ity - 1]};
    double ddr = std::sqrt(sum3(dr[0] * dr[0], dr[1] * dr[1], dr[2] * dr[2]));
    if (ddr > MAX_SHELL_DISPLACEMENT)
      for (int c = 0; c < 3; c++) dr[c] = dr[c] / ddr * MAX_SHELL_DISPLACEMENT;
    SX(r, i) = SX(r, i) + dr[0]; SY(r, i) = SY(r, i) + dr[1]; SZ(r, i) = SZ(r, i) + dr[2];
  }
}

// Diagnostic (pqeq_stale): redo the shell term of fpqeq the way a serial build of the reference does it -- `pqeqs` keeps the
// value of the last core-shell call that assigned it, across pairs and across atoms, in the loop order of qeq_initialize
// (cells c1, c2, c3; atoms of a cell in list order; neighbours in row order).  src/pqeq.F90:298-345.
void stale_fpqeq(const Params &P, Rank &r) {
  Grid &g = r.nbg;
  double pqeqs = 0.0, ffd[3];
  for (int c1 = 0; c1 < g.nc[0]; c1++)
    for (int c2 = 0; c2 < g.nc[1]; c2++)
      for (int c3 = 0; c3 < g.nc[2]; c3++) {
        size_t ci = g.idx(c1, c2, c3);
        int i = g.header[ci];
        for (int m = 0; m < g.nacell[ci]; m++, i = g.llist[i]) {
          int ity = nint(r.atype[i]);
          const int *nl = &r.nbplist[(size_t)i * r.W10];
          double fp = 0.0;
          for (int j1 = 0; j1 < r.nbpcnt[i]; j1++) {
            int j = nl[j1];
            int jty = nint(r.atype[j]);
            fp = fp + r.hessian[(size_t)i * r.W10 + j1] * P.Zpqeq[jty - 1];
            if (P.polarizable(jty)) {
              double rs[3] = {RX(r, i) - RX(r, j) - SX(r, j), RY(r, i) - RY(r, j) - SY(r, j), RZ(r, i) - RZ(r, j) - SZ(r, j)};
              get_clmb_pqeq(P, rs, pqeqs, P.inxnpqeq(jty, ity), P.TBL_psc, ffd);   // early return: pqeqs keeps its old value
              fp = fp - CCLMB0_QEQ * pqeqs * P.Zpqeq[jty - 1];
            }
          }
          r.fpqeq[i] = fp;
        }
      }
}

int PQEq(World &w) {   // src/pqeq.F90:2-182: the QEq driver with the PQEq kernels and the shell relaxation at the end
  const Params &P = w.P;
  const int nr_ = (int)w.R.size();
  int nmax;
  if (P.cfg.isQEq == 1) {
    for (auto &r : w.R) {
      for (int i = 0; i < r.natoms; i++) { r.qsfp[i] = r.q[i]; r.qsfv[i] = 0.0; }
      std::fill(r.qs.begin(), r.qs.end(), 0.0);
      std::fill(r.qt.begin(), r.qt.end(), 0.0);
      for (int i = 0; i < r.natoms; i++) r.qs[i] = r.q[i];
    }
    nmax = P.cfg.NMAXQEq;
  } else if (P.cfg.isQEq == 2) {
    for (auto &r : w.R)
      for (int i = 0; i < r.natoms; i++) {
        r.qs[i] = P.cfg.Lex_fqs * r.qsfp[i] + (1.0 - P.cfg.Lex_fqs) * r.q[i];
        r.qt[i] = 0.0;
      }
    nmax = 1;
  } else {
    return 0;
  }
  double QCopyDr[3] = {P.rctap / w.R[0].box.lata, P.rctap / w.R[0].box.latb, P.rctap / w.R[0].box.latc};
  int rc = COPYATOMS(w, MODE_COPY, QCopyDr);
  if (rc) return rc;
  for (auto &r : w.R) {
    rc = LINKEDLIST(w, r, r.nbg, r.box.nblcsize);
    if (rc) return rc;
    rc = PairList(w, r, true);
    if (rc) return rc;
    if (w.pqeq_stale) stale_fpqeq(P, r);
  }
  COPYATOMS(w, MODE_QCOPY1, QCopyDr);
  double Gnew[2] = {0, 0}, Gold[2];
  for (auto &r : w.R) { double gg[2]; get_gradient_pqeq(P, r, gg); Gnew[0] += gg[0]; Gnew[1] += gg[1]; }
  for (auto &r : w.R)
    for (int i = 0; i < r.natoms; i++) { r.hs[i] = r.gs[i]; r.ht[i] = r.gt[i]; }
  COPYATOMS(w, MODE_QCOPY2, QCopyDr);
  double GEst2 = 1e99;
  int nstep_qeq;
  for (nstep_qeq = 0; nstep_qeq < nmax; nstep_qeq++) {
    double GEst1 = 0, h_hsh[2] = {0, 0}, g_h[2] = {0, 0};
    for (int ir = 0; ir < nr_; ir++) {
      double Est, a, b;
      get_hsh_pqeq(P, w.R[ir], Est, a, b);
      GEst1 += Est; h_hsh[0] += a; h_hsh[1] += b;
    }
    if (0.5 * (std::fabs(GEst2) + std::fabs(GEst1)) < P.cfg.QEq_tol) break;
    if (std::fabs(GEst2) > 0.0 && std::fabs(GEst1 / GEst2 - 1.0) < P.cfg.QEq_tol) break;
    GEst2 = GEst1;
    for (auto &r : w.R) {
      double a = 0, b = 0;
      for (int i = 0; i < r.natoms; i++) { a += r.gs[i] * r.hs[i]; b += r.gt[i] * r.ht[i]; }
      g_h[0] += a; g_h[1] += b;
    }
    float lmin[2] = {(float)(g_h[0] / h_hsh[0]), (float)(g_h[1] / h_hsh[1])};   // real(4) :: lmin, src/pqeq.F90:27
    double ssum = 0, tsum = 0;
    for (auto &r : w.R) {
      double a = 0, b = 0;
      for (int i = 0; i < r.natoms; i++) {
        r.qs[i] = r.qs[i] + (double)lmin[0] * r.hs[i];
        r.qt[i] = r.qt[i] + (double)lmin[1] * r.ht[i];
      }
      for (int i = 0; i < r.natoms; i++) { a += r.qs[i]; b += r.qt[i]; }
      ssum += a; tsum += b;
    }
    double mu = ssum / tsum;
    for (auto &r : w.R)
      for (int i = 0; i < r.natoms; i++) r.q[i] = r.qs[i] - mu * r.qt[i];
    COPYATOMS(w, MODE_QCOPY1, QCopyDr);
    Gold[0] = Gnew[0]; Gold[1] = Gnew[1];
    Gnew[0] = Gnew[1] = 0;
    for (auto &r : w.R) { double gg[2]; get_gradient_pqeq(P, r, gg); Gnew[0] += gg[0]; Gnew[1] += gg[1]; }
    for (auto &r : w.R)
      for (int i = 0; i < r.natoms; i++) {
        r.hs[i] = r.gs[i] + (Gnew[0] / Gold[0]) * r.hs[i];
        r.ht[i] = r.gt[i] + (Gnew[1] / Gold[1]) * r.ht[i];
      }
    COPYATOMS(w, MODE_QCOPY2, QCopyDr);
  }
  for (auto &r : w.R) { r.nstep_qeq = nstep_qeq; update_shell_positions(P, r, w.pqeq_stale); }   // :171
  return 0;
}

// ---------------------------------------------------------------------------------------------
// BOCALC, src/bo.F90
#define SLOT(i, s) ((size_t)(i) * r.MAXN + (s))

void BOPRIM(const Params &P, Rank &r) {   // src/bo.F90:28-118
  const int n = r.copyptr[6];
  for (int i = 0; i < n; i++) r.deltap1[i] = -P.Val[nint(r.atype[i]) - 1];
  for (int i = 0; i < n; i++) {
    int ity = nint(r.atype[i]);
    for (int j1 = 0; j1 < r.nbrcnt[i]; j1++) {
      int j = r.nbrlist[SLOT(i, j1)];
      if (j < i) {
        int jty = nint(r.atype[j]);
        int inxn = P.inxn2(ity, jty);
        int i1 = r.nbrindx[SLOT(i, j1)];
        size_t a = SLOT(i, j1), b = SLOT(j, i1);
        double d0 = RX(r, i) - RX(r, j), d1 = RY(r, i) - RY(r, j), d2 = RZ(r, i) - RZ(r, j);
        double dr2 = sum3(d0 * d0, d1 * d1, d2 * d2);
        int x = inxn - 1;
        if (dr2 <= P.rc2[x]) {
          double arg[3] = {P.cBOp1[x] * std::pow(dr2, P.pbo2h[x]), P.cBOp3[x] * std::pow(dr2, P.pbo4h[x]),
                           P.cBOp5[x] * std::pow(dr2, P.pbo6h[x])};
          double bo[4];
          for (int c = 0; c < 3; c++) bo[c + 1] = P.swtch[c + 3 * x] * std::exp(arg[c]);
          bo[1] = (1.0 + P.cutoff_vpar30) * bo[1];
          if (sum3(bo[1], bo[2], bo[3]) > P.cutoff_vpar30) {
            double dl[3] = {P.swtch[0 + 3 * x] * P.pbo2[x] * arg[0], P.swtch[1 + 3 * x] * P.pbo4[x] * arg[1],
                            P.swtch[2 + 3 * x] * P.pbo6[x] * arg[2]};
            for (int c = 0; c < 3; c++) { dl[c] = dl[c] / dr2; r.dln_BOp[c][a] = dl[c]; r.dln_BOp[c][b] = dl[c]; }
            double db = sum3(bo[1] * dl[0], bo[2] * dl[1], bo[3] * dl[2]);
            r.dBOp[a] = db; r.dBOp[b] = db;
            bo[1] = bo[1] - P.cutoff_vpar30;
            bo[0] = sum3(bo[1], bo[2], bo[3]);
            for (int c = 0; c < 4; c++) { r.BO[c][a] = bo[c]; r.BO[c][b] = bo[c]; }
            r.deltap1[i] += bo[0];
            r.deltap1[j] += bo[0];
          } else {
            r.dBOp[a] = 0; r.dBOp[b] = 0;
            for (int c = 0; c < 3; c++) { r.dln_BOp[c][a] = 0; r.dln_BOp[c][b] = 0; }   // reference leaves stale values
            for (int c = 0; c < 4; c++) { r.BO[c][a] = 0; r.BO[c][b] = 0; }
          }
        }
      }
    }
  }
}

void BOFULL(const Params &P, Rank &r) {   // src/bo.F90:121-298
  const int n = r.copyptr[6];
  for (int i = 0; i < n; i++) {
    int ity = nint(r.atype[i]);
    r.deltap2[i] = r.deltap1[i] + P.Val[ity - 1] - P.Valval[ity - 1];
  }
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; i++) {
    int ity = nint(r.atype[i]);
    double exppboc1i = std::exp(-P.vpar1 * r.deltap1[i]);
    double exppboc2i = std::exp(-P.vpar2 * r.deltap1[i]);
    for (int j1 = 0; j1 < r.nbrcnt[i]; j1++) {
      int j = r.nbrlist[SLOT(i, j1)];
      if (j < i) {
        int jty = nint(r.atype[j]);
        double exppboc1j = std::exp(-P.vpar1 * r.deltap1[j]);
        double exppboc2j = std::exp(-P.vpar2 * r.deltap1[j]);
        int i1 = r.nbrindx[SLOT(i, j1)];
        size_t a = SLOT(i, j1), b = SLOT(j, i1);
        int x = P.inxn2(ity, jty) - 1;
        double Vi = P.Val[ity - 1], Vj = P.Val[jty - 1];
        double fn2 = exppboc1i + exppboc1j;
        double fn3 = (-1.0 / P.vpar2) * std::log(0.5 * (exppboc2i + exppboc2j));
        double fn23 = fn2 + fn3;
        double BOp0 = r.BO[0][a];
        double fn1 = 0.5 * ((Vi + fn2) / (Vi + fn23) + (Vj + fn2) / (Vj + fn23));
        if (P.ovc[x] < 1e-3) fn1 = 1.0;
        double BOpsqr = r.BO[0][a] * r.BO[0][a];
        double fn4 = 1.0 / (1.0 + std::exp(-P.pboc3[x] * (P.pboc4[x] * BOpsqr - r.deltap2[i]) + P.pboc5[x]));
        double fn5 = 1.0 / (1.0 + std::exp(-P.pboc3[x] * (P.pboc4[x] * BOpsqr - r.deltap2[j]) + P.pboc5[x]));
        if (P.v13cor[x] < 1e-3) { fn4 = 1.0; fn5 = 1.0; }
        double fn45 = fn4 * fn5, fn145 = fn1 * fn45, fn1145 = fn1 * fn145;
        double B0 = r.BO[0][a] * fn145, B2 = r.BO[2][a] * fn1145, B3 = r.BO[3][a] * fn1145;
        if (B0 < 1e-10) B0 = 0.0;
        if (B2 < 1e-10) B2 = 0.0;
        if (B3 < 1e-10) B3 = 0.0;
        double B1 = B0 - B2 - B3;
        r.BO[0][a] = B0; r.BO[1][a] = B1; r.BO[2][a] = B2; r.BO[3][a] = B3;
        r.BO[0][b] = B0; r.BO[1][b] = B1; r.BO[2][b] = B2; r.BO[3][b] = B3;
        double u1ij = Vi + fn23, u1ji = Vj + fn23;
        double u1ij_inv2 = 1.0 / (u1ij * u1ij), u1ji_inv2 = 1.0 / (u1ji * u1ji);
        double Cf1Aij = 0.5 * fn3 * (u1ij_inv2 + u1ji_inv2);
        double Cf1Bij = -0.5 * ((u1ij - fn3) * u1ij_inv2 + (u1ji - fn3) * u1ji_inv2);
        double exp_delt22 = exppboc2i + exppboc2j;
        double Cf1ij = (-Cf1Aij * P.pboc1[x] * exppboc1i) + (Cf1Bij * exppboc2i) / (exp_delt22);
        double Cf1ji = (-Cf1Aij * P.pboc1[x] * exppboc1j) + (Cf1Bij * exppboc2j) / (exp_delt22);
        double pboc34 = P.pboc3[x] * P.pboc4[x];
        double u45ij = P.pboc5[x] + P.pboc3[x] * r.deltap2[i] - pboc34 * BOpsqr;
        double u45ji = P.pboc5[x] + P.pboc3[x] * r.deltap2[j] - pboc34 * BOpsqr;
        double exph_45ij = std::exp(u45ij), exph_45ji = std::exp(u45ji);
        double exp1 = 1.0 / (1.0 + exph_45ij), exp2 = 1.0 / (1.0 + exph_45ji);
        double exp12 = exp1 * exp2;
        double Cf45ij = -exph_45ij * exp12 * exp1, Cf45ji = -exph_45ji * exp12 * exp2;
        if (P.ovc[x] < 1e-3) { Cf1ij = 0; Cf1ji = 0; }
        if (P.v13cor[x] < 1e-3) { Cf45ij = 0; Cf45ji = 0; }
        double fn45_inv = 1.0 / fn45;
        double Cf1ij_div1 = Cf1ij / fn1, Cf1ji_div1 = Cf1ji / fn1;
        r.A0[a] = fn145;
        r.A1[a] = -2.0 * pboc34 * BOp0 * (Cf45ij + Cf45ji) * fn45_inv;
        r.A2[a] = Cf1ij_div1 + (P.pboc3[x] * Cf45ij * fn45_inv);
        r.A3[a] = r.A2[a] + Cf1ij_div1;
        r.A0[b] = r.A0[a];
        r.A1[b] = r.A1[a];
        r.A2[b] = Cf1ji_div1 + (P.pboc3[x] * Cf45ji * fn45_inv);
        r.A3[b] = r.A2[b] + Cf1ji_div1;
      }
    }
  }
  for (int i = 0; i < n; i++) {
    int ity = nint(r.atype[i]);
    double s = 0;
    for (int j1 = 0; j1 < r.nbrcnt[i]; j1++) s += r.BO[0][SLOT(i, j1)];
    r.delta[i] = -P.Val[ity - 1] + s;
  }
}

// ---------------------------------------------------------------------------------------------
// force helpers, src/pot.F90:1230-1545
inline void addf(Rank &r, int i, double x, double y, double z) {
#pragma omp atomic
  FX(r, i) += x;
#pragma omp atomic
  FY(r, i) += y;
#pragma omp atomic
  FZ(r, i) += z;
}
inline void addc(std::vector<double> &a, int i, double x) {
#pragma omp atomic
  a[i] += x;
}

void ForceB(Rank &r, int i, int j1, int j, int i1, double coeff) {   // :1276-1316
  size_t a = SLOT(i, j1), b = SLOT(j, i1);
  double Cb1 = coeff * (r.A0[a] + r.BO[0][a] * r.A1[a]);
  double d0 = RX(r, i) - RX(r, j), d1 = RY(r, i) - RY(r, j), d2 = RZ(r, i) - RZ(r, j);
  double k = Cb1 * r.dBOp[a];
  addf(r, i, -(k * d0), -(k * d1), -(k * d2));
  addf(r, j, k * d0, k * d1, k * d2);
  addc(r.ccbnd, i, coeff * r.BO[0][a] * r.A2[a]);
  addc(r.ccbnd, j, coeff * r.BO[0][a] * r.A2[b]);
}

void ForceBbo(Rank &r, int i, int j1, int j, int i1, const double *coeff) {   // :1319-1365
  size_t a = SLOT(i, j1), b = SLOT(j, i1);
  double cf[3] = {coeff[0], coeff[1] - coeff[0], coeff[2] - coeff[0]};
  double Cb1 = cf[0] * (r.A0[a] + r.BO[0][a] * r.A1[a]) * r.dBOp[a] +
               cf[1] * r.BO[2][a] * (r.dln_BOp[1][a] + r.A1[a] * r.dBOp[a]) +
               cf[2] * r.BO[3][a] * (r.dln_BOp[2][a] + r.A1[a] * r.dBOp[a]);
  double d0 = RX(r, i) - RX(r, j), d1 = RY(r, i) - RY(r, j), d2 = RZ(r, i) - RZ(r, j);
  addf(r, i, -(Cb1 * d0), -(Cb1 * d1), -(Cb1 * d2));
  addf(r, j, Cb1 * d0, Cb1 * d1, Cb1 * d2);
  double cBO[3] = {cf[0] * r.BO[0][a], cf[1] * r.BO[2][a], cf[2] * r.BO[3][a]};
  addc(r.ccbnd, i, cBO[0] * r.A2[a] + (cBO[1] + cBO[2]) * r.A3[a]);
  addc(r.ccbnd, j, cBO[0] * r.A2[b] + (cBO[1] + cBO[2]) * r.A3[b]);
}

void ForceD(Rank &r, int i, double coeff) {   // :1230-1273 (serial caller)
  for (int j1 = 0; j1 < r.nbrcnt[i]; j1++) {
    int j = r.nbrlist[SLOT(i, j1)];
    int i1 = r.nbrindx[SLOT(i, j1)];
    size_t a = SLOT(i, j1), b = SLOT(j, i1);
    double Cb1 = coeff * (r.A0[a] + r.BO[0][a] * r.A1[a]);
    double d0 = RX(r, i) - RX(r, j), d1 = RY(r, i) - RY(r, j), d2 = RZ(r, i) - RZ(r, j);
    double k = Cb1 * r.dBOp[a];
    FX(r, i) -= k * d0; FY(r, i) -= k * d1; FZ(r, i) -= k * d2;
    FX(r, j) += k * d0; FY(r, j) += k * d1; FZ(r, j) += k * d2;
    r.ccbnd[i] += coeff * r.BO[0][a] * r.A2[a];
    r.ccbnd[j] += coeff * r.BO[0][a] * r.A2[b];
  }
}

// da0 = ri-rj, da1 = rj-rk ; element 0 = norm.  :1462-1521
void ForceA3(Rank &r, double coeff, int i, int j, int k, const double *da0, const double *da1) {
  double C00 = da0[0] * da0[0], C01 = sum3(da0[1] * da1[1], da0[2] * da1[2], da0[3] * da1[3]), C11 = da1[0] * da1[0];
  double CCisqr = 1.0 / (da0[0] * da1[0]);
  double coCC = coeff * CCisqr;
  double Ci1 = -(C01 / C00), Ci2 = 1.0, Ck1 = -1.0, Ck2 = C01 / C11;
  double fij[3], fjk[3];
  for (int c = 0; c < 3; c++) {
    fij[c] = coCC * (Ci1 * da0[c + 1] + Ci2 * da1[c + 1]);
    fjk[c] = -coCC * (Ck1 * da0[c + 1] + Ck2 * da1[c + 1]);
  }
  addf(r, i, fij[0], fij[1], fij[2]);
  addf(r, j, -fij[0] + fjk[0], -fij[1] + fjk[1], -fij[2] + fjk[2]);
  addf(r, k, -fjk[0], -fjk[1], -fjk[2]);
}

// da0 = ri-rj, da1 = rj-rk, da2 = rk-rl.  :1369-1459
void ForceA4(Rank &r, double coeff, int i, int j, int k, int l, const double *da0, const double *da1, const double *da2) {
  double C00 = da0[0] * da0[0], C01 = sum3(da0[1] * da1[1], da0[2] * da1[2], da0[3] * da1[3]),
         C02 = sum3(da0[1] * da2[1], da0[2] * da2[2], da0[3] * da2[3]);
  double C11 = da1[0] * da1[0], C12 = sum3(da1[1] * da2[1], da1[2] * da2[2], da1[3] * da2[3]), C22 = da2[0] * da2[0];
  double D0 = C00 * C11 - C01 * C01;
  double D1 = C11 * C22 - C12 * C12;
  double DDisqr = 1.0 / std::sqrt(D0 * D1);
  double coDD = coeff * DDisqr;
  double com = C01 * C12 - C02 * C11;
  double Cwi[3], Cwj[3], Cwl[3];
  Cwi[0] = C11 / D0 * com;
  Cwi[1] = -(C12 + C01 / D0 * com);
  Cwi[2] = C11;
  Cwj[0] = -(C12 + (C11 + C01) / D0 * com);
  Cwj[1] = -(-C12 - 2 * C02 - C22 / D1 * com - (C00 + C01) / D0 * com);
  Cwj[2] = -(C01 + C11 + C12 / D1 * com);
  Cwl[0] = -C11;
  Cwl[1] = (C01 + C12 / D1 * com);
  Cwl[2] = -(C11 / D1 * com);
  double fij[3], fjk[3], fkl[3];
  for (int c = 0; c < 3; c++) {
    fij[c] = coDD * (Cwi[0] * da0[c + 1] + Cwi[1] * da1[c + 1] + Cwi[2] * da2[c + 1]);
    fjk[c] = coDD * ((Cwj[0] + Cwi[0]) * da0[c + 1] + (Cwj[1] + Cwi[1]) * da1[c + 1] + (Cwj[2] + Cwi[2]) * da2[c + 1]);
    fkl[c] = -coDD * (Cwl[0] * da0[c + 1] + Cwl[1] * da1[c + 1] + Cwl[2] * da2[c + 1]);
  }
  addf(r, i, fij[0], fij[1], fij[2]);
  addf(r, j, -fij[0] + fjk[0], -fij[1] + fjk[1], -fij[2] + fjk[2]);
  addf(r, k, -fjk[0] + fkl[0], -fjk[1] + fkl[1], -fjk[2] + fkl[2]);
  addf(r, l, -fkl[0], -fkl[1], -fkl[2]);
}

void cross_product(const double *dr1, const double *dr2, double *crs) {   // :1524-1543
  double n1[3] = {dr1[1] / dr1[0], dr1[2] / dr1[0], dr1[3] / dr1[0]};
  double n2[3] = {dr2[1] / dr2[0], dr2[2] / dr2[0], dr2[3] / dr2[0]};
  crs[1] = n1[1] * n2[2] - n1[2] * n2[1];
  crs[2] = n1[2] * n2[0] - n1[0] * n2[2];
  crs[3] = n1[0] * n2[1] - n1[1] * n2[0];
  crs[0] = std::sqrt(sum3(crs[1] * crs[1], crs[2] * crs[2], crs[3] * crs[3]));
  if (crs[0] < NSMALL) crs[0] = NSMALL;
}

inline void vec(Rank &r, int a, int b, double *d) {   // d(1:3) = pos(a)-pos(b); d(0) = norm
  d[1] = RX(r, a) - RX(r, b); d[2] = RY(r, a) - RY(r, b); d[3] = RZ(r, a) - RZ(r, b);
  d[0] = std::sqrt(sum3(d[1] * d[1], d[2] * d[2], d[3] * d[3]));
}

// ---------------------------------------------------------------------------------------------
void ENbond(const Params &P, Rank &r) {   // src/pot.F90:676-781
  double pe11 = 0, pe12 = 0, pe13 = 0;
#pragma omp parallel for schedule(guided) reduction(+ : pe11, pe12, pe13)
  for (int i = 0; i < r.natoms; i++) {
    int ity = r.itype[i], iid = r.gtype[i];
    pe13 += CECHRGE * (P.chi[ity - 1] * r.q[i] + 0.5 * P.eta[ity - 1] * r.q[i] * r.q[i]);
    const int *nl = &r.nbplist[(size_t)i * r.W10];
    for (int j1 = 0; j1 < r.nbpcnt[i]; j1++) {
      int j = nl[j1];
      int jid = r.gtype[j];
      if (jid < iid) {
        double d0 = RX(r, i) - RX(r, j), d1 = RY(r, i) - RY(r, j), d2 = RZ(r, i) - RZ(r, j);
        double dr2 = sum3(d0 * d0, d1 * d1, d2 * d2);
        if (dr2 <= P.rctap2) {
          int jty = r.itype[j];
          int inxn = P.inxn2(ity, jty);
          int itb = (int)(dr2 * P.UDRi);
          int itb1 = itb + 1;
          if (inxn <= 0 || itb < 1 || itb1 > P.ntable) continue;   // out of bounds in the reference (Q9)
          double drtb = dr2 - itb * P.UDR;
          drtb = drtb * P.UDRi;
          double drtb1 = 1.0 - drtb;
          double PEvdw = drtb1 * P.evdw(0, itb, inxn) + drtb * P.evdw(0, itb1, inxn);
          double CEvdw = drtb1 * P.evdw(1, itb, inxn) + drtb * P.evdw(1, itb1, inxn);
          double qij = r.q[i] * r.q[j];
          double PEclmb = drtb1 * P.eclmb(0, itb, inxn) + drtb * P.eclmb(0, itb1, inxn);
          PEclmb = PEclmb * qij;
          double CEclmb = drtb1 * P.eclmb(1, itb, inxn) + drtb * P.eclmb(1, itb1, inxn);
          CEclmb = CEclmb * qij;
          pe11 += PEvdw;
          pe12 += PEclmb;
          double c = CEvdw + CEclmb;
          addf(r, i, -(c * d0), -(c * d1), -(c * d2));
          addf(r, j, c * d0, c * d1, c * d2);
        }
      }
    }
  }
  r.PE[11] += pe11; r.PE[12] += pe12; r.PE[13] += pe13;
}

void ENbond_PQEq(const Params &P, Rank &r) {   // src/pot.F90:784-923
  double pe11 = 0, pe12 = 0, pe13 = 0;
#pragma omp parallel for schedule(guided) reduction(+ : pe11, pe12, pe13)
  for (int i = 0; i < r.natoms; i++) {
    int ity = r.itype[i], iid = r.gtype[i];
    double Eshell = 0.0;
    if (P.polarizable(ity)) {
      double dr2 = sum3(SX(r, i) * SX(r, i), SY(r, i) * SY(r, i), SZ(r, i) * SZ(r, i));
      Eshell = 0.5 * P.Kspqeq[ity - 1] * dr2;
    }
    pe13 += CECHRGE * (P.chi[ity - 1] * r.q[i] + 0.5 * P.eta[ity - 1] * (r.q[i] * r.q[i])) + Eshell;
    double qic = r.q[i] + P.Zpqeq[ity - 1];
    const int *nl = &r.nbplist[(size_t)i * r.W10];
    for (int j1 = 0; j1 < r.nbpcnt[i]; j1++) {
      int j = nl[j1];
      int jid = r.gtype[j];
      if (iid < jid) {   // :838 (note: the opposite sense of ENbond's jid<iid)
        double dr[3] = {RX(r, i) - RX(r, j), RY(r, i) - RY(r, j), RZ(r, i) - RZ(r, j)};
        double dr2 = sum3(dr[0] * dr[0], dr[1] * dr[1], dr[2] * dr[2]);
        int jty = r.itype[j];
        int inxn = P.inxn2(ity, jty);
        int itb = (int)(dr2 * P.UDRi);
        int itb1 = itb + 1;
        double drtb = dr2 - itb * P.UDR;
        drtb = drtb * P.UDRi;
        double drtb1 = 1.0 - drtb;
        double PEvdw = 0.0, CEvdw = 0.0;
        if (inxn > 0 && itb >= 1 && itb1 <= P.ntable) {   // out of bounds in the reference otherwise (Q9)
          PEvdw = drtb1 * P.evdw(0, itb, inxn) + drtb * P.evdw(0, itb1, inxn);
          CEvdw = drtb1 * P.evdw(1, itb, inxn) + drtb * P.evdw(1, itb1, inxn);
        }
        double qjc = r.q[j] + P.Zpqeq[jty - 1];
        double qij = qic * qjc;
        double Ecc = 0, Esc = 0, Ecs = 0, Ess = 0;
        double fcc[3] = {0, 0, 0}, fsc[3] = {0, 0, 0}, fcs[3] = {0, 0, 0}, fss[3] = {0, 0, 0};
        get_clmb_pqeq(P, dr, Ecc, P.inxnpqeq(ity, jty), P.TBL_pcc, fcc);
        for (int c = 0; c < 3; c++) fcc[c] = CCLMB0 * qij * fcc[c];
        Ecc = CCLMB0 * Ecc * qij;
        if (P.polarizable(ity)) {
          double d[3] = {dr[0] + SX(r, i), dr[1] + SY(r, i), dr[2] + SZ(r, i)};
          get_clmb_pqeq(P, d, Esc, P.inxnpqeq(ity, jty), P.TBL_psc, fsc);
          for (int c = 0; c < 3; c++) fsc[c] = -CCLMB0 * P.Zpqeq[ity - 1] * qjc * fsc[c];
          Esc = -CCLMB0 * Esc * P.Zpqeq[ity - 1] * qjc;
        }
        if (P.polarizable(jty)) {
          double d[3] = {dr[0] - SX(r, j), dr[1] - SY(r, j), dr[2] - SZ(r, j)};
          get_clmb_pqeq(P, d, Ecs, P.inxnpqeq(jty, ity), P.TBL_psc, fcs);
          for (int c = 0; c < 3; c++) fcs[c] = -CCLMB0 * P.Zpqeq[jty - 1] * qic * fcs[c];
          Ecs = -CCLMB0 * Ecs * qic * P.Zpqeq[jty - 1];
        }
        if (P.polarizable(ity) && P.polarizable(jty)) {
          double d[3] = {dr[0] + SX(r, i) - SX(r, j), dr[1] + SY(r, i) - SY(r, j), dr[2] + SZ(r, i) - SZ(r, j)};
          get_clmb_pqeq(P, d, Ess, P.inxnpqeq(ity, jty), P.TBL_pss, fss);
          for (int c = 0; c < 3; c++) fss[c] = CCLMB0 * P.Zpqeq[ity - 1] * P.Zpqeq[jty - 1] * fss[c];
          Ess = CCLMB0 * Ess * P.Zpqeq[ity - 1] * P.Zpqeq[jty - 1];
        }
        double PEclmb = Ecc + Esc + Ecs + Ess;
        pe11 += PEvdw;
        pe12 += PEclmb;
        double ff[3];
        for (int c = 0; c < 3; c++) ff[c] = CEvdw * dr[c] + fcc[c] + fcs[c] + fsc[c] + fss[c];
        addf(r, i, -ff[0], -ff[1], -ff[2]);
        addf(r, j, ff[0], ff[1], ff[2]);
      }
    }
  }
  r.PE[11] += pe11; r.PE[12] += pe12; r.PE[13] += pe13;
}

void EEfield(const Params &P, Rank &r) {   // src/module.F90:359-383 (the energy is "to be determined" there: none added)
  for (int i = 0; i < r.natoms; i++) {
    int ity = nint(r.atype[i]);
    double qic = r.q[i] + P.Zpqeq[ity - 1];
    double Eforce = -qic * P.cfg.eFieldStrength * EEV_KCAL;
    r.f[(size_t)(P.cfg.eFieldDir - 1) * r.NB + i] += Eforce;
  }
}

void Ebond(const Params &P, Rank &r) {   // src/pot.F90:926-977
  double pe1 = 0;
#pragma omp parallel for schedule(static) reduction(+ : pe1)
  for (int i = 0; i < r.natoms; i++) {
    int ity = r.itype[i], iid = r.gtype[i];
    for (int j1 = 0; j1 < r.nbrcnt[i]; j1++) {
      int j = r.nbrlist[SLOT(i, j1)];
      if (r.gtype[j] < iid) {
        int x = P.inxn2(ity, r.itype[j]) - 1;
        size_t a = SLOT(i, j1);
        double bp = std::pow(r.BO[1][a], P.pbe2[x]);
        double exp_be12 = std::exp(P.pbe1[x] * (1.0 - bp));
        double PEbo = -P.Desig[x] * r.BO[1][a] * exp_be12 - P.Depi[x] * r.BO[2][a] - P.Depipi[x] * r.BO[3][a];
        pe1 += PEbo;
        double CEbo = -P.Desig[x] * exp_be12 * (1.0 - P.pbe1[x] * P.pbe2[x] * bp);
        double coeff[3] = {CEbo, -P.Depi[x], -P.Depipi[x]};
        ForceBbo(r, i, j1, j, r.nbrindx[a], coeff);
      }
    }
  }
  r.PE[1] += pe1;
}

void Elnpr(const Params &P, Rank &r, bool main_loop) {   // src/pot.F90:148-316
  const int n = r.copyptr[6];
  for (int i = 0; i < n; i++) {   // preparation :183-209
    int ity = r.itype[i];
    if (ity == 0) continue;
    int t = ity - 1;
    double deltaE = -P.Vale[t] + P.Val[t] + r.delta[i];
    double dEh = deltaE * 0.5;
    int idEh = (int)dEh;   // is_idEh = 1
    double u = 2.0 + deltaE - 2 * idEh;
    double explp1 = std::exp(-P.plp1[t] * (u * u));
    double Clp = 2.0 * P.plp1[t] * explp1 * u;
    r.dDlp[i] = Clp;
    r.nlp[i] = explp1 - (double)idEh;
    r.deltalp[i] = P.nlpopt[t] - r.nlp[i];
    if (P.mass[t] > 21.0) r.deltalp[i] = 0.0;
  }
  if (!main_loop) return;
  double pe2 = 0, pe3 = 0, pe4 = 0;
#pragma omp parallel for schedule(static) reduction(+ : pe2, pe3, pe4)
  for (int i = 0; i < r.natoms; i++) {
    int ity = r.itype[i];
    int t = ity - 1;
    double sum_ovun1 = 0, sum_ovun2 = 0;
    for (int j1 = 0; j1 < r.nbrcnt[i]; j1++) {
      int j = r.nbrlist[SLOT(i, j1)];
      int x = P.inxn2(ity, r.itype[j]) - 1;
      size_t a = SLOT(i, j1);
      sum_ovun1 = sum_ovun1 + P.povun1[x] * P.Desig[x] * r.BO[0][a];
      sum_ovun2 = sum_ovun2 + (r.delta[j] - r.deltalp[j]) * (r.BO[2][a] + r.BO[3][a]);
    }
    double expvd2 = std::exp(-75.0 * r.deltalp[i]);
    double dElp = P.plp2[t] * ((1.0 + expvd2) + 75.0 * r.deltalp[i] * expvd2) / ((1.0 + expvd2) * (1.0 + expvd2));
    double expovun1 = P.povun3[t] * std::exp(P.povun4[t] * sum_ovun2);
    double deltalpcorr = r.delta[i] - r.deltalp[i] / (1.0 + expovun1);
    double expovun2 = std::exp(P.povun2[t] * deltalpcorr);
    double DlpV_i = 1.0 / (deltalpcorr + P.Val[t] + 1e-8);
    double expovun2n = 1.0 / expovun2;
    double expovun6 = std::exp(P.povun6[t] * deltalpcorr);
    double expovun8 = P.povun7[t] * std::exp(P.povun8[t] * sum_ovun2);
    double div_expovun1 = 1.0 / (1.0 + expovun1);
    double div_expovun2 = 1.0 / (1.0 + expovun2);
    double div_expovun2n = 1.0 / (1.0 + expovun2n);
    double div_expovun8 = 1.0 / (1.0 + expovun8);
    double PElp = P.plp2[t] * r.deltalp[i] / (1.0 + expvd2);
    double PEover = sum_ovun1 * DlpV_i * deltalpcorr * div_expovun2;
    double PEunder = -P.povun5[t] * (1.0 - expovun6) * div_expovun2n * div_expovun8;
    pe2 += PElp; pe3 += PEover; pe4 += PEunder;
    double CElp1 = dElp * r.dDlp[i];
    double CEover[8], CEunder[7];
    CEover[1] = deltalpcorr * DlpV_i * div_expovun2;
    CEover[2] = sum_ovun1 * DlpV_i * div_expovun2 *
                (1.0 - deltalpcorr * DlpV_i - P.povun2[t] * deltalpcorr * div_expovun2n);
    CEover[3] = CEover[2] * (1.0 - r.dDlp[i] * div_expovun1);
    CEover[4] = CEover[2] * r.deltalp[i] * P.povun4[t] * expovun1 * (div_expovun1 * div_expovun1);
    CEunder[1] = (P.povun5[t] * P.povun6[t] * expovun6 * div_expovun8 + PEunder * P.povun2[t] * expovun2n) * div_expovun2n;
    CEunder[2] = -PEunder * P.povun8[t] * expovun8 * div_expovun8;
    CEunder[3] = CEunder[1] * (1.0 - r.dDlp[i] * div_expovun1);
    CEunder[4] = CEunder[1] * r.deltalp[i] * P.povun4[t] * expovun1 * (div_expovun1 * div_expovun1) + CEunder[2];
    for (int j1 = 0; j1 < r.nbrcnt[i]; j1++) {
      int j = r.nbrlist[SLOT(i, j1)];
      int x = P.inxn2(ity, r.itype[j]) - 1;
      size_t a = SLOT(i, j1);
      double bpp = r.BO[2][a] + r.BO[3][a];
      CEover[5] = CEover[1] * P.povun1[x] * P.Desig[x];
      CEover[6] = CEover[4] * (1.0 - r.dDlp[j]) * bpp;
      CEover[7] = CEover[4] * (r.delta[j] - r.deltalp[j]);
      CEunder[5] = CEunder[4] * (1.0 - r.dDlp[j]) * bpp;
      CEunder[6] = CEunder[4] * (r.delta[j] - r.deltalp[j]);
      double CElp_b = CElp1 + CEover[3] + CEover[5] + CEunder[3];
      double CElp_bpp = CEover[7] + CEunder[6];
      double coeff[3] = {CElp_b + 0.0, CElp_b + CElp_bpp, CElp_b + CElp_bpp};
      ForceBbo(r, i, j1, j, r.nbrindx[a], coeff);
      addc(r.cdbnd, j, CEover[6] + CEunder[5]);
    }
  }
  r.PE[2] += pe2; r.PE[3] += pe3; r.PE[4] += pe4;
}

void E3b(const Params &P, Rank &r) {   // src/pot.F90:319-557
  double pe5 = 0, pe6 = 0, pe7 = 0;
#pragma omp parallel for schedule(guided) reduction(+ : pe5, pe6, pe7)
  for (int j = 0; j < r.natoms; j++) {
    int jty = r.itype[j];
    int tj = jty - 1;
    double sum_BO8 = 0, sum_SBO1 = 0;
    for (int n1 = 0; n1 < r.nbrcnt[j]; n1++) {
      size_t a = SLOT(j, n1);
      sum_BO8 = sum_BO8 - std::pow(r.BO[0][a], 8.0);
      sum_SBO1 = sum_SBO1 + r.BO[2][a] + r.BO[3][a];
    }
    double prod_SBO = std::exp(sum_BO8);
    double delta_ang = r.delta[j] + P.Val[tj] - P.Valangle[tj];
    for (int i1 = 0; i1 < r.nbrcnt[j] - 1; i1++) {
      double BOij = r.BO[0][SLOT(j, i1)] - CUTOF2_ESUB;
      if (!(BOij > 0.0)) continue;
      int i = r.nbrlist[SLOT(j, i1)];
      int ity = r.itype[i];
      double rij[4];
      vec(r, i, j, rij);
      for (int k1 = i1 + 1; k1 < r.nbrcnt[j]; k1++) {
        double BOjk = r.BO[0][SLOT(j, k1)] - CUTOF2_ESUB;
        if (!(BOjk > 0.0)) continue;
        if (!(r.BO[0][SLOT(j, i1)] * r.BO[0][SLOT(j, k1)] > CUTOF2_ESUB)) continue;
        int k = r.nbrlist[SLOT(j, k1)];
        int kty = r.itype[k];
        double rjk[4];
        vec(r, j, k, rjk);
        double cos_ijk = -sum3(rij[1] * rjk[1], rij[2] * rjk[2], rij[3] * rjk[3]) / (rij[0] * rjk[0]);
        if (cos_ijk > MAXANGLE) cos_ijk = MAXANGLE;
        if (cos_ijk < MINANGLE) cos_ijk = MINANGLE;
        double theta_ijk = std::acos(cos_ijk);
        double sin_ijk = std::sin(theta_ijk);
        int inxn = P.inxn3(ity, jty, kty);
        if (inxn == 0) continue;
        int x = inxn - 1;
        // --- PEval
        double BOij_p4 = std::pow(BOij, P.pval4[x]);
        double exp3ij = std::exp(-P.pval3[tj] * BOij_p4);
        double fn7ij = 1.0 - exp3ij;
        double BOjk_p4 = std::pow(BOjk, P.pval4[x]);
        double exp3jk = std::exp(-P.pval3[tj] * BOjk_p4);
        double fn7jk = 1.0 - exp3jk;
        double exp6 = std::exp(P.pval6[x] * delta_ang);
        double exp7 = std::exp(-P.pval7[x] * delta_ang);
        double trm8 = 1.0 + exp6 + exp7;
        double fn8j = P.pval5[tj] - (P.pval5[tj] - 1.0) * (2.0 + exp6) / trm8;
        double SBO = sum_SBO1 + (1.0 - prod_SBO) * (-delta_ang - P.pval8[x] * r.nlp[j]);
        double SBO2 = 0.0;
        if (SBO <= 0) SBO2 = 0.0;
        if (SBO > 0) SBO2 = std::pow(SBO, P.pval9[x]);
        if (SBO > 1) SBO2 = 2.0 - std::pow(2.0 - SBO, P.pval9[x]);
        if (SBO > 2) SBO2 = 2.0;
        double theta0 = PI_RX - P.theta00[x] * (1.0 - std::exp(-P.pval10[x] * (2.0 - SBO2)));
        double theta_diff = theta0 - theta_ijk;
        double exp2 = std::exp(-P.pval2[x] * theta_diff * theta_diff);
        double PEval = fn7ij * fn7jk * fn8j * (P.pval1[x] - P.pval1[x] * exp2);
        double Cf7ij = P.pval3[tj] * P.pval4[x] * std::pow(BOij, P.pval4[x] - 1.0) * exp3ij;
        double Cf7jk = P.pval3[tj] * P.pval4[x] * std::pow(BOjk, P.pval4[x] - 1.0) * exp3jk;
        double Cf8j = (1.0 - P.pval5[tj]) / (trm8 * trm8) *
                      (P.pval6[x] * exp6 * trm8 - (2.0 + exp6) * (P.pval6[x] * exp6 - P.pval7[x] * exp7));
        double Ctheta0 = P.pval10[x] * P.theta00[x] * std::exp(-P.pval10[x] * (2.0 - SBO2));
        double CSBO2 = 0.0;
        if ((SBO <= 0) || (SBO > 2)) CSBO2 = 0.0;
        if ((SBO > 0) && (SBO <= 1)) CSBO2 = P.pval9[x] * std::pow(SBO, P.pval9[x] - 1.0);
        if ((SBO > 1) && (SBO <= 2)) CSBO2 = P.pval9[x] * std::pow(2.0 - SBO, P.pval9[x] - 1.0);
        double dSBO1 = -8.0 * prod_SBO * (delta_ang + P.pval8[x] * r.nlp[j]);
        double dSBO2 = (prod_SBO - 1.0) * (1.0 - P.pval8[x] * r.dDlp[j]);
        double CEval[9];
        CEval[1] = Cf7ij * fn7jk * fn8j * P.pval1[x] * (1.0 - exp2);
        CEval[2] = fn7ij * Cf7jk * fn8j * P.pval1[x] * (1.0 - exp2);
        CEval[3] = fn7ij * fn7jk * Cf8j * P.pval1[x] * (1.0 - exp2);
        CEval[4] = 2.0 * P.pval1[x] * P.pval2[x] * fn7ij * fn7jk * fn8j * exp2 * theta_diff;
        CEval[5] = CEval[4] * Ctheta0 * CSBO2;
        CEval[6] = CEval[5] * dSBO1;
        CEval[7] = CEval[5] * dSBO2;
        CEval[8] = CEval[4] / sin_ijk;
        // --- PEpen
        double exp_pen3 = std::exp(-P.ppen3[x] * r.delta[j]);
        double exp_pen4 = std::exp(P.ppen4[x] * r.delta[j]);
        double fn9 = (2.0 + exp_pen3) / (1.0 + exp_pen3 + exp_pen4);
        double exp_pen2ij = std::exp(-P.ppen2[x] * (BOij - 2.0) * (BOij - 2.0));
        double exp_pen2jk = std::exp(-P.ppen2[x] * (BOjk - 2.0) * (BOjk - 2.0));
        double PEpen = P.ppen1[x] * fn9 * exp_pen2ij * exp_pen2jk;
        double trm_pen34 = 1.0 + exp_pen3 + exp_pen4;
        double Cf9j = (-P.ppen3[x] * exp_pen3 * trm_pen34 -
                       (2.0 + exp_pen3) * (-P.ppen3[x] * exp_pen3 + P.ppen4[x] * exp_pen4)) / (trm_pen34 * trm_pen34);
        double CEpen[4];
        CEpen[1] = Cf9j / fn9;
        CEpen[2] = -2.0 * P.ppen2[x] * (BOij - 2.0);
        CEpen[3] = -2.0 * P.ppen2[x] * (BOjk - 2.0);
        for (int c = 1; c <= 3; c++) CEpen[c] = CEpen[c] * PEpen;
        // --- PEcoa
        int ti = ity - 1, tk = kty - 1;
        double sum_BOi = r.delta[i] + P.Val[ti];
        double sum_BOk = r.delta[k] + P.Val[tk];
        double delta_val = r.delta[j] + P.Val[tj] - P.Valval[tj];
        double exp_coa2 = std::exp(P.pcoa2[x] * delta_val);
        double ui = -BOij + sum_BOi, uk = -BOjk + sum_BOk;
        double exp_coa3i = std::exp(-P.pcoa3[x] * (ui * ui));
        double exp_coa3k = std::exp(-P.pcoa3[x] * (uk * uk));
        double exp_coa4i = std::exp(-P.pcoa4[x] * ((BOij - 1.5) * (BOij - 1.5)));
        double exp_coa4k = std::exp(-P.pcoa4[x] * ((BOjk - 1.5) * (BOjk - 1.5)));
        double PEcoa = P.pcoa1[x] / (1.0 + exp_coa2) * exp_coa3i * exp_coa3k * exp_coa4i * exp_coa4k;
        double CEcoa[6];
        CEcoa[1] = -2.0 * P.pcoa4[x] * (BOij - 1.5);
        CEcoa[2] = -2.0 * P.pcoa4[x] * (BOjk - 1.5);
        CEcoa[3] = -P.pcoa2[x] * exp_coa2 / (1.0 + exp_coa2);
        CEcoa[4] = -2.0 * P.pcoa3[x] * ui;
        CEcoa[5] = -2.0 * P.pcoa3[x] * uk;
        for (int c = 1; c <= 5; c++) CEcoa[c] = CEcoa[c] * PEcoa;
        pe5 += PEval; pe6 += PEpen; pe7 += PEcoa;
        double CE3body_b1 = CEpen[2] + CEcoa[1] - CEcoa[4] + CEval[1];
        double CE3body_b2 = CEpen[3] + CEcoa[2] - CEcoa[5] + CEval[2];
        double CE3body_d1 = CEpen[1] + CEcoa[3] + CEval[3] + CEval[7];
        double CE3body_d2 = CEcoa[4], CE3body_d3 = CEcoa[5];
        double CE3body_a = CEval[8];
        int j1 = r.nbrindx[SLOT(j, i1)];
        ForceB(r, i, j1, j, i1, CE3body_b1);
        j1 = r.nbrindx[SLOT(j, k1)];
        ForceB(r, j, k1, k, j1, CE3body_b2);
        for (int n1 = 0; n1 < r.nbrcnt[j]; n1++) {
          double c0 = CE3body_d1 + CEval[6] * std::pow(r.BO[0][SLOT(j, n1)], 7);
          double coeff[3] = {c0 + 0.0, c0 + CEval[5], c0 + CEval[5]};
          int n = r.nbrlist[SLOT(j, n1)];
          ForceBbo(r, j, n1, n, r.nbrindx[SLOT(j, n1)], coeff);
        }
        addc(r.cdbnd, i, CE3body_d2);
        addc(r.cdbnd, k, CE3body_d3);
        ForceA3(r, CE3body_a, i, j, k, rij, rjk);
      }
    }
  }
  r.PE[5] += pe5; r.PE[6] += pe6; r.PE[7] += pe7;
}

void Ehb(const Params &P, Rank &r) {   // src/pot.F90:559-673
  double pe10 = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : pe10)
  for (int i = 0; i < r.natoms; i++) {
    int ity = r.itype[i];
    for (int j1 = 0; j1 < r.nbrcnt[i]; j1++) {
      int j = r.nbrlist[SLOT(i, j1)];
      int jty = r.itype[j];
      size_t a = SLOT(i, j1);
      if (!((jty == 2) && (r.BO[0][a] > MINBO0))) continue;   // hydrogen hard-coded as type 2 (Q4)
      const int *nl = &r.nbplist[(size_t)i * r.W10];
      for (int kk = 0; kk < r.nbpcnt[i]; kk++) {
        int k = nl[kk];
        int kty = r.itype[k];
        int inxnhb = P.inxn3hb(ity, jty, kty);
        if (!((j != k) && (i != k) && (inxnhb != 0))) continue;
        double rik[3] = {RX(r, i) - RX(r, k), RY(r, i) - RY(r, k), RZ(r, i) - RZ(r, k)};
        double rik2 = sum3(rik[0] * rik[0], rik[1] * rik[1], rik[2] * rik[2]);
        if (!(rik2 < RCHB2)) continue;
        int x = inxnhb - 1;
        double rjk[4], rij[4];
        vec(r, j, k, rjk);
        vec(r, i, j, rij);
        double cos_ijk = -sum3(rij[1] * rjk[1], rij[2] * rjk[2], rij[3] * rjk[3]) / (rij[0] * rjk[0]);
        if (cos_ijk > MAXANGLE) cos_ijk = MAXANGLE;
        if (cos_ijk < MINANGLE) cos_ijk = MINANGLE;
        double theta_ijk = std::acos(cos_ijk);
        double sin_ijk_half = std::sin(0.5 * theta_ijk);
        double s2 = sin_ijk_half * sin_ijk_half;
        double sin_xhz4 = s2 * s2;
        double cos_xhz1 = (1.0 - cos_ijk);
        double exp_hb2 = std::exp(-P.phb2[x] * r.BO[0][a]);
        double exp_hb3 = std::exp(-P.phb3[x] * (P.r0hb[x] / rjk[0] + rjk[0] / P.r0hb[x] - 2.0));
        double PEhb = P.phb1[x] * (1.0 - exp_hb2) * exp_hb3 * sin_xhz4;
        pe10 += PEhb;
        double CEhb1 = P.phb1[x] * P.phb2[x] * exp_hb2 * exp_hb3 * sin_xhz4;
        double CEhb2 = -0.5 * P.phb1[x] * (1.0 - exp_hb2) * exp_hb3 * cos_xhz1;
        double CEhb3 = -PEhb * P.phb3[x] * (-P.r0hb[x] / (rjk[0] * rjk[0]) + 1.0 / P.r0hb[x]) * (1.0 / rjk[0]);
        ForceB(r, i, j1, j, r.nbrindx[a], CEhb1);
        ForceA3(r, CEhb2, i, j, k, rij, rjk);
        double ff[3] = {CEhb3 * rjk[1], CEhb3 * rjk[2], CEhb3 * rjk[3]};
        addf(r, j, -ff[0], -ff[1], -ff[2]);
        addf(r, k, ff[0], ff[1], ff[2]);
      }
    }
  }
  r.PE[10] += pe10;
}

void E4b(const Params &P, Rank &r) {   // src/pot.F90:980-1227
  double pe8 = 0, pe9 = 0;
#pragma omp parallel for schedule(guided) reduction(+ : pe8, pe9)
  for (int j = 0; j < r.natoms; j++) {
    int jty = r.itype[j];
    double delta_ang_j = r.delta[j] + P.Val[jty - 1] - P.Valangle[jty - 1];
    int jid = r.gtype[j];
    for (int k1 = 0; k1 < r.nbrcnt[j]; k1++) {
      double BOjk0 = r.BO[0][SLOT(j, k1)];
      double BOjk = BOjk0 - CUTOF2_ESUB;
      if (!(BOjk0 > CUTOF2_ESUB)) continue;
      int k = r.nbrlist[SLOT(j, k1)];
      int kid = r.gtype[k];
      if (!(jid < kid)) continue;
      int kty = r.itype[k];
      double delta_ang_k = r.delta[k] + P.Val[kty - 1] - P.Valangle[kty - 1];
      double delta_ang_jk = delta_ang_j + delta_ang_k;
      double rjk[4];
      vec(r, j, k, rjk);
      for (int i1 = 0; i1 < r.nbrcnt[j]; i1++) {
        double BOij0 = r.BO[0][SLOT(j, i1)];
        double BOij = BOij0 - CUTOF2_ESUB;
        if (!((BOij0 > CUTOF2_ESUB) && ((BOij0 * BOjk0) > CUTOF2_ESUB))) continue;
        int i = r.nbrlist[SLOT(j, i1)];
        if (i == k) continue;
        int ity = r.itype[i];
        double rij[4];
        vec(r, i, j, rij);
        double cos_ijk = -sum3(rij[1] * rjk[1], rij[2] * rjk[2], rij[3] * rjk[3]) / (rij[0] * rjk[0]);
        if (cos_ijk > MAXANGLE) cos_ijk = MAXANGLE;
        if (cos_ijk < MINANGLE) cos_ijk = MINANGLE;
        double theta_ijk = std::acos(cos_ijk);
        double sin_ijk = std::sin(theta_ijk);
        double tan_ijk_i = 1.0 / std::tan(theta_ijk);
        double crs_ijk[4];
        cross_product(rij, rjk, crs_ijk);
        for (int l1 = 0; l1 < r.nbrcnt[k]; l1++) {
          double BOkl0 = r.BO[0][SLOT(k, l1)];
          double BOkl = BOkl0 - CUTOF2_ESUB;
          if (!((BOkl0 > CUTOF2_ESUB) && (BOjk0 * BOkl0 > CUTOF2_ESUB))) continue;
          int l = r.nbrlist[SLOT(k, l1)];
          int lty = r.itype[l];
          int inxn = P.inxn4(ity, jty, kty, lty);
          if (!((inxn != 0) && (i != l) && (j != l))) continue;
          if (!((BOij0 * (BOjk0 * BOjk0) * BOkl0) > MINBO0)) continue;
          int x = inxn - 1;
          double rkl[4];
          vec(r, k, l, rkl);
          double exp_tor2[3] = {std::exp(-P.ptor2[x] * BOij), std::exp(-P.ptor2[x] * BOjk), std::exp(-P.ptor2[x] * BOkl)};
          double exp_tor3 = std::exp(-P.ptor3[x] * delta_ang_jk);
          double exp_tor4 = std::exp(P.ptor4[x] * delta_ang_jk);
          double exp_tor34_i = 1.0 / (1.0 + exp_tor3 + exp_tor4);
          double fn10 = (1.0 - exp_tor2[0]) * (1.0 - exp_tor2[1]) * (1.0 - exp_tor2[2]);
          double fn11 = (2.0 + exp_tor3) / (1.0 + exp_tor3 + exp_tor4);
          double fn12 = std::exp(-P.pcot2[x] * ((BOij - 1.5) * (BOij - 1.5) + (BOjk - 1.5) * (BOjk - 1.5) +
                                                (BOkl - 1.5) * (BOkl - 1.5)));
          double btb2 = 2.0 - r.BO[2][SLOT(j, k1)] - fn11;
          double exp_tor1 = std::exp(P.ptor1[x] * (btb2 * btb2));
          double cos_jkl = -sum3(rjk[1] * rkl[1], rjk[2] * rkl[2], rjk[3] * rkl[3]) / (rjk[0] * rkl[0]);
          if (cos_jkl > MAXANGLE) cos_jkl = MAXANGLE;
          if (cos_jkl < MINANGLE) cos_jkl = MINANGLE;
          double theta_jkl = std::acos(cos_jkl);
          double sin_jkl = std::sin(theta_jkl);
          double tan_jkl_i = 1.0 / std::tan(theta_jkl);
          double crs_jkl[4];
          cross_product(rjk, rkl, crs_jkl);
          double cos_ijkl[4];
          cos_ijkl[1] = sum3(crs_ijk[1] * crs_jkl[1], crs_ijk[2] * crs_jkl[2], crs_ijk[3] * crs_jkl[3]) / (crs_ijk[0] * crs_jkl[0]);
          if (cos_ijkl[1] > MAXANGLE) cos_ijkl[1] = MAXANGLE;
          if (cos_ijkl[1] < MINANGLE) cos_ijkl[1] = MINANGLE;
          double omega_ijkl = std::acos(cos_ijkl[1]);
          double cos_ijkl_sqr = cos_ijkl[1] * cos_ijkl[1];
          double cos_2ijkl = std::cos(2.0 * omega_ijkl);
          cos_ijkl[2] = 1.0 - cos_2ijkl;
          cos_ijkl[3] = 1.0 + std::cos(3.0 * omega_ijkl);
          double Vsum = P.V1[x] * (1.0 + cos_ijkl[1]) + P.V2[x] * exp_tor1 * cos_ijkl[2] + P.V3[x] * cos_ijkl[3];
          double PEtors = 0.5 * fn10 * sin_ijk * sin_jkl * Vsum;
          double PEconj = P.pcot1[x] * fn12 * (1.0 + (cos_ijkl_sqr - 1.0) * sin_ijk * sin_jkl);
          pe8 += PEtors; pe9 += PEconj;
          double CEtors[10];
          CEtors[1] = 0.5 * sin_ijk * sin_jkl * Vsum;
          CEtors[2] = -P.ptor1[x] * fn10 * sin_ijk * sin_jkl * P.V2[x] * exp_tor1 * btb2 * cos_ijkl[2];
          double dfn11 = (-P.ptor3[x] * exp_tor3 +
                          (P.ptor3[x] * exp_tor3 - P.ptor4[x] * exp_tor4) * (2.0 + exp_tor3) * exp_tor34_i) * exp_tor34_i;
          CEtors[3] = CEtors[2] * dfn11;
          CEtors[4] = CEtors[1] * P.ptor2[x] * exp_tor2[0] * (1.0 - exp_tor2[1]) * (1.0 - exp_tor2[2]);
          CEtors[5] = CEtors[1] * P.ptor2[x] * (1.0 - exp_tor2[0]) * exp_tor2[1] * (1.0 - exp_tor2[2]);
          CEtors[6] = CEtors[1] * P.ptor2[x] * (1.0 - exp_tor2[0]) * (1.0 - exp_tor2[1]) * exp_tor2[2];
          double cmn = -0.5 * fn10 * Vsum;
          CEtors[7] = cmn * sin_jkl * tan_ijk_i;
          CEtors[8] = cmn * sin_ijk * tan_jkl_i;
          CEtors[9] = fn10 * sin_ijk * sin_jkl *
                      (0.5 * P.V1[x] - 2.0 * P.V2[x] * exp_tor1 * cos_ijkl[1] + 1.5 * P.V3[x] * (cos_2ijkl + 2.0 * cos_ijkl_sqr));
          double Cconj = -2.0 * P.pcot2[x] * PEconj;
          double CEconj[7];
          CEconj[1] = Cconj * (BOij - 1.5);
          CEconj[2] = Cconj * (BOjk - 1.5);
          CEconj[3] = Cconj * (BOkl - 1.5);
          CEconj[4] = -P.pcot1[x] * fn12 * (cos_ijkl_sqr - 1.0) * tan_ijk_i * sin_jkl;
          CEconj[5] = -P.pcot1[x] * fn12 * (cos_ijkl_sqr - 1.0) * sin_ijk * tan_jkl_i;
          CEconj[6] = 2.0 * P.pcot1[x] * fn12 * cos_ijkl[1] * sin_ijk * sin_jkl;
          double C4body_b[3] = {CEconj[1] + CEtors[4], CEconj[2] + CEtors[5], CEconj[3] + CEtors[6]};
          double C4body_a[3] = {CEconj[4] + CEtors[7], CEconj[5] + CEtors[8], CEconj[6] + CEtors[9]};
          addc(r.cdbnd, j, CEtors[3]);
          addc(r.cdbnd, k, CEtors[3]);
          ForceB(r, i, r.nbrindx[SLOT(j, i1)], j, i1, C4body_b[0]);
          double C4body_b_jk[3] = {C4body_b[1] + 0.0, C4body_b[1] + CEtors[2], C4body_b[1] + 0.0};
          ForceBbo(r, j, k1, k, r.nbrindx[SLOT(j, k1)], C4body_b_jk);
          ForceB(r, k, l1, l, r.nbrindx[SLOT(k, l1)], C4body_b[2]);
          ForceA3(r, C4body_a[0], i, j, k, rij, rjk);
          ForceA3(r, C4body_a[1], j, k, l, rjk, rkl);
          ForceA4(r, C4body_a[2], i, j, k, l, rij, rjk, rkl);
        }
      }
    }
  }
  r.PE[8] += pe8; r.PE[9] += pe9;
}

// `corrected` (diagnostic only, never used for parity): run every ForceD before any ccbnd is consumed, so no
// contribution is discarded; the result is the exact gradient and can be checked by finite differences.
void ForceBondedTerms(Rank &r, bool corrected) {   // src/pot.F90:113-144, serial and order dependent
  if (corrected)
    for (int i = 0; i < r.copyptr[6]; i++) ForceD(r, i, r.cdbnd[i]);
  for (int i = 0; i < r.copyptr[6]; i++) {
    if (!corrected) ForceD(r, i, r.cdbnd[i]);
    for (int j1 = 0; j1 < r.nbrcnt[i]; j1++) {
      int j = r.nbrlist[SLOT(i, j1)];
      double k = r.ccbnd[i] * r.dBOp[SLOT(i, j1)];
      double d0 = RX(r, i) - RX(r, j), d1 = RY(r, i) - RY(r, j), d2 = RZ(r, i) - RZ(r, j);
      FX(r, i) -= k * d0; FY(r, i) -= k * d1; FZ(r, i) -= k * d2;
      FX(r, j) += k * d0; FY(r, j) += k * d1; FZ(r, j) += k * d2;
    }
    r.ccbnd[i] = 0.0;
  }
}

int FORCE(World &w) {   // src/pot.F90:2-90
  const Params &P = w.P;
  for (auto &r : w.R) {
    std::fill(r.ccbnd.begin(), r.ccbnd.end(), 0.0);
    std::fill(r.cdbnd.begin(), r.cdbnd.end(), 0.0);
    std::fill(r.f.begin(), r.f.end(), 0.0);
    for (int c = 0; c < 14; c++) r.PE[c] = 0.0;
  }
  double dr[3];
  for (int c = 0; c < 3; c++) dr[c] = P.cfg.nmincell * w.R[0].box.lcsize[c];
  int rc = COPYATOMS(w, MODE_COPY, dr);
  if (rc) return rc;
  for (auto &r : w.R) {
    if ((rc = LINKEDLIST(w, r, r.g, r.box.lcsize))) return rc;
    if ((rc = LINKEDLIST(w, r, r.nbg, r.box.nblcsize))) return rc;
    if ((rc = NEIGHBORLIST(w, r, P.cfg.nmincell))) return rc;
    if ((rc = PairList(w, r, false))) return rc;
    for (int i = 0; i < r.copyptr[6]; i++) { r.itype[i] = nint(r.atype[i]); r.gtype[i] = l2g(r.atype[i]); }
    BOPRIM(P, r);
    BOFULL(P, r);
    const int tm = w.term_mask;   // diagnostic term selection; parity runs use all terms (0x3f)
    if (tm & 1) { if (P.cfg.isPQEq) ENbond_PQEq(P, r); else ENbond(P, r); }
    if (tm & 2) Ebond(P, r);
    Elnpr(P, r, (tm & 4) != 0);   // the preparation loop feeds E3b and always runs
    if (tm & 8) Ehb(P, r);
    if (tm & 16) E3b(P, r);
    if (tm & 32) E4b(P, r);
    if (P.cfg.isEfield && P.cfg.isPQEq) EEfield(P, r);   // src/pot.F90:61 (reads Zpqeq: only meaningful with PQEq)
    ForceBondedTerms(r, w.corrected);
    for (int i = 0; i < r.copyptr[6]; i++) {   // :65-72
      r.astr[0] += RX(r, i) * FX(r, i); r.astr[1] += RY(r, i) * FY(r, i); r.astr[2] += RZ(r, i) * FZ(r, i);
      r.astr[3] += RY(r, i) * FZ(r, i); r.astr[4] += RZ(r, i) * FX(r, i); r.astr[5] += RX(r, i) * FY(r, i);
    }
  }
  double zero[3] = {0, 0, 0};
  return COPYATOMS(w, MODE_CPBK, zero);
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}   // namespace

// =================================================================================================
extern "C" {

int orc_create(const rxg_config *cfg, const rxg_ff *ff, const rxg_box *boxes, int nranks, orc_world *out) {
  World *w = new World();
  Params &P = w->P;
  P.cfg = *cfg;
  P.nso = ff->nso; P.nboty = ff->nboty; P.nvaty = ff->nvaty; P.ntoty = ff->ntoty; P.nhbty = ff->nhbty; P.ntable = ff->ntable;
  P.vpar1 = ff->vpar1; P.vpar2 = ff->vpar2; P.cutoff_vpar30 = ff->cutoff_vpar30;
  P.rctap = ff->rctap; P.rctap2 = ff->rctap2; P.UDR = ff->UDR; P.UDRi = ff->UDRi;
  const int ns = P.nso, nb = P.nboty, nv = P.nvaty, nt = P.ntoty, nh = P.nhbty;
#define C1(name, n) cpy(P.name, ff->name, n)
  C1(Val, ns); C1(Valval, ns); C1(Valangle, ns); C1(Vale, ns); C1(mass, ns); C1(plp1, ns); C1(plp2, ns); C1(nlpopt, ns);
  C1(povun2, ns); C1(povun3, ns); C1(povun4, ns); C1(povun5, ns); C1(povun6, ns); C1(povun7, ns); C1(povun8, ns);
  C1(pval3, ns); C1(pval5, ns); C1(chi, ns); C1(eta, ns);
  C1(cBOp1, nb); C1(cBOp3, nb); C1(cBOp5, nb); C1(pbo2h, nb); C1(pbo4h, nb); C1(pbo6h, nb); C1(pbo2, nb); C1(pbo4, nb);
  C1(pbo6, nb); C1(swtch, 3 * nb); C1(rc2, nb); C1(pboc1, nb); C1(pboc3, nb); C1(pboc4, nb); C1(pboc5, nb); C1(ovc, nb);
  C1(v13cor, nb); C1(Desig, nb); C1(Depi, nb); C1(Depipi, nb); C1(pbe1, nb); C1(pbe2, nb); C1(povun1, nb);
  C1(theta00, nv); C1(pval1, nv); C1(pval2, nv); C1(pval4, nv); C1(pval6, nv); C1(pval7, nv); C1(pval8, nv); C1(pval9, nv);
  C1(pval10, nv); C1(ppen1, nv); C1(ppen2, nv); C1(ppen3, nv); C1(ppen4, nv); C1(pcoa1, nv); C1(pcoa2, nv); C1(pcoa3, nv);
  C1(pcoa4, nv);
  C1(ptor1, nt); C1(ptor2, nt); C1(ptor3, nt); C1(ptor4, nt); C1(V1, nt); C1(V2, nt); C1(V3, nt); C1(pcot1, nt); C1(pcot2, nt);
  C1(phb1, nh); C1(phb2, nh); C1(phb3, nh); C1(r0hb, nh);
#undef C1
  cpyi(P.inxn2v, ff->inxn2, (size_t)ns * ns);
  cpyi(P.inxn3v, ff->inxn3, (size_t)ns * ns * ns);
  cpyi(P.inxn3hbv, ff->inxn3hb, (size_t)ns * ns * ns);
  cpyi(P.inxn4v, ff->inxn4, (size_t)ns * ns * ns * ns);
  cpy(P.TBL_Evdw, ff->TBL_Evdw, (size_t)2 * P.ntable * nb);
  cpy(P.TBL_Eclmb, ff->TBL_Eclmb, (size_t)2 * P.ntable * nb);
  cpy(P.TBL_Eclmb_QEq, ff->TBL_Eclmb_QEq, (size_t)P.ntable * nb);
  if (cfg->isPQEq) {
    if (ff->ntype_pqeq < 1 || !ff->isPolarizable || !ff->TBL_Eclmb_pcc) { delete w; return RXG_ERR_ARG; }
    const size_t np = ff->ntype_pqeq, nt2 = np * np * (size_t)P.ntable * 2;
    P.ntype_pqeq = (int)np;
    cpyi(P.isPolarizable, ff->isPolarizable, np); cpyi(P.inxnpqeqv, ff->inxnpqeq, np * np);
    cpy(P.Zpqeq, ff->Zpqeq, np); cpy(P.Kspqeq, ff->Kspqeq, np);
    cpy(P.TBL_pcc, ff->TBL_Eclmb_pcc, nt2); cpy(P.TBL_psc, ff->TBL_Eclmb_psc, nt2); cpy(P.TBL_pss, ff->TBL_Eclmb_pss, nt2);
  }
  w->R.resize(nranks);
  for (int ir = 0; ir < nranks; ir++) {
    Rank &r = w->R[ir];
    r.box = boxes[ir];
    r.nbmesh.assign(boxes[ir].nbmesh, boxes[ir].nbmesh + 3 * (size_t)boxes[ir].nbnmesh);
    r.box.nbmesh = nullptr;
    r.NB = cfg->nbuffer; r.MAXN = cfg->maxneighbs; r.W10 = cfg->maxneighbs10;
    size_t nb_ = r.NB;
    for (auto *p : {&r.atype, &r.q, &r.qs, &r.qt, &r.gs, &r.gt, &r.hs, &r.ht, &r.qsfp, &r.qsfv, &r.frcindx, &r.delta,
                    &r.deltap1, &r.deltap2, &r.nlp, &r.dDlp, &r.deltalp, &r.ccbnd, &r.cdbnd})
      p->assign(nb_, 0.0);
    for (auto *p : {&r.pos, &r.v, &r.f, &r.spos}) p->assign(3 * nb_, 0.0);
    r.fpqeq.assign(nb_, 0.0);
    r.g.setup(r.box.cc, MAXLAYERS, r.NB);
    r.nbg.setup(r.box.nbcc, MAXLAYERS_NB, r.NB);
    r.nbrcnt.assign(nb_, 0);
    r.nbrlist.assign(nb_ * r.MAXN, 0);
    r.nbrindx.assign(nb_ * r.MAXN, 0);
    for (int c = 0; c < 4; c++) r.BO[c].assign(nb_ * r.MAXN, 0.0);
    for (int c = 0; c < 3; c++) r.dln_BOp[c].assign(nb_ * r.MAXN, 0.0);
    for (auto *p : {&r.dBOp, &r.A0, &r.A1, &r.A2, &r.A3}) p->assign(nb_ * r.MAXN, 0.0);
    r.itype.assign(nb_, 0); r.gtype.assign(nb_, 0);
    for (int c = 0; c < 14; c++) r.PE[c] = 0;
    for (int c = 0; c < 6; c++) r.astr[c] = 0;
  }
  *out = w;
  return 0;
}

int orc_set_corrected(orc_world h, int on) { ((World *)h)->corrected = on != 0; return 0; }
int orc_set_terms(orc_world h, int mask) { ((World *)h)->term_mask = mask; return 0; }
int orc_set_pqeq_stale(orc_world h, int on) { ((World *)h)->pqeq_stale = on != 0; return 0; }
int orc_destroy(orc_world h) { delete (World *)h; return 0; }
const char *orc_last_error(orc_world h) { return ((World *)h)->err.c_str(); }

int orc_set_atoms(orc_world h, int rank, int natoms, const double *atype, const double *pos, const double *v,
                  const double *q, const double *qsfp, const double *qsfv) {
  World *w = (World *)h;
  Rank &r = w->R[rank];
  if (natoms > r.NB) { w->err = "natoms > nbuffer"; return RXG_ERR_ARG; }
  r.natoms = natoms;
  for (int c = 0; c < 7; c++) r.copyptr[c] = natoms;
  // pair-list rows exist for residents only (the reference sizes them NBUFFER wide but fills resident rows)
  size_t rows = (size_t)natoms + (natoms / 8) + 64;
  if (rows > (size_t)r.NB) rows = r.NB;
  r.nbpcnt.assign(r.NB, 0);
  r.nbplist.assign(rows * r.W10, 0);
  r.hessian.assign(rows * r.W10, 0.0);
  for (int i = 0; i < natoms; i++) {
    r.atype[i] = atype[i];
    for (int c = 0; c < 3; c++) {
      r.pos[(size_t)c * r.NB + i] = pos[(size_t)c * natoms + i];
      r.v[(size_t)c * r.NB + i] = v ? v[(size_t)c * natoms + i] : 0.0;
    }
    r.q[i] = q ? q[i] : 0.0;
    r.qsfp[i] = qsfp ? qsfp[i] : 0.0;
    r.qsfv[i] = qsfv ? qsfv[i] : 0.0;
  }
  return 0;
}

int orc_natoms(orc_world h, int rank) { return ((World *)h)->R[rank].natoms; }

int orc_set_spos(orc_world h, int rank, const double *spos) {   // compact [3][natoms]
  World *w = (World *)h;
  Rank &r = w->R[rank];
  for (int c = 0; c < 3; c++)
    for (int i = 0; i < r.natoms; i++) r.spos[(size_t)c * r.NB + i] = spos[(size_t)c * r.natoms + i];
  return 0;
}

static int check_rows(World *w) {
  for (auto &r : w->R)
    if ((size_t)r.natoms * r.W10 > r.nbplist.size()) {
      size_t rows = (size_t)r.natoms + r.natoms / 8 + 64;
      r.nbplist.assign(rows * r.W10, 0);
      r.hessian.assign(rows * r.W10, 0.0);
    }
  return 0;
}

int orc_qeq(orc_world h) {
  World *w = (World *)h;
  double t0 = now();
  check_rows(w);
  int rc = w->P.cfg.isPQEq ? PQEq(*w) : QEq(*w);   // src/main.F90:27-31,78-82
  w->t_qeq += now() - t0;
  return rc;
}
int orc_force(orc_world h) {
  World *w = (World *)h;
  double t0 = now();
  check_rows(w);
  int rc = FORCE(*w);
  w->t_force += now() - t0;
  return rc;
}
int orc_move(orc_world h) {
  World *w = (World *)h;
  double t0 = now();
  double zero[3] = {0, 0, 0};
  int rc = COPYATOMS(*w, MODE_MOVE, zero);
  w->t_move += now() - t0;
  return rc;
}

// main loop body, src/main.F90:64-98 (mdmode 1: no thermostat); vkick src/main.F90:192-207
int orc_md_run(orc_world h, int nsteps, double dt, int qstep, double Lex_w2, int step0) {
  World *w = (World *)h;
  const Params &P = w->P;
  for (int nstep = step0; nstep < step0 + nsteps; nstep++) {
    for (auto &r : w->R) {
      for (int i = 0; i < r.natoms; i++) {
        double dthm = dt * 0.5 / P.mass[nint(r.atype[i]) - 1];
        for (int c = 0; c < 3; c++) r.v[(size_t)c * r.NB + i] = r.v[(size_t)c * r.NB + i] + 1.0 * dthm * r.f[(size_t)c * r.NB + i];
      }
      for (int i = 0; i < r.natoms; i++) r.qsfv[i] = r.qsfv[i] + 0.5 * dt * Lex_w2 * (r.q[i] - r.qsfp[i]);
      for (int i = 0; i < r.natoms; i++) r.qsfp[i] = r.qsfp[i] + dt * r.qsfv[i];
    }
    if (P.cfg.isEfield) {   // "always correct the linear momentum when electric field is applied", src/main.F90:70-71; LinearMomentum :773-803
      double mm = 0, vcm[3] = {0, 0, 0};
      for (auto &r : w->R)
        for (int i = 0; i < r.natoms; i++) {
          double m = P.mass[nint(r.atype[i]) - 1];
          for (int c = 0; c < 3; c++) vcm[c] = vcm[c] + m * r.v[(size_t)c * r.NB + i];
          mm = mm + m;
        }
      for (int c = 0; c < 3; c++) vcm[c] = vcm[c] / mm;
      for (auto &r : w->R)
        for (int c = 0; c < 3; c++)
          for (int i = 0; i < r.natoms; i++) r.v[(size_t)c * r.NB + i] = r.v[(size_t)c * r.NB + i] - vcm[c];
    }
    for (auto &r : w->R)
      for (int c = 0; c < 3; c++)
        for (int i = 0; i < r.natoms; i++) r.pos[(size_t)c * r.NB + i] = r.pos[(size_t)c * r.NB + i] + dt * r.v[(size_t)c * r.NB + i];
    int rc = orc_move(h);
    if (rc) return rc;
    if (nstep % qstep == 0 && (rc = orc_qeq(h))) return rc;
    if ((rc = orc_force(h))) return rc;
    for (auto &r : w->R) {
      for (int i = 0; i < r.natoms; i++) {
        double m = P.mass[nint(r.atype[i]) - 1];
        double vx = r.v[i], vy = r.v[(size_t)r.NB + i], vz = r.v[2 * (size_t)r.NB + i];
        r.astr[0] += vx * vx * m; r.astr[1] += vy * vy * m; r.astr[2] += vz * vz * m;
        r.astr[3] += vy * vz * m; r.astr[4] += vz * vx * m; r.astr[5] += vx * vy * m;
      }
      for (int i = 0; i < r.natoms; i++) {
        double dthm = dt * 0.5 / P.mass[nint(r.atype[i]) - 1];
        for (int c = 0; c < 3; c++) r.v[(size_t)c * r.NB + i] = r.v[(size_t)c * r.NB + i] + 1.0 * dthm * r.f[(size_t)c * r.NB + i];
      }
      for (int i = 0; i < r.natoms; i++) r.qsfv[i] = r.qsfv[i] + 0.5 * dt * Lex_w2 * (r.q[i] - r.qsfp[i]);
    }
  }
  return 0;
}

static long long put(const std::vector<double> &v, size_t n, double *out, long long cap) {
  if (out) { if ((long long)n > cap) return -1; memcpy(out, v.data(), n * sizeof(double)); }
  return (long long)n;
}
// per-atom 3-vectors are returned compact: [3][n]
static long long put3(const std::vector<double> &v, size_t NB, size_t n, double *out, long long cap) {
  if (out) {
    if ((long long)(3 * n) > cap) return -1;
    for (int c = 0; c < 3; c++) memcpy(out + c * n, v.data() + c * NB, n * sizeof(double));
  }
  return (long long)(3 * n);
}

long long orc_get_f64(orc_world h, int rank, const char *name, double *out, long long cap) {
  World *w = (World *)h;
  Rank &r = w->R[rank];
  std::string s(name);
  size_t n = r.copyptr[6], ns = (size_t)r.copyptr[6] * r.MAXN;
#define G1(nm) if (s == #nm) return put(r.nm, n, out, cap)
  G1(atype); G1(q); G1(qs); G1(qt); G1(gs); G1(gt); G1(hs); G1(ht); G1(qsfp); G1(qsfv); G1(frcindx); G1(delta);
  G1(deltap1); G1(deltap2); G1(nlp); G1(dDlp); G1(deltalp); G1(ccbnd); G1(cdbnd);
#undef G1
#define GS(nm) if (s == #nm) return put(r.nm, ns, out, cap)
  GS(dBOp); GS(A0); GS(A1); GS(A2); GS(A3);
#undef GS
  if (s == "pos") return put3(r.pos, r.NB, n, out, cap);
  if (s == "v") return put3(r.v, r.NB, n, out, cap);
  if (s == "f") return put3(r.f, r.NB, n, out, cap);
  if (s == "spos") return put3(r.spos, r.NB, n, out, cap);
  if (s == "fpqeq") return put(r.fpqeq, r.natoms, out, cap);
  if (s == "BO0") return put(r.BO[0], ns, out, cap);
  if (s == "BO1") return put(r.BO[1], ns, out, cap);
  if (s == "BO2") return put(r.BO[2], ns, out, cap);
  if (s == "BO3") return put(r.BO[3], ns, out, cap);
  if (s == "dln_BOp1") return put(r.dln_BOp[0], ns, out, cap);
  if (s == "dln_BOp2") return put(r.dln_BOp[1], ns, out, cap);
  if (s == "dln_BOp3") return put(r.dln_BOp[2], ns, out, cap);
  if (s == "hessian") return put(r.hessian, (size_t)r.natoms * r.W10, out, cap);
  if (s == "PE") { if (out) { if (cap < 14) return -1; memcpy(out, r.PE, 14 * 8); } return 14; }
  if (s == "astr") { if (out) { if (cap < 6) return -1; memcpy(out, r.astr, 6 * 8); } return 6; }
  return -1;
}

long long orc_get_i32(orc_world h, int rank, const char *name, int *out, long long cap) {
  World *w = (World *)h;
  Rank &r = w->R[rank];
  std::string s(name);
  auto puti = [&](const int *p, size_t n) -> long long {
    if (out) { if ((long long)n > cap) return -1; memcpy(out, p, n * sizeof(int)); }
    return (long long)n;
  };
  if (s == "copyptr") return puti(r.copyptr, 7);
  if (s == "nbrcnt") return puti(r.nbrcnt.data(), r.copyptr[6]);
  if (s == "nbrlist") return puti(r.nbrlist.data(), (size_t)r.copyptr[6] * r.MAXN);
  if (s == "nbrindx") return puti(r.nbrindx.data(), (size_t)r.copyptr[6] * r.MAXN);
  if (s == "nbpcnt") return puti(r.nbpcnt.data(), r.natoms);
  if (s == "nbplist") return puti(r.nbplist.data(), (size_t)r.natoms * r.W10);
  if (s == "nstep_qeq") return puti(&r.nstep_qeq, 1);
  if (s == "natoms") return puti(&r.natoms, 1);
  if (s == "pqeq_skips") { int k = (int)r.pqeq_skips; return puti(&k, 1); }
  return -1;
}

int orc_observe(orc_world h, double *PE, double *KE, double *qsum, int *nstep_qeq) {
  World *w = (World *)h;
  const Params &P = w->P;
  double pe[14] = {0}, ke = 0, qq = 0;
  for (auto &r : w->R) {
    r.PE[0] = 0;
    for (int c = 1; c < 14; c++) r.PE[0] += r.PE[c];      // PRINTE, src/main.F90:232
    for (int c = 0; c < 14; c++) pe[c] += r.PE[c];
    for (int i = 0; i < r.natoms; i++) {
      double hm = 0.5 * P.mass[nint(r.atype[i]) - 1];
      double vx = r.v[i], vy = r.v[(size_t)r.NB + i], vz = r.v[2 * (size_t)r.NB + i];
      ke += hm * sum3(vx * vx, vy * vy, vz * vz);
      qq += r.q[i];
    }
  }
  if (PE) memcpy(PE, pe, sizeof(pe));
  if (KE) *KE = ke;
  if (qsum) *qsum = qq;
  if (nstep_qeq) *nstep_qeq = w->R[0].nstep_qeq;
  return 0;
}

// OpenMP team size of the oracle's loops.  A launcher may export OMP_NUM_THREADS=1 (torch.distributed.run does), and the
// runtime reads the environment only once, so bench.py sets the team size explicitly and reports what the runtime then uses.
int orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

int orc_timers(orc_world h, double *a, double *b, double *c) {
  World *w = (World *)h;
  if (a) *a = w->t_qeq;
  if (b) *b = w->t_force;
  if (c) *c = w->t_move;
  return 0;
}

}   // extern "C"
