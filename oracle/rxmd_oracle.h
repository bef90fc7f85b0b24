/*
 * rxmd_oracle.h -- C API of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * The oracle is a loop-for-loop CPU restatement of the reference's per-timestep hot path
 * (QEq + FORCE + COPYATOMS, reference src/qeq.F90, src/pot.F90, src/bo.F90, src/comm.F90,
 * src/main.F90:277-477).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product library (rxmd_b200/csrc) never does.
 *
 * It simulates all ranks of a `vprocs` decomposition inside one process (messages are memcpy
 * between simulated ranks), so multi-GPU runs can be compared against the oracle run with the
 * same decomposition.
 */
#ifndef RXMD_ORACLE_H
#define RXMD_ORACLE_H
#include "../include/rxmd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef void *orc_world;

/* one rxg_box per rank (boxes[0..nranks)); cfg/ff shared by all ranks */
int orc_create(const rxg_config *cfg, const rxg_ff *ff, const rxg_box *boxes, int nranks, orc_world *out);
int orc_destroy(orc_world w);
/* diagnostic: 1 = keep the ccbnd contributions the reference discards; forces become the exact
 * gradient (finite-difference checkable).  Parity always uses 0 (literal). */
int orc_set_corrected(orc_world w, int on);
/* diagnostic: bit mask of energy terms FORCE evaluates: 1 ENbond, 2 Ebond, 4 Elnpr, 8 Ehb, 16 E3b, 32 E4b (default all) */
int orc_set_terms(orc_world w, int mask);
/* diagnostic: 1 = where get_coulomb_and_dcoulomb_pqeq returns early (src/module.F90:402) keep the output variable's previous
 * value, as a serial build of the reference does (src/pqeq.F90:219-231,340-343), instead of the zero contribution that parity
 * uses.  Quantifies the size of that deviation (tests/test_oracle_pqeq.py). */
int orc_set_pqeq_stale(orc_world w, int on);
const char *orc_last_error(orc_world w);

/* resident state of one rank; pos is pos(NBUFFER,3) compact: double[3*n] = x[n] y[n] z[n] (REAL coordinates) */
int orc_set_atoms(orc_world w, int rank, int natoms, const double *atype, const double *pos, const double *v,
                  const double *q, const double *qsfp, const double *qsfv);
int orc_natoms(orc_world w, int rank);
/* PQEq shell displacements spos(NBUFFER,3) of the residents, compact double[3*natoms] (zero after orc_set_atoms) */
int orc_set_spos(orc_world w, int rank, const double *spos);

/* the reference entry points, executed on every simulated rank */
int orc_qeq(orc_world w);                    /* subroutine QEq src/qeq.F90:2, or PQEq src/pqeq.F90:2 when cfg.isPQEq */
int orc_force(orc_world w);                  /* subroutine FORCE src/pot.F90:2   */
int orc_move(orc_world w);                   /* COPYATOMS(MODE_MOVE) src/main.F90:75 */
/* nsteps iterations of the main loop body src/main.F90:64-98 (mdmode 1); call orc_qeq+orc_force first */
int orc_md_run(orc_world w, int nsteps, double dt, int qstep, double Lex_w2, int step0);

/* fetch a per-rank array; returns element count (or -1). out may be NULL to query the count.
 * double names: atype q pos v f qs qt gs gt hs ht qsfp qsfv hessian BO(4 planes) dBOp dln_BOp A0 A1 A2 A3 delta
 *               deltap1 deltap2 nlp dDlp deltalp ccbnd cdbnd PE(14) astr(6) frcindx spos fpqeq
 * int names   : copyptr(7) nbrcnt nbrlist nbrindx nbpcnt nbplist nstep_qeq natoms pqeq_skips */
long long orc_get_f64(orc_world w, int rank, const char *name, double *out, long long cap);
long long orc_get_i32(orc_world w, int rank, const char *name, int *out, long long cap);
/* global sums over ranks: PE(0:13), KE, sum q */
int orc_observe(orc_world w, double *PE, double *KE, double *qsum, int *nstep_qeq);
/* wall-clock seconds spent inside orc_qeq / orc_force / orc_move since creation */
int orc_timers(orc_world w, double *t_qeq, double *t_force, double *t_move);
/* sets the OpenMP team size (n > 0) and returns the size the runtime will use */
int orc_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
