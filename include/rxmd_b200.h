/*
 * rxmd_b200.h -- C ABI of the B200-native ReaxFF + QEq hot path behind RXMD's entry points.
 *
 * The reference (USCCACS/RXMD) has no FFI: the per-timestep hot path sits behind four external
 * Fortran subroutines that share module-global state.  Each function below names the reference
 * interface it replaces (file:line under the reference's src/).  A Fortran host binds them with
 * `bind(C)` interfaces (see rxmd_b200/gpu_shim.F90 and INTEGRATION.md); the Python harness binds
 * the same symbols through ctypes.  Only plain pointers, ints and doubles cross the boundary.
 *
 * Array conventions (identical to the Fortran host's memory):
 *   - per-atom 1-D arrays: double[nbuffer]            (atype, q, qsfp, qsfv, ...)
 *   - per-atom 3-vectors : double[3*nbuffer], x[0..nbuffer) y[..] z[..]   == pos(NBUFFER,3)
 *   - parameter arrays   : 1-based Fortran arrays passed by their first element, column-major
 *   - residents are elements 0..natoms-1 (Fortran 1..NATOMS)
 *   - `pos` is in/out in QEq/FORCE exactly as in the reference: COPYATOMS maps positions to normalised
 *     coordinates and back on every call (src/comm.F90:222-227,260-264), which perturbs them by ~1 ulp
 * All calls are synchronous at return (the host reads `f` right after FORCE, src/main.F90:86-97).
 * One host thread per handle; a handle owns one CUDA device (one MPI rank == one GPU).
 */
#ifndef RXMD_B200_H
#define RXMD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* status codes; 1..3 map 1:1 onto the reference's three overflow traps */
#define RXG_OK                 0
#define RXG_ERR_MAXNEIGHBS     1   /* src/main.F90:403  "overflow of max # in neighbor list"      */
#define RXG_ERR_MAXNEIGHBS10   2   /* src/qeq.F90:248   "nbplist greater then MAXNEIGHBS10"       */
#define RXG_ERR_NBUFFER        3   /* src/comm.F90:467  "over capacity in append_atoms"           */
#define RXG_ERR_CUDA           4
#define RXG_ERR_NCCL           5
#define RXG_ERR_ARG            6
#define RXG_ERR_STATE          7

typedef void *rxg_handle;

/* Run-time configuration: the reference's compile-time capacities (src/module.F90:80-84) and the
 * rxmd.in / command-line switches the hot path reads (src/cmdline.F90:255-297). */
typedef struct rxg_config {
  int device;         /* CUDA device ordinal for this rank                                   */
  int nbuffer;        /* NBUFFER: capacity of every per-atom array (residents + ghosts)      */
  int maxneighbs;     /* MAXNEIGHBS   (30)   bonded-list row width                           */
  int maxneighbs10;   /* MAXNEIGHBS10 (1500) trap threshold for 10 A rows                    */
  int nmincell;       /* NMINCELL (4): FORCE halo depth in bonded cells                      */
  int isQEq;          /* 0 skip, 1 CG, 2 extended Lagrangian (one CG step)                   */
  int NMAXQEq;        /* max CG iterations                                                   */
  int isPQEq;         /* 1: PQEq / ENbond_PQEq variants                                      */
  int isEfield;       /* 1: EEfield add-on (src/module.F90:359)                              */
  int eFieldDir;      /* 1..3                                                                */
  double QEq_tol;     /* CG stop tolerance                                                   */
  double Lex_fqs;     /* extended-Lagrangian mixing (src/qeq.F90:53)                         */
  double eFieldStrength;
} rxg_config;

/* Flattened `module parameters` (src/module.F90:620-723) + derived tables, exactly as the host
 * computed them in GETPARAMS (src/param.F90), CUTOFFLENGTH and POTENTIALTABLE (src/init.F90). */
typedef struct rxg_ff {
  int nso, nboty, nvaty, ntoty, nhbty, ntable;
  double vpar1, vpar2, cutoff_vpar30;
  double rctap, rctap2, UDR, UDRi;
  /* per atom type, [nso] */
  const double *Val, *Valval, *Valangle, *Vale, *mass, *plp1, *plp2, *nlpopt;
  const double *povun2, *povun3, *povun4, *povun5, *povun6, *povun7, *povun8;
  const double *pval3, *pval5, *chi, *eta;
  /* per bond type, [nboty]; swtch is switch(1:3,nboty) */
  const double *cBOp1, *cBOp3, *cBOp5, *pbo2h, *pbo4h, *pbo6h, *pbo2, *pbo4, *pbo6, *swtch;
  const double *rc2, *pboc1, *pboc3, *pboc4, *pboc5, *ovc, *v13cor;
  const double *Desig, *Depi, *Depipi, *pbe1, *pbe2, *povun1;
  /* per valence-angle type, [nvaty] */
  const double *theta00, *pval1, *pval2, *pval4, *pval6, *pval7, *pval8, *pval9, *pval10;
  const double *ppen1, *ppen2, *ppen3, *ppen4, *pcoa1, *pcoa2, *pcoa3, *pcoa4;
  /* per torsion type, [ntoty] */
  const double *ptor1, *ptor2, *ptor3, *ptor4, *V1, *V2, *V3, *pcot1, *pcot2;
  /* per hydrogen-bond type, [nhbty] */
  const double *phb1, *phb2, *phb3, *r0hb;
  /* lookup tables, column-major: inxn2(nso,nso) inxn3(nso,nso,nso) inxn3hb(..) inxn4(nso^4) */
  const int *inxn2, *inxn3, *inxn3hb, *inxn4;
  /* r^2-space tables: TBL_Evdw(0:1,NTABLE,nboty), TBL_Eclmb(0:1,NTABLE,nboty), TBL_Eclmb_QEq(NTABLE,nboty) */
  const double *TBL_Evdw, *TBL_Eclmb, *TBL_Eclmb_QEq;
  /* PQEq (module pqeq_vars, src/module.F90:336-615); NULL / 0 when isPQEq == 0 */
  int ntype_pqeq;
  const int *isPolarizable;          /* [ntype_pqeq] 0/1 */
  const double *Zpqeq, *Kspqeq;      /* [ntype_pqeq] */
  const int *inxnpqeq;               /* (ntype_pqeq,ntype_pqeq) */
  const double *TBL_Eclmb_pcc, *TBL_Eclmb_psc, *TBL_Eclmb_pss; /* (ntype_pqeq^2,NTABLE,0:1) */
} rxg_ff;

/* Box, cell grids and rank topology (src/init.F90:74-100,525-607,636-668). */
typedef struct rxg_box {
  double HH[9];        /* HH(3,3,0) column-major                                             */
  double HHi[9];       /* matinv(HH), column-major                                           */
  double lata, latb, latc;
  double LBOX[3];      /* 1/vprocs                                                           */
  double OBOX[3];      /* LBOX*vID                                                           */
  double lcsize[3];    /* bonded cell size, normalised                                       */
  double nblcsize[3];  /* non-bonded cell size, normalised                                   */
  int cc[3];           /* bonded cells per domain                                            */
  int nbcc[3];         /* non-bonded cells per domain                                        */
  int nbnmesh;         /* stencil size                                                       */
  int vprocs[3];
  int vID[3];
  int myparity[3];
  int target_node[6];  /* +x,-x,+y,-y,+z,-z neighbour ranks                                  */
  int myid, nprocs;
  const int *nbmesh;   /* nbmesh(3,nbnmesh) column-major                                     */
} rxg_box;

/* ---- life cycle ------------------------------------------------------------------------- */
/* replaces the allocation part of INITSYSTEM (src/init.F90:110-201) for the device mirrors   */
int rxg_create(const rxg_config *cfg, rxg_handle *out);
/* replaces reading `module parameters` + TBL_* globals inside FORCE/QEq (src/pot.F90:3, src/qeq.F90:3) */
int rxg_set_forcefield(rxg_handle h, const rxg_ff *ff);
/* replaces reading HH/HHi/LBOX/OBOX/cc/lcsize/nbcc/nblcsize/nbmesh/target_node globals       */
int rxg_set_box(rxg_handle h, const rxg_box *box);
/* communicator: nccl_unique_id is the 128-byte ncclUniqueId broadcast by the host (MPI_Bcast in
 * the Fortran shim, TCPStore in the harness).  Not needed when nprocs == 1.
 * Replaces MPI_SEND/MPI_RECV/MPI_ALLREDUCE inside COPYATOMS/QEq (src/comm.F90:291-364, src/qeq.F90:107-144,357). */
int rxg_comm_init(rxg_handle h, int rank, int nranks, const void *nccl_unique_id);
/* bit 0: the per-iteration ghost refreshes go through peer memory (cudaIpc windows written over NVLink by the neighbours'
 * kernels) rather than ncclSend/ncclRecv; bit 1: the CG's scalar all-reduces go through the windows too (summed in rank
 * order).  Decided collectively in rxg_comm_init; RXG_PEER_HALO=0 / RXG_PEER_ALLREDUCE=0 force NCCL. */
int rxg_comm_peer_halo(rxg_handle h);
/* rank 0 creates the 128-byte id and the host broadcasts it (MPI_Bcast / TCPStore) before rxg_comm_init */
int rxg_comm_unique_id(void *out128);
int rxg_destroy(rxg_handle h);
const char *rxg_last_error(rxg_handle h);

/* ---- optional promises of the host about the NEXT entry-point call (cleared by that call) ----------------
 * The reference's main loop calls COPYATOMS(MODE_MOVE), QEq, FORCE back to back on the same arrays (src/main.F90:75-84) and
 * touches none of them in between.  A shim that knows this may say so, and the library then skips the PCIe copies that
 * would only move bytes it already holds.  Without hints every call is literal: all inputs up, all outputs down.
 *   RXG_HINT_ATOMS_ON_DEVICE  atype/pos (and natoms) of the next call are exactly what the previous call left on the device
 *                             (the host did not write them since): do not upload them; rxg_force then also reuses the halo
 *                             and 10 A list of the rxg_qeq before it without re-verifying the atoms (RXG_FUSE_API=1)
 *   RXG_HINT_Q_ON_DEVICE      q of the next call is what the previous rxg_qeq / rxg_pqeq returned: do not upload it
 *   RXG_HINT_DEFER_POS        the next call need not copy pos back (its only change is the ulp-level normalise/de-normalise
 *                             round trip of COPYATOMS): the following call is hinted ATOMS_ON_DEVICE and a later
 *                             un-deferred call (rxg_force) returns the final positions.  Ignored by rxg_move when atoms migrate.
 *   RXG_HINT_CHARGES_STAY     (rxg_move) q, qs and qt of this call are what the library itself last produced and the host reads
 *                             none of them before the next rxg_qeq / rxg_pqeq returns q: they migrate with their atoms on the
 *                             device (src/comm.F90:164-171) but cross PCIe in neither direction -- qs/qt are QEq's own CG
 *                             vectors, which no host loop of the reference touches (src/main.F90:64-98)
 * A promise that does not hold gives wrong results; hints are dropped when natoms differs from the device's. */
#define RXG_HINT_ATOMS_ON_DEVICE 1
#define RXG_HINT_Q_ON_DEVICE     2
#define RXG_HINT_DEFER_POS       4
#define RXG_HINT_CHARGES_STAY    8
int rxg_hint(rxg_handle h, int flags);

/* ---- the drop-in entry points (host buffers in, host buffers out) -------------------------- */
/* subroutine QEq(atype,pos,q)   src/qeq.F90:2     (also writes qsfp,qsfv when isQEq==1, :42-43) */
int rxg_qeq(rxg_handle h, const int *natoms, const double *atype, double *pos, double *q,
            double *qsfp, double *qsfv, int *nstep_qeq);
/* subroutine PQEq(atype,pos,q)  src/pqeq.F90:2    (spos = shell displacements spos(NBUFFER,3), in/out: relaxed one capped
 * step after the CG, :171,187-259).  Needs a handle created with isPQEq = 1 and the PQEq members of rxg_ff. */
int rxg_pqeq(rxg_handle h, const int *natoms, const double *atype, double *pos, double *q,
             double *spos, double *qsfp, double *qsfv, int *nstep_qeq);
/* module-global spos(NBUFFER,3) (src/module.F90:286) is shared state of PQEq, FORCE (ENbond_PQEq, src/pot.F90:784) and
 * COPYATOMS(MODE_MOVE) (src/comm.F90:153,165-167).  The device keeps its own copy: rxg_pqeq refreshes it in both directions;
 * a host that changes spos elsewhere (restart, MODE_MOVE through rxg_move) brackets the call with these two.
 * rxg_pqeq_skips: how often get_coulomb_and_dcoulomb_pqeq returned early where the reference then reads an unassigned
 * variable (src/pqeq.F90:219-231,340-343); this library adds nothing for those pairs. */
int rxg_spos_upload(rxg_handle h, int natoms, const double *spos);
int rxg_spos_download(rxg_handle h, int natoms, double *spos);
long long rxg_pqeq_skips(rxg_handle h);
/* subroutine FORCE(atype,pos,f,q) src/pot.F90:2   PE(0:13) overwritten, astr(1:6) incremented (:65-72) */
int rxg_force(rxg_handle h, const int *natoms, const double *atype, double *pos, double *f,
              const double *q, double *PE, double *astr);
/* call COPYATOMS(MODE_MOVE,[0,0,0],atype,pos,v,f,q)  src/main.F90:75, src/comm.F90:2,238-256
 * natoms is in/out; every listed array is compacted exactly like the reference's finalize(). */
int rxg_move(rxg_handle h, int *natoms, double *atype, double *pos, double *v, double *q,
             double *qs, double *qt, double *qsfp, double *qsfv);
/* WriteBND's inputs (src/fileio.F90:56-121): nbrlist(NBUFFER,0:MAXNEIGHBS) and BO(0,:,:) as
 * double[nbuffer*maxneighbs] (atom index fastest), valid after rxg_force */
int rxg_fetch_bonds(rxg_handle h, int *nbrlist, double *BO0);
/* 30 doubles of accumulated timings (the analogue of it_timer(1:30), src/module.F90:215-217), milliseconds unless noted:
 * [0] rxg_qeq  [1] rxg_force  [2] rxg_move (device part, CUDA events)   [3] rxg_md_run total (CUDA events)
 * [4] QEq inside md_run  [5] FORCE inside md_run  [6] MOVE inside md_run  [7] md steps (count)
 * [10] get_hsh SpMV kernel total (CUDA events)  [11] its launches  [12] get_gradient SpMV kernel  [13] its launches
 * [14] nnz of the last QEq matrix (entries, without row padding)  [15] residents  [16] residents+ghosts at the last QEq
 * [17] CG iterations (count)  [18] nnz including row padding  [19] entries of the SpMV's union stream (k_spmv_items)
 * [20] bytes copied host->device by the entry points so far  [21] bytes copied device->host
 * [22] rxg_force calls that reused the halo and 10 A list of the preceding rxg_qeq (RXG_FUSE_API=1)
 * [23] 10 A list builds without a count pass (rows laid out from the previous step's counts)  [24] of those, how many
 *      overflowed a row and were rebuilt with the count pass
 * [25] sparse products taken by k_spmv_win  [26] by k_spmv_rows  [27] cells per group of the current list's window-relative
 *      column stream  [28] entries of its largest window */
int rxg_timers(rxg_handle h, double *it_timer_ms);
/* the reference's it_timer(1:30) itself (src/module.F90:215-217; table printed at src/main.F90:144-180), in SECONDS
 * (the reference keeps system_clock ticks and divides by the clock rate when printing), measured with CUDA events at the
 * phase boundaries and resolved lazily: [0] QEq  [2] LINKEDLIST  [3] COPYATOMS  [4] NEIGHBORLIST  [5] BOCALC  [6] ENbond
 * [7] Ebond  [8] Elnpr  [9] Ehb  [10] E3b  [11] E4b  [12] ForceBondedTerms  [14] GetNonbondingPairList
 * [15] qeq_initialize  [17] get_hsh (the whole CG: the single-pass CG folds get_gradient's product into it)  [23] QEq
 * iterations (a count).  Slots the hot path does not own (file I/O 20-23, send_rec/store/append 25-27, total 30) stay 0
 * and remain the host's.  Fused kernels are booked to one slot: Ebond + Elnpr's main loop -> Ebond, Elnpr's preparation
 * loop -> Elnpr. */
int rxg_it_timer(rxg_handle h, double *it_timer_sec);

/* ---- device-resident stepping: the reference main-loop body src/main.F90:64-98
 * executed nsteps times without host round trips (mdmode 1 NVE; vkick src/main.F90:192-207).   */
int rxg_state_upload(rxg_handle h, int natoms, const double *atype, const double *pos,
                     const double *v, const double *q, const double *qsfp, const double *qsfv);
/* dt in reduced time units (dt[fs]/UTIME, src/init.F90:66); qstep as in rxmd.in; Lex_w2 src/init.F90:69 */
int rxg_md_run(rxg_handle h, int nsteps, double dt, int qstep, double Lex_w2, int step0);
/* initial QEq+FORCE of src/main.F90:27-32 on the resident state */
int rxg_md_prime(rxg_handle h);
int rxg_state_download(rxg_handle h, int *natoms, double *atype, double *pos, double *v, double *f,
                       double *q, double *qsfp, double *qsfv);
/* PE(0:13) of the last FORCE, kinetic energy sum_i hmas*v^2 (PRINTE src/main.F90:225-229), sum q,
 * last nstep_qeq, accumulated astr(1:6) */
int rxg_md_observe(rxg_handle h, double *PE, double *KE, double *qsum, int *nstep_qeq, double *astr);

/* thermostat hooks for resident velocities: the thermostats stay host logic (mdmode 4/5/7/8 and LinearMomentum,
 * src/main.F90:49-62,684-803); these two calls are the O(N) parts they need.
 * stats[6*t .. 6*t+5], t = type-1 < nso: {atom count, sum 1/2 m v^2, sum m, sum m vx, sum m vy, sum m vz} of THIS rank's
 * residents (the host sums over ranks like the reference's MPI_ALLREDUCE);  affine: v(i) = scale[type(i)-1]*v(i) - shift. */
int rxg_md_velocity_stats(rxg_handle h, double *stats);
int rxg_md_velocity_affine(rxg_handle h, const double *scale, const double *shift);

/* ---- introspection used by the parity tests (device -> host copies of hot-path products) ---- */
/* name in: "copyptr"(7 ints) "nbrlist" "nbrindx" "nbp_rowptr"(int64) "nbp_col" "qeq_rowptr" "qeq_col"
 * "qeq_val" "BO"(4 planes) "delta" "deltap" "atype" "pos" "q" "f" "qs" "qt" "gs" "gt" "hs" "ht" "cdbnd" "ccbnd"
 * "nlp" "dDlp" "deltalp" "cell_bonded" "cell_nb" ...; returns element count through *count;
 * out may be NULL to query the count only. */
int rxg_debug_fetch(rxg_handle h, const char *name, void *out, long long capacity_bytes, long long *count);
/* the CG's sparse product H.(x1,x2) of the last QEq matrix as the production path launches it (get_hsh's inner sum,
 * src/qeq.F90:290-306): x2 = {x1,x2} interleaved per atom (residents + ghosts), out4 = {sum H x1, sum H x2, the same sums over
 * ghost columns only} per resident.  Test and timing hook (tests/test_gpu_spmv.py, tools/spmv_bench.py). */
int rxg_debug_spmv(rxg_handle h, const double *x2, double *out4, int reps, double *ms_avg);
/* kernels launched so far by this handle (bench.py's gpu_launches) */
long long rxg_launch_count(rxg_handle h);

#ifdef __cplusplus
}
#endif
#endif /* RXMD_B200_H */
