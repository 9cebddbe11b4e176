/*
 * wolfd2_b200.h -- C ABI of the B200-native wolfd2 time-step hot path.
 *
 * Two layers are exported by libwolfd2_b200.so:
 *
 *  (1) LITERAL SHIMS: the gfortran calling convention of the reference's
 *      hot-path subroutines (lower-case name + '_', every argument a pointer,
 *      INTEGER -> int32_t*, REAL -> double*, LOGICAL -> int32_t*; no CHARACTER
 *      arguments on this path).  Arrays are caller-owned HOST buffers laid out
 *      exactly as the reference declares them: REAL*8 f(0:mnx,0:mny), i fastest,
 *      pitch mnx+1 (reference include/config.f:19-26, include/wolfd2.h:5-8).
 *      Each shim uploads its arguments, runs the CUDA kernels and downloads the
 *      in/out arrays.  They let an unmodified main.f link against this library.
 *
 *  (2) RESIDENCY API (wolfd2_b200_*): fields and metrics stay in HBM across time
 *      steps; the whole step body of main.f:690-981 runs on the device and only a
 *      small status tuple returns to the host per step.
 *
 * There is NO CPU fallback: every entry point fails loudly (message on stderr and
 * a non-zero status / abort for the literal shims, which have no error channel in
 * the reference either -- it uses `stop`) when no CUDA device is usable.
 *
 * Every declaration cites the reference interface it replaces (paths relative to
 * the reference tree).
 */
#ifndef WOLFD2_B200_H
#define WOLFD2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enum constants: include/wolfd2.h:15-59 ------------------------------------ */
enum { W2_RM_BLOCKG = 0, W2_RM_INTERN = 1, W2_RM_POROUS = 2 };
enum { W2_BM_INTERN = 0, W2_BM_WALL1 = 1, W2_BM_WALL2 = 2, W2_BM_INLET = 3,
       W2_BM_OUTLT1 = 4, W2_BM_OUTLT2 = 5 };
enum { W2_WEST = 1, W2_EAST = 2, W2_SOUTH = 3, W2_NORTH = 4 };
enum { W2_U = 1, W2_V = 2, W2_P = 3, W2_T = 4 };
/* ppe_solver ids: src/parse.f:440-465 */
enum { W2_PPE_SOR = 1, W2_PPE_LSOR = 2, W2_PPE_RB_LSOR = 3, W2_PPE_PAR_RB_LSOR = 4,
       W2_PPE_RB_SOR = 5, W2_PPE_PAR_RB_SOR = 6 };

/* ---- plain-data descriptors shared by the residency API ------------------------ */

/* Scalars of `section input_parameters` that the hot path reads
 * (src/parse.f:117-193 defaults; src/file_manip.f:134-159 for dk, re, fr). */
typedef struct wolfd2_params {
    int32_t nx, ny;            /* grid points (src/grid.f:88-89)                     */
    int32_t mqiter;            /* max_ql_iter            (default 20)                */
    int32_t nmeiter;           /* max_me_iter; forced to 1 without thermal energy    */
    int32_t nPpeSolver;        /* ppe_solver id 1..6                                 */
    int32_t msorit;            /* max_sor_iter           (default 2000)              */
    int32_t lCartesGrid;       /* grid_file ... cartesian_grid                       */
    int32_t nfiltu, nfiltv;    /* filter_u / filter_v given                          */
    int32_t reserved_;
    double  dk;                /* non-dimensional time step                          */
    double  re, fr;            /* Reynolds, Froude                                   */
    double  qtol;              /* ql_tolerance           (default 1e-4)              */
    double  sortol, sorrel;    /* sor_tolerance 1e-8, sor_relaxation 1.0             */
    double  fpu, fpv;          /* Shuman filter parameters                           */
} wolfd2_params;

/* Region / boundary tables as built by SetUpBCs (src/bound_cond.f:34-443),
 * Fortran layout: nRegBrd(mgri,mgrj,4), nRegType(mgri,mgrj), nMomBdTp(mgri,mgrj,4),
 * dBCVal(mgri,mgrj,4,4), dPRporos/c1/c2(mgri,mgrj); nReg(2).                        */
typedef struct wolfd2_regions {
    const int32_t *nReg;
    const int32_t *nRegBrd;
    const int32_t *nRegType;
    const int32_t *nMomBdTp;
    const double  *dBCVal;
    const double  *dPRporos;
    const double  *dPRporc1;
    const double  *dPRporc2;
} wolfd2_regions;

/* Thermal energy equation: what main.f passes to ThermEnergy / TempBoundCond / EqState /
 * Filter(_T_) (src/main.f:840-894, 955) that is not in wolfd2_params / wolfd2_regions.  Tables in the Fortran
 * layout: nTRgType(mgri,mgrj) {RT_NOSRCE 0, RT_HEATGN 1, RT_TEMPER 2}, nTemBdTp(mgri,mgrj,4) {BT_INTERN 0,
 * BT_TEMPER 1, BT_HTFLUX 2} (include/wolfd2.h:26-32), dTRgVal, dHGSTval(mgri,mgrj); the temperature BC values
 * are dBCVal(.,.,.,_T_) of wolfd2_regions. */
typedef struct wolfd2_thermal {
    int32_t nthermen;          /* 1: solve the thermal energy equation (thermal_energy)   */
    int32_t neqstate;          /* 1: density from the ideal-gas law after every M-E iteration */
    int32_t nfiltt;            /* 1: Shuman filter on t (filter_t <fpt>)                  */
    int32_t reserved_;
    double  pe;                /* Peclet number re*Pr                                     */
    double  dmeittol;          /* tolerance of the momentum-energy iterations             */
    double  fpt;
    double  uref, densref, tmax, tref, rconst;   /* EqState, src/thermal.f:283-327        */
    const int32_t *nTRgType;
    const int32_t *nTemBdTp;
    const double  *dTRgVal;
    const double  *dHGSTval;
} wolfd2_thermal;
enum { W2_RT_NOSRCE = 0, W2_RT_HEATGN = 1, W2_RT_TEMPER = 2 };
enum { W2_BT_INTERN = 0, W2_BT_TEMPER = 1, W2_BT_HTFLUX = 2 };

/* ATD small-scale model: the arguments of SmallScale (src/small_scale.f:31-49) that main.f
 * passes at src/main.f:647-665 and :912-930 and that are not in wolfd2_params / wolfd2_regions / wolfd2_thermal.
 * Defaults: src/parse.f:144-178.  The thermal region tables (nTRgType, nTemBdTp, dTRgVal) must have been given
 * with wolfd2_b200_set_thermal (nthermen may be 0): SmallScale calls TempBoundCond and Filter(_T_) in any case. */
typedef struct wolfd2_smallscale {
    int32_t nsmallscl;         /* 1: small_scale given                                     */
    int32_t nssPpeSlvr;        /* atd_ppe_solver id 1..6 (default 1)                       */
    int32_t mssSorIt;          /* atd_max_sor_iter (default 2000)                          */
    int32_t reserved_;
    double  dlref, uref, tref, tmax;
    double  pe;
    double  ssSorTol, ssSorRel;
    double  ssFiltPar[4];      /* Fortran ssFiltPar(1:4): (_U_), (_V_), unused, (_T_)      */
    double  ssCu0, ssTsCoef, ssHsCoef, ssTemCoef, ssBnCrit, ssRMpMax, ssRMpExp;
} wolfd2_smallscale;

/* Lagrangian particle trajectories: scalar arguments of Traject (src/traject.f:154-164).
 * nTrMethod 1 = HeunTrap, 2 = FwdEuler; nTrCdEq 1 Stokes, 2 Chein, 3 White, 4 Tilly. */
typedef struct wolfd2_traject {
    int32_t ntr;               /* number of particles                                      */
    int32_t ntsubstp;          /* sub-steps per flow time step                             */
    int32_t nTrMethod, nTrCdEq, mTrHTmit;
    int32_t reserved_;
    double  densref, dTrHTtol, dTrHTdel;
} wolfd2_traject;

/* The 30 metric arrays produced by Metric (src/grid.f:368-535), each
 * REAL*8 (0:mnx,0:mny), zero outside 1..nx,1..ny (static storage in main.f). */
typedef struct wolfd2_metrics {
    const double *rau, *rbu, *rbv, *rgv;
    const double *ran, *rbn, *rgn;
    const double *rac, *rbc, *rgc;
    const double *dju, *djv, *djc, *djn;
    const double *xen, *yen, *xzn, *yzn;
    const double *xec, *yec, *xzc, *yzc;
    const double *xeu, *yeu, *xzv, *yzv;
    const double *xzu, *yzu, *xev, *yev;
} wolfd2_metrics;

/* Per-step tuple printed by PrintDiff (src/string.f:547-559) plus a status word. */
typedef struct wolfd2_step_log {
    int32_t nQLiter;      /* return value of nAuxMomentum (-1: not converged)       */
    int32_t nSorConv;     /* SOR iterations (msorit if not converged, pressure.f:242)*/
    int32_t sor_converged;/* 0 when the msorit cap was hit                           */
    int32_t diverged;     /* difmax > 1e12 (src/main.f:969-972)                      */
    double  dif[4];       /* |p1-pn|, |u1-un|, |v1-vn|, |t1-tn|  (main.f:962-965)    */
} wolfd2_step_log;

enum { W2_OK = 0, W2_ERR_NO_DEVICE = 1, W2_ERR_BAD_ARG = 2, W2_ERR_UNSUPPORTED = 3,
       W2_ERR_CUDA = 4, W2_ERR_DIVERGED = 5 };

/* field selectors for wolfd2_b200_upload_field / download_field */
enum { W2_F_U = 0, W2_F_V = 1, W2_F_P = 2, W2_F_US = 3, W2_F_VS = 4, W2_F_UN = 5,
       W2_F_VN = 6, W2_F_PN = 7, W2_F_D = 8, W2_F_DN = 9, W2_F_B = 10,
       W2_F_T = 11, W2_F_TS = 12, W2_F_TN = 13,
       /* small-scale fields; they exist once wolfd2_b200_set_smallscale has been called */
       W2_F_USS = 14, W2_F_VSS = 15, W2_F_PSS = 16, W2_F_TSS = 17, W2_F_COUNT = 18 };
#define W2_F_CORE 14   /* fields every context holds */

/* ---- library-wide configuration ------------------------------------------------ */

/* The reference fixes mnx,mny,mgri,mgrj at compile time (include/config.f:19-26);
 * the library takes them once at run time.  Must be called before any literal shim.
 * Returns W2_OK or an error code. */
int wolfd2_b200_config(int32_t mnx, int32_t mny, int32_t mgri, int32_t mgrj);
/* Select the CUDA device (default 0). */
int wolfd2_b200_set_device(int32_t device);
/* Tuning knobs that do not change results (every setting yields the same bits; they exist for A/B timing and tests):
 *   "sor_fused_T"      0 = one kernel per colour half-sweep, 1 or 2 = red+black and T iterations fused into one pass
 *                      (default 2); -1 = default / environment W2_SOR_T
 *   "sor_resident"     1 (default) = grids up to 1024^2 solve the PPE in one cooperative, shared-memory-resident launch
 *   "sor_slab_inpass"  several GPUs: 1 = the pass kernel stores its edge rows to the neighbours itself, 0 (default) =
 *                      a follow-up kernel does
 *   "mom_np_cache"     1 (default) = the time-level-n explicit terms are cached across the QL iterations of a step
 *   "mom_cart"         1 (default) = metric arrays that are bitwise one-dimensional / constant (checked) are read as such
 *   "mom_two_streams"  1 (default) = one GPU: XMomentum and YMomentum of a QL iteration run on two streams */
int wolfd2_b200_set_option(const char *name, int32_t value);
const char *wolfd2_b200_last_error(void);
const char *wolfd2_b200_version(void);

/* ---- (2) residency API ---------------------------------------------------------- */
typedef struct wolfd2_ctx wolfd2_ctx;

/* Create a device context for an nx*ny grid whose host arrays have pitch mnx+1 and
 * mny+1 rows (set by wolfd2_b200_config).  Uploads metrics and region tables
 * (what main.f holds after Grid/SetUpBCs, src/main.f:434-448). */
int wolfd2_b200_create(wolfd2_ctx **out, const wolfd2_params *par,
                       const wolfd2_regions *reg, const wolfd2_metrics *met);
void wolfd2_b200_destroy(wolfd2_ctx *ctx);
int wolfd2_b200_set_params(wolfd2_ctx *ctx, const wolfd2_params *par);
/* Switch the thermal energy equation on (or, with th->nthermen == 0, off) for the following steps: the
 * momentum-energy iteration loop of src/main.f:736-880 then runs up to par->nmeiter times per step with
 * ThermEnergy, EqState and the t-norm inside, Filter(_T_) and TempBoundCond after it.  One GPU only. */
int wolfd2_b200_set_thermal(wolfd2_ctx *ctx, const wolfd2_thermal *th);

/* ATD small-scale model on (ss->nsmallscl == 1) or off for the following steps: the blocks of src/main.f:706-727 and
 * :896-940 then run inside wolfd2_b200_step.  The thermal region tables must have been given with
 * wolfd2_b200_set_thermal before (nthermen may be 0).  One GPU only. */
int wolfd2_b200_set_smallscale(wolfd2_ctx *ctx, const wolfd2_smallscale *ss);
/* src/main.f:643-665: SmallScale with initflg = 0 on the context's current u, v, t (seeds the maps, computes the
 * first uss, vss, pss, tss); call after uploading the initial / restart fields, before the first step. */
int wolfd2_b200_smallscale_init(wolfd2_ctx *ctx);
/* One plane (1..3) of the saved map iterates of family 0/1/2 = umap/vmap/tmap: download (upload = 0) or upload. */
int wolfd2_b200_smallscale_map(wolfd2_ctx *ctx, int32_t family, int32_t plane, double *host, int32_t upload);

/* Lagrangian particles (tr->ntr > 0) on or off: x, y are the grid nodes as main.f holds them (REAL*8 (0:mnx,0:mny),
 * src/grid.f), the particle arrays have tr->ntr entries (src/main.f:242-260); nTOutBnd may be NULL (all zero).
 * Every following step ends with the block of src/main.f:1000-1024 (VelAvg, PTDAvg, Traject) on the device;
 * wolfd2_b200_get_particles reads the particles back (any pointer may be NULL).  One GPU only. */
int wolfd2_b200_set_trajectories(wolfd2_ctx *ctx, const wolfd2_traject *tr, const double *x, const double *y,
                                 const double *cpartx, const double *cparty, const double *repc,
                                 const double *xp, const double *yp, const double *up, const double *vp,
                                 const int32_t *nTOutBnd);
int wolfd2_b200_get_particles(wolfd2_ctx *ctx, double *xp, double *yp, double *up, double *vp, int32_t *nTOutBnd);

/* Node averages of the resident fields for output dumps: VelAvg and PTDAvg (src/utility.f:513-647) as main.f calls
 * them before SaveStdVarsP3D / SaveTimeSrs (src/main.f:1053-1062), computed on the device; only the averaged arrays
 * come back, in the host layout (0:mnx,0:mny), zero outside the nodes 1..nx, 1..ny.  set 0: (u, v, p); set 1: the
 * small-scale fields (uss, vss, pss).  tav: TAveraged (src/utility.f:668-740) of t (set 0) or tss (set 1, nScale = 1);
 * it needs the thermal region tables (wolfd2_b200_set_thermal).  Any output pointer may be NULL.  One GPU only. */
int wolfd2_b200_node_averages(wolfd2_ctx *ctx, int32_t set, double *util, double *vbar, double *pav, double *tav);

/* Time-series monitor points: SaveTimeSrs case 1 (src/file_manip.f:806-834) as called from src/main.f:984-995.  After
 * set_probes every wolfd2_b200_step samples u, v, p, t, uss, vss, pss, tss at the points (iTS, jTS) (1-based, as in
 * the reference; points outside 1..nx, 1..ny become (1,1) with the reference's warning) whenever the number of steps
 * taken since set_probes is a multiple of freq (nTSFreq).  get_probe_records returns the records gathered since the
 * last call, out[rec][point][8], with the step number of each; *nrec = records copied (at most maxrec; with out NULL:
 * the number waiting).  The `.ts` files themselves are written by the host (wolfd2_b200/timeseries.py).  One GPU. */
int wolfd2_b200_set_probes(wolfd2_ctx *ctx, int32_t npoints, const int32_t *iTS, const int32_t *jTS, int32_t freq);
int wolfd2_b200_get_probe_records(wolfd2_ctx *ctx, int32_t maxrec, double *out, int32_t *steps, int32_t *nrec);

/* Time averaging, the reference's -D_TIMEAVG_ mode, on the resident fields.  wolfd2_b200_timeavg(ctx, op): op 0 = begin
 * (src/main.f:510-541: the nineteen arrays zeroed), 1 / 2 = accumulate pass 1 / pass 2 at the end of every following
 * wolfd2_b200_step (:1107-1208: node averages via VelAvg / TAveraged / PTDAvg, then the sums), 3 = stop, 4 = release.
 * timeavg_finish(pass, nts, uref, dlref): :1239-1297 (division by the steps taken, turbulence kinetic energy,
 * dissipation rate (...)*uref*dlref/re).  As in the reference the run is made twice: pass 1 for the means, then --
 * from the same initial state -- pass 2 for the fluctuations about them.  timeavg_get: array `which` = 0..18 in the order
 * ubar vbar tbar pbar upb vpb tpb upupb vpvpb upvpb uptpb vptpb upxsb upysb vpxsb vpysb trbke dssrt dtdyb, host layout
 * (0:mnx,0:mny); wolfd2_b200/plot3d.py save_tmavg_p3d writes SaveTmAvgP3D's file from them.  One GPU. */
int wolfd2_b200_timeavg(wolfd2_ctx *ctx, int32_t op);
int wolfd2_b200_timeavg_finish(wolfd2_ctx *ctx, int32_t pass, int32_t nts, double uref, double dlref);
int wolfd2_b200_timeavg_get(wolfd2_ctx *ctx, int32_t which, double *host);

/* Host <-> device copies of one field, host layout (0:mnx,0:mny). */
int wolfd2_b200_upload_field(wolfd2_ctx *ctx, int32_t which, const double *host);
int wolfd2_b200_download_field(wolfd2_ctx *ctx, int32_t which, double *host);
/* Rows jfirst..jfirst+nrows-1 (global row indices, within the rows the context holds) of metric array `which`
 * (0-based position in wolfd2_metrics) from a host block of nrows x (mnx+1) doubles: lets the host build the metrics
 * of a large grid window by window (src/grid.f:368-535 is a local formula of the node rows j-2..j+2) and pass NULL
 * metric pointers to wolfd2_b200_create. */
int wolfd2_b200_upload_metric_rows(wolfd2_ctx *ctx, int32_t which, int32_t jfirst, int32_t nrows, const double *host);

/* Cold-start projection, src/main.f:606-641. */
int wolfd2_b200_coldstart(wolfd2_ctx *ctx, int32_t *nSorConv);
/* nsteps iterations of the step body src/main.f:690-981 (cold flow: no thermal
 * energy, no ATD) entirely on the device.  logs may be NULL or hold nsteps entries. */
int wolfd2_b200_step(wolfd2_ctx *ctx, int32_t nsteps, wolfd2_step_log *logs);
/* Same, but the state crosses the boundary in HOST buffers every call:
 * u,v,p are uploaded, nsteps run, u,v,p downloaded (the e2e path of bench.py). */
int wolfd2_b200_step_host(wolfd2_ctx *ctx, int32_t nsteps, double *u, double *v,
                          double *p, wolfd2_step_log *logs);

/* Device-timed sections of the last wolfd2_b200_step call (CUDA events on the
 * library's stream), milliseconds: [0] total, [1] momentum (QL loop), [2] PPE
 * (div+rhs+SOR), [3] projection + BC fills + norms.  Kernel launch counters since
 * context creation in launches[0..3] likewise. */
int wolfd2_b200_last_timing(wolfd2_ctx *ctx, double ms[4], int64_t launches[4]);
/* Average device time of the SOR sweep kernel launches inside the last step (ms)
 * and the number of SOR iterations they covered. */
int wolfd2_b200_last_sor_timing(wolfd2_ctx *ctx, double *ms_total, int64_t *iterations);
/* Host waits (stream / event synchronisations) issued by the last wolfd2_b200_step call: the loop control of the QL and
 * SOR loops lives on the device, the host only reads the step's status tuple and throttles how far it runs ahead. */
int wolfd2_b200_last_host_syncs(wolfd2_ctx *ctx, int64_t *syncs);
int wolfd2_b200_sync(wolfd2_ctx *ctx);
/* Page-locked host buffers for the e2e path (cudaMallocHost / cudaFreeHost). */
void *wolfd2_b200_host_alloc(uint64_t bytes);
void wolfd2_b200_host_free(void *p);

/* ---- (3) several GPUs of one node: row slabs, one process per GPU ------------------
 * The reference is serial (its only parallel constructs are OpenMP loops inside SorRBP/SlorRBP,
 * src/pressure.f:599-649); this is the decomposition SURVEY.md section 8(e) derives from it.  The unknown
 * pressure rows j = 2..ny are cut into `world` contiguous slabs.  Rank r holds every array on rows
 * A0..A1 (its rows J0..J1 plus HG halo rows, clipped to 0..ny+1); the arrays it passes to create_slab /
 * upload / download / step_host hold those rows only, host row 0 = global row A0, so mny >= A1-A0.
 * params and region tables stay global.  Results are bit-identical to the one-GPU run.
 * Supported: ppe_solver 5/6, Cartesian grid, nx >= 254; every face type (the OUTLT2 recurrences along west / east faces
 * cross the slabs through peer-mapped gather buffers and need CUDA IPC between the GPUs). */
/* out = {J0, J1, A0, A1, HG} for `rank` of `world` */
int wolfd2_b200_slab_layout(int32_t nx, int32_t ny, int32_t world, int32_t rank, int32_t out[5]);
/* NCCL communicator over the ranks: rank 0 makes the 128-byte id, the launcher hands it to every rank
 * (MPI_Bcast, torch.distributed, a file ...), every rank calls comm_init after wolfd2_b200_set_device. */
int wolfd2_b200_comm_unique_id(unsigned char id[128]);
int wolfd2_b200_comm_init(int32_t rank, int32_t world, const unsigned char id[128]);
int wolfd2_b200_comm_finalize(void);
int wolfd2_b200_create_slab(wolfd2_ctx **out, const wolfd2_params *par, const wolfd2_regions *reg,
                            const wolfd2_metrics *met, int32_t rank, int32_t world);
/* Verification of a slab run against one GPU (collective over the ranks): gather_global copies every rank's rows of
 * the 30 metric arrays (what & 1) and of u, v, p (what & 2) into `global`, a one-GPU context of the same grid on rank
 * 0's device (NULL on the other ranks); compare_global gathers field `which` of the slab run and counts the cells
 * 0..nx+1 x 0..ny+1 whose bit patterns differ from `global`'s own field (results on rank 0). */
int wolfd2_b200_gather_global(wolfd2_ctx *slab, wolfd2_ctx *global, int32_t what);
int wolfd2_b200_compare_global(wolfd2_ctx *slab, wolfd2_ctx *global, int32_t which, uint64_t *ndiff, double *maxabs);

/* ---- (1) literal shims: gfortran ABI of the reference subroutines --------------- */

/* src/momentum.f:33-48  INTEGER function nAuxMomentum(...) */
int32_t nauxmomentum_(const int32_t *nx, const int32_t *ny, const int32_t *mqiter,
    const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
    const int32_t *nMomBdTp,
    const double *dk, const double *re, const double *fr, const double *qtol,
    const double *dPRporos, const double *dPRporc1, const double *dPRporc2,
    const double *dBCVal,
    const double *ran, const double *rbn, const double *rgn,
    const double *rac, const double *rbc, const double *rgc,
    const double *dju, const double *djv,
    const double *xec, const double *yec, const double *xzn, const double *yzn,
    const double *xen, const double *yen, const double *xzc, const double *yzc,
    const double *xeu, const double *yeu, const double *xzu, const double *yzu,
    const double *xev, const double *yev, const double *xzv, const double *yzv,
    const double *d, const double *dn,
    const double *un, const double *vn, double *us, double *vs);

/* src/momentum.f:199-208 */
void xmomentum_(const int32_t *nx, const int32_t *ny,
    const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
    const int32_t *nMomBdTp, const double *dk, const double *re,
    const double *dPRporos, const double *dPRporc1, const double *dPRporc2,
    const double *rbn, const double *rgn, const double *rac, const double *rbc,
    const double *dju,
    const double *xec, const double *yec, const double *xzn, const double *yzn,
    const double *xeu, const double *yeu, const double *xzu, const double *yzu,
    const double *us, const double *vs, const double *un, const double *vn,
    double *dus);

/* src/momentum.f:520-530 */
void ymomentum_(const int32_t *nx, const int32_t *ny,
    const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
    const int32_t *nMomBdTp, const double *dk, const double *re, const double *fr,
    const double *dPRporos, const double *dPRporc1, const double *dPRporc2,
    const double *ran, const double *rbn, const double *rbc, const double *rgc,
    const double *djv,
    const double *xen, const double *yen, const double *xzc, const double *yzc,
    const double *xev, const double *yev, const double *xzv, const double *yzv,
    const double *d, const double *dn,
    const double *us, const double *vs, const double *un, const double *vn,
    double *dvs);

/* src/momentum.f:1307  subroutine AltTridLU(n, a, b): a(3,n) AoS, solution in b */
void alttridlu_(const int32_t *n, double *a, double *b);

/* src/pressure.f:30-38 */
void ppe_(const int32_t *nx, const int32_t *ny,
    const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
    const int32_t *lCartesGrid,
    const int32_t *nPpeSolver, const int32_t *msorit, int32_t *nSorConv,
    const double *dk, const double *sortol, const double *sorrel,
    const double *rau, const double *rbu, const double *rbv, const double *rgv,
    const double *xeu, const double *yeu, const double *xzv, const double *yzv,
    const double *u, const double *v, double *p);

/* src/pressure.f:265-267 */
void divergence_(const int32_t *nx, const int32_t *ny, const int32_t *nloc,
    const double *xet, const double *yet, const double *xzi, const double *yzi,
    const double *u, const double *v, double *div);

/* src/utility.f:253-259 */
void project_(const int32_t *nx, const int32_t *ny,
    const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
    const int32_t *nMomBdTp, const double *dk,
    const double *dju, const double *djv,
    const double *yeu, const double *xzv, const double *yzu, const double *xev,
    const double *p, double *u, double *v);

/* src/bound_cond.f:511-513 */
void velboundcond_(const int32_t *nx, const int32_t *ny, const int32_t *nReg,
    const int32_t *nRegBrd, const int32_t *nMomBdTp, const double *dBCVal,
    double *u, double *v);
/* src/bound_cond.f:853-856 */
void presboundcond_(const int32_t *nx, const int32_t *ny, const int32_t *nReg,
    const int32_t *nRegBrd, const int32_t *nRegType, const int32_t *nMomBdTp,
    const double *dBCVal, double *p);
/* src/bound_cond.f:1656-1659 */
void veloutflowbcs_(const int32_t *nx, const int32_t *ny, const int32_t *nReg,
    const int32_t *nRegBrd, const int32_t *nMomBdTp, const double *dBCVal,
    double *u, double *v);
/* src/utility.f:33-36 */
void filter_(const int32_t *nx, const int32_t *ny, const int32_t *ncomp,
    const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
    const int32_t *nMomBdTp, const int32_t *nTRgType, const double *fp, double *qu);
/* src/utility.f:446, 479 */
double diffmaxnorm_(const int32_t *nx, const int32_t *ny, const double *un,
                    const double *u);
double dmaxnorm_(const int32_t *nx, const int32_t *ny, const double *u);

/* Routines below XMomentum / YMomentum / Ppe in the reference's call tree, exported for unit parity.
 * On the production path they are evaluated per unknown inside the momentum kernels; these entry points run
 * the same device functions one operator at a time.  Output arrays are in/out: only the reference's loop
 * ranges are written.
 * ConvCoef src/momentum.f:864-866; DConvU :987; DDiffU :1015; DConvV :1051; DDiffV :1079; PorosCoef :1115-1118;
 * RhsPpe src/pressure.f:329-330 (b is the vector b(mn), entries 1..(nx-1)(ny-1)).
 * Not exported: Sor, SorRB, SorRBP, Slor, SlorRB, SlorRBP (src/pressure.f:384-1138) take the assembled matrix
 * a(mn,5), which this implementation never forms (coefficients are re-formed from rau, rgv); they are reached
 * through ppe_ with nPpeSolver = 1..6. */
void convcoef_(const int32_t *nx, const int32_t *ny, const int32_t *ncomp, const int32_t *njacob,
               const double *xzi, const double *xet, const double *yzi, const double *yet,
               const double *u, const double *v, double *cc1, double *cc2);
void dconvu_(const int32_t *nx, const int32_t *ny, const double *c1, const double *c2, const double *u, double *c);
void ddiffu_(const int32_t *nx, const int32_t *ny, const double *ac, const double *bc, const double *bn,
             const double *gn, const double *u, double *d);
void dconvv_(const int32_t *nx, const int32_t *ny, const double *c1, const double *c2, const double *v, double *c);
void ddiffv_(const int32_t *nx, const int32_t *ny, const double *an, const double *bc, const double *bn,
             const double *gc, const double *v, double *d);
void poroscoef_(const int32_t *nx, const int32_t *ny, const int32_t *ncomp, const int32_t *njacob,
                const int32_t *nReg, const int32_t *nRegBrd, const int32_t *nRegType,
                const double *dPRporos, const double *dPRporc1, const double *dPRporc2,
                const double *u, const double *v, double *cp);
void rhsppe_(const int32_t *nx, const int32_t *ny, const int32_t *lCartesGrid, const double *dk,
             const double *rbu, const double *rbv, const double *div, const double *p, double *b);

/* src/bound_cond.f:1030-1034 */
void tempboundcond_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
    const int32_t *nTRgType, const int32_t *nTemBdTp, const double *dTRgVal, const double *dBCVal, double *t);
/* src/thermal.f:24-33 */
void thermenergy_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
    const int32_t *nTRgType, const int32_t *nTemBdTp, const double *dk, const double *pe,
    const double *dTRgVal, const double *dHGSTval, const double *dBCVal,
    const double *rau, const double *rbu, const double *rbv, const double *rgv, const double *djc,
    const double *xeu, const double *yeu, const double *xzv, const double *yzv,
    const double *xec, const double *yec, const double *xzc, const double *yzc,
    const double *un, const double *vn, const double *u, const double *v, const double *tn, double *t);
/* src/thermal.f:283-285 */
void eqstate_(const int32_t *nx, const int32_t *ny, const double *uref, const double *densref,
    const double *tmax, const double *tref, const double *rconst, const double *p, const double *t, double *den);


/* src/small_scale.f:31-49.  The maps, cell areas and work arrays are `save`d locals in the reference: the shim keeps
 * them in the library's shim context between calls (initflg <= 0 re-seeds them). */
void smallscale_(const int32_t *nx, const int32_t *ny, const int32_t *initflg, const int32_t *nthermen,
    const int32_t *lCartesGrid, const int32_t *nReg, const int32_t *nRegBrd,
    const int32_t *nRegType, const int32_t *nTRgType, const int32_t *nMomBdTp, const int32_t *nTemBdTp,
    const int32_t *nPpeSolver, const int32_t *msorit,
    const double *dlref, const double *uref, const double *tref, const double *tmax,
    const double *dka, const double *re, const double *pe, const double *sortol, const double *sorrel,
    const double *fp, const double *cu0, const double *TsCoef, const double *HsCoef, const double *TemCoef,
    const double *bnumc, const double *rmax, const double *rlc, const double *dTRgVal, const double *dBCVal,
    const double *rau, const double *rbu, const double *rbv, const double *rgv,
    const double *dju, const double *djv, const double *djc,
    const double *xeu, const double *yeu, const double *xzv, const double *yzv,
    const double *xzu, const double *yzu, const double *xev, const double *yev,
    const double *xec, const double *yec, const double *xzc, const double *yzc,
    const double *u1, const double *v1, const double *t1,
    double *uss, double *vss, double *pss, double *tss);
/* src/bound_cond.f:1209-1213 */
void smlsclbc_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
    const int32_t *nRegType, const int32_t *nMomBdTp, const int32_t *nTRgType, const int32_t *nTemBdTp,
    const double *dBCVal, double *u, double *v, double *p, double *t);
/* src/utility.f:512, 574 */
void ptdavg_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
    const int32_t *nRegType, const double *p, double *pav);
/* src/utility.f:668-671 */
void taveraged_(const int32_t *nx, const int32_t *ny, const int32_t *nScale, const int32_t *nReg,
    const int32_t *nRegBrd, const int32_t *nTRgType, const double *dTRgVal, const double *t, double *tav);
void velavg_(const int32_t *nx, const int32_t *ny, const int32_t *nReg, const int32_t *nRegBrd,
    const int32_t *nRegType, const double *u, const double *v, double *util, double *vbar);
/* src/traject.f:154-164.  HeunTrap's step is the sub-step size (the reference passes the INTEGER sub-step counter
 * there). */
void traject_(const int32_t *nx, const int32_t *ny, const int32_t *ntr, const int32_t *ntsubstp,
    const int32_t *nTrMethod, const int32_t *nTrCdEq, const int32_t *mTrHTmit, int32_t *nTOutBnd,
    const double *dkflow, const double *densref, const double *fr, const double *dTrHTtol, const double *dTrHTdel,
    const double *cpartx, const double *cparty, const double *repc,
    const double *x, const double *y, const double *u, const double *v, const double *un, const double *vn,
    const double *dens, const double *densn, double *xp, double *yp, double *up, double *vp);

#ifdef __cplusplus
}
#endif
#endif /* WOLFD2_B200_H */
