/*
 * pinc_b200.h — C-ABI of libpinc_b200.so: the B200 (sm_100a) implementation of PINC's
 * per-timestep particle-in-cell loop behind PINC's own C entry points.
 *
 * Every function below with a PINC name has the reference's signature and argument
 * meaning (cited as /root/reference/src/<file>:<line> of the declaration it replaces),
 * so a PINC build links this library in place of the corresponding objects
 * (see INTEGRATION.md).  Struct layouts restate core.h / multigrid.h byte for byte.
 *
 * Memory model.  The host structs stay owned by the caller (plain malloc, exactly as
 * gAlloc/pAlloc/gAllocMpi/gCreateNeighborhood/mgAllocSolver of the reference produce
 * them).  The first time a Grid* / Population* is seen it gets a device-resident
 * mirror keyed by the host pointer and its host contents are uploaded.  After that all
 * PINC-named functions work on the mirror and the large host arrays (grid->val,
 * pop->pos, pop->vel) are STALE until pincSync*ToHost() is called; host code that
 * writes them (initial conditions) calls pincSync*ToDevice().  The small host scalars
 * the reference's driver reads (pop->iStop[], pop->kinEnergy[], pop->potEnergy[],
 * mpiInfo->nEmigrants[], mpiInfo->nImmigrants[]) are always kept current.
 *
 * There is no CPU fallback: every entry point aborts through pincFatal() when no CUDA
 * device is usable.
 */
#ifndef PINC_B200_H
#define PINC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* hid_t of the HDF5 the host was built with (HDF5 >= 1.10: int64_t; 1.8: int). */
#ifndef PINC_HID_T
#define PINC_HID_T int64_t
#endif
typedef PINC_HID_T pinc_hid_t;

/* ---------------------------------------------------------------------------------
 * Data model (restates src/core.h:72-86, 112-138, 146-151, 261-277, 467 and
 * src/grid.h:22-25, src/multigrid.h:27-57)
 * ------------------------------------------------------------------------------- */
typedef void (*funPtr)();                       /* core.h:467 */
#ifndef _DICTIONARY_H_
typedef struct _dictionary_ dictionary;         /* lib/iniparser/src/dictionary.h:41-47; opaque here: read through the host's iniGet* */
#endif

typedef enum { PERIODIC = 0x01, DIRICHLET = 0x02, NEUMANN = 0x03, NONE = 0x10 } bndType; /* core.h:146-151 */
typedef enum { TOHALO = 0, FROMHALO = 1 } opDirection;                                  /* grid.h:22-25 */

typedef struct {                                /* core.h:72-86 */
	double *pos;            /* AoS: pos[3*i+d], local grid units */
	double *vel;
	long int *iStart;       /* nSpecies+1 */
	long int *iStop;        /* nSpecies   */
	long int *objVicinity;
	long int *collisions;
	double *charge;
	double *mass;
	double *kinEnergy;      /* nSpecies+1 */
	double *potEnergy;      /* nSpecies+1 */
	int nSpecies;
	int nDims;
	pinc_hid_t h5;
} Population;

typedef struct {                                /* core.h:112-138 */
	int mpiRank;
	int mpiSize;
	int nDims;
	int *subdomain;
	int *nSubdomains;
	int *nSubdomainsProd;
	int *offset;
	double *posToSubdomain;
	int nSpecies;
	int nNeighbors;
	int neighborhoodCenter;
	long int **migrants;
	long int **migrantsDummy;
	long int *nEmigrants;       /* [ne*nSpecies+s] */
	long int *nEmigrantsAlloc;  /* [ne] */
	long int *nImmigrants;
	long int nImmigrantsAlloc;
	double **emigrants;
	double **emigrantsDummy;
	double *immigrants;
	double *thresholds;         /* 2*nDims: lower x,y,z then upper x,y,z */
	void *send;                 /* MPI_Request* in the reference */
	void *recv;
} MpiInfo;

typedef struct {                                /* core.h:261-277 */
	double *val;
	int rank;
	int *size;
	int *trueSize;
	long int *sizeProd;
	int *nGhostLayers;
	double *sendSlice;
	double *recvSlice;
	double *bndSlice;
	pinc_hid_t h5;
	pinc_hid_t h5MemSpace;
	pinc_hid_t h5FileSpace;
	bndType *bnd;
} Grid;

typedef struct {                                /* multigrid.h:27-50 */
	Grid **grids;
	int nLevels;
	int nMGCycles;
	int nPreSmooth;
	int nPostSmooth;
	int nCoarseSolve;
	void (*coarseSolv)(Grid *phi, const Grid *rho, const int nCycles, const MpiInfo *mpiInfo);
	void (*postSmooth)(Grid *phi, const Grid *rho, const int nCycles, const MpiInfo *mpiInfo);
	void (*preSmooth)(Grid *phi, const Grid *rho, const int nCycles, const MpiInfo *mpiInfo);
	void (*restrictor)(const Grid *fine, Grid *coarse);
	void (*prolongator)(Grid *fine, const Grid *coarse, const MpiInfo *mpiInfo);
} Multigrid;

typedef struct {                                /* multigrid.h:52-58 */
	Grid *res;
	Multigrid *mgRho;
	Multigrid *mgPhi;
	Multigrid *mgRes;
	funPtr mgAlgo;
} MultigridSolver;

typedef struct Object Object;                   /* object.h: opaque here; puMove ignores it (quirk Q4) */

/* ---------------------------------------------------------------------------------
 * Particle path (src/pusher.h)
 * ------------------------------------------------------------------------------- */
void puMove(Population *pop, Object *obj);                                      /* pusher.h:24  (pusher.c:86)  */
void puAcc3D1(Population *pop, Grid *E);                                        /* pusher.h:119 (pusher.c:147) */
void puAcc3D1KE(Population *pop, Grid *E);                                      /* pusher.h:120 (pusher.c:178) */
void puBoris3D1(Population *pop, Grid *E, const double *T, const double *S);    /* pusher.h:125 (pusher.c:394) */
void puBoris3D1KE(Population *pop, Grid *E, const double *T, const double *S);  /* pusher.h:126 (pusher.c:433) */
void puDistr3D1(const Population *pop, Grid *rho);                              /* pusher.h:163 (pusher.c:512) */
/* the N-dimensional / zeroth-order select() targets of src/main.c:55-70, for nDims = 3: first order with the operation order of
 * the reference's recursion (puInterpND1Inner, puDistrND1Inner), zeroth order = nearest grid point */
void puAccND1(Population *pop, Grid *E);                                         /* pusher.h (pusher.c:269) */
void puAccND1KE(Population *pop, Grid *E);                                       /* pusher.c:219 */
void puAccND0(Population *pop, Grid *E);                                         /* pusher.c:357 */
void puAccND0KE(Population *pop, Grid *E);                                       /* pusher.c:311 */
void puDistrND1(const Population *pop, Grid *rho);                               /* pusher.c:578 */
void puDistrND0(const Population *pop, Grid *rho);                               /* pusher.c:644 */
void puExtractEmigrantsND(Population *pop, MpiInfo *mpiInfo);                    /* pusher.c:864 */
void puExtractEmigrants3D(Population *pop, MpiInfo *mpiInfo);                   /* pusher.h:180 (pusher.c:782) */
void puMigrate(Population *pop, MpiInfo *mpiInfo, Grid *grid);                  /* pusher.h:184 (pusher.c:1030) */
int  puRankToNeighbor(MpiInfo *mpiInfo, int rank);                              /* pusher.h:186 (pusher.c:1214) */
int  puNeighborToRank(MpiInfo *mpiInfo, int neighbor);                          /* pusher.h:187 (pusher.c:1194) */
int  puNeighborToReciprocal(int neighbor, int nDims);                           /* pusher.h:188 (pusher.c:1181) */
void pPosAssertInLocalFrame(const Population *pop, const Grid *grid);           /* population.h (population.c:316) */
void pVelAssertMax(const Population *pop, double max);                          /* population.h (population.c:343) */
void pSumKinEnergy(Population *pop);                                            /* population.h (population.c:700) */
void pNew(Population *pop, int s, const double *pos, const double *vel);        /* population.h:108 (population.c:430) */
void pCut(Population *pop, int s, long int p, double *pos, double *vel);        /* population.h:128 (population.c:452) */
/* plain-argument form of puGet3DRotationParameters (pusher.c:485; the reference reads
 * BExt/charge/mass from the ini dictionary, which stays host code) */
void pincGet3DRotationParameters(int nSpecies, const double *BExt, const double *charge,
                                 const double *mass, double *T, double *S);
/* plain-argument form of the validator behind every puXxx_set (pusher.c:1047 puSanity);
 * returns 0 if the configuration is acceptable, else an error code and a message. */
int  pincPuSanity(const char *name, int nDims, const int *nGhostLayers, const double *thresholds,
                  int dim, int order, char *errbuf, int errlen);
/* The X_set(ini) selectors that io.h:105 select() calls, mgSolver_set and mgAllocSolver, puGet3DRotationParameters: same
 * names, signatures and checks as the reference.  They read `ini` through the host's own iniGetInt / iniGetStr /
 * iniGetIntArr / iniGetDoubleArr (src/io.h:228-240, weak references resolved from the PINC executable's io.o); a
 * host without PINC's ini layer uses the pinc* plain-argument forms above and below. */
funPtr puAcc3D1_set(dictionary *ini);                                           /* pusher.h:119 (pusher.c:143) */
funPtr puAcc3D1KE_set(dictionary *ini);                                         /* pusher.h:120 (pusher.c:174) */
funPtr puDistr3D1_set(dictionary *ini);                                         /* pusher.h:163 (pusher.c:508) */
funPtr puExtractEmigrants3D_set(const dictionary *ini);                         /* pusher.h:180 (pusher.c:777) */
funPtr puAccND1_set(dictionary *ini); funPtr puAccND1KE_set(dictionary *ini); funPtr puAccND0_set(dictionary *ini); funPtr puAccND0KE_set(dictionary *ini);
funPtr puDistrND1_set(dictionary *ini); funPtr puDistrND0_set(dictionary *ini); funPtr puExtractEmigrantsND_set(const dictionary *ini);   /* pusher.c:215-391, 574-646, 860 */
void puGet3DRotationParameters(dictionary *ini, double *T, double *S);          /* pusher.h:128 (pusher.c:485) */

/* ---------------------------------------------------------------------------------
 * Grid path (src/grid.h)
 * ------------------------------------------------------------------------------- */
void getSlice(double *slice, const Grid *grid, int d, int offset);              /* grid.h:194 (grid.c:85)  */
void setSlice(const double *slice, Grid *grid, int d, int offset);              /* grid.h:226 (grid.c:111) */
void addSlice(const double *slice, Grid *grid, int d, int offset);              /* grid.h:239 (grid.c:137) */
void gHaloOp(funPtr sliceOp, Grid *grid, const MpiInfo *mpiInfo, opDirection dir);             /* grid.h:156 (grid.c:340) */
void gHaloOpDim(funPtr sliceOp, Grid *grid, const MpiInfo *mpiInfo, int d, opDirection dir);   /* grid.h:140 (grid.c:349) */
void gFinDiff1st(const Grid *scalar, Grid *field);                              /* grid.h:316 (grid.c:226) */
void gFinDiff2nd3D(Grid *result, const Grid *object);                           /* grid.h:325 (grid.c:296) */
void gZero(Grid *grid);                                                         /* grid.h:247 (grid.c:699) */
void gMul(Grid *grid, double num);                                              /* grid.h:275 (grid.c:668) */
void gAdd(Grid *grid, double num);                                              /* grid.h:283 (grid.c:675) */
void gSub(Grid *grid, double num);                                              /* grid.h:291 (grid.c:682) */
void gSquare(Grid *grid);                                                       /* grid.h:299 (grid.c:689) */
void gCopy(const Grid *original, Grid *copy);                                   /* grid.h:267 (grid.c:718) */
void gAddTo(Grid *result, Grid *addition);                                      /* grid.h:367 (grid.c:783) */
void gSubFrom(Grid *result, const Grid *subtraction);                           /* grid.h:378 (grid.c:793) */
double gSumTruegrid(const Grid *grid);                                          /* grid.h:388 (grid.c:833) */
long int gTotTruesize(const Grid *grid, const MpiInfo *mpiInfo);                /* grid.h:308 (grid.c:849) */
void gNeutralizeGrid(Grid *grid, const MpiInfo *mpiInfo);                       /* grid.h:357 (grid.c:730) */
void gBnd(Grid *grid, const MpiInfo *mpiInfo);                                  /* grid.h:402 (grid.c:992): PERIODIC, DIRICHLET, NEUMANN */
void gDirichlet(Grid *grid, const int boundary, const MpiInfo *mpiInfo);        /* grid.c:929 (no prototype in grid.h) */
void gNeumann(Grid *grid, const int boundary, const MpiInfo *mpiInfo);          /* grid.c:958 (no prototype in grid.h) */
void gSetBndSlices(Grid *grid, MpiInfo *mpiInfo);                               /* grid.h:66 (grid.c:608); host arrays */
void gPotEnergy(const Grid *rho, const Grid *phi, Population *pop);             /* grid.h:527 (grid.c:1276) */

/* ---------------------------------------------------------------------------------
 * Multigrid Poisson solver (src/multigrid.h)
 * ------------------------------------------------------------------------------- */
void mgSolve(const MultigridSolver *solver, const Grid *rho, const Grid *phi, const MpiInfo *mpiInfo); /* multigrid.h:95 (multigrid.c:403) */
void mgSolveRaw(funPtr mgAlgo, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo); /* multigrid.c:1688 */
void mgRestrictBnd(Multigrid *mgGrid);                                          /* multigrid.h:366 (multigrid.c:1314); host arrays */
void mgVRecursive(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi,
                  Multigrid *mgRes, const MpiInfo *mpiInfo);                    /* multigrid.c:1550 */
void mgVRegular(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo); /* multigrid.c:1559 */
void mgW(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo);        /* multigrid.c:1675 */
void mgFMG(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo);      /* multigrid.c:1652; fails loudly: the reference's destroys mgRho->grids[0] */
void mgJacob3D(Grid *phi, const Grid *rho, const int nCycles, const MpiInfo *mpiInfo);                                          /* multigrid.c:500 ("jacobian") */
void mgGS3D(Grid *phi, const Grid *rho, int nCycles, const MpiInfo *mpiInfo);  /* multigrid.c:683 */
void mgHalfRestrict3D(const Grid *fine, Grid *coarse);                          /* multigrid.c:844 */
void mgBilinProl3D(Grid *fine, const Grid *coarse, const MpiInfo *mpiInfo);     /* multigrid.c:1127 */
void mgResidual(Grid *res, const Grid *rho, const Grid *phi, const MpiInfo *mpiInfo); /* multigrid.c:1385 */
double mgSumTrueSquared(Grid *error, const MpiInfo *mpiInfo);                   /* multigrid.c:1471 */
void mgSolver(void (**solve)(), MultigridSolver *(**solverAlloc)(), void (**solverFree)()); /* multigrid.c:392: hands out mgSolve, mgAllocSolver, mgFreeSolver */
funPtr mgSolver_set(const dictionary *ini);                                     /* multigrid.h:92 (multigrid.c:398) */
MultigridSolver *mgAllocSolver(const dictionary *ini, Grid *rho, Grid *phi);    /* multigrid.h:93 (multigrid.c:364): reads [multigrid] through the host's iniGet* */
/* plain-argument form of mgAllocSolver/mgAlloc/mgAllocSubGrids (multigrid.c:128-382);
 * the reference reads the five integers from [multigrid] of the ini file. */
MultigridSolver *pincMgAllocSolver(Grid *rho, Grid *phi, int mgLevels, int mgCycles,
                                   int nPreSmooth, int nPostSmooth, int nCoarseSolve);
void mgFreeSolver(MultigridSolver *solver);                                     /* multigrid.h:94 (multigrid.c:384) */
/* residual history of the most recent mgSolve on this rank: returns the V-cycle count and
 * copies up to `cap` values of barRes (multigrid.c:1700-1704), one per V-cycle. */
int pincMgLastHistory(double *barRes, int cap);
/* barRes of the last V-cycle of the most recent mgSolve on this rank.  The reference's tolerance loop is unbounded
 * (multigrid.c:1697, a stalled solve hangs); here it is bounded by $PINC_B200_MG_MAXCYCLES (default 10000) and a solve
 * that ends above 1e-10 is a fatal "PINC-B200 ERROR" raised at the first synchronisation after the solve. */
double pincMgLastBarRes(void);
/* multi-rank solves replicated (1, default) or distributed over the ranks (0); overrides
 * $PINC_B200_MG_REPLICA.  Process-wide: call it on every rank before the next mgSolve. */
void pincMgSetReplica(int on);
/* replicated multi-rank solves keep the finest level distributed (1, default: hybrid - every rank smooths its own sub-domain,
 * block faces across sub-domain boundaries travel through peer memory; the coarser levels are replicated) or replicate the
 * whole hierarchy (0); overrides $PINC_B200_MG_HYBRID.  Process-wide. */
void pincMgSetHybrid(int on);
/* A/B switch of the all-SM kernel's block-resident smoother: 0 never the row smoother (mgrows.cuh), 1 (default) for blocks
 * too big for the register-descriptor path (levels of >= 1 M nodes), 2 whenever the block shape allows it. */
void pincMgSetRowMode(int mode);
/* which implementation ran the most recent mgSolve on this rank: 0 one kernel per reference call (distributed), 1 the
 * all-SM persistent kernel, 2 the cluster kernel; +4: a multi-rank solve done by replication (every rank gathers rho and
 * phi, solves the global problem with the single-GPU kernel and keeps its own sub-domain) */
int pincMgLastPath(void);
/* Cell-slotted particle storage: while the host loops pincAccMove3D1KE -> puExtractEmigrants3D -> puMigrate
 * -> puDistr3D1, every cell keeps its particles in slots of its own and only the particles that change cell move (instead of a
 * counting sort of the whole population every step); any other entry point first restores the contiguous cell-ordered planes.
 * on = 0 disables it (also $PINC_B200_SLOTTED=0); headroomPercent / extraSlots (>= 0, default 25 / 16): free slots kept per cell on
 * top of the fullest cell's count - a cell that outgrows them sends the population back to the sort for one step.
 * pincPopLayout: 1 while `pop` is slotted, else 0.  pincSlottedOverflows: how often that fallback ran (this process). */
void pincSetSlotted(int on, int headroomPercent, int extraSlots);
int pincPopLayout(const Population *pop);
long pincSlottedOverflows(void);
/* execution mode of single-rank periodic solves (same arithmetic per node in all of them):
 *   0 ops            one kernel per reference call, ghost layers exchanged as the reference does;
 *   1 fused          one persistent cooperative kernel over all SMs, a grid barrier and gBnd after every half-sweep;
 *   2 auto           (default) the same kernel with gBnd's mean subtraction applied once per smoother call, grid-wide
 *                    levels smoothed block-resident in shared memory (faces exchanged through tagged mailboxes in L2)
 *                    and levels of <= 4096 nodes inside one CTA; grids of <= 65536 nodes whose small levels are not
 *                    the cubic 16/8/4 pyramid run on one 16-CTA cluster with phi in distributed shared memory;
 *   3 cluster-exact  the cluster kernel with gBnd after every half-sweep;
 *   4 cluster-always as 2, but the cluster kernel is used whenever the levels fit its shared memory;
 *   5 allsm          as 2, but always the all-SM kernel.
 * Multi-rank solves: 2 = replicated (every rank gathers rho and phi of all ranks in one exchange, solves the global problem
 * with the single-GPU kernel and keeps its sub-domain; global grids of <= 16 M nodes), else and with
 * $PINC_B200_MG_REPLICA=0 distributed with smoother and ghost fills over NVLink peer memory and the V-cycle replayed as a
 * CUDA graph; 0/1/3 = distributed, one kernel and one exchange per reference call.
 * $PINC_B200_MG = ops | fused | cluster | cluster-exact | cluster-always | allsm selects the start-up value. */
void pincMgSetMode(int mode);

/* ---------------------------------------------------------------------------------
 * Host-struct constructors with plain arguments (restating gAlloc grid.c:413,
 * gAllocMpi grid.c:502, gCreateNeighborhood grid.c:1029, pAlloc population.c:42) for
 * hosts that do not link the reference's ini layer (tests, bench, other bindings).
 * All arrays are zero-initialised (quirk Q5).
 * ------------------------------------------------------------------------------- */
Grid *pincGridAlloc(int nDims, const int *trueSize, const int *nGhostLayers /*2*nDims*/,
                    int nValues, const int *bnd /*2*nDims, bndType*/);
void  pincGridFree(Grid *grid);
MpiInfo *pincMpiAlloc(int nDims, int nSpecies, const int *nSubdomains, const int *nGhostLayers,
                      const int *trueSize, int mpiRank, int mpiSize);
void  pincMpiFree(MpiInfo *mpiInfo);
void  pincCreateNeighborhood(MpiInfo *mpiInfo, const Grid *grid, const long int *nEmigrantsAlloc,
                             int nAllocEntries /*1, nDims or 3^nDims*/, const double *thresholds /*2*nDims*/);
Population *pincPopAlloc(int nSpecies, int nDims, const long int *nAllocPerRank,
                         const double *charge, const double *mass);
void  pincPopFree(Population *pop);

/* ---------------------------------------------------------------------------------
 * Initial conditions on the device (src/population.c:110-276, 367-428 with plain arguments; the reference reads
 * nParticles / trueSize / perturbAmplitude / perturbMode / drift / thermalVelocity from the ini).  Every rank
 * walks all global particles and keeps its own, so the particle set does not depend on the decomposition.
 * Random numbers: Philox4x32-10 (GSL's generators are not reproducible without GSL).  Host arrays are stale
 * afterwards; pop->iStop[] is current.
 * ------------------------------------------------------------------------------- */
void pincPosLattice(Population *pop, const MpiInfo *mpiInfo, const long int *nParticles, const int *trueSize);      /* population.c:172 */
void pincPosUniform(Population *pop, const MpiInfo *mpiInfo, const long int *nParticles, const int *trueSize,
                    unsigned long long seed);                                                                       /* population.c:110 */
void pincPosPerturb(Population *pop, const MpiInfo *mpiInfo, const double *amplitude, const double *mode,
                    const int *trueSize);                                                                           /* population.c:242 */
void pincVelMaxwell(Population *pop, const MpiInfo *mpiInfo, const double *drift, const double *thermalVelocity,
                    unsigned long long seed);                                                                       /* population.c:367 */
void pincVelZero(Population *pop);                                                                                  /* population.c:412 */

/* ---------------------------------------------------------------------------------
 * Device context, coherence, transport, timing
 * ------------------------------------------------------------------------------- */
typedef struct PincCtx PincCtx;
/* One context per rank.  In a PINC build there is one rank per process and the context
 * is created implicitly on the device given by $PINC_B200_DEVICE / $LOCAL_RANK / 0. */
PincCtx *pincCtxCreate(int device, int rank, int size);
void pincCtxMakeCurrent(PincCtx *ctx);          /* binds ctx to the calling host thread */
void pincCtxDestroy(PincCtx *ctx);
/* transports for size>1: (a) ranks are host threads of this process (tests on one GPU);
 * (b) ranks are processes, one GPU each, NCCL over NVLink. */
void pincCommInitThreads(PincCtx **ctxs, int n);
int  pincNcclUniqueId(char *out128);
void pincCommInitNccl(PincCtx *ctx, const char *uniqueId128);
const char *pincTransportName(void);            /* "self", "threads" or "nccl" */

void pincSyncGridToDevice(Grid *grid);
void pincSyncGridToHost(Grid *grid);
void pincSyncPopToDevice(Population *pop);
void pincSyncPopToHost(Population *pop);
int  pincHostRegister(void *ptr, size_t bytes); /* page-lock a caller-owned host array; 0 on success */
int  pincHostUnregister(void *ptr);
void pincForget(void *hostStruct);              /* drop the mirror of a Grid*/ /*Population* */
void pincDeviceSynchronize(void);
/* fused step helpers (same arithmetic as the separate entry points, one pass over the particles):
 * puAcc3D1KE immediately followed by puMove + puExtractEmigrants3D classification. */
void pincAccMove3D1KE(Population *pop, Grid *E, MpiInfo *mpiInfo);
/* ... and additionally the next step's puDistr3D1 for the particles that stay on this rank: their weights go into
 * rho's integer accumulators in the same pass; the following puDistr3D1(pop, rho) only adds the immigrants and
 * converts (bit-identical rho, the accumulation is order independent). */
void pincAccMoveDistr3D1KE(Population *pop, Grid *E, Grid *rho, MpiInfo *mpiInfo);

/* CUDA-event timing on the context's stream (what bench.py brackets the step with) */
void   pincTimerStart(void);
double pincTimerStopMs(void);
/* per-kernel-class accumulated device time since the last reset (names/ms), for roofline */
void   pincProfEnable(int on);
void   pincProfReset(void);
int    pincProfGet(int idx, char *name, int namelen, double *ms, long int *launches, double *algBytes);
long int pincLaunchCount(void);
const char *pincVersion(void);
int   pincLastError(char *buf, int len);

#ifdef __cplusplus
}
#endif
#endif
