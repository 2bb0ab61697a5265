/*
 * pinc_oracle.h — CPU restatement of PINC's per-timestep PIC loop.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is compared with; it is
 * never linked, imported or executed by the product (pinc_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check every function here against
 *   (i)  the reference's own known-answer vectors (test/pusher.test.c, test/grid.test.c,
 *        committed as tests/golden/kat_*.json), and
 *   (ii) outputs of the reference's own sources compiled in place (oracle/_ref, see
 *        oracle/Makefile) on seeded inputs, committed as tests/golden/ref_*.npz with the
 *        generating script tests/golden/make_ref_fixtures.py.
 *
 * Multi-rank semantics.  The reference is one MPI rank per sub-domain.  The oracle holds a
 * "world" of R sub-domains in one process and executes every phase for all ranks in
 * lock-step; exchanges are direct copies.  Arithmetic per rank is the reference's.
 *
 * Layout (as core.h): scalar grid val[j + sx*(k + sy*l)], vector grid val[c + 3*(j + sx*(k + sy*l))]
 * with ghost-inclusive sizes s = true + 2; particles AoS pos[3*i+d], species s in [iStart[s], iStop[s]).
 */
#ifndef PINC_ORACLE_H
#define PINC_ORACLE_H

typedef struct {
	int nRanks;
	int nSub[3];        /* nSubdomains */
	int trueSize[3];    /* per-rank true size */
} OrcTopo;

/* ---- single-rank particle kernels ---- */
void orc_move(double *pos, const double *vel, int nSpecies, const long *iStart, const long *iStop);
void orc_acc3d1(double *pos, double *vel, int nSpecies, const long *iStart, const long *iStop,
                const double *charge, const double *mass, double *E, const int *size,
                double *kinEnergy /* nSpecies, or NULL for the non-KE variant */);
void orc_boris3d1(double *pos, double *vel, int nSpecies, const long *iStart, const long *iStop,
                  const double *charge, const double *mass, double *E, const int *size,
                  const double *T, const double *S, double *kinEnergy, int bugCompatible);
void orc_rotation_parameters(int nSpecies, const double *BExt, const double *charge,
                             const double *mass, double *T, double *S);
void orc_distr3d1(const double *pos, int nSpecies, const long *iStart, const long *iStop,
                  const double *charge, double *rho, const int *size);
/* the N-dimensional / zeroth-order select() targets for nDims = 3 (src/pusher.c:215-391, 574-678); order = 1 or 0 */
void orc_acc_nd(double *pos, double *vel, int nSpecies, const long *iStart, const long *iStop,
                const double *charge, const double *mass, double *E, const int *size, int order, double *kinEnergy);
void orc_distr_nd(const double *pos, int nSpecies, const long *iStart, const long *iStop,
                  const double *charge, double *rho, const int *size, int order);
void orc_extract3d(double *pos, double *vel, int nSpecies, const long *iStart, long *iStop,
                   const double *thresholds, double **emigrants /*27*/, long *nEmigrants /*27*nSpecies*/);
int  orc_pnew(double *pos, double *vel, const long *iStart, long *iStop, int s, const double *p3, const double *v3);
void orc_pcut(double *pos, double *vel, long *iStop, int s, long p, double *p3, double *v3);
int  orc_neighbor_to_rank(const OrcTopo *t, int rank, int ne);
int  orc_neighbor_to_reciprocal(int ne);
int  orc_rank_to_neighbor(const OrcTopo *t, int rank, int other);
void orc_thresholds(const int *size, const double *thrIn, double *thrOut);

/* ---- world (all ranks, lock-step) ---- */
/* emigrants[r][ne], nEmigrants[r][ne*nSpecies+s]; particles are imported in ascending
 * neighbour index of the RECEIVER (the reference's arrival order is not deterministic). */
void orc_migrate(const OrcTopo *t, double **pos, double **vel, int nSpecies, long **iStop,
                 double ***emigrants, long **nEmigrants, long **nImmigrants);
/* op: 0 = set, 1 = add; dir: 0 = TOHALO, 1 = FROMHALO; d in 0..2; nValues 1 or 3 */
void orc_halo_dim(const OrcTopo *t, double **val, const int *size, int nValues, int d, int op, int dir);
void orc_halo(const OrcTopo *t, double **val, const int *size, int nValues, int op, int dir);
void orc_neutralize(const OrcTopo *t, double **val, const int *size);
void orc_findiff1st(const double *phi, double *E, const int *size);
void orc_gmul(double *val, long n, double num);
void orc_gs3d(const OrcTopo *t, double **phi, double **rho, const int *size, int nCycles);
/* boundary conditions (src/grid.c:608-662, 921-1023): bnd[8] = bndType per boundary index of a rank-4 grid (1..3 lower, 5..7
 * upper; 1 PERIODIC, 2 DIRICHLET, 3 NEUMANN); bndSlice[rank] = 8 slices of orc_slice_max(size) doubles */
long orc_slice_max(const int *size);
void orc_set_bnd_slices(const OrcTopo *t, int rank, const int *size, const int *bnd, double *bndSlice);
void orc_bnd(const OrcTopo *t, double **val, const int *size, const int *bnd, double **bndSlice);
void orc_gs3d_bnd(const OrcTopo *t, double **phi, double **rho, const int *size, int nCycles, const int *bnd, double **bndSlice);
void orc_residual(double *res, const double *rho, const double *phi, const int *size);
void orc_half_restrict3d(const double *fine, const int *fsize, double *coarse, const int *csize);
void orc_bilin_prol3d(const OrcTopo *t, double **fine, const int *fsize, double **coarse, const int *csize);
double orc_sum_true(const double *val, const int *size);
double orc_pot_energy(const double *rho, const double *phi, const int *size);

/* ---- multigrid solver with persistent coarse levels (quirk Q5: never re-zeroed) ---- */
typedef struct OrcMg OrcMg;
OrcMg *orc_mg_alloc(const OrcTopo *t, int nLevels, int nPre, int nPost, int nCoarse);
void   orc_mg_free(OrcMg *mg);
/* non-periodic edges: boundary types for every level; the (zeroed) slices of level q, rank r; mgRestrictBnd (multigrid.c:1314) */
void   orc_mg_set_bnd(OrcMg *mg, const int *bnd);
double *orc_mg_bnd_slice(OrcMg *mg, int level, int rank);
void   orc_mg_restrict_bnd(OrcMg *mg);
/* rho0/phi0/res0: per-rank finest-level arrays owned by the caller.  Returns the number of
 * V-cycles; barRes[c] receives the residual norm after V-cycle c (up to cap). */
int    orc_mg_solve(OrcMg *mg, double **rho0, double **phi0, double **res0, double tol,
                    int maxCycles, double *barRes, int cap);
void   orc_mg_vcycle(OrcMg *mg, double **rho0, double **phi0, double **res0);
/* the other select() targets of the solver: mgW (multigrid.c:1675), mgVRegular (:1559), and the smoothers of mgSetSolver (:28-83):
 * 0 gaussSeidelRB (mgGS3D), 1 jacobian (mgJacob3D :500), for the pre-, post- and coarse-level smoother */
void   orc_mg_wcycle(OrcMg *mg, double **rho0, double **phi0, double **res0);
void   orc_mg_vregular(OrcMg *mg, double **rho0, double **phi0, double **res0);
void   orc_mg_set_smoothers(OrcMg *mg, int pre, int post, int coarse);
void   orc_jacobi3d(const OrcTopo *t, double **phi, double **rho, const int *size, int nCycles, const int *bnd, double **bndSlice);
/* access to a coarse level (for parity checks): which = 0 rho, 1 phi, 2 res */
double *orc_mg_level(OrcMg *mg, int which, int level, int rank, int *sizeOut);

/* ---- whole step in the canonical order (src/main.c:197-274 minus objects/H5) ---- */
typedef struct {
	OrcTopo topo;
	int nSpecies;
	double charge[8], mass[8];
	double thresholds[6];      /* already converted (upper = size-1-thr) */
	int size[3];
	double **pos, **vel;       /* [rank] */
	long **iStart, **iStop;    /* [rank][nSpecies(+1)] */
	double **rho, **phi, **res, **E;
	double ***emigrants;       /* [rank][27] */
	long **nEmigrants, **nImmigrants;
	OrcMg *mg;
	double kinEnergy[9], potEnergy;
	int lastCycles;
	double lastBarRes[256];
} OrcSim;
void orc_step(OrcSim *s);
/* everything of a step except the trailing accelerate (used for the t=0 set-up, main.c:155-180) */
void orc_field_solve(OrcSim *s);
void orc_accelerate(OrcSim *s, double scaleE);

#endif
