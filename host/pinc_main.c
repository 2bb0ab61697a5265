/* pinc_main.c — C host of libpinc_b200: PINC's `regular()` run mode (/root/reference/src/main.c:50-304) on
 * top of the PINC-named entry points of include/pinc_b200.h, in the canonical order of SURVEY 8c (one rho fold
 * and one solve per step; no object calls).  Output: the per-step diagnostics on stdout and, when `files:output` is set, the
 * reference's HDF5 files (<prefix>rho/phi/E.grid.h5, <prefix>pop.pop.h5, <prefix>history.xy.h5; src/main.c:120-131, 262-266)
 * through the format-level writer of pinc_h5.c (libhdf5 is not in this image); single rank, see writeGrid below.
 *
 *   pinc_b200 input.ini [section:key=value ...]          (same command line as the reference's `pinc`)
 *
 * Reads the reference's .ini keys (pinc_ini.c), normalises like uAlloc/uNormalize (src/units.c:61-252), builds the
 * host structs with the plain-argument constructors, places particles (pPosLattice + pPosPerturb + pVelZero as
 * main.c:144-152 does, or uniform + Maxwellian when the ini gives thermal velocities), then time-steps.
 * Extensions (not PINC keys): population:thermalVelocityCells (sigma in cells/step), methods:fused = 1 (run
 * puAcc3D1KE + puMove + classification as one pass; 2: the deposition of the staying particles joins the pass), time:report = N, population:icOnDevice = 1 (initial conditions by the library's kernels), population:seed.
 * Multi-rank without MPI: RANK / WORLD_SIZE / LOCAL_RANK from the environment (as torchrun sets them) and the NCCL
 * id passed through the file $PINC_B200_ID_FILE. */
#define _POSIX_C_SOURCE 200809L
#include "pinc_ini.h"
#include "pinc_h5.h"
#include "../include/pinc_b200.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#define ELEMENTARY_CHARGE 1.60217733e-19       /* src/units.c:30-32 */
#define ELECTRON_MASS 9.10938188e-31
#define VACUUM_PERMITTIVITY 8.854187817e-12
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static void fail(const char *msg){ fprintf(stderr, "ERROR: %s\n", msg); exit(EXIT_FAILURE); }
static double nowSec(void){ struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9*t.tv_nsec; }

/* ---- uAlloc + uNormalize (src/units.c:61-252); same operation order as pinc_b200/config.py:normalize ---- */
static double g_unitLength = 1, g_unitVelocity = 1;          /* units->length, units->velocity (src/units.c:225, 244) */
static void normalize(Ini *ini){
	int nD = iniGetInt(ini, "grid:nDims"), nS = iniGetInt(ini, "population:nSpecies");
	int nSub[3], ts[3];
	iniGetInts(ini, "grid:nSubdomains", nD, nSub);
	iniGetInts(ini, "grid:trueSize", nD, ts);
	double L[3], V = 1, invL[3];
	for(int d = 0; d < nD; d++){ L[d] = (double)(nSub[d]*ts[d]); V *= L[d]; invL[d] = 1.0/L[d]; }
	iniApplySuffix(ini, "population:nParticles", "pc", &V, 1);           /* parseIndirectInput :138-158 */
	iniApplySuffix(ini, "population:nAlloc", "pc", &V, 1);
	iniApplySuffix(ini, "grid:nEmigrantsAlloc", "pc", &V, 1);
	iniApplySuffix(ini, "grid:stepSize", "tot", invL, nD);

	double charge[8], mass[8], dens[8], nPart[8], step[3];
	const char *method = iniRaw(ini, "methods:normalization");
	if(!strcmp(method, "semiSI")){                                        /* :159-189 */
		iniGetDoubles(ini, "population:charge", nS, charge);
		iniGetDoubles(ini, "population:mass", nS, mass);
		iniGetDoubles(ini, "population:density", nS, dens);
		double wpe = sqrt(ELEMENTARY_CHARGE*ELEMENTARY_CHARGE*dens[0]/(VACUUM_PERMITTIVITY*ELECTRON_MASS));
		for(int s = 0; s < nS; s++){ charge[s] = charge[s]*ELEMENTARY_CHARGE; mass[s] = mass[s]*ELECTRON_MASS; }
		iniSetDoubles(ini, "population:charge", nS, charge);
		iniSetDoubles(ini, "population:mass", nS, mass);
		double T = iniGetDouble(ini, "time:timeStep")/wpe;
		iniSetDoubles(ini, "time:timeStep", 1, &T);
	} else if(strcmp(method, "SI")) fail("methods:normalization not valid (must be SI or semiSI)");
	/* uSI :191-231 + derived units :233-252 */
	double T = iniGetDouble(ini, "time:timeStep");
	iniGetDoubles(ini, "grid:stepSize", nD, step);
	iniGetDoubles(ini, "population:nParticles", nS, nPart);
	iniGetDoubles(ini, "population:density", nS, dens);
	iniGetDoubles(ini, "population:charge", nS, charge);
	double Vphys = V*pow(step[0], nD);
	double w[8];
	for(int s = 0; s < nS; s++) w[s] = dens[s]*Vphys/(double)(long)nPart[s];
	double X = step[0];
	double Q = w[0]*fabs(charge[0]);
	double M = pow(T*Q, 2)/(VACUUM_PERMITTIVITY*pow(X, nD));
	double uVel = X/T, uDens = 1.0/pow(X, nD), uE = X*M/(T*T*Q), uB = M/(T*Q);
	g_unitLength = X; g_unitVelocity = uVel;
	iniGetDoubles(ini, "population:charge", nS, charge);
	iniGetDoubles(ini, "population:mass", nS, mass);
	iniGetDoubles(ini, "population:density", nS, dens);
	for(int s = 0; s < nS; s++){                          /* adScale(.., 1.0/unit): multiply by the reciprocal */
		charge[s] = charge[s]*w[s]*(1.0/Q);
		mass[s] = mass[s]*w[s]*(1.0/M);
		dens[s] = dens[s]/w[s]*(1.0/uDens);
	}
	iniSetDoubles(ini, "population:charge", nS, charge);
	iniSetDoubles(ini, "population:mass", nS, mass);
	iniSetDoubles(ini, "population:density", nS, dens);
	const char *keys[] = { "population:thermalVelocity", "population:drift", "population:perturbAmplitude", "fields:BExt", "fields:EExt" };
	double units[] = { uVel, uVel, X, uB, uE };
	for(int i = 0; i < 5; i++) if(iniHas(ini, keys[i])){
		int n = iniNElements(ini, keys[i]);
		double v[64];
		iniGetDoubles(ini, keys[i], n, v);
		for(int j = 0; j < n; j++) v[j] = v[j]*(1.0/units[i]);
		iniSetDoubles(ini, keys[i], n, v);
	}
}

/* ---- initial conditions (host side, once) ---- */
static unsigned long long rngS[4];
static unsigned long long rotl(unsigned long long x, int k){ return (x << k) | (x >> (64-k)); }
static unsigned long long rngNext(void){                 /* xoshiro256** */
	unsigned long long r = rotl(rngS[1]*5, 7)*9, t = rngS[1] << 17;
	rngS[2] ^= rngS[0]; rngS[3] ^= rngS[1]; rngS[1] ^= rngS[2]; rngS[0] ^= rngS[3]; rngS[2] ^= t; rngS[3] = rotl(rngS[3], 45);
	return r;
}
static void rngSeed(unsigned long long s){               /* splitmix64 */
	for(int i = 0; i < 4; i++){ s += 0x9e3779b97f4a7c15ULL; unsigned long long z = s; z = (z ^ (z >> 30))*0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27))*0x94d049bb133111ebULL; rngS[i] = z ^ (z >> 31); }
}
static double rngUniform(void){ return (rngNext() >> 11)*(1.0/9007199254740992.0); }
static double rngNormal(void){ double u = 1.0 - rngUniform(), v = rngUniform(); return sqrt(-2.0*log(u))*cos(2.0*M_PI*v); }

typedef struct {
	int nD, nS, nSub[3], ts[3], gl[6], rank, size, sub[3], off[3];
	long nPart[8], nAlloc[8];
	double charge[8], mass[8], thr[6], vth[8], drift[8], pertA[24], pertM[24];
} Cfg;

static void pNewLocal(Population *pop, int s, const double *pos, const double *vel){          /* population.c:430-450 */
	long p = pop->iStop[s];
	if(p >= pop->iStart[s+1]) fail("population:nAlloc too small for the initial particles of this rank");
	for(int d = 0; d < 3; d++){ pop->pos[3*p+d] = pos[d]; pop->vel[3*p+d] = vel[d]; }
	pop->iStop[s]++;
}
/* pPosLattice + pVelZero + pPosPerturb (population.c:172-276, 412): every rank walks the global lattice and keeps
 * the particles of its own sub-domain, as the reference does */
static void icLattice(const Cfg *c, Population *pop){
	double L[3], V = 1;
	for(int d = 0; d < 3; d++){ L[d] = (double)(c->nSub[d]*c->ts[d]); V *= L[d]; }
	for(int s = 0; s < c->nS; s++){
		long n = c->nPart[s];
		double l = pow(V/(double)n, 1.0/3.0);
		for(long i = 0; i < n; i++){
			double lin = l*(double)i, pos[3], vel[3] = {0, 0, 0};
			int mine = 1;
			for(int d = 0; d < 3; d++){
				pos[d] = fmod(lin, L[d]);
				lin = lin/L[d];
				int sd = (int)(pos[d]*(1.0/c->ts[d]));
				if(sd != c->sub[d]) mine = 0;
			}
			if(!mine) continue;
			/* perturbation in the global frame (pPosPerturb :242-276 = pToGlobalFrame, perturb, pToLocalFrame);
			 * positions are stored in the local frame, so the frame round trip comes first, as in the reference */
			for(int d = 0; d < 3; d++){ pos[d] -= (double)c->off[d]; pos[d] += (double)c->off[d]; }
			for(int d = 0; d < 3; d++){
				double theta = 2.0*M_PI*c->pertM[s*3+d]*pos[d]/L[d];
				pos[d] += c->pertA[s*3+d]*cos(theta);
			}
			for(int d = 0; d < 3; d++) pos[d] -= (double)c->off[d];
			pNewLocal(pop, s, pos, vel);
		}
	}
}
/* pPosUniform + pVelMaxwell (population.c:110-170, 367-392) with this file's RNG; each rank draws its own share */
static void icMaxwell(const Cfg *c, Population *pop, unsigned long long seed){
	rngSeed(seed*1000003ULL + (unsigned long long)c->rank);
	for(int s = 0; s < c->nS; s++){
		long n = c->nPart[s]/c->size;
		for(long i = 0; i < n; i++){
			double pos[3], vel[3];
			for(int d = 0; d < 3; d++){
				pos[d] = c->gl[d] + c->ts[d]*rngUniform();
				do vel[d] = c->vth[s]*rngNormal() + c->drift[s]; while(fabs(vel[d]) >= 1.0);     /* maxVel = 1 cell/step */
			}
			pNewLocal(pop, s, pos, vel);
		}
	}
}

/* ---- HDF5 output as src/main.c:120-131, 262-266 writes it (format-level writer, pinc_h5.c) ---------------------------------
 * openH5File's name rule (src/io.c:566-604): prefix "." -> "./", a prefix that does not end in '/' gets '_' appended. */
static PH5 *openH5(Ini *ini, const char *fName, const char *sub){
	const char *prefix = iniRaw(ini, "files:output");
	char path[1024];
	size_t n = strlen(prefix);
	const char *sep = !strcmp(prefix, ".") ? "/" : (n > 0 && prefix[n-1] != '/' ? "_" : "");
	snprintf(path, sizeof path, "%s%s%s.%s.h5", prefix, sep, fName, sub);
	PH5 *f = ph5Create(path);
	if(!f){ fprintf(stderr, "ERROR: Could not open or create '%s'\n", path); exit(EXIT_FAILURE); }
	return f;
}
/* gOpenH5 (src/grid.c:1210-1270): the two attributes; gWriteH5 (:1161): dataset "/n=%.1f" of the TRUE nodes, extents reversed
 * (z, y, x, component).  One rank writes the whole grid here (the reference writes every rank's hyperslab through MPI-IO). */
static PH5 *gridOpenH5(Ini *ini, const char *name){
	PH5 *f = openH5(ini, name, "grid");
	double denorm = 1.;
	ph5Attr(f, "Axis denormalization factor", &g_unitLength, 1);
	ph5Attr(f, "Quantity denormalization factor", &denorm, 1);
	return f;
}
static void gridWriteH5(PH5 *f, Grid *g, double n){
	pincSyncGridToHost(g);
	const int nv = g->size[0], *sz = g->size, *ts = g->trueSize, *gl = g->nGhostLayers;
	unsigned long long dims[4] = { (unsigned long long)ts[3], (unsigned long long)ts[2], (unsigned long long)ts[1], (unsigned long long)nv };
	char name[64];
	snprintf(name, sizeof name, "/n=%.1f", n);
	int d = ph5Dataset(f, name, 4, dims);
	if(d < 0){ fprintf(stderr, "ERROR: could not create dataset %s\n", name); exit(EXIT_FAILURE); }
	unsigned long long at = 0;
	const unsigned long long row = (unsigned long long)ts[1]*nv;          /* one x-row of true nodes with its components is contiguous in memory */
	for(int l = 0; l < ts[3]; l++) for(int k = 0; k < ts[2]; k++){
		const double *src = g->val + (long)nv*(gl[1] + (long)sz[1]*((k + gl[2]) + (long)sz[2]*(l + gl[3])));
		ph5Write(f, d, at, row, src);
		at += row;
	}
}
/* pOpenH5 / pWriteH5 (src/population.c:497-651): positions in the GLOBAL frame (pToGlobalFrame: + offset), one (N, 3) dataset per
 * species and step; the reference converts in place and back (quirk Q6), here a copy is converted so the particles keep their bits */
static PH5 *popOpenH5(Ini *ini, int nS){
	PH5 *f = openH5(ini, "pop", "pop");
	char name[64];
	for(int s = 0; s < nS; s++){
		snprintf(name, sizeof name, "/pos/specie %i", s); ph5Group(f, name);
		snprintf(name, sizeof name, "/vel/specie %i", s); ph5Group(f, name);
	}
	ph5Attr(f, "Position denormalization factor", &g_unitLength, 1);
	ph5Attr(f, "Velocity denormalization factor", &g_unitVelocity, 1);
	return f;
}
static void popWriteH5(PH5 *f, Population *pop, const int *offset, double posN, double velN){
	pincSyncPopToHost(pop);
	char name[64];
	for(int s = 0; s < pop->nSpecies; s++){
		long n = pop->iStop[s] - pop->iStart[s];
		if(n <= 0){ fprintf(stderr, "WARNING: No particles of specie %i to store in .h5-file\n", s); continue; }
		unsigned long long dims[2] = { (unsigned long long)n, 3 };
		double *glob = malloc((size_t)n*3*sizeof *glob);
		const double *p = pop->pos + 3*pop->iStart[s];
		for(long i = 0; i < n; i++) for(int d = 0; d < 3; d++) glob[3*i+d] = p[3*i+d] + (double)offset[d];
		snprintf(name, sizeof name, "/pos/specie %i/n=%.1f", s, posN);
		int d = ph5Dataset(f, name, 2, dims);
		if(d >= 0) ph5Write(f, d, 0, (unsigned long long)n*3, glob);
		free(glob);
		snprintf(name, sizeof name, "/vel/specie %i/n=%.1f", s, velN);
		d = ph5Dataset(f, name, 2, dims);
		if(d >= 0) ph5Write(f, d, 0, (unsigned long long)n*3, pop->vel + 3*pop->iStart[s]);
	}
}

int main(int argc, char **argv){
	if(argc < 2){ fprintf(stderr, "usage: %s input.ini [section:key=value ...]\n", argv[0]); return EXIT_FAILURE; }
	Ini *ini = iniLoad(argv[1]);
	if(!ini) fail("Failed to open input file");
	for(int i = 2; i < argc; i++) if(!iniApplyOverride(ini, argv[i])) fprintf(stderr, "WARNING: ignoring argument %s\n", argv[i]);

	/* plug-in selection by name, as select() does (src/io.h:105, src/main.c:55-79) */
	if(strcmp(iniRaw(ini, "methods:mode"), "regular")) fail("methods:mode must be regular");
	const char *acc = iniRaw(ini, "methods:acc");
	int withKE = !strcmp(acc, "puAcc3D1KE");
	if(!withKE && strcmp(acc, "puAcc3D1")) fail("methods:acc must be puAcc3D1 or puAcc3D1KE");
	if(strcmp(iniRaw(ini, "methods:distr"), "puDistr3D1")) fail("methods:distr must be puDistr3D1");
	if(strcmp(iniRaw(ini, "methods:migrate"), "puExtractEmigrants3D")) fail("methods:migrate must be puExtractEmigrants3D");
	if(strcmp(iniRaw(ini, "methods:poisson"), "mgSolver")) fail("methods:poisson must be mgSolver");
	int fused = iniHas(ini, "methods:fused") ? iniGetInt(ini, "methods:fused") : 0;

	normalize(ini);
	Cfg c; memset(&c, 0, sizeof c);
	c.nD = iniGetInt(ini, "grid:nDims"); c.nS = iniGetInt(ini, "population:nSpecies");
	if(c.nD != 3) fail("only grid:nDims=3 is supported");
	iniGetInts(ini, "grid:nSubdomains", 3, c.nSub); iniGetInts(ini, "grid:trueSize", 3, c.ts);
	iniGetInts(ini, "grid:nGhostLayers", 6, c.gl); iniGetDoubles(ini, "grid:thresholds", 6, c.thr);
	iniGetLongs(ini, "population:nParticles", c.nS, c.nPart); iniGetLongs(ini, "population:nAlloc", c.nS, c.nAlloc);
	iniGetDoubles(ini, "population:charge", c.nS, c.charge); iniGetDoubles(ini, "population:mass", c.nS, c.mass);
	if(iniHas(ini, "population:thermalVelocityCells")) iniGetDoubles(ini, "population:thermalVelocityCells", c.nS, c.vth);
	else if(iniHas(ini, "population:thermalVelocity")) iniGetDoubles(ini, "population:thermalVelocity", c.nS, c.vth);
	if(iniHas(ini, "population:drift")) iniGetDoubles(ini, "population:drift", c.nS, c.drift);
	if(iniHas(ini, "population:perturbAmplitude")) iniGetDoubles(ini, "population:perturbAmplitude", 3*c.nS, c.pertA);
	if(iniHas(ini, "population:perturbMode")) iniGetDoubles(ini, "population:perturbMode", 3*c.nS, c.pertM);
	char tok[64];
	for(int i = 0; i < 6; i++){ iniGetStr(ini, "grid:boundaries", i, 6, tok, sizeof tok); if(strcmp(tok, "PERIODIC")) fail("only PERIODIC boundaries are implemented"); }
	int nSteps = iniGetInt(ini, "time:nTimeSteps");
	int report = iniHas(ini, "time:report") ? iniGetInt(ini, "time:report") : 1;

	/* ranks */
	c.rank = getenv("RANK") ? atoi(getenv("RANK")) : 0;
	c.size = getenv("WORLD_SIZE") ? atoi(getenv("WORLD_SIZE")) : 1;
	if(c.size != c.nSub[0]*c.nSub[1]*c.nSub[2]) fail("The product of grid:nSubdomains does not match the number of ranks (WORLD_SIZE)");
	int dev = getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : (getenv("PINC_B200_DEVICE") ? atoi(getenv("PINC_B200_DEVICE")) : 0);
	PincCtx *ctx = pincCtxCreate(dev, c.rank, c.size);
	if(c.size > 1){
		const char *idFile = getenv("PINC_B200_ID_FILE");
		if(!idFile) fail("multi-rank runs need $PINC_B200_ID_FILE (path all ranks can read) for the NCCL id");
		char id[128], tmp[600];
		if(c.rank == 0){
			pincNcclUniqueId(id);
			snprintf(tmp, sizeof tmp, "%s.tmp", idFile);
			FILE *f = fopen(tmp, "wb"); if(!f) fail("cannot write the NCCL id file");
			fwrite(id, 1, 128, f); fclose(f); rename(tmp, idFile);
		} else {
			FILE *f = NULL;
			for(int tries = 0; tries < 6000 && !(f = fopen(idFile, "rb")); tries++){ struct timespec ts = {0, 10000000}; nanosleep(&ts, NULL); }
			if(!f || fread(id, 1, 128, f) != 128) fail("cannot read the NCCL id file");
			fclose(f);
		}
		pincCommInitNccl(ctx, id);
	}

	/* allocation: src/main.c:84-99 */
	int bnd[6] = { PERIODIC, PERIODIC, PERIODIC, PERIODIC, PERIODIC, PERIODIC };
	MpiInfo *mpiInfo = pincMpiAlloc(3, c.nS, c.nSub, c.gl, c.ts, c.rank, c.size);
	for(int d = 0; d < 3; d++){ c.sub[d] = mpiInfo->subdomain[d]; c.off[d] = mpiInfo->offset[d]; }
	long perRank[8];
	for(int s = 0; s < c.nS; s++) perRank[s] = (c.nAlloc[s] + c.size - 1)/c.size;
	Population *pop = pincPopAlloc(c.nS, 3, perRank, c.charge, c.mass);
	Grid *E = pincGridAlloc(3, c.ts, c.gl, -1, bnd), *rho = pincGridAlloc(3, c.ts, c.gl, 1, bnd), *phi = pincGridAlloc(3, c.ts, c.gl, 1, bnd);
	MultigridSolver *solver = pincMgAllocSolver(rho, phi, iniGetInt(ini, "multigrid:mgLevels"), iniGetInt(ini, "multigrid:mgCycles"),
		iniGetInt(ini, "multigrid:nPreSmooth"), iniGetInt(ini, "multigrid:nPostSmooth"), iniGetInt(ini, "multigrid:nCoarseSolve"));
	int nEA = iniNElements(ini, "grid:nEmigrantsAlloc");
	long nEmAlloc[27];
	iniGetLongs(ini, "grid:nEmigrantsAlloc", nEA, nEmAlloc);
	pincCreateNeighborhood(mpiInfo, rho, nEmAlloc, nEA, c.thr);
	char err[256];
	const char *plugins[] = { acc, "puDistr3D1", "puExtractEmigrants3D" };
	for(int i = 0; i < 3; i++) if(pincPuSanity(plugins[i], 3, c.gl, c.thr, 3, 1, err, sizeof err)) fail(err);

	/* initial conditions: src/main.c:144-152 */
	int anyPert = 0, anyVth = 0;
	for(int i = 0; i < 3*c.nS; i++) if(c.pertA[i] != 0) anyPert = 1;
	for(int s = 0; s < c.nS; s++) if(c.vth[s] != 0) anyVth = 1;
	double tIC = nowSec();
	unsigned long long seed = iniHas(ini, "population:seed") ? (unsigned long long)iniGetInt(ini, "population:seed") : 1;
	if(iniHas(ini, "population:icOnDevice") && iniGetInt(ini, "population:icOnDevice")){
		/* the same initial conditions generated by the library's kernels */
		if(anyVth && !anyPert){ pincPosUniform(pop, mpiInfo, c.nPart, c.ts, seed); pincVelMaxwell(pop, mpiInfo, c.drift, c.vth, seed + 1); }
		else { pincPosLattice(pop, mpiInfo, c.nPart, c.ts); pincPosPerturb(pop, mpiInfo, c.pertA, c.pertM, c.ts); pincVelZero(pop); }
		pincDeviceSynchronize();
	} else {
		if(anyVth && !anyPert) icMaxwell(&c, pop, seed);
		else icLattice(&c, pop);
		pincSyncPopToDevice(pop);
	}
	tIC = nowSec() - tIC;

	/* src/main.c:155-186 */
	puExtractEmigrants3D(pop, mpiInfo);
	puMigrate(pop, mpiInfo, rho);
	puDistr3D1(pop, rho);
	gHaloOp((funPtr)addSlice, rho, mpiInfo, FROMHALO);
	mgSolve(solver, rho, phi, mpiInfo);
	gHaloOp((funPtr)setSlice, phi, mpiInfo, TOHALO);
	gFinDiff1st(phi, E);
	gHaloOp((funPtr)setSlice, E, mpiInfo, TOHALO);
	gMul(E, -1.);
	gMul(E, 0.5);
	if(withKE) puAcc3D1KE(pop, E); else puAcc3D1(pop, E);
	gMul(E, 2.0);

	/* files: src/main.c:120-131 (opt-in here: only when files:output is given; particles only with files:particles = 1) */
	const int writeH5 = iniHas(ini, "files:output");
	const int writePop = writeH5 && iniHas(ini, "files:particles") && iniGetInt(ini, "files:particles");
	PH5 *h5Rho = NULL, *h5Phi = NULL, *h5E = NULL, *h5Pop = NULL, *h5Hist = NULL;
	int xyPot[9], xyKin[9];
	if(writeH5){
		if(c.size != 1) fail("files:output: the format-level HDF5 writer is single-rank (no MPI-IO); run one rank or leave files:output out");
		if(writePop) h5Pop = popOpenH5(ini, c.nS);
		h5Rho = gridOpenH5(ini, "rho"); h5Phi = gridOpenH5(ini, "phi"); h5E = gridOpenH5(ini, "E");
		h5Hist = openH5(ini, "history", "xy");
		char name[64];                                      /* pCreateEnergyDatasets (src/population.c:658-676) */
		xyPot[c.nS] = ph5XYCreate(h5Hist, "/energy/potential/total");
		xyKin[c.nS] = ph5XYCreate(h5Hist, "/energy/kinetic/total");
		for(int s = 0; s < c.nS; s++){
			snprintf(name, sizeof name, "/energy/potential/specie %i", s); xyPot[s] = ph5XYCreate(h5Hist, name);
			snprintf(name, sizeof name, "/energy/kinetic/specie %i", s); xyKin[s] = ph5XYCreate(h5Hist, name);
		}
	}

	long nLocal = 0;
	for(int s = 0; s < c.nS; s++) nLocal += pop->iStop[s] - pop->iStart[s];
	if(c.rank == 0) printf("STATUS: %ld particles on rank 0 of %d, grid %dx%dx%d per rank, %d steps%s\n", nLocal, c.size, c.ts[0], c.ts[1], c.ts[2], nSteps, fused ? " (fused particle pass)" : "");

	/* time loop: src/main.c:197-274 */
	int moved = 0;
	pincDeviceSynchronize();
	double t0 = nowSec();
	pincTimerStart();
	for(int n = 1; n <= nSteps; n++){
		if(!moved) puMove(pop, NULL);
		puExtractEmigrants3D(pop, mpiInfo);
		puMigrate(pop, mpiInfo, rho);
		puDistr3D1(pop, rho);
		gHaloOp((funPtr)addSlice, rho, mpiInfo, FROMHALO);
		mgSolve(solver, rho, phi, mpiInfo);
		gHaloOp((funPtr)setSlice, phi, mpiInfo, TOHALO);
		gFinDiff1st(phi, E);
		gHaloOp((funPtr)setSlice, E, mpiInfo, TOHALO);
		gMul(E, -1.);
		if(fused == 2 && withKE){ pincAccMoveDistr3D1KE(pop, E, rho, mpiInfo); moved = 1; }
		else if(fused && withKE){ pincAccMove3D1KE(pop, E, mpiInfo); moved = 1; }
		else if(withKE) puAcc3D1KE(pop, E);
		else puAcc3D1(pop, E);
		if(withKE) pSumKinEnergy(pop);
		gPotEnergy(rho, phi, pop);
		if(report > 0 && n % report == 0 && c.rank == 0)
			printf("n=%d kinetic=%.17g potential=%.17g particles=%ld\n", n, pop->kinEnergy[c.nS], pop->potEnergy[c.nS], (long)(pop->iStop[0]-pop->iStart[0]));
		if(writeH5){                                        /* src/main.c:262-266 */
			gridWriteH5(h5E, E, (double)n); gridWriteH5(h5Rho, rho, (double)n); gridWriteH5(h5Phi, phi, (double)n);
			if(writePop){
				if(moved) fail("files:particles with methods:fused: the fused pass leaves the positions one puMove ahead of step n");
				popWriteH5(h5Pop, pop, c.off, (double)n, (double)n + 0.5);
			}
			for(int s = 0; s <= c.nS; s++){                   /* pWriteEnergy (src/population.c:678-700) */
				ph5XYAppend(h5Hist, xyPot[s], (double)n, pop->potEnergy[s]);
				ph5XYAppend(h5Hist, xyKin[s], (double)n, pop->kinEnergy[s]);
			}
		}
	}
	double devMs = pincTimerStopMs();
	double wall = nowSec() - t0;
	double hist[256];
	int cyc = pincMgLastHistory(hist, 256);
	if(c.rank == 0){
		printf("TIMER: Time spent: %.3f s for %d steps (device %.3f ms/step), initial conditions %.2f s\n", wall, nSteps, devMs/nSteps, tIC);
		printf("{\"particle_steps_per_s\": %.6g, \"ms_per_step\": %.6g, \"particles_rank0\": %ld, \"ranks\": %d, \"vcycles_last_solve\": %d, \"launches\": %ld, \"transport\": \"%s\"}\n",
			(double)nLocal*c.size*nSteps/(devMs*1e-3), devMs/nSteps, nLocal, c.size, cyc, pincLaunchCount(), pincTransportName());
	}
	if(writeH5){
		int bad = 0;
		if(h5Pop) bad |= ph5Close(h5Pop);
		bad |= ph5Close(h5Rho); bad |= ph5Close(h5Phi); bad |= ph5Close(h5E); bad |= ph5Close(h5Hist);
		if(bad) fail("writing the .h5 files failed");
	}
	mgFreeSolver(solver);
	pincGridFree(E); pincGridFree(rho); pincGridFree(phi);
	pincPopFree(pop); pincMpiFree(mpiInfo);
	pincCtxDestroy(ctx);
	iniFree(ini);
	return 0;
}
