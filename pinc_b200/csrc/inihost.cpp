// inihost.cpp — the entry points of the path that take PINC's `dictionary *ini`: the X_set(ini) selectors
// behind select() (src/io.h:105, src/pusher.c:143, 174, 508, 777), mgSolver_set (src/multigrid.c:398), mgAllocSolver
// (src/multigrid.c:364) and puGet3DRotationParameters (src/pusher.c:485).
//
// The ini layer (iniparser + src/io.c) stays host code.  These functions read the dictionary
// through the host's own accessors iniGetInt / iniGetStr / iniGetIntArr / iniGetDoubleArr (src/io.h:228-240), which are
// weak references here: a PINC build that links this library resolves them from its io.o; a host without PINC's ini
// layer (the Python driver, host/pinc_main.c) uses the plain-argument forms (pincPuSanity, pincMgAllocSolver,
// pincGet3DRotationParameters) and gets a PINC-B200 ERROR if it calls these.
#include "common.h"
#include <cstring>
#include <cstdlib>

extern "C" {
int iniGetInt(const dictionary *ini, const char *key) __attribute__((weak));
char *iniGetStr(const dictionary *ini, const char *key) __attribute__((weak));
int *iniGetIntArr(const dictionary *ini, const char *key, int nElements) __attribute__((weak));
double *iniGetDoubleArr(const dictionary *ini, const char *key, int nElements) __attribute__((weak));
}

using namespace pinc;

static void needIni(const char *who){
	if(!iniGetInt || !iniGetStr || !iniGetIntArr || !iniGetDoubleArr)
		fatal("%s(ini) needs the host program's ini layer (iniGetInt, iniGetStr, iniGetIntArr, iniGetDoubleArr of src/io.c); "
		      "hosts without it call the plain-argument form (pincPuSanity, pincMgAllocSolver, pincGet3DRotationParameters)", who);
}
// puSanity (src/pusher.c:1240-1283) on the values of the dictionary
static funPtr selector(const dictionary *ini, const char *name, int dim, int order, funPtr f){
	needIni(name);
	int nDims = iniGetInt(ini, "grid:nDims");
	int *gl = iniGetIntArr(ini, "grid:nGhostLayers", 2*nDims);
	double *th = iniGetDoubleArr(ini, "grid:thresholds", 2*nDims);
	char err[256];
	int bad = pincPuSanity(name, nDims, gl, th, dim, order, err, sizeof err);
	free(gl); free(th);
	if(bad) fatal("%s", err);
	return f;
}
static void wantStr(const dictionary *ini, const char *key, const char *only){
	char *v = iniGetStr(ini, key);
	bool ok = v && !strcmp(v, only);
	if(!ok) fatal("%s = %s: libpinc_b200 implements %s only (SURVEY 8f-3)", key, v ? v : "(missing)", only);
	free(v);
}

extern "C" {

funPtr puAcc3D1_set(dictionary *ini){ return selector(ini, "puAcc3D1", 3, 1, (funPtr)puAcc3D1); }                 /* pusher.c:143 */
funPtr puAcc3D1KE_set(dictionary *ini){ return selector(ini, "puAcc3D1KE", 3, 1, (funPtr)puAcc3D1KE); }           /* pusher.c:174 */
funPtr puDistr3D1_set(dictionary *ini){ return selector(ini, "puDistr3D1", 3, 1, (funPtr)puDistr3D1); }           /* pusher.c:508 */
funPtr puExtractEmigrants3D_set(const dictionary *ini){                                                              /* pusher.c:777 */
	needIni("puExtractEmigrants3D_set");
	if(iniGetInt(ini, "grid:nDims") != 3) fatal("puExtractEmigrants3D requires grid:nDims=3");
	return (funPtr)puExtractEmigrants3D;
}
/* the N-dimensional and zeroth-order select() targets (pusher.c:215-391, 574-678, 857-862); nDims = 3 only */
static funPtr selectorND(dictionary *ini, const char *name, int order, funPtr f){
	needIni(name);
	if(iniGetInt(ini, "grid:nDims") != 3) fatal("%s: libpinc_b200 is 3-D only (grid:nDims=%d)", name, iniGetInt(ini, "grid:nDims"));
	return selector(ini, name, 0, order, f);
}
funPtr puAccND1_set(dictionary *ini){ return selectorND(ini, "puAccND1", 1, (funPtr)puAccND1); }
funPtr puAccND1KE_set(dictionary *ini){ return selectorND(ini, "puAccND1KE", 1, (funPtr)puAccND1KE); }
funPtr puAccND0_set(dictionary *ini){ return selectorND(ini, "puAccND0", 0, (funPtr)puAccND0KE); }       /* (the reference returns the KE form here, pusher.c:355) */
funPtr puAccND0KE_set(dictionary *ini){ return selectorND(ini, "puAccND0KE", 0, (funPtr)puAccND0KE); }
funPtr puDistrND1_set(dictionary *ini){ return selectorND(ini, "puDistrND1", 1, (funPtr)puDistrND1); }
funPtr puDistrND0_set(dictionary *ini){ return selectorND(ini, "puDistrND0", 0, (funPtr)puDistrND0); }
funPtr puExtractEmigrantsND_set(const dictionary *ini){ (void)ini; return (funPtr)puExtractEmigrantsND; }
funPtr mgSolver_set(const dictionary *ini){ (void)ini; return (funPtr)mgSolver; }                                    /* multigrid.c:398 */

/* multigrid.c:364-382 with mgAlloc's reads and checks (:297-349) and the method names of mgSetSolver,
 * mgSetRestrictProlong, getMgAlgo (:26-125) */
MultigridSolver *mgAllocSolver(const dictionary *ini, Grid *rho, Grid *phi){
	needIni("mgAllocSolver");
	int nLevels = iniGetInt(ini, "multigrid:mgLevels");
	int nMGCycles = iniGetInt(ini, "multigrid:mgCycles");
	int nPre = iniGetInt(ini, "multigrid:nPreSmooth");
	int nPost = iniGetInt(ini, "multigrid:nPostSmooth");
	int nCoarse = iniGetInt(ini, "multigrid:nCoarseSolve");
	wantStr(ini, "multigrid:restrictor", "halfWeight");
	wantStr(ini, "multigrid:prolongator", "bilinear");
	MultigridSolver *s = pincMgAllocSolver(rho, phi, nLevels, nMGCycles, nPre, nPost, nCoarse);      /* the same sanity checks as mgAlloc */
	/* mgSetSolver (multigrid.c:28-83) and getMgAlgo (:113-125) for what the library provides */
	typedef void (*Smooth)(Grid*, const Grid*, const int, const MpiInfo*);
	auto smoother = [&](const char *key) -> Smooth {
		char *v = iniGetStr(ini, key);
		Smooth f = nullptr;
		if(v && !strcmp(v, "gaussSeidelRB")) f = mgGS3D;
		else if(v && !strcmp(v, "jacobian")) f = mgJacob3D;
		else fatal("%s = %s: libpinc_b200 provides gaussSeidelRB and jacobian (3-D)", key, v ? v : "(missing)");
		free(v);
		return f;
	};
	Smooth pre = smoother("multigrid:preSmooth"), post = smoother("multigrid:postSmooth"), coarse = smoother("multigrid:coarseSolver");
	Multigrid *mgs[3] = { s->mgRho, s->mgPhi, s->mgRes };
	for(Multigrid *mg : mgs){ mg->preSmooth = pre; mg->postSmooth = post; mg->coarseSolv = coarse; }
	char *cyc = iniGetStr(ini, "multigrid:cycle");
	if(cyc && !strcmp(cyc, "mgVRecursive")) s->mgAlgo = (funPtr)mgVRecursive;
	else if(cyc && !strcmp(cyc, "mgVRegular")) s->mgAlgo = (funPtr)mgVRegular;
	else if(cyc && !strcmp(cyc, "mgW")) s->mgAlgo = (funPtr)mgW;
	else fatal("multigrid:cycle = %s: libpinc_b200 provides mgVRecursive, mgVRegular and mgW (the reference's mgFMG destroys mgRho->grids[0])", cyc ? cyc : "(missing)");
	free(cyc);
	return s;
}

void puGet3DRotationParameters(dictionary *ini, double *T, double *S){                                                /* pusher.c:485 */
	needIni("puGet3DRotationParameters");
	int nDims = iniGetInt(ini, "grid:nDims");
	int nSpecies = iniGetInt(ini, "grid:nSpecies");
	double *BExt = iniGetDoubleArr(ini, "fields:BExt", nDims);
	double *charge = iniGetDoubleArr(ini, "population:charge", nSpecies);
	double *mass = iniGetDoubleArr(ini, "population:mass", nSpecies);
	pincGet3DRotationParameters(nSpecies, BExt, charge, mass, T, S);
	free(BExt); free(charge); free(mass);
}

} // extern "C"
