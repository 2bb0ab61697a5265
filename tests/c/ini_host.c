/* A PINC-style host with an ini layer of its own: defines the four accessors of src/io.h:228-240 over a fixed table
 * (what io.c does over iniparser) and drives the entry points of libpinc_b200.so that take `dictionary *ini`
 * (X_set selectors, mgSolver triple, mgAllocSolver, puGet3DRotationParameters).  No GPU needed.
 * argv[1] = "ok" | "badcycle" | "baddims". */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pinc_b200.h"

struct _dictionary_ { const char *mode; };
static const char *KV[][2] = {
	{"grid:nDims", "3"}, {"grid:nSpecies", "2"}, {"grid:nGhostLayers", "1,1,1,1,1,1"}, {"grid:thresholds", "0.1,0.1,0.1,0.1,0.1,0.1"},
	{"fields:BExt", "0,0,2"}, {"population:charge", "-1,1"}, {"population:mass", "1,4"},
	{"multigrid:mgLevels", "3"}, {"multigrid:mgCycles", "1"}, {"multigrid:nPreSmooth", "10"}, {"multigrid:nPostSmooth", "9"},
	{"multigrid:nCoarseSolve", "8"}, {"multigrid:preSmooth", "gaussSeidelRB"}, {"multigrid:postSmooth", "gaussSeidelRB"},
	{"multigrid:coarseSolver", "gaussSeidelRB"}, {"multigrid:restrictor", "halfWeight"}, {"multigrid:prolongator", "bilinear"},
	{"multigrid:cycle", "mgVRecursive"}, {NULL, NULL}};
static const char *look(const dictionary *ini, const char *key){
	if(!strcmp(ini->mode, "badcycle") && !strcmp(key, "multigrid:cycle")) return "mgFMG";
	if(!strcmp(ini->mode, "baddims") && !strcmp(key, "grid:nDims")) return "2";
	for(int i = 0; KV[i][0]; i++) if(!strcmp(KV[i][0], key)) return KV[i][1];
	fprintf(stderr, "missing key %s\n", key); exit(3);
}
int iniGetInt(const dictionary *ini, const char *key){ return atoi(look(ini, key)); }
char *iniGetStr(const dictionary *ini, const char *key){ const char *v = look(ini, key); char *r = malloc(strlen(v)+1); strcpy(r, v); return r; }
static int split(const char *v, double *out, int n){ int k = 0; char *c = malloc(strlen(v)+1), *t; strcpy(c, v);
	for(t = strtok(c, ","); t && k < n; t = strtok(NULL, ",")) out[k++] = atof(t); for(int i = k; i < n; i++) out[i] = out[i % (k ? k : 1)]; free(c); return k; }
int *iniGetIntArr(const dictionary *ini, const char *key, int n){ double tmp[16]; split(look(ini, key), tmp, n); int *r = malloc(n*sizeof *r); for(int i = 0; i < n; i++) r[i] = (int)tmp[i]; return r; }
double *iniGetDoubleArr(const dictionary *ini, const char *key, int n){ double *r = malloc(n*sizeof *r); split(look(ini, key), r, n); return r; }

int main(int argc, char **argv){
	dictionary ini = { argc > 1 ? argv[1] : "ok" };
	/* main.c:55-61: the selectors return the compute functions */
	if(puAcc3D1_set(&ini) != (funPtr)puAcc3D1 || puAcc3D1KE_set(&ini) != (funPtr)puAcc3D1KE) return 10;
	if(puDistr3D1_set(&ini) != (funPtr)puDistr3D1 || puExtractEmigrants3D_set(&ini) != (funPtr)puExtractEmigrants3D) return 11;
	/* main.c:58-70: the N-dimensional / zeroth-order targets (puAccND0_set hands out the KE form, as pusher.c:355 does) */
	if(puAccND1_set(&ini) != (funPtr)puAccND1 || puAccND1KE_set(&ini) != (funPtr)puAccND1KE || puAccND0_set(&ini) != (funPtr)puAccND0KE
	   || puAccND0KE_set(&ini) != (funPtr)puAccND0KE || puDistrND1_set(&ini) != (funPtr)puDistrND1 || puDistrND0_set(&ini) != (funPtr)puDistrND0
	   || puExtractEmigrantsND_set(&ini) != (funPtr)puExtractEmigrantsND) return 15;
	/* main.c:63-70: solver interface */
	void (*solverInterface)() = mgSolver_set(&ini);
	void (*solve)() = NULL; MultigridSolver *(*solverAlloc)() = NULL; void (*solverFree)() = NULL;
	solverInterface(&solve, &solverAlloc, &solverFree);
	if(solve != (void(*)())mgSolve || solverAlloc != (MultigridSolver*(*)())mgAllocSolver || solverFree != (void(*)())mgFreeSolver) return 12;
	int ts[3] = {16, 8, 8}, gl[6] = {1,1,1,1,1,1}, bnd[6] = {PERIODIC,PERIODIC,PERIODIC,PERIODIC,PERIODIC,PERIODIC};
	Grid *rho = pincGridAlloc(3, ts, gl, 1, bnd), *phi = pincGridAlloc(3, ts, gl, 1, bnd);
	MultigridSolver *solver = solverAlloc(&ini, rho, phi);                      /* main.c:99 */
	if(solver->mgRho->nLevels != 3 || solver->mgRho->nPreSmooth != 10 || solver->mgPhi->nPostSmooth != 9 || solver->mgRes->nCoarseSolve != 8) return 13;
	if(solver->mgRho->grids[0] != rho || solver->mgPhi->grids[0] != phi || solver->mgAlgo != (funPtr)mgVRecursive) return 14;
	if(solver->mgPhi->grids[2]->trueSize[1] != 4 || solver->mgPhi->grids[2]->size[3] != 4) return 15;
	solverFree(solver);
	double T[6], S[6];
	puGet3DRotationParameters(&ini, T, S);
	if(T[2] != -1.0 || T[5] != 0.25 || S[2] != -1.0 || S[5] != 0.5/1.0625 || T[0] != 0 || S[4] != 0) return 16;
	printf("ini-host-ok\n");
	return 0;
}
