/* Drop-in proof (TEST INFRASTRUCTURE; built by oracle/Makefile `make dropin`, only where /root/reference exists).
 *
 * A PINC host in C compiled against the REFERENCE'S OWN headers (src/core.h, pusher.h, multigrid.h) and linked against the reference's own
 * io.o, units.o, aux.o, population.o, grid.o and iniparser - i.e. its ini layer (iniOpen, select(), iniGet*), unit
 * normalisation, allocators (gAllocMpi, pAlloc, gAlloc, gCreateNeighborhood, gSetBndSlices) and initial conditions
 * (pPosLattice, pPosPerturb, pVelZero) - while every compute entry point of the time step resolves to
 * libpinc_b200.so (the reference's pusher.o and multigrid.o are not linked; the compute functions of grid.o and
 * population.o that the library replaces are made file-local with objcopy, which is what compiling those files with
 * the -DPINC_B200 guards of INTEGRATION.md section 1 amounts to).  The sequence is regular() of src/main.c:50-290 minus
 * the object/capacitance lines (object.c does not compile, SURVEY finding 1) and the HDF5 writers.
 *
 * The same source built with -DDROPIN_REFERENCE links the reference's pusher.o and multigrid.o instead of the library:
 * the CPU twin whose energies the test compares with.
 *
 *   dropin_host <file.ini> [section:key=value ...]       prints one line per time step: n, kinetic, potential, particles
 */
#include "core.h"
#include "pusher.h"
#include "multigrid.h"

#ifndef DROPIN_REFERENCE
/* the library's own additions used here (include/pinc_b200.h; its struct definitions are the reference's, so only prototypes) */
const char *pincVersion(void);
int pincMgLastPath(void);
void pincSyncPopToHost(Population *pop);
void pincDeviceSynchronize(void);
#endif

int main(int argc, char *argv[]){
	MPI_Init(&argc, &argv);
	dictionary *ini = iniOpen(argc, argv);                       /* src/main.c:31, src/io.c */

	/* src/main.c:55-80: method selection through the reference's select() (src/io.c:115-168) */
	void (*acc)() = select(ini, "methods:acc", puAcc3D1_set, puAcc3D1KE_set, puAccND1_set, puAccND1KE_set, puAccND0_set, puAccND0KE_set);
	void (*distr)() = select(ini, "methods:distr", puDistr3D1_set, puDistrND1_set, puDistrND0_set);
	void (*extractEmigrants)() = select(ini, "methods:migrate", puExtractEmigrants3D_set, puExtractEmigrantsND_set);
	void (*solverInterface)() = select(ini, "methods:poisson", mgSolver_set);
	void (*solve)() = NULL;
	void *(*solverAlloc)() = NULL;
	void (*solverFree)() = NULL;
	solverInterface(&solve, &solverAlloc, &solverFree);

	/* src/main.c:85-107: the reference's allocators */
	Units *units = uAlloc(ini);
	uNormalize(ini, units);
	MpiInfo *mpiInfo = gAllocMpi(ini);
	Population *pop = pAlloc(ini);
	Grid *E = gAlloc(ini, VECTOR);
	Grid *rho = gAlloc(ini, SCALAR);
	Grid *phi = gAlloc(ini, SCALAR);
	void *solver = solverAlloc(ini, rho, phi);
	gCreateNeighborhood(ini, mpiInfo, rho);
	gSetBndSlices(phi, mpiInfo);

	/* src/main.c:144-160: initial conditions on the host arrays, then migration */
	pPosLattice(ini, pop, mpiInfo);
	pVelZero(pop);
	pPosPerturb(ini, pop, mpiInfo);
	double maxVel = iniGetDouble(ini, "population:maxVel");
	extractEmigrants(pop, mpiInfo);
	puMigrate(pop, mpiInfo, rho);

	/* src/main.c:172-187 */
	distr(pop, rho);
	gHaloOp(addSlice, rho, mpiInfo, FROMHALO);
	solve(solver, rho, phi, mpiInfo);
	gFinDiff1st(phi, E);
	gHaloOp(setSlice, E, mpiInfo, TOHALO);
	gMul(E, -1.);
	gMul(E, 0.5);
	acc(pop, E);
	gMul(E, 2.0);

	int nTimeSteps = iniGetInt(ini, "time:nTimeSteps");
	for(int n = 1; n <= nTimeSteps; n++){                        /* src/main.c:197-274 */
		pVelAssertMax(pop, maxVel);
		puMove(pop, NULL);
		extractEmigrants(pop, mpiInfo);
		puMigrate(pop, mpiInfo, rho);
		pPosAssertInLocalFrame(pop, rho);
		distr(pop, rho);
		gHaloOp(addSlice, rho, mpiInfo, FROMHALO);
		solve(solver, rho, phi, mpiInfo);
		gHaloOp(setSlice, phi, mpiInfo, TOHALO);
		gFinDiff1st(phi, E);
		gHaloOp(setSlice, E, mpiInfo, TOHALO);
		gMul(E, -1.);
		acc(pop, E);
		pSumKinEnergy(pop);
		gPotEnergy(rho, phi, pop);
		int nS = pop->nSpecies;
		long np = 0;
		for(int s = 0; s < nS; s++) np += pop->iStop[s] - pop->iStart[s];
		printf("n=%d kinetic=%.17g potential=%.17g particles=%ld\n", n, pop->kinEnergy[nS], pop->potEnergy[nS], np);
	}
#ifndef DROPIN_REFERENCE
	pincSyncPopToHost(pop);                                      /* what a host does before pWriteH5 */
	printf("library=%s mg_path=%d\n", pincVersion(), pincMgLastPath());
#else
	printf("library=reference\n");
#endif
	double cs = 0;                                               /* checksum of the final phase space as the host sees it */
	for(int s = 0; s < pop->nSpecies; s++)
		for(long i = pop->iStart[s]*3; i < pop->iStop[s]*3; i++) cs += pop->pos[i] + 1e3*pop->vel[i];
	printf("checksum=%.15g\n", cs);
	solverFree(solver);
	gFreeMpi(mpiInfo);
	iniClose(ini);
	MPI_Finalize();
	return 0;
}
