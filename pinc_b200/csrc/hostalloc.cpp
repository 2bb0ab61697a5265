// hostalloc.cpp — host-side pieces of the C-ABI that need no GPU: neighbour maps (src/pusher.c:1181-1231),
// the validator behind puXxx_set (src/pusher.c:1047), and plain-argument constructors of the reference's
// host structs (src/grid.c:413-545, 1029-1132; src/population.c:42-92; src/multigrid.c:128-382) for hosts
// that do not link the reference's ini layer.  All arrays are zero-initialised (quirk Q5).
#include "common.h"
#include <cmath>

namespace pinc {

int neighborToRank(const MpiInfo *m, int ne){
	int rank = 0;
	for(int d = 0; d < m->nDims; d++){
		int n = (ne % 3) - 1;
		ne /= 3;
		n = (m->subdomain[d] + n + m->nSubdomains[d]) % m->nSubdomains[d];
		rank += n*m->nSubdomainsProd[d];
	}
	return rank;
}
int neighborToReciprocal(int ne, int nDims){
	int rec = 0, pw = 1;
	for(int d = 0; d < nDims; d++){ rec += (2 - (ne % 3))*pw; ne /= 3; pw *= 3; }
	return rec;
}
int rankToNeighbor(const MpiInfo *m, int rank){
	int ne = 0, pw = 1;
	for(int d = 0; d < m->nDims; d++){
		int n = rank % m->nSubdomains[d];
		n = (n - m->subdomain[d] + 1 + m->nSubdomains[d]) % m->nSubdomains[d];
		rank /= m->nSubdomains[d];
		ne += n*pw; pw *= 3;
	}
	return ne;
}

template<class T> static T *zalloc(size_t n){ return (T*)calloc(n ? n : 1, sizeof(T)); }

} // namespace pinc

using namespace pinc;

extern "C" {

int puNeighborToRank(MpiInfo *mpiInfo, int neighbor){ return neighborToRank(mpiInfo, neighbor); }
int puRankToNeighbor(MpiInfo *mpiInfo, int rank){ return rankToNeighbor(mpiInfo, rank); }
int puNeighborToReciprocal(int neighbor, int nDims){ return neighborToReciprocal(neighbor, nDims); }

// src/pusher.c:1047-1087 with the ini lookups replaced by arguments; returns 0 if acceptable
int pincPuSanity(const char *name, int nDims, const int *nGhostLayers, const double *thresholds,
                 int dim, int order, char *errbuf, int errlen){
	int minLayers = nGhostLayers[0];
	double minThr = thresholds[0], maxThr = thresholds[0];
	for(int i = 1; i < 2*nDims; i++){
		if(nGhostLayers[i] < minLayers) minLayers = nGhostLayers[i];
		if(thresholds[i] < minThr) minThr = thresholds[i];
		if(thresholds[i] > maxThr) maxThr = thresholds[i];
	}
	int reqLayers = order == 0 ? 0 : 1;
	double reqMinThr = order == 0 ? -0.5 : (order == 2 ? 0.5 : 0.0);
	auto err = [&](int code, const char *fmt, double v){ if(errbuf && errlen > 0) snprintf(errbuf, errlen, fmt, name, v); return code; };
	if(nDims != dim && dim != 0) return err(1, "%s only supports grid:nDims=%.0f", (double)dim);
	if(minLayers < 1) return err(2, "%s requires grid:nGhostLayers >=%.0f", (double)reqLayers);
	if(minThr < reqMinThr) return err(3, "%s requires grid:thresholds >=%.1f", reqMinThr);
	if(maxThr > minLayers - 0.5) return err(4, "%s requires grid:thresholds <= grid:nGhostLayers - 0.5 (%.1f)", minLayers - 0.5);
	if(errbuf && errlen > 0) errbuf[0] = 0;
	return 0;
}

// src/grid.c:413-500
Grid *pincGridAlloc(int nDims, const int *trueSizeIn, const int *nGhostLayersIn, int nValues, const int *bndIn){
	int rank = nDims + 1;
	Grid *g = zalloc<Grid>(1);
	g->rank = rank;
	g->size = zalloc<int>(rank); g->trueSize = zalloc<int>(rank); g->nGhostLayers = zalloc<int>(2*rank);
	if(nValues < 0) nValues = nDims;                          // VECTOR == -1
	g->size[0] = nValues; g->trueSize[0] = nValues;
	for(int d = 1; d < rank; d++){
		g->trueSize[d] = trueSizeIn[d-1];
		g->nGhostLayers[d] = nGhostLayersIn[d-1];
		g->nGhostLayers[d+rank] = nGhostLayersIn[d+nDims-1];
		g->size[d] = g->trueSize[d] + g->nGhostLayers[d] + g->nGhostLayers[d+rank];
	}
	g->sizeProd = zalloc<long int>(rank+1);
	g->sizeProd[0] = 1;
	for(int d = 0; d < rank; d++) g->sizeProd[d+1] = g->sizeProd[d]*g->size[d];
	long nSliceMax = 0;
	for(int d = 0; d < rank; d++){ long ns = g->sizeProd[rank]/g->size[d]; if(ns > nSliceMax) nSliceMax = ns; }
	g->val = zalloc<double>(g->sizeProd[rank]);
	g->sendSlice = zalloc<double>(nSliceMax);
	g->recvSlice = zalloc<double>(nSliceMax);
	g->bndSlice = zalloc<double>(2*rank*nSliceMax);           // src/grid.c:467 (there uninitialised; zero here)
	g->bnd = (bndType*)zalloc<int>(2*rank);
	int b = 0;
	for(int r = 0; r < 2*rank; r++){
		if(r % rank == 0) g->bnd[r] = NONE;
		else g->bnd[r] = bndIn ? (bndType)bndIn[b++] : PERIODIC;
	}
	return g;
}
void pincGridFree(Grid *g){
	if(!g) return;
	pincForget(g);
	free(g->size); free(g->trueSize); free(g->nGhostLayers); free(g->sizeProd);
	free(g->val); free(g->sendSlice); free(g->recvSlice); free(g->bnd); free(g->bndSlice);
	free(g);
}

// src/grid.c:502-545 (+ getSubdomain :149-176)
MpiInfo *pincMpiAlloc(int nDims, int nSpecies, const int *nSubdomains, const int *nGhostLayers,
                      const int *trueSize, int mpiRank, int mpiSize){
	int prod = 1;
	for(int d = 0; d < nDims; d++) prod *= nSubdomains[d];
	if(prod != mpiSize) fatal("The product of grid:nSubdomains does not match the number of ranks");
	MpiInfo *m = zalloc<MpiInfo>(1);
	m->mpiRank = mpiRank; m->mpiSize = mpiSize; m->nDims = nDims; m->nSpecies = nSpecies;
	m->subdomain = zalloc<int>(nDims); m->nSubdomains = zalloc<int>(nDims);
	m->nSubdomainsProd = zalloc<int>(nDims+1); m->offset = zalloc<int>(nDims);
	m->posToSubdomain = zalloc<double>(nDims);
	int r = mpiRank;
	m->nSubdomainsProd[0] = 1;
	for(int d = 0; d < nDims; d++){
		m->nSubdomains[d] = nSubdomains[d];
		m->nSubdomainsProd[d+1] = m->nSubdomainsProd[d]*nSubdomains[d];
		m->subdomain[d] = r % nSubdomains[d];
		r /= nSubdomains[d];
		m->offset[d] = m->subdomain[d]*trueSize[d] - nGhostLayers[d];
		m->posToSubdomain[d] = (double)1/trueSize[d];
	}
	m->nNeighbors = 0;
	return m;
}
void pincMpiFree(MpiInfo *m){
	if(!m) return;
	free(m->subdomain); free(m->nSubdomains); free(m->nSubdomainsProd); free(m->offset); free(m->posToSubdomain);
	free(m->nEmigrants); free(m->nEmigrantsAlloc); free(m->nImmigrants); free(m->thresholds);
	free(m->emigrants); free(m->migrants);
	free(m);
}

// src/grid.c:1029-1132.  The per-neighbour host buffers of the reference (emigrants[ne], immigrants) are not
// allocated: migrants stay on the device (DevPop::d_emig / d_immig, sized on demand).
void pincCreateNeighborhood(MpiInfo *m, const Grid *grid, const long int *nAllocIn, int nEntries, const double *thresholdsIn){
	int nDims = m->nDims;
	int nNeighbors = 1, center = 0, pw = 1;
	for(int d = 0; d < nDims; d++){ nNeighbors *= 3; center += pw; pw *= 3; }
	if(nEntries != nNeighbors && nEntries != 1 && nEntries != nDims)
		fatal("grid:nEmigrantsAlloc must consist of 1, nDims=%i or 3^nDims=%i elements", nDims, nNeighbors);
	m->nEmigrantsAlloc = zalloc<long int>(nNeighbors);
	for(int ne = 0; ne < nNeighbors; ne++){
		if(ne == center){ m->nEmigrantsAlloc[ne] = 0; continue; }
		if(nEntries == 1) m->nEmigrantsAlloc[ne] = nAllocIn[0];
		else if(nEntries == nNeighbors) m->nEmigrantsAlloc[ne] = nAllocIn[ne];
		else {
			int t = ne, interfaceDims = nDims;
			for(int d = nDims-1; d >= 0; d--){
				int power = 1; for(int i = 0; i < d; i++) power *= 3;
				if(t/power != 1) interfaceDims--;
				t %= power;
			}
			m->nEmigrantsAlloc[ne] = nAllocIn[interfaceDims];
		}
	}
	m->thresholds = zalloc<double>(2*nDims);
	for(int i = 0; i < 2*nDims; i++) m->thresholds[i] = thresholdsIn[i];
	for(int i = nDims; i < 2*nDims; i++) m->thresholds[i] = (grid->size[i%nDims+1]-1) - m->thresholds[i];
	m->nEmigrants = zalloc<long int>((size_t)nNeighbors*m->nSpecies);
	m->nImmigrants = zalloc<long int>((size_t)nNeighbors*m->nSpecies);
	long mx = 0;
	for(int ne = 0; ne < nNeighbors; ne++) if(m->nEmigrantsAlloc[ne] > mx) mx = m->nEmigrantsAlloc[ne];
	m->nImmigrantsAlloc = 2*nDims*mx;
	m->emigrants = zalloc<double*>(nNeighbors);
	m->migrants = zalloc<long int*>(nNeighbors);
	m->immigrants = nullptr; m->emigrantsDummy = nullptr; m->migrantsDummy = nullptr;
	m->send = nullptr; m->recv = nullptr;
	m->nNeighbors = nNeighbors;
	m->neighborhoodCenter = center;
}

// src/population.c:42-92; nAllocPerRank is what pAlloc derives as ceil(nAlloc/size)
Population *pincPopAlloc(int nSpecies, int nDims, const long int *nAllocPerRank, const double *charge, const double *mass){
	Population *p = zalloc<Population>(1);
	p->nSpecies = nSpecies; p->nDims = nDims;
	p->iStart = zalloc<long int>(nSpecies+1); p->iStop = zalloc<long int>(nSpecies);
	for(int s = 1; s <= nSpecies; s++) p->iStart[s] = p->iStart[s-1] + nAllocPerRank[s-1];
	for(int s = 0; s < nSpecies; s++) p->iStop[s] = p->iStart[s];
	size_t n = (size_t)nDims*p->iStart[nSpecies];
	p->pos = zalloc<double>(n); p->vel = zalloc<double>(n);
	if(!p->pos || !p->vel) fatal("pincPopAlloc: out of host memory for %zu doubles", 2*n);
	p->objVicinity = nullptr; p->collisions = nullptr;       // object code is out of scope
	p->kinEnergy = zalloc<double>(nSpecies+1); p->potEnergy = zalloc<double>(nSpecies+1);
	p->charge = zalloc<double>(nSpecies); p->mass = zalloc<double>(nSpecies);
	for(int s = 0; s < nSpecies; s++){ p->charge[s] = charge[s]; p->mass[s] = mass[s]; }
	return p;
}
void pincPopFree(Population *p){
	if(!p) return;
	pincForget(p);
	free(p->pos); free(p->vel); free(p->iStart); free(p->iStop); free(p->kinEnergy); free(p->potEnergy);
	free(p->charge); free(p->mass);
	free(p);
}

// src/multigrid.c:128-206 (mgAllocSubGrids), :297-349 (mgAlloc), :364-382 (mgAllocSolver)
static Multigrid *mgAllocPlain(Grid *grid, int nLevels, int nCycles, int nPre, int nPost, int nCoarse){
	if(nLevels < 1) fatal("Multi Grid levels is 0, need 1 grid level");
	if(!nCycles) fatal("MG cycles is 0");
	int nDims = grid->rank - 1, rank = grid->rank;
	int power = 1;
	for(int i = 0; i < nLevels; i++) power *= 2;
	for(int d = 0; d < nDims; d++)
		if(grid->trueSize[d+1] % power) fatal("All elements in grid:trueSize must be a multiple of 2^mgLevels=%d", power);
	Multigrid *mg = zalloc<Multigrid>(1);
	mg->grids = zalloc<Grid*>(nLevels);
	mg->grids[0] = grid;
	for(int q = 1; q < nLevels; q++){
		int ts[3], gl[6], bnd[6];
		for(int d = 0; d < nDims; d++){
			ts[d] = grid->trueSize[d+1] >> q;
			gl[d] = grid->nGhostLayers[d+1]; gl[d+nDims] = grid->nGhostLayers[d+1+rank];
			bnd[d] = grid->bnd[d+1]; bnd[d+nDims] = grid->bnd[d+1+rank];
		}
		mg->grids[q] = pincGridAlloc(nDims, ts, gl, grid->size[0], bnd);
	}
	mg->nLevels = nLevels; mg->nMGCycles = nCycles;
	mg->nPreSmooth = nPre; mg->nPostSmooth = nPost; mg->nCoarseSolve = nCoarse;
	mg->coarseSolv = mgGS3D; mg->preSmooth = mgGS3D; mg->postSmooth = mgGS3D;
	mg->restrictor = mgHalfRestrict3D; mg->prolongator = mgBilinProl3D;
	return mg;
}
static void mgFreePlain(Multigrid *mg){
	for(int q = 1; q < mg->nLevels; q++) pincGridFree(mg->grids[q]);
	free(mg->grids);
	free(mg);
}

MultigridSolver *pincMgAllocSolver(Grid *rho, Grid *phi, int mgLevels, int mgCycles, int nPreSmooth, int nPostSmooth, int nCoarseSolve){
	MultigridSolver *s = zalloc<MultigridSolver>(1);
	int nDims = rho->rank - 1, rank = rho->rank;
	int ts[3], gl[6], bnd[6];
	for(int d = 0; d < nDims; d++){
		ts[d] = rho->trueSize[d+1];
		gl[d] = rho->nGhostLayers[d+1]; gl[d+nDims] = rho->nGhostLayers[d+1+rank];
		bnd[d] = rho->bnd[d+1]; bnd[d+nDims] = rho->bnd[d+1+rank];
	}
	s->res = pincGridAlloc(nDims, ts, gl, 1, bnd);
	s->mgRho = mgAllocPlain(rho, mgLevels, mgCycles, nPreSmooth, nPostSmooth, nCoarseSolve);
	s->mgRes = mgAllocPlain(s->res, mgLevels, mgCycles, nPreSmooth, nPostSmooth, nCoarseSolve);
	s->mgPhi = mgAllocPlain(phi, mgLevels, mgCycles, nPreSmooth, nPostSmooth, nCoarseSolve);
	s->mgAlgo = (funPtr)mgVRecursive;
	return s;
}
void mgFreeSolver(MultigridSolver *s){
	if(!s) return;
	mgFreePlain(s->mgRho); mgFreePlain(s->mgPhi); mgFreePlain(s->mgRes);
	pincGridFree(s->res);
	free(s);
}

} // extern "C"
