// mgcluster.cu — the multigrid tolerance loop (src/multigrid.c:1688-1706) as ONE kernel on ONE thread-block
// cluster, with phi of EVERY level resident in (distributed) shared memory.
//
// Why: the reference's V-cycle is ~45 dependent half-sweeps per level on grids of at most a few hundred
// thousand nodes, and its convergence factor on 64^3 is ~0.7, i.e. ~55 V-cycles per time step.  The work per
// half-sweep is microscopic; what costs is the dependency between half-sweeps.  Measured on B200
// (tools/ubench_sync.cu): a grid-wide barrier through L2 ~3700 cycles, a dependent L2 load ~300 cycles, a
// cluster barrier ~470 cycles (512 threads/CTA, any cluster size), __syncthreads ~45 cycles.  Hence:
//
//   big levels   (> 4096 true nodes) are split by z-planes over the 16 CTAs of the cluster: CTA r owns planes
//                [r*ppc+1, (r+1)*ppc] (true nodes only, no ghosts) at the same shared-memory offset in every
//                CTA, so a neighbour plane is read straight out of the neighbouring SM's shared memory
//                (map_shared_rank); one cluster barrier per half-sweep.  64^3 = 128 KB of phi per CTA.
//   small levels (<= 4096 true nodes) live entirely (phi and rho) in CTA 0, which runs their part of the
//                V-cycle alone on __syncthreads while the other CTAs wait at one cluster barrier.
//   rho          of big levels >= 1 is kept in shared memory when it fits, else read from L2 one batch of rows
//                ahead of use; res is only ever written (global memory).
//
// Arithmetic per node is that of multigrid.cu (same expression order).  Periodic wrap replaces ghost reads.
// gBnd's mean subtraction inside mgGS3D: EXACT applies it after every half-sweep as the reference does
// (pending-shift formulation, one block/cluster-wide sum per half-sweep); the default applies it once at the end
// of the smoother call, which is the same function in exact arithmetic because the Gauss-Seidel update commutes
// with adding a constant to phi (and sum(rho)=0 keeps the mean bounded); the two differ by rounding only
// (~1e-16 relative per half-sweep) and both are tested against the oracle (V-cycle counts, residual history, phi).
#include "common.h"
#include <cmath>
#include "mgsmem.cuh"

namespace pinc {

template<bool EXACT> __device__ void vcycle(const CPlan &P, CK &K){
	const int b = P.nLevels - 1, nb = P.nBig;           // levels [0,nb) are big, [nb,b] small
	for(int q = 0; q < nb && q < b; q++) cDown<false,EXACT>(P, q, K);
	if(nb > b){
		cBottom<false,EXACT>(P, K);
	} else {
		if(K.rank == 0){
			for(int q = nb; q < b; q++) cDown<true,EXACT>(P, q, K);
			cBottom<true,EXACT>(P, K);
			for(int q = b-1; q >= nb; q--) cUp<true,EXACT>(P, q, K);
		}
		K.cl.sync();
	}
	for(int q = (nb > b ? b : nb) - 1; q >= 0; q--) cUp<false,EXACT>(P, q, K);
}

__global__ void __launch_bounds__(MC_BLOCK, 1) k_mg_cluster(CPlan P){
	__shared__ double red[40];
	CK K{ cg::this_cluster(), 0, P.nc, mgS, red, 0, P.prof, -1 };
	K.rank = (int)K.cl.block_rank();
	const int b = P.nLevels - 1;
	// phi of every level, and rho of the levels that keep it in shared memory: global -> shared (own planes)
	for(int q = 0; q <= b; q++){
		const CLvl &L = P.L[q];
		int l0, nl; ownPlanes(L, K, l0, nl);
		int n = L.nx*L.ny*nl;
		for(int i = threadIdx.x; i < n; i += blockDim.x){
			int j,k,l; ownNode(L,l0,i,j,k,l);
			mgS[L.offPhi + i] = __ldcg(L.phiG + gix(L,j,k,l));
			if(L.offRho >= 0) mgS[L.offRho + i] = __ldcg(L.rho + gix(L,j,k,l));
		}
	}
	K.cl.sync();
	double barRes = 2.;
	int cycles = 0;
	while(barRes > P.tol && cycles < P.maxCycles){
		if(P.exact) vcycle<true>(P, K); else vcycle<false>(P, K);
		// mgSolveRaw :1700-1704: residual of level 0, squared in place, true-grid sum, RMS
		ProfScope psn(K, PS_NORM);
		const CLvl &L = P.L[0];
		int l0, nl; ownPlanes(L, K, l0, nl);
		int n = L.nx*L.ny*nl;
		double acc = 0;
		for(int i = threadIdx.x; i < n; i += blockDim.x){
			int j,k,l; ownNode(L,l0,i,j,k,l);
			double r = cResAt(L, K, j, k, l);
			r = r*r;
			L.res[gix(L,j,k,l)] = r;
			acc += r;
		}
		barRes = sumAll<false>(K, acc);
		barRes /= P.totTrue;
		barRes = sqrt(barRes);
		if(K.rank == 0 && threadIdx.x == 0 && cycles < 250) P.hist[1+cycles] = barRes;
		cycles++;
	}
	if(K.rank == 0 && threadIdx.x == 0){ P.hist[0] = (double)cycles; P.hist[251] = barRes; }
	// back to global (phi of every level, rho where it lived in shared memory), then the ghost layers of every
	// array: the state the reference leaves behind
	for(int q = 0; q <= b; q++){
		const CLvl &L = P.L[q];
		int l0, nl; ownPlanes(L, K, l0, nl);
		int n = L.nx*L.ny*nl;
		for(int i = threadIdx.x; i < n; i += blockDim.x){
			int j,k,l; ownNode(L,l0,i,j,k,l);
			L.phiG[gix(L,j,k,l)] = mgS[L.offPhi + i];
			if(L.offRho >= 0) L.rho[gix(L,j,k,l)] = mgS[L.offRho + i];
		}
	}
	__threadfence();
	K.cl.sync();
	for(int q = 0; q <= b; q++){
		cGhosts(P.L[q].phiG, P.L[q], K);
		cGhosts(P.L[q].rho, P.L[q], K);
		cGhosts(P.L[q].res, P.L[q], K);
	}
}

// cycle accounting of the multigrid kernels ($PINC_B200_MGPROF=1): per-context device buffer of 64 long longs, or null
void *mgProfBuffer(Ctx *c){
	static const bool on = getenv("PINC_B200_MGPROF") != nullptr;
	if(!on) return nullptr;
	if(!c->d_mgProf){ PINC_CUDA(cudaMalloc(&c->d_mgProf, 64*sizeof(long long))); PINC_CUDA(cudaMemset(c->d_mgProf, 0, 64*sizeof(long long))); }
	return c->d_mgProf;
}
// host side: returns false if this solve does not fit the cluster kernel (the caller falls back)
bool clusterSolve(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, double tol, int maxCycles, int exact){
	int &maxNc = c->clNc;                        // per context: attributes and schedulability belong to the device
	size_t &maxSmem = c->clSmem;
	if(maxNc < 0){
		int v = 0;
		cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
		maxSmem = (size_t)v - 2048;              // room for the kernel's static shared memory
		maxNc = 0;
		cudaError_t ea = cudaFuncSetAttribute((const void*)k_mg_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxSmem);
		if(ea != cudaSuccess) fprintf(stderr, "PINC-B200 WARNING: cluster multigrid unavailable (%s)\n", cudaGetErrorString(ea));
		if(ea == cudaSuccess){
			cudaFuncSetAttribute((const void*)k_mg_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
			for(int nc : {16, 8}){
				cudaLaunchConfig_t cfg = {};
				cfg.gridDim = dim3(nc); cfg.blockDim = dim3(MC_BLOCK); cfg.dynamicSmemBytes = maxSmem;
				cudaLaunchAttribute at[1];
				at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
				cfg.attrs = at; cfg.numAttrs = 1;
				int nClusters = 0;
				cudaError_t eo = cudaOccupancyMaxActiveClusters(&nClusters, (const void*)k_mg_cluster, &cfg);
				if(eo == cudaSuccess && nClusters >= 1){ maxNc = nc; break; }
				if(getenv("PINC_B200_VERBOSE")) fprintf(stderr, "pinc-b200: cluster size %d not schedulable (%s, %d clusters)\n", nc, cudaGetErrorString(eo), nClusters);
			}
		}
		cudaGetLastError();
	}
	if(maxNc < 8) return false;
	const int nL = mgRho->nLevels, nc = maxNc;
	if(nL > MC_MAXLEV) return false;
	CPlan P{};
	long off = 0;
	int nBig = 0;
	for(int q = 0; q < nL; q++){
		DevGrid *r = devGrid(c, mgRho->grids[q]), *p = devGrid(c, mgPhi->grids[q]), *e = devGrid(c, mgRes->grids[q]);
		if(r->n != p->n || r->n != e->n || r->nv != 1) fatal("multigrid level %d: rho/phi/res differ in shape", q);
		CLvl &L = P.L[q];
		L.phiG = p->d; L.rho = r->d; L.res = e->d;
		L.nx = r->tsize[0]; L.ny = r->tsize[1]; L.nz = r->tsize[2];
		L.s0 = r->size[0]; L.s1 = r->size[1];
		long nt = (long)L.nx*L.ny*L.nz;
		L.small = nt <= MC_SMALL;
		if(!L.small){ if(nBig != q) return false; nBig = q+1; }      // levels shrink monotonically
		if(L.small && (L.nx/2)*L.ny*L.nz > MC_U*MC_BLOCK) return false;
		L.ppc = L.small ? L.nz : (L.nz + nc - 1)/nc;
		L.offPhi = (int)off; off += (long)L.nx*L.ny*L.ppc;
		L.offRho = -1;
		if(L.small){ L.offRho = (int)off; off += nt; }
	}
	// rho of the big levels >= 1 joins phi in shared memory while it fits (coarsest first: most latency-bound)
	for(int q = nBig-1; q >= 1; q--){
		CLvl &L = P.L[q];
		long slab = (long)L.nx*L.ny*L.ppc;
		if((size_t)(off + slab)*sizeof(double) <= maxSmem){ L.offRho = (int)off; off += slab; }
	}
	size_t need = (size_t)off*sizeof(double);
	if(need > maxSmem) return false;
	P.nLevels = nL; P.nBig = nBig; P.nPre = mgRho->nPreSmooth; P.nPost = mgRho->nPostSmooth; P.nCoarse = mgRho->nCoarseSolve;
	P.maxCycles = maxCycles; P.exact = exact; P.nc = nc; P.tol = tol;
	P.totTrue = (double)((long)P.L[0].nx*P.L[0].ny*P.L[0].nz);
	if(!c->d_mgHist){
		PINC_CUDA(cudaMalloc(&c->d_mgHist, 256*sizeof(double)));
		PINC_CUDA(cudaMallocHost(&c->h_mgHist, 256*sizeof(double)));
	}
	P.hist = c->d_mgHist;
	P.prof = (long long*)mgProfBuffer(c);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(nc); cfg.blockDim = dim3(MC_BLOCK); cfg.dynamicSmemBytes = need; cfg.stream = c->stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
	cfg.attrs = at; cfg.numAttrs = 1;
	double work = 0;
	for(int q = 0; q < nL; q++) work += 24.0*P.L[q].nx*P.L[q].ny*P.L[q].nz*(q == nL-1 ? P.nCoarse : P.nPre + P.nPost);
	{
		LaunchScope ls(c, K_MGFUSED, work);
		PINC_CUDA(cudaLaunchKernelEx(&cfg, k_mg_cluster, P));
	}
	PINC_CUDA(cudaMemcpyAsync(c->h_mgHist, c->d_mgHist, 256*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	c->mgHistPending = true;
	c->mgCheckPending = true; c->mgTol = tol; c->mgMaxCycles = maxCycles;
	return true;
}

} // namespace pinc
