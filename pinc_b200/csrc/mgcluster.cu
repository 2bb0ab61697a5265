// mgcluster.cu — the multigrid tolerance loop (src/multigrid.c:1688-1706) as ONE kernel on ONE thread-block
// cluster, with phi of EVERY level resident in distributed shared memory.
//
// Why: the reference's V-cycle is ~45 dependent half-sweeps per level on grids of at most a few hundred
// thousand nodes, and its convergence factor on 64^3 is ~0.7, i.e. ~55 V-cycles per time step.  The work per
// half-sweep is microscopic; what costs is the dependency between half-sweeps.  A grid-wide barrier through
// L2 costs ~2-3 us; a cluster barrier costs ~0.2 us and neighbour planes are read straight out of the
// neighbouring SM's shared memory (DSMEM, ~0.1 us) instead of going through L2.
//
// Layout: level q has nz_q true z-planes; CTA r of the cluster owns planes [r*ppc_q+1, (r+1)*ppc_q] (ppc_q =
// ceil(nz_q/NC)) and keeps them (true nodes only, no ghosts) at the same shared-memory offset in every CTA, so
// the address of a remote node is map_shared_rank(own address, owner).  64^3 over 16 CTAs = 128 KB of phi per
// CTA, coarser levels add 18 KB.  rho and res stay in global memory (L2-resident; read once per update).
//
// Arithmetic per node is that of multigrid.cu (same expression order).  Periodic wrap replaces ghost reads.
// gBnd's mean subtraction inside mgGS3D: mode `exact` applies it after every half-sweep as the reference does
// (pending-shift formulation, one cluster-wide sum per half-sweep); the default applies it once at the end of
// the smoother call, which is the same function in exact arithmetic because the Gauss-Seidel update commutes
// with adding a constant to phi (and sum(rho)=0 keeps the mean bounded); the two differ by rounding only
// (~1e-16 relative per half-sweep) and are both tested against the oracle.
#include "common.h"
#include <cmath>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace pinc {

#define MC_MAXLEV 10
#define MC_BLOCK 512
struct CLvl { double *phiG, *rho, *res; int nx, ny, nz, s0, s1, ppc, off; };
struct CPlan {
	CLvl L[MC_MAXLEV];
	int nLevels, nPre, nPost, nCoarse, maxCycles, exact, nc;
	double tol, totTrue;
	double *hist;
};

struct CK {
	cg::cluster_group cl;
	int rank, nc;
	double *sm;          // dynamic shared memory (phi slabs of all levels)
	double *red;         // [2][1] cluster-sum slots + [32] block scratch (static shared)
	int flip;
};

__device__ __forceinline__ int upW(int j, int n){ return j == n ? 1 : j+1; }
__device__ __forceinline__ int dnW(int j, int n){ return j == 1 ? n : j-1; }
__device__ __forceinline__ long gix(const CLvl &L, int j, int k, int l){ return j + (long)L.s0*(k + (long)L.s1*l); }

// phi of level L at true node (j,k,l), any owner
__device__ __forceinline__ double rdPhi(const CLvl &L, const CK &K, int j, int k, int l){
	int r = (l-1)/L.ppc;
	int lp = (l-1) - r*L.ppc;
	double *base = K.sm + L.off;
	if(r != K.rank) base = K.cl.map_shared_rank(base, r);
	return base[(lp*L.ny + (k-1))*L.nx + (j-1)];
}

__device__ __forceinline__ double blockSumC(CK &K, double v){
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
	for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if(lane == 0) K.red[4 + w] = v;
	__syncthreads();
	if(w == 0){
		double t = lane < nw ? K.red[4 + lane] : 0.0;
		for(int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
		if(lane == 0) K.red[2] = t;
	}
	__syncthreads();
	return K.red[2];
}
// sum over the whole cluster, identical bits in every thread; acts as a cluster barrier
__device__ __forceinline__ double clusterSum(CK &K, double v){
	double b = blockSumC(K, v);
	if(threadIdx.x == 0) K.red[K.flip] = b;
	K.cl.sync();
	double tot = 0;
	for(int r = 0; r < K.nc; r++){
		const double *p = (r == K.rank) ? K.red : K.cl.map_shared_rank(K.red, r);
		tot += p[K.flip];
	}
	K.flip ^= 1;
	return tot;
}

__device__ __forceinline__ void ownPlanes(const CLvl &L, const CK &K, int &l0, int &nl){
	l0 = K.rank*L.ppc + 1;
	nl = L.nz - K.rank*L.ppc;
	if(nl > L.ppc) nl = L.ppc;
	if(nl < 0) nl = 0;
}

// gNeutralizeGrid on a global array (rho): own planes
__device__ void cNeutralizeG(const CLvl &L, double *v, CK &K){
	int l0, nl; ownPlanes(L, K, l0, nl);
	long n = (long)L.nx*L.ny*nl;
	double acc = 0;
	for(long i = threadIdx.x; i < n; i += blockDim.x){
		int j = (int)(i % L.nx) + 1; long t = i / L.nx; int k = (int)(t % L.ny) + 1; int l = l0 + (int)(t / L.ny);
		acc += __ldcg(v + gix(L,j,k,l));
	}
	double avg = clusterSum(K, acc)/((double)L.nx*L.ny*L.nz);
	for(long i = threadIdx.x; i < n; i += blockDim.x){
		int j = (int)(i % L.nx) + 1; long t = i / L.nx; int k = (int)(t % L.ny) + 1; int l = l0 + (int)(t / L.ny);
		long g = gix(L,j,k,l);
		v[g] = __ldcg(v + g) - avg;
	}
	K.cl.sync();
}
// gNeutralizeGrid on phi (shared memory): own planes
__device__ void cNeutralizeS(const CLvl &L, CK &K, double extra){
	int l0, nl; ownPlanes(L, K, l0, nl);
	long n = (long)L.nx*L.ny*nl;
	double *P = K.sm + L.off;
	double acc = 0;
	for(long i = threadIdx.x; i < n; i += blockDim.x) acc += P[i];
	double avg = clusterSum(K, acc)/((double)L.nx*L.ny*L.nz);
	(void)extra;
	for(long i = threadIdx.x; i < n; i += blockDim.x) P[i] -= avg;
	K.cl.sync();
}

// mgGS3D (src/multigrid.c:683-767).  sIn: mean shift still pending on every value at entry.
__device__ void cGS(const CLvl &L, int nCycles, double sIn, int exact, CK &K){
	int l0, nl; ownPlanes(L, K, l0, nl);
	const int nx = L.nx, ny = L.ny, nz = L.nz;
	double *P = K.sm + L.off;
	const long nOwn = (long)nx*ny*nl;
	const double nTot = (double)nx*ny*nz;
	if(nCycles <= 0 || !exact){
		if(sIn != 0.0){
			for(long i = threadIdx.x; i < nOwn; i += blockDim.x) P[i] -= sIn;
			K.cl.sync();
		}
		if(nCycles <= 0) return;
		sIn = 0.0;
	}
	const int half = nx/2;
	const long items = (long)half*ny*nl;
	double sR = sIn, sPrev = 0;
	for(int h = 0; h < 2*nCycles; h++){
		const int parity = (h & 1) ? 0 : 1;
		double acc = 0;
		for(long i = threadIdx.x; i < items; i += blockDim.x){
			int m = (int)(i % half); long t = i / half; int k = (int)(t % ny) + 1; int lp = (int)(t / ny);
			int l = l0 + lp;
			int ja = 2*m + 1;
			int j = (((ja+k+l)&1) == parity) ? ja : ja+1;
			long row = ((long)lp*ny + (k-1))*nx;
			double a = P[row + upW(j,nx)-1] - sR;
			double b = P[row + dnW(j,nx)-1] - sR;
			double c = P[((long)lp*ny + upW(k,ny)-1)*nx + j-1] - sR;
			double d = P[((long)lp*ny + dnW(k,ny)-1)*nx + j-1] - sR;
			double e = ((lp+1 < nl) ? P[row + (long)ny*nx + j-1] : rdPhi(L, K, j, k, upW(l,nz))) - sR;
			double f = ((lp > 0)    ? P[row - (long)ny*nx + j-1] : rdPhi(L, K, j, k, dnW(l,nz))) - sR;
			const double coeff = 1./6.;
			double vn = coeff*(a + b + c + d + e + f + __ldcg(L.rho + gix(L,j,k,l)));
			if(exact){
				int jo = 2*ja + 1 - j;
				acc += vn; acc += P[row + jo-1] - sR;
			}
			P[row + j-1] = vn;
		}
		if(exact){
			double avg = clusterSum(K, acc)/nTot;
			sPrev = sR; sR = avg;
		} else {
			K.cl.sync();
		}
	}
	if(exact){
		for(long i = threadIdx.x; i < nOwn; i += blockDim.x){
			int j = (int)(i % nx) + 1; long t = i / nx; int k = (int)(t % ny) + 1; int l = l0 + (int)(t / ny);
			double v = P[i];
			if((j+k+l)&1) v -= sPrev;
			v -= sR;
			P[i] = v;
		}
		K.cl.sync();
	} else {
		// the 2*nCycles mean subtractions of gBnd, applied once (see the header comment)
		cNeutralizeS(L, K, 0.0);
	}
}

// residual of level L at true node (j,k,l), any owner: -6 phi; += six neighbours; += rho
__device__ __forceinline__ double cResAt(const CLvl &L, const CK &K, int j, int k, int l){
	double r = -6.*rdPhi(L, K, j, k, l);
	r += rdPhi(L,K,upW(j,L.nx),k,l) + rdPhi(L,K,dnW(j,L.nx),k,l)
	   + rdPhi(L,K,j,upW(k,L.ny),l) + rdPhi(L,K,j,dnW(k,L.ny),l)
	   + rdPhi(L,K,j,k,upW(l,L.nz)) + rdPhi(L,K,j,k,dnW(l,L.nz));
	r += __ldcg(L.rho + gix(L,j,k,l));
	return r;
}

__device__ void cDown(const CPlan &P, int q, CK &K){
	const CLvl &L = P.L[q], &C = P.L[q+1];
	cNeutralizeG(L, L.rho, K);
	cGS(L, P.nPre, 0.0, P.exact, K);
	// mgResidual + mgHalfRestrict3D fused: the owner of a coarse node evaluates the seven fine residuals it needs
	int l0, nl; ownPlanes(C, K, l0, nl);
	long n = (long)C.nx*C.ny*nl;
	for(long i = threadIdx.x; i < n; i += blockDim.x){
		int J = (int)(i % C.nx) + 1; long t = i / C.nx; int Kk = (int)(t % C.ny) + 1; int Lz = l0 + (int)(t / C.ny);
		int j = 2*J-1, k = 2*Kk-1, l = 2*Lz-1;
		const double coeff = 1./12.;
		double v = coeff*(6*cResAt(L,K,j,k,l)
			+ cResAt(L,K,upW(j,L.nx),k,l) + cResAt(L,K,dnW(j,L.nx),k,l)
			+ cResAt(L,K,j,upW(k,L.ny),l) + cResAt(L,K,j,dnW(k,L.ny),l)
			+ cResAt(L,K,j,k,upW(l,L.nz)) + cResAt(L,K,j,k,dnW(l,L.nz)));
		C.rho[gix(C,J,Kk,Lz)] = v;
	}
	K.cl.sync();
}
__device__ void cBottom(const CPlan &P, CK &K){
	const CLvl &L = P.L[P.nLevels-1];
	cNeutralizeG(L, L.rho, K);
	cGS(L, P.nCoarse, 0.0, P.exact, K);
	cNeutralizeS(L, K, 0.0);
}
__device__ __forceinline__ double cProlZ(const CLvl &C, const CK &K, int J, int Kk, int l){
	if(l & 1) return rdPhi(C, K, J, Kk, (l+1)/2);
	return 0.5*(rdPhi(C, K, J, Kk, l/2) + rdPhi(C, K, J, Kk, upW(l/2, C.nz)));
}
__device__ __forceinline__ double cProlY(const CLvl &C, const CK &K, int J, int k, int l){
	if(k & 1) return cProlZ(C, K, J, (k+1)/2, l);
	return 0.5*(cProlZ(C, K, J, k/2, l) + cProlZ(C, K, J, upW(k/2, C.ny), l));
}
__device__ __forceinline__ double cProl(const CLvl &C, const CK &K, int j, int k, int l){
	if(j & 1) return cProlY(C, K, (j+1)/2, k, l);
	return 0.5*(cProlY(C, K, j/2, k, l) + cProlY(C, K, upW(j/2, C.nx), k, l));
}
// res(q) := P(phi(q+1)); phi(q) += res(q); gBnd; post-smooth; gBnd
__device__ void cUp(const CPlan &P, int q, CK &K){
	const CLvl &L = P.L[q], &C = P.L[q+1];
	int l0, nl; ownPlanes(L, K, l0, nl);
	long n = (long)L.nx*L.ny*nl;
	double *S = K.sm + L.off;
	double acc = 0;
	for(long i = threadIdx.x; i < n; i += blockDim.x){
		int j = (int)(i % L.nx) + 1; long t = i / L.nx; int k = (int)(t % L.ny) + 1; int l = l0 + (int)(t / L.ny);
		double p = cProl(C, K, j, k, l);
		L.res[gix(L,j,k,l)] = p;
		double v = S[i]; v += p;
		S[i] = v;
		acc += v;
	}
	double avg = clusterSum(K, acc)/((double)L.nx*L.ny*L.nz);
	cGS(L, P.nPost, avg, P.exact, K);
	cNeutralizeS(L, K, 0.0);
}
__device__ void cGhosts(double *v, const CLvl &L, const CK &K){
	int s0 = L.s0, s1 = L.s1, s2 = L.nz + 2;
	long n = (long)s0*s1*s2;
	for(long i = K.rank*(long)blockDim.x + threadIdx.x; i < n; i += (long)K.nc*blockDim.x){
		int j = (int)(i % s0); long r = i / s0; int k = (int)(r % s1); int l = (int)(r / s1);
		int jw = j == 0 ? s0-2 : (j == s0-1 ? 1 : j);
		int kw = k == 0 ? s1-2 : (k == s1-1 ? 1 : k);
		int lw = l == 0 ? s2-2 : (l == s2-1 ? 1 : l);
		if(jw != j || kw != k || lw != l) v[i] = __ldcg(v + (jw + (long)s0*(kw + (long)s1*lw)));
	}
}

__global__ void __launch_bounds__(MC_BLOCK, 1) k_mg_cluster(CPlan P){
	extern __shared__ double dyn[];
	__shared__ double red[40];
	CK K{ cg::this_cluster(), 0, P.nc, dyn, red, 0 };
	K.rank = (int)K.cl.block_rank();
	const int b = P.nLevels - 1;
	// phi of every level: global -> shared (own planes)
	for(int q = 0; q <= b; q++){
		const CLvl &L = P.L[q];
		int l0, nl; ownPlanes(L, K, l0, nl);
		long n = (long)L.nx*L.ny*nl;
		double *S = K.sm + L.off;
		for(long i = threadIdx.x; i < n; i += blockDim.x){
			int j = (int)(i % L.nx) + 1; long t = i / L.nx; int k = (int)(t % L.ny) + 1; int l = l0 + (int)(t / L.ny);
			S[i] = __ldcg(L.phiG + gix(L,j,k,l));
		}
	}
	K.cl.sync();
	double barRes = 2.;
	int cycles = 0;
	while(barRes > P.tol && cycles < P.maxCycles){
		for(int q = 0; q < b; q++) cDown(P, q, K);
		cBottom(P, K);
		for(int q = b-1; q >= 0; q--) cUp(P, q, K);
		// mgSolveRaw :1700-1704
		const CLvl &L = P.L[0];
		int l0, nl; ownPlanes(L, K, l0, nl);
		long n = (long)L.nx*L.ny*nl;
		double acc = 0;
		for(long i = threadIdx.x; i < n; i += blockDim.x){
			int j = (int)(i % L.nx) + 1; long t = i / L.nx; int k = (int)(t % L.ny) + 1; int l = l0 + (int)(t / L.ny);
			double r = cResAt(L, K, j, k, l);
			r = r*r;
			L.res[gix(L,j,k,l)] = r;
			acc += r;
		}
		barRes = clusterSum(K, acc);
		barRes /= P.totTrue;
		barRes = sqrt(barRes);
		if(K.rank == 0 && threadIdx.x == 0 && cycles < 250) P.hist[1+cycles] = barRes;
		cycles++;
	}
	if(K.rank == 0 && threadIdx.x == 0) P.hist[0] = (double)cycles;
	// phi back to global, then the ghost layers of every array (the state the reference leaves behind)
	for(int q = 0; q <= b; q++){
		const CLvl &L = P.L[q];
		int l0, nl; ownPlanes(L, K, l0, nl);
		long n = (long)L.nx*L.ny*nl;
		double *S = K.sm + L.off;
		for(long i = threadIdx.x; i < n; i += blockDim.x){
			int j = (int)(i % L.nx) + 1; long t = i / L.nx; int k = (int)(t % L.ny) + 1; int l = l0 + (int)(t / L.ny);
			L.phiG[gix(L,j,k,l)] = S[i];
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

// host side: returns false if this solve does not fit the cluster kernel (caller falls back)
bool clusterSolve(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, double tol, int maxCycles, int exact){
	static int maxNc = -1;
	static size_t maxSmem = 0;
	if(maxNc < 0){
		int dev = c->device, v = 0;
		cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
		maxSmem = (size_t)v;
		maxNc = 0;
		if(cudaFuncSetAttribute((const void*)k_mg_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxSmem) == cudaSuccess){
			cudaFuncSetAttribute((const void*)k_mg_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
			for(int nc : {16, 8}){
				cudaLaunchConfig_t cfg = {};
				cfg.gridDim = dim3(nc); cfg.blockDim = dim3(MC_BLOCK); cfg.dynamicSmemBytes = maxSmem - 1024;
				cudaLaunchAttribute at[1];
				at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
				cfg.attrs = at; cfg.numAttrs = 1;
				int nClusters = 0;
				if(cudaOccupancyMaxActiveClusters(&nClusters, (const void*)k_mg_cluster, &cfg) == cudaSuccess && nClusters >= 1){ maxNc = nc; break; }
			}
		}
		cudaGetLastError();
	}
	if(maxNc < 8) return false;
	int nL = mgRho->nLevels;
	if(nL > MC_MAXLEV) return false;
	CPlan P{};
	int nc = maxNc;
	for(int attempt = 0; attempt < 2; attempt++){
		long off = 0;
		for(int q = 0; q < nL; q++){
			DevGrid *r = devGrid(c, mgRho->grids[q]), *p = devGrid(c, mgPhi->grids[q]), *e = devGrid(c, mgRes->grids[q]);
			if(r->n != p->n || r->n != e->n || r->nv != 1) fatal("multigrid level %d: rho/phi/res differ in shape", q);
			CLvl &L = P.L[q];
			L.phiG = p->d; L.rho = r->d; L.res = e->d;
			L.nx = r->tsize[0]; L.ny = r->tsize[1]; L.nz = r->tsize[2];
			L.s0 = r->size[0]; L.s1 = r->size[1];
			L.ppc = (L.nz + nc - 1)/nc;
			L.off = (int)off;
			off += (long)L.nx*L.ny*L.ppc;
		}
		size_t need = (size_t)off*sizeof(double);
		if(need <= maxSmem - 2048){
			P.nLevels = nL; P.nPre = mgRho->nPreSmooth; P.nPost = mgRho->nPostSmooth; P.nCoarse = mgRho->nCoarseSolve;
			P.maxCycles = maxCycles; P.exact = exact; P.nc = nc; P.tol = tol;
			DevGrid *r0 = devGrid(c, mgRho->grids[0]);
			P.totTrue = (double)((long)r0->tsize[0]*r0->tsize[1]*r0->tsize[2]);
			if(!c->d_mgHist){
				PINC_CUDA(cudaMalloc(&c->d_mgHist, 256*sizeof(double)));
				PINC_CUDA(cudaMallocHost(&c->h_mgHist, 256*sizeof(double)));
			}
			P.hist = c->d_mgHist;
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
			return true;
		}
		if(nc == 8) return false;
		return false;          // does not fit even over the largest cluster
	}
	return false;
}

} // namespace pinc
