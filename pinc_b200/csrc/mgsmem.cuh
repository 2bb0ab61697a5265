// mgsmem.cuh — device code of the shared-memory multigrid: level descriptors, distributed-shared-memory accessors,
// the smoother / transfer / neutralise routines for "big" levels (z-planes split over the CTAs of a cluster) and
// "small" levels (whole level in one CTA).  Included by mgcluster.cu (cluster kernel) and multigrid.cu (the
// all-SM persistent kernel runs its small levels through the same code in CTA 0).  See mgcluster.cu for the design.
#pragma once
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace pinc {

// the dynamic shared memory of the kernels that include this header (every extern __shared__ array of a kernel
// starts at the same address); indexing it directly keeps the accesses in the shared state space (LDS/STS)
extern __shared__ double mgS[];

#define MC_MAXLEV 10
#define MC_BLOCK 512
#define MC_SMALL 4096        // a level with at most this many true nodes runs inside CTA 0
#define MC_U 4               // nodes per thread and colour on a small level (4096/2/512)
#define MC_UB 2              // rows in flight per thread on a big level
struct CLvl {
	double *phiG, *rho, *res;     // global arrays, ghost-inclusive layout of the reference
	int nx, ny, nz, s0, s1;
	int ppc;                      // z-planes per CTA; = nz for a small level (everything in CTA 0)
	int offPhi, offRho;           // shared-memory offsets in doubles; offRho < 0: rho is read from global
	int small;
};
struct CPlan {
	CLvl L[MC_MAXLEV];
	int nLevels, nBig, nPre, nPost, nCoarse, maxCycles, exact, nc;
	double tol, totTrue;
	double *hist;
	long long *prof;
};

struct CK {
	cg::cluster_group cl;
	int rank, nc;
	double *sm;          // dynamic shared memory
	double *red;         // static shared: [0..1] cluster-sum slots, [2] block total, [4..35] per-warp scratch
	int flip;
	long long *prof;     // optional cycle accounting (thread 0 of CTA 0): [2*slot] cycles, [2*slot+1] calls
	int offZ;            // exchange buffer of sGS16 (doubles into mgS), or < 0
};

struct ProfScope {
	long long *p; int slot; long long t0;
	__device__ __forceinline__ ProfScope(const CK &K, int s) : p((K.prof && K.rank == 0 && threadIdx.x == 0) ? K.prof : nullptr), slot(s), t0(0) { if(p) t0 = clock64(); }
	__device__ __forceinline__ ~ProfScope(){ if(p){ p[2*slot] += clock64() - t0; p[2*slot+1] += 1; } }
};
static __device__ __forceinline__ int psLvl(const CLvl &L, int kind){ int t = L.nx >= 16 ? 0 : (L.nx >= 8 ? 1 : (L.nx >= 4 ? 2 : 3)); return 16 + 4*t + kind; }
enum { PS_NEUT_RHO = 0, PS_GS_BIG, PS_GS_SMALL, PS_RESTRICT, PS_PROLONG, PS_NEUT_PHI, PS_NORM, PS_GS_BIG_SYNC, PS_LEVEL0 = 8 };

static __device__ __forceinline__ int upW(int j, int n){ return j == n ? 1 : j+1; }
static __device__ __forceinline__ int dnW(int j, int n){ return j == 1 ? n : j-1; }
static __device__ __forceinline__ long gix(const CLvl &L, int j, int k, int l){ return j + (long)L.s0*(k + (long)L.s1*l); }

// pointer to node (1,1,l) of plane l of an array that is distributed like phi (offset `off`), any owner
static __device__ __forceinline__ double *planePtr(const CLvl &L, const CK &K, int off, int l){
	if(L.small && K.rank == 0) return mgS + off + (l-1)*L.ny*L.nx;        // whole level lives in this CTA: no owner arithmetic
	int r = (l-1)/L.ppc;
	int lp = (l-1) - r*L.ppc;
	double *base = mgS + off;
	if(r != K.rank) base = K.cl.map_shared_rank(base, r);
	return base + lp*L.ny*L.nx;
}
static __device__ __forceinline__ double rdPhi(const CLvl &L, const CK &K, int j, int k, int l){
	return planePtr(L, K, L.offPhi, l)[(k-1)*L.nx + (j-1)];
}
static __device__ __forceinline__ double rdRho(const CLvl &L, const CK &K, int j, int k, int l){
	if(L.offRho < 0) return __ldcg(L.rho + gix(L,j,k,l));
	return planePtr(L, K, L.offRho, l)[(k-1)*L.nx + (j-1)];
}
static __device__ __forceinline__ void wrRho(const CLvl &L, const CK &K, int j, int k, int l, double v){
	if(L.offRho < 0) L.rho[gix(L,j,k,l)] = v;
	else planePtr(L, K, L.offRho, l)[(k-1)*L.nx + (j-1)] = v;
}

static __device__ __noinline__ double blockSumC(CK &K, double v){
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
	double tot = K.red[2];
	__syncthreads();
	return tot;
}
// sum over all participating threads (CTA 0 for a small level, the cluster otherwise); identical bits in every
// thread; acts as the scope's barrier
template<bool SMALL> static __device__ __forceinline__ double sumAll(CK &K, double v){
	double b = blockSumC(K, v);
	if(SMALL) return b;
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
template<bool SMALL> static __device__ __forceinline__ void syncAll(CK &K){
	if(SMALL) __syncthreads(); else K.cl.sync();
}

static __device__ __forceinline__ void ownPlanes(const CLvl &L, const CK &K, int &l0, int &nl){
	l0 = K.rank*L.ppc + 1;
	nl = L.nz - K.rank*L.ppc;
	if(nl > L.ppc) nl = L.ppc;
	if(nl < 0) nl = 0;
}
// flat index over the own true nodes -> (j,k,l) and the slab offset
static __device__ __forceinline__ void ownNode(const CLvl &L, int l0, int i, int &j, int &k, int &l){
	unsigned u = (unsigned)i, nx = (unsigned)L.nx, ny = (unsigned)L.ny;
	unsigned t = u / nx; j = (int)(u - t*nx) + 1;
	unsigned lp = t / ny; k = (int)(t - lp*ny) + 1; l = l0 + (int)lp;
}

// gNeutralizeGrid of rho (src/grid.c:730-779) on the own planes
template<bool SMALL> static __device__ __noinline__ void cNeutralizeRho(const CLvl &L, CK &K){
	ProfScope ps(K, PS_NEUT_RHO);
	int l0, nl; ownPlanes(L, K, l0, nl);
	int n = L.nx*L.ny*nl;
	double avg;
	if(L.offRho >= 0){
		double *R = mgS + L.offRho;
		double acc = 0;
		for(int i = threadIdx.x; i < n; i += blockDim.x) acc += R[i];
		avg = sumAll<SMALL>(K, acc)/((double)L.nx*L.ny*L.nz);
		for(int i = threadIdx.x; i < n; i += blockDim.x) R[i] -= avg;
	} else {
		double acc = 0;
		for(int i = threadIdx.x; i < n; i += blockDim.x){ int j,k,l; ownNode(L,l0,i,j,k,l); acc += __ldcg(L.rho + gix(L,j,k,l)); }
		avg = sumAll<SMALL>(K, acc)/((double)L.nx*L.ny*L.nz);
		for(int i = threadIdx.x; i < n; i += blockDim.x){ int j,k,l; ownNode(L,l0,i,j,k,l); long g = gix(L,j,k,l); L.rho[g] = __ldcg(L.rho + g) - avg; }
	}
	syncAll<SMALL>(K);
}
// gNeutralizeGrid of phi on the own planes
template<bool SMALL> static __device__ __noinline__ void cNeutralizePhi(const CLvl &L, CK &K){
	ProfScope ps(K, PS_NEUT_PHI);
	int l0, nl; ownPlanes(L, K, l0, nl);
	int n = L.nx*L.ny*nl;
	double *P = mgS + L.offPhi;
	double acc = 0;
	for(int i = threadIdx.x; i < n; i += blockDim.x) acc += P[i];
	double avg = sumAll<SMALL>(K, acc)/((double)L.nx*L.ny*L.nz);
	for(int i = threadIdx.x; i < n; i += blockDim.x) P[i] -= avg;
	syncAll<SMALL>(K);
}

} // namespace pinc
#include "mgsmall.cuh"
namespace pinc {

// one Gauss-Seidel node: 1/6 * (x+ + x- + y+ + y- + z+ + z- + rho), summed left to right (src/multigrid.c:711-714)
template<bool EXACT> static __device__ __forceinline__ double gsVal(double a, double b, double c, double d, double e, double f, double rho, double sR){
	if(EXACT){ a -= sR; b -= sR; c -= sR; d -= sR; e -= sR; f -= sR; }
	const double coeff = 1./6.;
	return coeff*(a + b + c + d + e + f + rho);
}

// mgGS3D (src/multigrid.c:683-767) on a big level.  sIn: mean shift still pending on every value at entry.
template<bool EXACT> static __device__ __noinline__ void cGSBig(const CLvl &Lref, int nCycles, double sIn, CK &K){
	const CLvl L = Lref;
	ProfScope ps(K, PS_GS_BIG);
	int l0, nl; ownPlanes(L, K, l0, nl);
	const int nx = L.nx, ny = L.ny, nz = L.nz, pl = nx*ny;
	double *P = mgS + L.offPhi;
	const int nOwn = pl*nl;
	const double nTot = (double)nx*ny*nz;
	if(nCycles <= 0 || !EXACT){
		if(sIn != 0.0){
			for(int i = threadIdx.x; i < nOwn; i += blockDim.x) P[i] -= sIn;
			K.cl.sync();
		}
		if(nCycles <= 0) return;
		sIn = 0.0;
	}
	const int half = nx/2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
	const bool rhoS = L.offRho >= 0;
	double sR = sIn, sPrev = 0;
	for(int h = 0; h < 2*nCycles; h++){
		const int parity = (h & 1) ? 0 : 1;
		double acc = 0;
		for(int lp = 0; lp < nl; lp++){
			const int l = l0 + lp;
			double *Pl = P + lp*pl;
			const double *Pzu = (lp+1 < nl) ? Pl + pl : planePtr(L, K, L.offPhi, upW(l,nz));
			const double *Pzd = (lp > 0)    ? Pl - pl : planePtr(L, K, L.offPhi, dnW(l,nz));
			// rho row base, indexed by the 0-based x index
			const double *Rl = rhoS ? (mgS + L.offRho + lp*pl) : (L.rho + gix(L,1,0,l));
			const int rStride = rhoS ? nx : L.s0;
			const int rFirst = rhoS ? 0 : 1;              // global rows start at k=0 (ghost), shared ones at k=1
			for(int k0 = warp + 1; k0 <= ny; k0 += MC_UB*nWarps){
				for(int m = lane; m < half; m += 32){
					double a[MC_UB], b[MC_UB], c[MC_UB], d[MC_UB], e[MC_UB], f[MC_UB], rh[MC_UB], oth[MC_UB];
					int xo[MC_UB];
					#pragma unroll
					for(int u = 0; u < MC_UB; u++){
						int k = k0 + u*nWarps;
						if(k > ny) continue;
						int j = ((((1+k+l)&1) == parity) ? 1 : 2) + 2*m;          // 1-based own-colour node
						int row = (k-1)*nx;
						xo[u] = row + j-1;
						a[u] = Pl[row + upW(j,nx)-1];
						b[u] = Pl[row + dnW(j,nx)-1];
						c[u] = Pl[(upW(k,ny)-1)*nx + j-1];
						d[u] = Pl[(dnW(k,ny)-1)*nx + j-1];
						e[u] = Pzu[row + j-1];
						f[u] = Pzd[row + j-1];
						rh[u] = rhoS ? Rl[row + j-1] : __ldcg(Rl + (long)(k-1+rFirst)*rStride + j-1);
						if(EXACT) oth[u] = Pl[row + ((j-1)^1)];                  // the other node of the pair (2m, 2m+1)
					}
					#pragma unroll
					for(int u = 0; u < MC_UB; u++){
						int k = k0 + u*nWarps;
						if(k > ny) continue;
						double vn = gsVal<EXACT>(a[u], b[u], c[u], d[u], e[u], f[u], rh[u], sR);
						if(EXACT){ acc += vn; acc += oth[u] - sR; }
						Pl[xo[u]] = vn;
					}
				}
			}
		}
		if(EXACT){
			double avg = sumAll<false>(K, acc)/nTot;
			sPrev = sR; sR = avg;
		} else {
			ProfScope pss(K, PS_GS_BIG_SYNC);
			K.cl.sync();
		}
	}
	if(EXACT){
		for(int i = threadIdx.x; i < nOwn; i += blockDim.x){
			int j,k,l; ownNode(L,l0,i,j,k,l);
			double v = P[i];
			if((j+k+l)&1) v -= sPrev;
			v -= sR;
			P[i] = v;
		}
		K.cl.sync();
	} else {
		cNeutralizePhi<false>(L, K);     // the 2*nCycles mean subtractions of gBnd, applied once (header comment)
	}
}

// mgGS3D on a small level: everything in CTA 0's shared memory, at most MC_U nodes per thread and colour,
// node coordinates decoded once per call
template<bool EXACT> static __device__ __noinline__ void cGSSmall(const CLvl &Lref, int nCycles, double sIn, CK &K){
	const CLvl L = Lref;
	ProfScope ps(K, PS_GS_SMALL);
	ProfScope psl(K, psLvl(L, 0));
	if(!EXACT && nCycles > 0 && L.nx == L.ny && L.ny == L.nz && blockDim.x == 512 && K.rank == 0){
		// the benchmark pyramid's levels have routines of their own (mgsmall.cuh)
		if(L.nx == 16 && K.offZ >= 0){ sGS16(mgS + L.offPhi, mgS + L.offRho, mgS + K.offZ, nCycles, sIn, K); return; }
		if(L.nx == 8){ sGSOne<8>(mgS + L.offPhi, mgS + L.offRho, nCycles, sIn, K); return; }
		if(L.nx == 4){ sGSOne<4>(mgS + L.offPhi, mgS + L.offRho, nCycles, sIn, K); return; }
	}
	const int nx = L.nx, ny = L.ny, nz = L.nz, pl = nx*ny;
	double *P = mgS + L.offPhi;
	const double *R = mgS + L.offRho;
	const int nOwn = pl*nz;
	const double nTot = (double)nOwn;
	if(nCycles <= 0 || !EXACT){
		if(sIn != 0.0){
			for(int i = threadIdx.x; i < nOwn; i += blockDim.x) P[i] -= sIn;
			__syncthreads();
		}
		if(nCycles <= 0) return;
		sIn = 0.0;
	}
	const int half = nx/2, items = half*ny*nz;
	int rowO[MC_U], yu[MC_U], yd[MC_U], zu[MC_U], zd[MC_U], m2[MC_U], par[MC_U];
	#pragma unroll
	for(int u = 0; u < MC_U; u++){
		int i = threadIdx.x + u*blockDim.x;
		rowO[u] = -1;
		if(i < items){
			unsigned t = (unsigned)i/(unsigned)half; int m = i - (int)t*half;
			unsigned lp = t/(unsigned)ny; int k = (int)(t - lp*ny) + 1; int l = (int)lp + 1;
			rowO[u] = ((l-1)*ny + (k-1))*nx;
			yu[u] = ((l-1)*ny + (upW(k,ny)-1))*nx; yd[u] = ((l-1)*ny + (dnW(k,ny)-1))*nx;
			zu[u] = ((upW(l,nz)-1)*ny + (k-1))*nx; zd[u] = ((dnW(l,nz)-1)*ny + (k-1))*nx;
			m2[u] = 2*m; par[u] = (1+k+l)&1;
		}
	}
	double sR = sIn, sPrev = 0;
	for(int h = 0; h < 2*nCycles; h++){
		const int parity = (h & 1) ? 0 : 1;
		double acc = 0;
		// two nodes at a time: all loads of the pair, then the arithmetic, then the stores (own-colour stores never
		// alias other-colour loads of the same half-sweep, but the compiler cannot know that)
		#pragma unroll
		for(int u0 = 0; u0 < MC_U; u0 += 2){
			double vn[2], oth[2];
			int xo[2];
			#pragma unroll
			for(int w = 0; w < 2; w++){
				const int u = u0 + w;
				if(rowO[u] < 0) continue;
				int j = ((par[u] == parity) ? 1 : 2) + m2[u];
				int jm = j-1;
				xo[w] = rowO[u] + jm;
				double a = P[rowO[u] + upW(j,nx)-1], b = P[rowO[u] + dnW(j,nx)-1];
				double c = P[yu[u] + jm], d = P[yd[u] + jm], e = P[zu[u] + jm], f = P[zd[u] + jm];
				vn[w] = gsVal<EXACT>(a, b, c, d, e, f, R[xo[w]], sR);
				if(EXACT) oth[w] = P[rowO[u] + (jm^1)];
			}
			#pragma unroll
			for(int w = 0; w < 2; w++){
				if(rowO[u0 + w] < 0) continue;
				if(EXACT){ acc += vn[w]; acc += oth[w] - sR; }
				P[xo[w]] = vn[w];
			}
		}
		if(EXACT){
			double avg = blockSumC(K, acc)/nTot;
			sPrev = sR; sR = avg;
		} else {
			__syncthreads();
		}
	}
	if(EXACT){
		for(int i = threadIdx.x; i < nOwn; i += blockDim.x){
			int j,k,l; ownNode(L,1,i,j,k,l);
			double v = P[i];
			if((j+k+l)&1) v -= sPrev;
			v -= sR;
			P[i] = v;
		}
		__syncthreads();
	} else {
		cNeutralizePhi<true>(L, K);
	}
}
template<bool SMALL, bool EXACT> static __device__ __forceinline__ void cGS(const CLvl &L, int n, double sIn, CK &K){
	if(SMALL) cGSSmall<EXACT>(L, n, sIn, K); else cGSBig<EXACT>(L, n, sIn, K);
}

// residual of level L at true node (j,k,l), any owner: -6 phi; += six neighbours; += rho
static __device__ __forceinline__ double cResAt(const CLvl &L, const CK &K, int j, int k, int l){
	double r = -6.*rdPhi(L, K, j, k, l);
	r += rdPhi(L,K,upW(j,L.nx),k,l) + rdPhi(L,K,dnW(j,L.nx),k,l)
	   + rdPhi(L,K,j,upW(k,L.ny),l) + rdPhi(L,K,j,dnW(k,L.ny),l)
	   + rdPhi(L,K,j,k,upW(l,L.nz)) + rdPhi(L,K,j,k,dnW(l,L.nz));
	r += rdRho(L, K, j, k, l);
	return r;
}

// pre-smoothing leg of level q: gBnd(rho); mgGS3D; mgResidual + mgHalfRestrict3D fused into rho(q+1).
// A coarse node is evaluated by the CTA that owns its fine centre plane (2L-1), wherever the coarse level lives.
template<bool SMALL, bool EXACT> static __device__ __noinline__ void cDown(const CPlan &P, int q, CK &K){
	const CLvl &L = P.L[q], &C = P.L[q+1];
	cNeutralizeRho<SMALL>(L, K);
	cGS<SMALL,EXACT>(L, P.nPre, 0.0, K);
	ProfScope ps(K, PS_RESTRICT);
	ProfScope psl(K, SMALL ? psLvl(L, 1) : 15);
	int l0, nl; ownPlanes(L, K, l0, nl);
	int Lz0 = l0/2 + 1;                       // first coarse plane with 2*Lz-1 >= l0
	int Lz1 = (l0 + nl)/2;                    // last coarse plane with 2*Lz-1 <= l0+nl-1
	int nC = Lz1 - Lz0 + 1;
	if(nl <= 0) nC = 0;
	int n = C.nx*C.ny*(nC > 0 ? nC : 0);
	for(int i = threadIdx.x; i < n; i += blockDim.x){
		unsigned u = (unsigned)i, cx = (unsigned)C.nx, cy = (unsigned)C.ny;
		unsigned t = u/cx; int J = (int)(u - t*cx) + 1; unsigned lz = t/cy; int Kk = (int)(t - lz*cy) + 1; int Lz = Lz0 + (int)lz;
		int j = 2*J-1, k = 2*Kk-1, l = 2*Lz-1;
		const double coeff = 1./12.;
		double v = coeff*(6*cResAt(L,K,j,k,l)
			+ cResAt(L,K,upW(j,L.nx),k,l) + cResAt(L,K,dnW(j,L.nx),k,l)
			+ cResAt(L,K,j,upW(k,L.ny),l) + cResAt(L,K,j,dnW(k,L.ny),l)
			+ cResAt(L,K,j,k,upW(l,L.nz)) + cResAt(L,K,j,k,dnW(l,L.nz)));
		wrRho(C, K, J, Kk, Lz, v);
	}
	syncAll<SMALL>(K);
}
template<bool SMALL, bool EXACT> __device__ __noinline__ void cBottom(const CPlan &P, CK &K){
	const CLvl &L = P.L[P.nLevels-1];
	cNeutralizeRho<SMALL>(L, K);
	cGS<SMALL,EXACT>(L, P.nCoarse, 0.0, K);
	if(EXACT || P.nCoarse <= 0) cNeutralizePhi<SMALL>(L, K);      // batched mode: cGS just ended with this gBnd
}
// trilinear prolongation in the nesting of the reference's three passes (z, then y, then x; multigrid.c:1127-1238)
// (branch-free: for an odd fine index both coarse neighbours are the same node and 0.5*(a + a) == a exactly, see prolPoint)
static __device__ __forceinline__ void cProlPair(int i, int n, int &a, int &b){ a = (i+1) >> 1; b = (i+2) >> 1; if(b > n) b = 1; }
static __device__ __forceinline__ double cProl(const CLvl &C, const CK &K, int j, int k, int l){
	int Ja, Jb, Ka, Kb, La, Lb;
	cProlPair(j, C.nx, Ja, Jb); cProlPair(k, C.ny, Ka, Kb); cProlPair(l, C.nz, La, Lb);
	const double v000 = rdPhi(C, K, Ja, Ka, La), v001 = rdPhi(C, K, Ja, Ka, Lb), v010 = rdPhi(C, K, Ja, Kb, La), v011 = rdPhi(C, K, Ja, Kb, Lb);
	const double v100 = rdPhi(C, K, Jb, Ka, La), v101 = rdPhi(C, K, Jb, Ka, Lb), v110 = rdPhi(C, K, Jb, Kb, La), v111 = rdPhi(C, K, Jb, Kb, Lb);
	const double ya = 0.5*(0.5*(v000 + v001) + 0.5*(v010 + v011));
	const double yb = 0.5*(0.5*(v100 + v101) + 0.5*(v110 + v111));
	return 0.5*(ya + yb);
}
// post-smoothing leg of level q: res(q) := P(phi(q+1)); phi(q) += res(q); gBnd; mgGS3D; gBnd
template<bool SMALL, bool EXACT> static __device__ __noinline__ void cUp(const CPlan &P, int q, CK &K){
	const CLvl &L = P.L[q], &C = P.L[q+1];
	ProfScope ps(K, PS_PROLONG);
	long long tP = clock64();
	int l0, nl; ownPlanes(L, K, l0, nl);
	int n = L.nx*L.ny*nl;
	double *S = mgS + L.offPhi;
	double acc = 0;
	for(int i = threadIdx.x; i < n; i += blockDim.x){
		int j,k,l; ownNode(L,l0,i,j,k,l);
		double p = cProl(C, K, j, k, l);
		L.res[gix(L,j,k,l)] = p;
		double v = S[i]; v += p;
		S[i] = v;
		acc += v;
	}
	double avg = sumAll<SMALL>(K, acc)/((double)L.nx*L.ny*L.nz);
	ps.~ProfScope(); ps.p = nullptr;
	if(SMALL && K.prof && K.rank == 0 && threadIdx.x == 0){ K.prof[2*psLvl(L, 2)] += clock64() - tP; K.prof[2*psLvl(L, 2)+1] += 1; }
	cGS<SMALL,EXACT>(L, P.nPost, avg, K);
	if(EXACT || P.nPost <= 0) cNeutralizePhi<SMALL>(L, K);
}
static __device__ __noinline__ void cGhosts(double *v, const CLvl &L, const CK &K){
	int s0 = L.s0, s1 = L.s1, s2 = L.nz + 2;
	int n = s0*s1*s2;
	for(int i = K.rank*blockDim.x + threadIdx.x; i < n; i += K.nc*blockDim.x){
		int j = i % s0; int r = i / s0; int k = r % s1; int l = r / s1;
		int jw = j == 0 ? s0-2 : (j == s0-1 ? 1 : j);
		int kw = k == 0 ? s1-2 : (k == s1-1 ? 1 : k);
		int lw = l == 0 ? s2-2 : (l == s2-1 ? 1 : l);
		if(jw != j || kw != k || lw != l) v[i] = __ldcg(v + (jw + (long)s0*(kw + (long)s1*lw)));
	}
}


} // namespace pinc
