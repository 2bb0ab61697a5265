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
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace pinc {

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
};

struct CK {
	cg::cluster_group cl;
	int rank, nc;
	double *sm;          // dynamic shared memory
	double *red;         // static shared: [0..1] cluster-sum slots, [2] block total, [4..35] per-warp scratch
	int flip;
};

__device__ __forceinline__ int upW(int j, int n){ return j == n ? 1 : j+1; }
__device__ __forceinline__ int dnW(int j, int n){ return j == 1 ? n : j-1; }
__device__ __forceinline__ long gix(const CLvl &L, int j, int k, int l){ return j + (long)L.s0*(k + (long)L.s1*l); }

// pointer to node (1,1,l) of plane l of an array that is distributed like phi (offset `off`), any owner
__device__ __forceinline__ double *planePtr(const CLvl &L, const CK &K, int off, int l){
	int r = (l-1)/L.ppc;
	int lp = (l-1) - r*L.ppc;
	double *base = K.sm + off;
	if(r != K.rank) base = K.cl.map_shared_rank(base, r);
	return base + lp*L.ny*L.nx;
}
__device__ __forceinline__ double rdPhi(const CLvl &L, const CK &K, int j, int k, int l){
	return planePtr(L, K, L.offPhi, l)[(k-1)*L.nx + (j-1)];
}
__device__ __forceinline__ double rdRho(const CLvl &L, const CK &K, int j, int k, int l){
	if(L.offRho < 0) return __ldcg(L.rho + gix(L,j,k,l));
	return planePtr(L, K, L.offRho, l)[(k-1)*L.nx + (j-1)];
}
__device__ __forceinline__ void wrRho(const CLvl &L, const CK &K, int j, int k, int l, double v){
	if(L.offRho < 0) L.rho[gix(L,j,k,l)] = v;
	else planePtr(L, K, L.offRho, l)[(k-1)*L.nx + (j-1)] = v;
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
	double tot = K.red[2];
	__syncthreads();
	return tot;
}
// sum over all participating threads (CTA 0 for a small level, the cluster otherwise); identical bits in every
// thread; acts as the scope's barrier
template<bool SMALL> __device__ __forceinline__ double sumAll(CK &K, double v){
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
template<bool SMALL> __device__ __forceinline__ void syncAll(CK &K){
	if(SMALL) __syncthreads(); else K.cl.sync();
}

__device__ __forceinline__ void ownPlanes(const CLvl &L, const CK &K, int &l0, int &nl){
	l0 = K.rank*L.ppc + 1;
	nl = L.nz - K.rank*L.ppc;
	if(nl > L.ppc) nl = L.ppc;
	if(nl < 0) nl = 0;
}
// flat index over the own true nodes -> (j,k,l) and the slab offset
__device__ __forceinline__ void ownNode(const CLvl &L, int l0, int i, int &j, int &k, int &l){
	unsigned u = (unsigned)i, nx = (unsigned)L.nx, ny = (unsigned)L.ny;
	unsigned t = u / nx; j = (int)(u - t*nx) + 1;
	unsigned lp = t / ny; k = (int)(t - lp*ny) + 1; l = l0 + (int)lp;
}

// gNeutralizeGrid of rho (src/grid.c:730-779) on the own planes
template<bool SMALL> __device__ __noinline__ void cNeutralizeRho(const CLvl &L, CK &K){
	int l0, nl; ownPlanes(L, K, l0, nl);
	int n = L.nx*L.ny*nl;
	double avg;
	if(L.offRho >= 0){
		double *R = K.sm + L.offRho;
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
template<bool SMALL> __device__ __noinline__ void cNeutralizePhi(const CLvl &L, CK &K){
	int l0, nl; ownPlanes(L, K, l0, nl);
	int n = L.nx*L.ny*nl;
	double *P = K.sm + L.offPhi;
	double acc = 0;
	for(int i = threadIdx.x; i < n; i += blockDim.x) acc += P[i];
	double avg = sumAll<SMALL>(K, acc)/((double)L.nx*L.ny*L.nz);
	for(int i = threadIdx.x; i < n; i += blockDim.x) P[i] -= avg;
	syncAll<SMALL>(K);
}

// one Gauss-Seidel node: 1/6 * (x+ + x- + y+ + y- + z+ + z- + rho), summed left to right (src/multigrid.c:711-714)
template<bool EXACT> __device__ __forceinline__ double gsVal(double a, double b, double c, double d, double e, double f, double rho, double sR){
	if(EXACT){ a -= sR; b -= sR; c -= sR; d -= sR; e -= sR; f -= sR; }
	const double coeff = 1./6.;
	return coeff*(a + b + c + d + e + f + rho);
}

// mgGS3D (src/multigrid.c:683-767) on a big level.  sIn: mean shift still pending on every value at entry.
template<bool EXACT> __device__ __noinline__ void cGSBig(const CLvl &L, int nCycles, double sIn, CK &K){
	int l0, nl; ownPlanes(L, K, l0, nl);
	const int nx = L.nx, ny = L.ny, nz = L.nz, pl = nx*ny;
	double *P = K.sm + L.offPhi;
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
			const double *Rl = rhoS ? (K.sm + L.offRho + lp*pl) : (L.rho + gix(L,1,0,l));
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
template<bool EXACT> __device__ __noinline__ void cGSSmall(const CLvl &L, int nCycles, double sIn, CK &K){
	const int nx = L.nx, ny = L.ny, nz = L.nz, pl = nx*ny;
	double *P = K.sm + L.offPhi;
	const double *R = K.sm + L.offRho;
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
		double vn[MC_U], oth[MC_U];
		int xo[MC_U];
		#pragma unroll
		for(int u = 0; u < MC_U; u++){
			if(rowO[u] < 0) continue;
			int j = ((par[u] == parity) ? 1 : 2) + m2[u];
			int jm = j-1;
			xo[u] = rowO[u] + jm;
			double a = P[rowO[u] + upW(j,nx)-1], b = P[rowO[u] + dnW(j,nx)-1];
			double c = P[yu[u] + jm], d = P[yd[u] + jm], e = P[zu[u] + jm], f = P[zd[u] + jm];
			vn[u] = gsVal<EXACT>(a, b, c, d, e, f, R[xo[u]], sR);
			if(EXACT) oth[u] = P[rowO[u] + (jm^1)];
		}
		double acc = 0;
		#pragma unroll
		for(int u = 0; u < MC_U; u++){
			if(rowO[u] < 0) continue;
			if(EXACT){ acc += vn[u]; acc += oth[u] - sR; }
			P[xo[u]] = vn[u];
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
template<bool SMALL, bool EXACT> __device__ __forceinline__ void cGS(const CLvl &L, int n, double sIn, CK &K){
	if(SMALL) cGSSmall<EXACT>(L, n, sIn, K); else cGSBig<EXACT>(L, n, sIn, K);
}

// residual of level L at true node (j,k,l), any owner: -6 phi; += six neighbours; += rho
__device__ __forceinline__ double cResAt(const CLvl &L, const CK &K, int j, int k, int l){
	double r = -6.*rdPhi(L, K, j, k, l);
	r += rdPhi(L,K,upW(j,L.nx),k,l) + rdPhi(L,K,dnW(j,L.nx),k,l)
	   + rdPhi(L,K,j,upW(k,L.ny),l) + rdPhi(L,K,j,dnW(k,L.ny),l)
	   + rdPhi(L,K,j,k,upW(l,L.nz)) + rdPhi(L,K,j,k,dnW(l,L.nz));
	r += rdRho(L, K, j, k, l);
	return r;
}

// pre-smoothing leg of level q: gBnd(rho); mgGS3D; mgResidual + mgHalfRestrict3D fused into rho(q+1).
// A coarse node is evaluated by the CTA that owns its fine centre plane (2L-1), wherever the coarse level lives.
template<bool SMALL, bool EXACT> __device__ __noinline__ void cDown(const CPlan &P, int q, CK &K){
	const CLvl &L = P.L[q], &C = P.L[q+1];
	cNeutralizeRho<SMALL>(L, K);
	cGS<SMALL,EXACT>(L, P.nPre, 0.0, K);
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
template<bool SMALL, bool EXACT> __device__ void cBottom(const CPlan &P, CK &K){
	const CLvl &L = P.L[P.nLevels-1];
	cNeutralizeRho<SMALL>(L, K);
	cGS<SMALL,EXACT>(L, P.nCoarse, 0.0, K);
	cNeutralizePhi<SMALL>(L, K);
}
// trilinear prolongation in the nesting of the reference's three passes (z, then y, then x; multigrid.c:1127-1238)
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
// post-smoothing leg of level q: res(q) := P(phi(q+1)); phi(q) += res(q); gBnd; mgGS3D; gBnd
template<bool SMALL, bool EXACT> __device__ __noinline__ void cUp(const CPlan &P, int q, CK &K){
	const CLvl &L = P.L[q], &C = P.L[q+1];
	int l0, nl; ownPlanes(L, K, l0, nl);
	int n = L.nx*L.ny*nl;
	double *S = K.sm + L.offPhi;
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
	cGS<SMALL,EXACT>(L, P.nPost, avg, K);
	cNeutralizePhi<SMALL>(L, K);
}
__device__ __noinline__ void cGhosts(double *v, const CLvl &L, const CK &K){
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
	extern __shared__ double dyn[];
	__shared__ double red[40];
	CK K{ cg::this_cluster(), 0, P.nc, dyn, red, 0 };
	K.rank = (int)K.cl.block_rank();
	const int b = P.nLevels - 1;
	// phi of every level, and rho of the levels that keep it in shared memory: global -> shared (own planes)
	for(int q = 0; q <= b; q++){
		const CLvl &L = P.L[q];
		int l0, nl; ownPlanes(L, K, l0, nl);
		int n = L.nx*L.ny*nl;
		for(int i = threadIdx.x; i < n; i += blockDim.x){
			int j,k,l; ownNode(L,l0,i,j,k,l);
			K.sm[L.offPhi + i] = __ldcg(L.phiG + gix(L,j,k,l));
			if(L.offRho >= 0) K.sm[L.offRho + i] = __ldcg(L.rho + gix(L,j,k,l));
		}
	}
	K.cl.sync();
	double barRes = 2.;
	int cycles = 0;
	while(barRes > P.tol && cycles < P.maxCycles){
		if(P.exact) vcycle<true>(P, K); else vcycle<false>(P, K);
		// mgSolveRaw :1700-1704: residual of level 0, squared in place, true-grid sum, RMS
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
	if(K.rank == 0 && threadIdx.x == 0) P.hist[0] = (double)cycles;
	// back to global (phi of every level, rho where it lived in shared memory), then the ghost layers of every
	// array: the state the reference leaves behind
	for(int q = 0; q <= b; q++){
		const CLvl &L = P.L[q];
		int l0, nl; ownPlanes(L, K, l0, nl);
		int n = L.nx*L.ny*nl;
		for(int i = threadIdx.x; i < n; i += blockDim.x){
			int j,k,l; ownNode(L,l0,i,j,k,l);
			L.phiG[gix(L,j,k,l)] = K.sm[L.offPhi + i];
			if(L.offRho >= 0) L.rho[gix(L,j,k,l)] = K.sm[L.offRho + i];
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

// host side: returns false if this solve does not fit the cluster kernel (the caller falls back)
bool clusterSolve(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, double tol, int maxCycles, int exact){
	static int maxNc = -1;
	static size_t maxSmem = 0;
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

} // namespace pinc
