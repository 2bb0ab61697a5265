// multigrid.cu — geometric multigrid Poisson solver (src/multigrid.c): red-black Gauss-Seidel (mgGS3D :683),
// residual (:1385), half-weight restriction (:844), trilinear prolongation (:1127), recursive V-cycle (:1496)
// and the tolerance loop of mgSolveRaw (:1688).
//
// Two execution modes with the SAME per-point arithmetic (the device functions below):
//
//  ops    one kernel per reference call, ghost layers kept current by gHaloOp/gBnd exactly as the
//         reference does.  Used for the individual entry points and for distributed multi-rank solves (the fallback:
//         by default a multi-rank solve is REPLICATED - every rank gathers rho/phi and runs the fused kernel on the
//         global problem, replicaSolve below).
//  fused  single-rank periodic solves: ONE persistent cooperative kernel (k_mg_solve) runs the whole tolerance loop.
//         Ghost layers are not touched inside (neighbours are read at the periodic image of the true node, which is
//         bit-identical to reading a ghost that setSlice just filled) and are filled once at the end, so the arrays are
//         left in the state the reference leaves them in.  Grid-wide levels are smoothed block-resident in shared
//         memory with the faces exchanged through tagged mailboxes in L2 (bGS); levels of <= 4096 nodes run inside
//         CTA 0 (mgsmem.cuh, mgsmall.cuh); transfers are grid-stride phases between grid barriers.  gBnd's mean
//         subtraction is applied once per smoother call (mode 2) or carried bit-faithfully as a pending shift that is
//         applied when a value is next read (modes 1, 3).
//
// The grids of a level are tiny (<= 2.2 M nodes): the solver is bound by the latency of 180 dependent half-sweeps per
// V-cycle and ~50 V-cycles per solve, not by HBM, which is why the fused mode exists.
#include "common.h"
#include <cmath>
#include "mgsmem.cuh"

namespace pinc {

struct Lvl { double *phi, *rho, *res; int s0, s1, s2; };

// loads of data other CTAs wrote in an earlier phase of the same kernel go to L2 (measured: L1-cached loads, which the
// release/acquire grid barrier would allow, do not shorten the sweep: it is latency-, not L2-bandwidth-bound)
__device__ __forceinline__ double ldg2(const double *p){ return __ldcg(p); }
template<bool WRAP> __device__ __forceinline__ int upI(int j, int s){ return (WRAP && j == s-2) ? 1 : j+1; }
template<bool WRAP> __device__ __forceinline__ int dnI(int j, int s){ return (WRAP && j == 1) ? s-2 : j-1; }
__device__ __forceinline__ long ix(int j, int k, int l, int s0, int s1){ return j + (long)s0*(k + (long)s1*l); }

// phi_new = 1/6 * (phi[+j] + phi[-j] + phi[+k] + phi[-k] + phi[+l] + phi[-l] + rho), summed left to right
// (src/multigrid.c:711-714); `sh` is the pending mean shift of the neighbours' colour.
template<bool WRAP> __device__ __forceinline__ double gsPoint(const double *phi, const double *rho, int j, int k, int l,
		int s0, int s1, int s2, double sh){
	long g = ix(j,k,l,s0,s1);
	double a = ldg2(phi + ix(upI<WRAP>(j,s0),k,l,s0,s1)) - sh;
	double b = ldg2(phi + ix(dnI<WRAP>(j,s0),k,l,s0,s1)) - sh;
	double c = ldg2(phi + ix(j,upI<WRAP>(k,s1),l,s0,s1)) - sh;
	double d = ldg2(phi + ix(j,dnI<WRAP>(k,s1),l,s0,s1)) - sh;
	double e = ldg2(phi + ix(j,k,upI<WRAP>(l,s2),s0,s1)) - sh;
	double f = ldg2(phi + ix(j,k,dnI<WRAP>(l,s2),s0,s1)) - sh;
	const double coeff = 1./6.;
	return coeff*(a + b + c + d + e + f + ldg2(rho + g));
}
// res = -6 phi; res += (sum of six neighbours); res += rho  (src/grid.c:318-322, src/multigrid.c:1400)
// sh: a mean shift that is still pending on every stored value of phi (the block smoother leaves its gBnd to the readers, see
// bGS): the value the reference would have stored is (stored - sh), formed here before anything else is done with it
template<bool WRAP> __device__ __forceinline__ double resPoint(const double *phi, const double *rho, int j, int k, int l,
		int s0, int s1, int s2, double sh = 0.0){
	long g = ix(j,k,l,s0,s1);
	double r = -6.*(ldg2(phi + g) - sh);
	r += (ldg2(phi + ix(upI<WRAP>(j,s0),k,l,s0,s1)) - sh) + (ldg2(phi + ix(dnI<WRAP>(j,s0),k,l,s0,s1)) - sh)
	   + (ldg2(phi + ix(j,upI<WRAP>(k,s1),l,s0,s1)) - sh) + (ldg2(phi + ix(j,dnI<WRAP>(k,s1),l,s0,s1)) - sh)
	   + (ldg2(phi + ix(j,k,upI<WRAP>(l,s2),s0,s1)) - sh) + (ldg2(phi + ix(j,k,dnI<WRAP>(l,s2),s0,s1)) - sh);
	r += ldg2(rho + g);
	return r;
}
// coarse(J,K,L) = 1/12 * (6 f + f[+j] + f[-j] + f[+k] + f[-k] + f[+l] + f[-l]) centred on fine (2J-1,2K-1,2L-1)
// (src/multigrid.c:844-911)
template<bool WRAP> __device__ __forceinline__ double restrictPoint(const double *f, int J, int K, int L, int s0, int s1, int s2){
	int j = 2*J-1, k = 2*K-1, l = 2*L-1;
	const double coeff = 1./12.;
	return coeff*(6*ldg2(f + ix(j,k,l,s0,s1))
		+ ldg2(f + ix(upI<WRAP>(j,s0),k,l,s0,s1)) + ldg2(f + ix(dnI<WRAP>(j,s0),k,l,s0,s1))
		+ ldg2(f + ix(j,upI<WRAP>(k,s1),l,s0,s1)) + ldg2(f + ix(j,dnI<WRAP>(k,s1),l,s0,s1))
		+ ldg2(f + ix(j,k,upI<WRAP>(l,s2),s0,s1)) + ldg2(f + ix(j,k,dnI<WRAP>(l,s2),s0,s1)));
}
// Trilinear prolongation of the coarse grid to fine true node (j,k,l), in the nesting the reference's three
// passes produce (z first, then y, then x; src/multigrid.c:1127-1238).  Periodic wrap on the coarse index.
// Branch-free: for an odd fine index both coarse neighbours are the same node, and 0.5*(a + a) == a exactly, so every node
// takes the eight-load form - no divergence between the eight parity classes of a warp's nodes, and the eight loads (L2 or
// shared memory) are in flight together instead of one dependent load per taken branch.
__device__ __forceinline__ void prolPair(int i, int cN, int &a, int &b){ a = (i+1) >> 1; b = (i+2) >> 1; if(b == cN-1) b = 1; }
__device__ __forceinline__ double prolPoint(const double *c, int j, int k, int l, int c0, int c1, int c2, double sh = 0.0){
	int Ja, Jb, Ka, Kb, La, Lb;
	prolPair(j, c0, Ja, Jb); prolPair(k, c1, Ka, Kb); prolPair(l, c2, La, Lb);
	const double v000 = ldg2(c + ix(Ja,Ka,La,c0,c1)) - sh, v001 = ldg2(c + ix(Ja,Ka,Lb,c0,c1)) - sh;       // (sh: see resPoint)
	const double v010 = ldg2(c + ix(Ja,Kb,La,c0,c1)) - sh, v011 = ldg2(c + ix(Ja,Kb,Lb,c0,c1)) - sh;
	const double v100 = ldg2(c + ix(Jb,Ka,La,c0,c1)) - sh, v101 = ldg2(c + ix(Jb,Ka,Lb,c0,c1)) - sh;
	const double v110 = ldg2(c + ix(Jb,Kb,La,c0,c1)) - sh, v111 = ldg2(c + ix(Jb,Kb,Lb,c0,c1)) - sh;
	const double ya = 0.5*(0.5*(v000 + v001) + 0.5*(v010 + v011));       // prolY at Ja: 0.5*(prolZ(Ka) + prolZ(Kb))
	const double yb = 0.5*(0.5*(v100 + v101) + 0.5*(v110 + v111));
	return 0.5*(ya + yb);
}

// (levels have fewer than 2^31 nodes: 32-bit divisions, ~10x cheaper than 64-bit ones)
__device__ __forceinline__ void truePoint(long i, int t0, int t1, int &j, int &k, int &l){
	unsigned u = (unsigned)i, r = u / (unsigned)t0;
	j = (int)(u - r*(unsigned)t0) + 1;
	unsigned q = r / (unsigned)t1;
	k = (int)(r - q*(unsigned)t1) + 1; l = (int)q + 1;
}

// =================================================================================================
// ops mode: one kernel per reference call (ghost layers are read, not wrapped)
// =================================================================================================
__global__ void k_gs_colour(double *__restrict__ phi, const double *__restrict__ rho, int s0, int s1, int s2, int parity){
	int t0 = s0-2, t1 = s1-2, t2 = s2-2;
	long nt = (long)t0*t1*t2;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < nt; i += st){
		int j, k, l; truePoint(i, t0, t1, j, k, l);
		if(((j+k+l)&1) != parity) continue;
		phi[ix(j,k,l,s0,s1)] = gsPoint<false>(phi, rho, j, k, l, s0, s1, s2, 0.0);
	}
}
__global__ void k_residual(double *__restrict__ res, const double *__restrict__ rho, const double *__restrict__ phi, int s0, int s1, int s2){
	int t0 = s0-2, t1 = s1-2, t2 = s2-2;
	long nt = (long)t0*t1*t2;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < nt; i += st){
		int j, k, l; truePoint(i, t0, t1, j, k, l);
		res[ix(j,k,l,s0,s1)] = resPoint<false>(phi, rho, j, k, l, s0, s1, s2);
	}
}
__global__ void k_restrict(const double *__restrict__ f, int s0, int s1, int s2, double *__restrict__ cgrid, int c0, int c1, int c2){
	int t0 = c0-2, t1 = c1-2, t2 = c2-2;
	long nt = (long)t0*t1*t2;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < nt; i += st){
		int J, K, L; truePoint(i, t0, t1, J, K, L);
		cgrid[ix(J,K,L,c0,c1)] = restrictPoint<false>(f, J, K, L, s0, s1, s2);
	}
}
// the reference's three prolongation passes; pass 0 injects, pass 1..3 interpolate along z, y, x
__global__ void k_prolong_pass(double *__restrict__ f, int s0, int s1, int s2, const double *__restrict__ cgrid, int c0, int c1, int pass){
	int t0 = s0-2, t1 = s1-2, t2 = s2-2;
	long nt = (long)t0*t1*t2;
	long sx = s0, sxy = (long)s0*s1;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < nt; i += st){
		int j, k, l; truePoint(i, t0, t1, j, k, l);
		long g = ix(j,k,l,s0,s1);
		if(pass == 0){ if((j&1) && (k&1) && (l&1)) f[g] = cgrid[ix((j+1)/2,(k+1)/2,(l+1)/2,c0,c1)]; }
		else if(pass == 1){ if((j&1) && (k&1) && !(l&1)) f[g] = 0.5*(f[g-sxy] + f[g+sxy]); }
		else if(pass == 2){ if((j&1) && !(k&1)) f[g] = 0.5*(f[g-sx] + f[g+sx]); }
		else { if(!(j&1)) f[g] = 0.5*(f[g-1] + f[g+1]); }
	}
}


static inline int tGrid(Ctx *c, long nt){ return gridFor(nt, 256, c->numSMs*8); }
static inline long trueCount(const DevGrid *g){ return (long)g->tsize[0]*g->tsize[1]*g->tsize[2]; }

// one colour with per-dimension periodic wrap (bit d of wrapMask: dimension d is not decomposed, read the periodic
// image instead of the ghost); the other dimensions read ghosts that gridHaloFaces keeps current
__global__ void k_gs_colour_w(double *__restrict__ phi, const double *__restrict__ rho, int s0, int s1, int s2, int parity, int wrapMask){
	int t0 = s0-2, t1 = s1-2, t2 = s2-2;
	long nt = (long)t0*t1*t2;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < nt; i += st){
		int j, k, l; truePoint(i, t0, t1, j, k, l);
		if(((j+k+l)&1) != parity) continue;
		int ju = (wrapMask&1) ? upI<true>(j,s0) : j+1, jd = (wrapMask&1) ? dnI<true>(j,s0) : j-1;
		int ku = (wrapMask&2) ? upI<true>(k,s1) : k+1, kd = (wrapMask&2) ? dnI<true>(k,s1) : k-1;
		int lu = (wrapMask&4) ? upI<true>(l,s2) : l+1, ld = (wrapMask&4) ? dnI<true>(l,s2) : l-1;
		const double coeff = 1./6.;
		phi[ix(j,k,l,s0,s1)] = coeff*(ldg2(phi + ix(ju,k,l,s0,s1)) + ldg2(phi + ix(jd,k,l,s0,s1)) + ldg2(phi + ix(j,ku,l,s0,s1))
			+ ldg2(phi + ix(j,kd,l,s0,s1)) + ldg2(phi + ix(j,k,lu,s0,s1)) + ldg2(phi + ix(j,k,ld,s0,s1)) + ldg2(rho + ix(j,k,l,s0,s1)));
	}
}

// ---- multi-rank smoother over peer memory (NVLink): one kernel per half-sweep, no pack/unpack, no NCCL call --------
// A rank's ghost values of phi live in its six mailbox planes (P2P in common.h).  A half-sweep kernel (1) waits until
// the neighbours' previous half-sweep has arrived (their arrival counters in MY arena), (2) updates its colour,
// reading ghosts of the decomposed dimensions from the mailbox, and stores every boundary node it updates straight
// into the neighbour's mailbox plane, (3) the last block to finish fences and bumps the neighbours' counters.  Nodes of
// one colour only read the other colour, so a neighbour that is one half-sweep ahead never overwrites what is being read.
static int mgModeResolved();
struct P2PArgs {
	const double *myMail[6];        // [2*dd + side]: neighbour's boundary layer on my lower (0) / upper (1) side of dim dd
	double *peerMail[6];            // [2*dd + 0]: upper neighbour's lower-side plane, [2*dd + 1]: lower neighbour's upper-side plane
	const unsigned long long *myFlag;     // [6] arrival counters in my arena
	unsigned long long *peerFlag[6];      // counter of peerMail[i] in the neighbour's arena
	unsigned long long *ticket;           // block counter (local)
	const unsigned long long *seqBase;    // device word: sequence number = *seqBase + seq (lets a captured CUDA graph be replayed)
	unsigned long long seq;               // offset from *seqBase
	int active[3];
};
__device__ __forceinline__ void p2pWait(const P2PArgs &A, unsigned long long need, int *flags){
	if(threadIdx.x == 0){
		long long t0 = clock64();
		for(int i = 0; i < 6; i++){
			if(!A.active[i>>1]) continue;
			while(*((volatile const unsigned long long*)&A.myFlag[i]) < need){
				if(clock64() - t0 > 6000000000LL){ atomicOr(flags, ERR_P2P_TIMEOUT); break; }     // ~3 s: report, do not hang
			}
		}
		__threadfence_system();
	}
	__syncthreads();
}
__device__ __forceinline__ void p2pSignal(const P2PArgs &A){
	__syncthreads();
	if(threadIdx.x == 0){
		__threadfence_system();
		unsigned long long old = atomicAdd(A.ticket, 1ULL);
		if(old == gridDim.x - 1){
			atomicExch(A.ticket, 0ULL);
			__threadfence_system();
			const unsigned long long sq = *A.seqBase + A.seq;
			for(int i = 0; i < 6; i++) if(A.active[i>>1]) *((volatile unsigned long long*)A.peerFlag[i]) = sq;
		}
	}
}
__device__ __forceinline__ double ldv(const double *p){ return *((volatile const double*)p); }
// publish both boundary layers of every decomposed dimension (start of a smoother call)
__device__ __forceinline__ void p2pPublishPlanes(const double *__restrict__ phi, int s0, int s1, int s2, const P2PArgs &A, int dimMask){
	long st = (long)gridDim.x*blockDim.x, i0 = blockIdx.x*(long)blockDim.x + threadIdx.x;
	int sz[3] = {s0, s1, s2};
	for(int dd = 0; dd < 3; dd++){
		if(!A.active[dd] || !(dimMask & (1 << dd))) continue;
		int a = dd == 0 ? s1 : s0, b = dd == 2 ? s1 : s2;          // extents of the plane's two axes (ghost-inclusive)
		long np = (long)a*b;
		for(long i = i0; i < np; i += st){
			int u = (int)(i % a), v = (int)(i / a);
			int jU[3], jL[3];
			if(dd == 0){ jU[0] = sz[0]-2; jU[1] = u; jU[2] = v; } else if(dd == 1){ jU[0] = u; jU[1] = sz[1]-2; jU[2] = v; } else { jU[0] = u; jU[1] = v; jU[2] = sz[2]-2; }
			jL[0] = jU[0]; jL[1] = jU[1]; jL[2] = jU[2]; jL[dd] = 1;
			A.peerMail[2*dd][i]   = ldg2(phi + ix(jU[0],jU[1],jU[2],s0,s1));
			A.peerMail[2*dd+1][i] = ldg2(phi + ix(jL[0],jL[1],jL[2],s0,s1));
		}
	}
}
// Every peer-memory operation has a sequence number S that is the same on all ranks (same call sequence).  It starts
// when the neighbours have finished operation S-1 (so whatever they wrote for us has arrived and whatever they read
// from our last writes is done), may store into the neighbours' planes only what they do not read during their own
// operation S, and ends by publishing S in the neighbours' arrival counters.
__global__ void k_p2p_publish(const double *__restrict__ phi, int s0, int s1, int s2, P2PArgs A, int dimMask, int *flags){
	p2pWait(A, *A.seqBase + A.seq - 1, flags);
	p2pPublishPlanes(phi, s0, s1, s2, A, dimMask);
	p2pSignal(A);
}
// ghost layers of dimension dd := the planes the neighbours published (full planes, rims included)
__global__ void k_p2p_copy(double *__restrict__ phi, int s0, int s1, int s2, int dd, P2PArgs A, int *flags){
	p2pWait(A, *A.seqBase + A.seq - 1, flags);
	int a = dd == 0 ? s1 : s0, b = dd == 2 ? s1 : s2;
	long np = (long)a*b, st = (long)gridDim.x*blockDim.x;
	int sz = dd == 0 ? s0 : (dd == 1 ? s1 : s2);
	for(long i = blockIdx.x*(long)blockDim.x + threadIdx.x; i < np; i += st){
		int u = (int)(i % a), v = (int)(i / a);
		long gLo, gHi;
		if(dd == 0){ gLo = ix(0,u,v,s0,s1); gHi = ix(sz-1,u,v,s0,s1); }
		else if(dd == 1){ gLo = ix(u,0,v,s0,s1); gHi = ix(u,sz-1,v,s0,s1); }
		else { gLo = ix(u,v,0,s0,s1); gHi = ix(u,v,sz-1,s0,s1); }
		phi[gLo] = ldv(A.myMail[2*dd] + i);
		phi[gHi] = ldv(A.myMail[2*dd+1] + i);
	}
	p2pSignal(A);
}
__global__ void k_gs_p2p(double *__restrict__ phi, const double *__restrict__ rho, int s0, int s1, int s2, int parity, int wrapMask, P2PArgs A, int *flags){
	p2pWait(A, *A.seqBase + A.seq - 1, flags);
	int t0 = s0-2, t1 = s1-2, t2 = s2-2;
	long nt = (long)t0*t1*t2;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < nt; i += st){
		int j, k, l; truePoint(i, t0, t1, j, k, l);
		if(((j+k+l)&1) != parity) continue;
		double a, b, c, d, e, f;
		if(wrapMask&1){ a = ldg2(phi + ix(upI<true>(j,s0),k,l,s0,s1)); b = ldg2(phi + ix(dnI<true>(j,s0),k,l,s0,s1)); }
		else { a = j == t0 ? ldv(A.myMail[1] + (k + (long)s1*l)) : ldg2(phi + ix(j+1,k,l,s0,s1)); b = j == 1 ? ldv(A.myMail[0] + (k + (long)s1*l)) : ldg2(phi + ix(j-1,k,l,s0,s1)); }
		if(wrapMask&2){ c = ldg2(phi + ix(j,upI<true>(k,s1),l,s0,s1)); d = ldg2(phi + ix(j,dnI<true>(k,s1),l,s0,s1)); }
		else { c = k == t1 ? ldv(A.myMail[3] + (j + (long)s0*l)) : ldg2(phi + ix(j,k+1,l,s0,s1)); d = k == 1 ? ldv(A.myMail[2] + (j + (long)s0*l)) : ldg2(phi + ix(j,k-1,l,s0,s1)); }
		if(wrapMask&4){ e = ldg2(phi + ix(j,k,upI<true>(l,s2),s0,s1)); f = ldg2(phi + ix(j,k,dnI<true>(l,s2),s0,s1)); }
		else { e = l == t2 ? ldv(A.myMail[5] + (j + (long)s0*k)) : ldg2(phi + ix(j,k,l+1,s0,s1)); f = l == 1 ? ldv(A.myMail[4] + (j + (long)s0*k)) : ldg2(phi + ix(j,k,l-1,s0,s1)); }
		const double coeff = 1./6.;
		double v = coeff*(a + b + c + d + e + f + ldg2(rho + ix(j,k,l,s0,s1)));
		phi[ix(j,k,l,s0,s1)] = v;
		if(!(wrapMask&1)){ if(j == t0) A.peerMail[0][k + (long)s1*l] = v; if(j == 1) A.peerMail[1][k + (long)s1*l] = v; }
		if(!(wrapMask&2)){ if(k == t1) A.peerMail[2][j + (long)s0*l] = v; if(k == 1) A.peerMail[3][j + (long)s0*l] = v; }
		if(!(wrapMask&4)){ if(l == t2) A.peerMail[4][j + (long)s0*k] = v; if(l == 1) A.peerMail[5][j + (long)s0*k] = v; }
	}
	p2pSignal(A);
}
// The whole smoother call (publish + 2*nCycles half-sweeps) as ONE cooperative kernel: between half-sweeps a block
// fences its peer stores, takes a ticket, the last block publishes the sequence number to the neighbours and to the
// local generation word, and every block waits until both the local generation and the neighbours' counters have
// reached it.  Sequence numbers seq0 .. seq0 + 2*nCycles.
__global__ void k_gs_p2p_loop(double *__restrict__ phi, const double *__restrict__ rho, int s0, int s1, int s2, int nHalf, int wrapMask,
		P2PArgs A, unsigned long long *localGen, int *flags, long long *prof){
	const unsigned long long seq0 = *A.seqBase + A.seq;
	const bool pr = prof && blockIdx.x == 0 && threadIdx.x == 0;
	long long tA = 0, tB = 0, tC = 0, tD = 0;
	int t0 = s0-2, t1 = s1-2, t2 = s2-2;
	long nt = (long)t0*t1*t2;
	p2pWait(A, seq0 - 1, flags);
	p2pPublishPlanes(phi, s0, s1, s2, A, 7);
	for(int h = 0; h <= nHalf; h++){
		// end of operation seq0+h (h = 0: the publish): fence, ticket, last block signals
		const unsigned long long seq = seq0 + h;
		if(pr) tA = clock64();
		__syncthreads();
		if(threadIdx.x == 0){
			if(pr){ tB = clock64(); prof[20] += tB - tA; }        // sweep tail: waiting for the block's other warps
			__threadfence_system();
			if(pr){ tC = clock64(); prof[22] += tC - tB; prof[21] += 1; }   // system fence
			unsigned long long old = atomicAdd(A.ticket, 1ULL);
			if(old == gridDim.x - 1){
				atomicExch(A.ticket, 0ULL);
				__threadfence();          // the flag stores below are issued after the ticket is known (control dependency)
				for(int i = 0; i < 6; i++) if(A.active[i>>1]) *((volatile unsigned long long*)A.peerFlag[i]) = seq;
				*((volatile unsigned long long*)localGen) = seq;
			}
			if(h < nHalf){
				long long c0 = clock64();
				bool ok = false;
				while(!ok){
					ok = *((volatile unsigned long long*)localGen) >= seq;
					for(int i = 0; i < 6 && ok; i++) if(A.active[i>>1] && *((volatile const unsigned long long*)&A.myFlag[i]) < seq) ok = false;
					if(!ok && clock64() - c0 > 6000000000LL){ atomicOr(flags, ERR_P2P_TIMEOUT); break; }
				}
				__threadfence_system();       // acquire: the mailbox loads below must not be satisfied before the counters were seen
				if(pr){ tD = clock64(); prof[24] += tD - tC; }      // ticket + local generation + neighbours' counters
			}
		}
		__syncthreads();
		if(h == nHalf) break;
		if(pr) tA = clock64();
		const int parity = (h & 1) ? 0 : 1;
		// two passes: first the nodes on a decomposed boundary (their peer stores are in flight while the interior is
		// swept, so the system fence at the end finds them acknowledged), then the rest
		for(int pass = 0; pass < 2; pass++){
			long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
			for(; i < nt; i += st){
				int j, k, l; truePoint(i, t0, t1, j, k, l);
				if(((j+k+l)&1) != parity) continue;
				bool edge = (!(wrapMask&1) && (j == 1 || j == t0)) || (!(wrapMask&2) && (k == 1 || k == t1)) || (!(wrapMask&4) && (l == 1 || l == t2));
				if(edge != (pass == 0)) continue;
				double a, b, c, d, e, f;
				if(wrapMask&1){ a = ldg2(phi + ix(upI<true>(j,s0),k,l,s0,s1)); b = ldg2(phi + ix(dnI<true>(j,s0),k,l,s0,s1)); }
				else { a = j == t0 ? ldv(A.myMail[1] + (k + (long)s1*l)) : ldg2(phi + ix(j+1,k,l,s0,s1)); b = j == 1 ? ldv(A.myMail[0] + (k + (long)s1*l)) : ldg2(phi + ix(j-1,k,l,s0,s1)); }
				if(wrapMask&2){ c = ldg2(phi + ix(j,upI<true>(k,s1),l,s0,s1)); d = ldg2(phi + ix(j,dnI<true>(k,s1),l,s0,s1)); }
				else { c = k == t1 ? ldv(A.myMail[3] + (j + (long)s0*l)) : ldg2(phi + ix(j,k+1,l,s0,s1)); d = k == 1 ? ldv(A.myMail[2] + (j + (long)s0*l)) : ldg2(phi + ix(j,k-1,l,s0,s1)); }
				if(wrapMask&4){ e = ldg2(phi + ix(j,k,upI<true>(l,s2),s0,s1)); f = ldg2(phi + ix(j,k,dnI<true>(l,s2),s0,s1)); }
				else { e = l == t2 ? ldv(A.myMail[5] + (j + (long)s0*k)) : ldg2(phi + ix(j,k,l+1,s0,s1)); f = l == 1 ? ldv(A.myMail[4] + (j + (long)s0*k)) : ldg2(phi + ix(j,k,l-1,s0,s1)); }
				const double coeff = 1./6.;
				double v = coeff*(a + b + c + d + e + f + ldg2(rho + ix(j,k,l,s0,s1)));
				phi[ix(j,k,l,s0,s1)] = v;
				if(edge){
					if(!(wrapMask&1)){ if(j == t0) A.peerMail[0][k + (long)s1*l] = v; if(j == 1) A.peerMail[1][k + (long)s1*l] = v; }
					if(!(wrapMask&2)){ if(k == t1) A.peerMail[2][j + (long)s0*l] = v; if(k == 1) A.peerMail[3][j + (long)s0*l] = v; }
					if(!(wrapMask&4)){ if(l == t2) A.peerMail[4][j + (long)s0*k] = v; if(l == 1) A.peerMail[5][j + (long)s0*k] = v; }
				}
			}
		}
		if(pr){ prof[26] += clock64() - tA; }                     // thread 0's own share of the sweep
	}
}
static int dimNb(const MpiInfo *m, int dd, int dir){
	int nb[3];
	for(int d = 0; d < 3; d++) nb[d] = m->subdomain[d];
	nb[dd] = (nb[dd] + dir + m->nSubdomains[dd]) % m->nSubdomains[dd];
	return nb[0] + m->nSubdomains[0]*(nb[1] + m->nSubdomains[1]*nb[2]);
}
static bool p2pArgs(Ctx *c, DevGrid *phi, const MpiInfo *m, P2PArgs &A){
	P2P *p = c->tp->p2p();
	if(!p) return false;
	for(int dd = 0; dd < 3; dd++){
		long face = phi->n / phi->size[dd];
		if(face > P2P_PLANE_CAP) return false;
		A.active[dd] = m->nSubdomains[dd] > 1;
		char *up = p->peerArena[dimNb(m, dd, +1)], *lo = p->peerArena[dimNb(m, dd, -1)];
		A.myMail[2*dd] = P2P::plane(p->arena, 2*dd); A.myMail[2*dd+1] = P2P::plane(p->arena, 2*dd+1);
		A.peerMail[2*dd] = P2P::plane(up, 2*dd);          // my upper layer is the upper neighbour's lower-side ghost
		A.peerMail[2*dd+1] = P2P::plane(lo, 2*dd+1);      // my lower layer is the lower neighbour's upper-side ghost
		A.peerFlag[2*dd] = P2P::flag(up, 2*dd); A.peerFlag[2*dd+1] = P2P::flag(lo, 2*dd+1);
	}
	A.myFlag = P2P::flag(p->arena, 0);
	A.ticket = P2P::flag(p->arena, 8);
	A.seqBase = P2P::flag(p->arena, 10);
	return true;
}

// gHaloOp(setSlice, TOHALO) of a scalar grid over peer memory: per decomposed dimension one kernel that stores the two
// boundary planes (rims included, so edges and corners propagate dimension by dimension as in src/grid.c:340-347)
// into the neighbours' mailbox planes and one that copies the planes received into the ghost layers
bool gridHaloP2P(Ctx *c, DevGrid *g, const MpiInfo *m){
	if(g->nv != 1 || m->mpiSize < 2) return false;
	if(mgModeResolved() != 2) return false;
	P2PArgs A{};
	if(!p2pArgs(c, g, m, A)) return false;
	P2P *p = c->tp->p2p();
	int s0 = g->size[0], s1 = g->size[1], s2 = g->size[2];
	for(int dd = 0; dd < 3; dd++){
		if(m->nSubdomains[dd] == 1){ gridHaloDim(c, g, m, dd+1, 0, 0); continue; }
		long np = g->n / g->size[dd];
		int blocks = gridFor(np, 256, c->numSMs);
		A.seq = ++p->seq - p->base;
		PINC_LAUNCH(c, K_HALO, 16.0*np, (k_p2p_publish<<<blocks,256,0,c->stream>>>(g->d, s0, s1, s2, A, 1 << dd, c->d_flags)));
		A.seq = ++p->seq - p->base;
		PINC_LAUNCH(c, K_HALO, 16.0*np, (k_p2p_copy<<<blocks,256,0,c->stream>>>(g->d, s0, s1, s2, dd, A, c->d_flags)));
	}
	return true;
}

extern int g_mgMode, g_mgForceCluster, g_mgNoCluster, g_mgReplica, g_mgRowMode, g_mgHybrid;
// 0 ops, 1 fused-exact, 2 auto, 3 auto-exact (resolved from $PINC_B200_MG at first use; pincMgSetMode overrides)
static int mgMode(){
	if(g_mgMode < 0){
		const char *e = getenv("PINC_B200_MG");
		g_mgMode = !e ? 2 : !strcmp(e, "ops") ? 0 : !strcmp(e, "fused") ? 1 : !strcmp(e, "cluster-exact") ? 3 : 2;
		if(e && !strcmp(e, "cluster-always")) g_mgForceCluster = 1;
		if(e && !strcmp(e, "allsm")) g_mgNoCluster = 1;
	}
	return g_mgMode;
}
static int mgModeResolved(){ return mgMode(); }
static void opGS(Ctx *c, DevGrid *phi, DevGrid *rho, int nCycles, const MpiInfo *m){
	long nt = trueCount(phi);
	if(m->mpiSize > 1 && mgMode() == 2 && nCycles > 0 && !phi->nonPeriodic){
		// multi-rank lean path: between half-sweeps only the faces of the decomposed dimensions are exchanged (one
		// grouped exchange), gBnd's mean subtraction is applied once at the end (see mgcluster.cu for why that is
		// the same function), then the full dimension-by-dimension halo restores edge and corner ghosts
		int wrapMask = (m->nSubdomains[0] == 1 ? 1 : 0) | (m->nSubdomains[1] == 1 ? 2 : 0) | (m->nSubdomains[2] == 1 ? 4 : 0);
		P2PArgs A{};
		if(p2pArgs(c, phi, m, A)){
			P2P *p = c->tp->p2p();
			int blocks = gridFor(nt/2, 256, c->numSMs);          // co-resident: the blocks wait for each other's tickets
			A.seq = p->seq + 1 - p->base;                         // publish = seq, half-sweep h = seq + h
			p->seq += 1 + 2*nCycles;
			int nHalf = 2*nCycles, s0 = phi->size[0], s1 = phi->size[1], s2 = phi->size[2];
			unsigned long long *gen = P2P::flag(p->arena, 9);
			int *dflags = c->d_flags;
			long long *prof = (long long*)mgProfBuffer(c);
			void *args[] = { &phi->d, &rho->d, &s0, &s1, &s2, &nHalf, &wrapMask, &A, &gen, &dflags, &prof };
			{
				LaunchScope ls(c, K_GS, 12.0*nt*nHalf);
				PINC_CUDA(cudaLaunchCooperativeKernel((void*)k_gs_p2p_loop, dim3(blocks), dim3(256), args, 0, c->stream));
			}
			gridHalo(c, phi, m, 0, 0);
			gridNeutralize(c, phi, m);
			return;
		}
		for(int h = 0; h < 2*nCycles; h++){
			PINC_LAUNCH(c, K_GS, 12.0*nt, (k_gs_colour_w<<<tGrid(c,nt),256,0,c->stream>>>(phi->d, rho->d, phi->size[0], phi->size[1], phi->size[2], (h&1) ? 0 : 1, wrapMask)));
			gridHaloFaces(c, phi, m);
		}
		gridHalo(c, phi, m, 0, 0);
		gridNeutralize(c, phi, m);
		return;
	}
	for(int cyc = 0; cyc < nCycles; cyc++)
		for(int parity = 1; parity >= 0; parity--){
			PINC_LAUNCH(c, K_GS, 12.0*nt, (k_gs_colour<<<tGrid(c,nt),256,0,c->stream>>>(phi->d, rho->d, phi->size[0], phi->size[1], phi->size[2], parity)));
			gridHalo(c, phi, m, 0, 0);
			gridBnd(c, phi, m);
		}
}
// "jacobian" smoother.  The reference's mgJacob3D (src/multigrid.c:500-551) is undefined behaviour as written: its loop
// never advances the write index (`tempVal[g]` with g fixed), its neighbour indices start at +-sizeProd[d] instead of
// g +- sizeProd[d] (the first read is phiVal[-1]), and it then copies the uninitialised scratch array over phi.  Provided
// here is the iteration its comments describe - every true node from the OLD values of its six neighbours, then gHaloOp and
// gBnd - so that `multigrid:preSmooth = jacobian` selects something usable; parity for this one function is against the
// oracle's restatement of the same intent only (tests/cycles_common.py).
__global__ void k_jacobi(const double *__restrict__ phi, const double *__restrict__ rho, double *__restrict__ out, int s0, int s1, int s2){
	int t0 = s0-2, t1 = s1-2, t2 = s2-2;
	long nt = (long)t0*t1*t2;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	const double coeff = 1./6;
	for(; i < nt; i += st){
		int j, k, l; truePoint(i, t0, t1, j, k, l);
		long g = ix(j,k,l,s0,s1); const long sx = s0, sxy = (long)s0*s1;
		out[g] = coeff*(phi[g+1] + phi[g-1] + phi[g+sx] + phi[g-sx] + phi[g+sxy] + phi[g-sxy] + rho[g]);
	}
}
__global__ void k_copy_true(double *__restrict__ dst, const double *__restrict__ src, int s0, int s1, int s2){
	int t0 = s0-2, t1 = s1-2, t2 = s2-2;
	long nt = (long)t0*t1*t2;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < nt; i += st){ int j, k, l; truePoint(i, t0, t1, j, k, l); long g = ix(j,k,l,s0,s1); dst[g] = src[g]; }
}
static void opJacobi(Ctx *c, DevGrid *phi, DevGrid *rho, int nCycles, const MpiInfo *m){
	long nt = trueCount(phi);
	double *tmp = (double*)tmpBuffer(c, (size_t)phi->n*sizeof(double));
	for(int cyc = 0; cyc < nCycles; cyc++){
		PINC_LAUNCH(c, K_GS, 24.0*nt, (k_jacobi<<<tGrid(c,nt),256,0,c->stream>>>(phi->d, rho->d, tmp, phi->size[0], phi->size[1], phi->size[2])));
		PINC_LAUNCH(c, K_GS, 16.0*nt, (k_copy_true<<<tGrid(c,nt),256,0,c->stream>>>(phi->d, tmp, phi->size[0], phi->size[1], phi->size[2])));
		gridHalo(c, phi, m, 0, 0);
		gridBnd(c, phi, m);
	}
}
} // namespace pinc
extern "C" void mgJacob3D(Grid *phi, const Grid *rho, const int nCycles, const MpiInfo *mpiInfo);
namespace pinc {
typedef void (*SmoothFn)(Grid*, const Grid*, const int, const MpiInfo*);
// the smoother a Multigrid names (mgSetSolver, src/multigrid.c:28-83): red-black Gauss-Seidel (default) or Jacobi
static void opSmooth(Ctx *c, SmoothFn fn, DevGrid *phi, DevGrid *rho, int nCycles, const MpiInfo *m){
	if(!fn || fn == (SmoothFn)mgGS3D) opGS(c, phi, rho, nCycles, m);
	else if(fn == (SmoothFn)mgJacob3D) opJacobi(c, phi, rho, nCycles, m);
	else fatal("multigrid: smoother is neither mgGS3D (gaussSeidelRB) nor mgJacob3D (jacobian)");
}
static void opResidual(Ctx *c, DevGrid *res, DevGrid *rho, DevGrid *phi){
	long nt = trueCount(phi);
	PINC_LAUNCH(c, K_RESIDUAL, 24.0*nt, (k_residual<<<tGrid(c,nt),256,0,c->stream>>>(res->d, rho->d, phi->d, phi->size[0], phi->size[1], phi->size[2])));
}
static void opRestrict(Ctx *c, DevGrid *fine, DevGrid *coarse){
	for(int d = 0; d < 3; d++) if(coarse->tsize[d]*2 != fine->tsize[d]) fatal("mgHalfRestrict3D: coarse grid is not half the fine grid");
	long nt = trueCount(coarse);
	PINC_LAUNCH(c, K_RESTRICT, 9.0*trueCount(fine), (k_restrict<<<tGrid(c,nt),256,0,c->stream>>>(fine->d, fine->size[0], fine->size[1], fine->size[2], coarse->d, coarse->size[0], coarse->size[1], coarse->size[2])));
}
static void opProlong(Ctx *c, DevGrid *fine, DevGrid *coarse, const MpiInfo *m){
	for(int d = 0; d < 3; d++) if(coarse->tsize[d]*2 != fine->tsize[d]) fatal("mgBilinProl3D: coarse grid is not half the fine grid");
	long nt = trueCount(fine);
	for(int pass = 0; pass < 4; pass++){
		if(pass > 0) gridHaloDim(c, fine, m, 4-pass, 0, 0);          // z, then y, then x (multigrid.c:1170,1194,1215)
		PINC_LAUNCH(c, K_PROLONG, 4.0*nt, (k_prolong_pass<<<tGrid(c,nt),256,0,c->stream>>>(fine->d, fine->size[0], fine->size[1], fine->size[2], coarse->d, coarse->size[0], coarse->size[1], pass)));
	}
}

// src/multigrid.c:1496-1548
static void opVCycle(Ctx *c, int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *m){
	DevGrid *phi = devGrid(c, mgPhi->grids[level]), *rho = devGrid(c, mgRho->grids[level]), *res = devGrid(c, mgRes->grids[level]);
	if(level == bottom){
		gridHalo(c, phi, m, 0, 0);
		gridHalo(c, rho, m, 0, 0);
		gridNeutralize(c, rho, m);
		opSmooth(c, (SmoothFn)mgRho->coarseSolv, phi, rho, mgRho->nCoarseSolve, m);
		gridBnd(c, phi, m);
		if(level > 0) opProlong(c, devGrid(c, mgRes->grids[level-1]), phi, m);
		return;
	}
	gridHalo(c, rho, m, 0, 0);
	gridNeutralize(c, rho, m);
	opSmooth(c, (SmoothFn)mgRho->preSmooth, phi, rho, mgRho->nPreSmooth, m);
	opResidual(c, res, rho, phi);
	gridHalo(c, res, m, 0, 0);
	opRestrict(c, res, devGrid(c, mgRho->grids[level+1]));
	opVCycle(c, level+1, bottom, top, mgRho, mgPhi, mgRes, m);
	gridAddTo(c, phi, res);
	gridHalo(c, phi, m, 0, 0);
	gridBnd(c, phi, m);
	opSmooth(c, (SmoothFn)mgRho->postSmooth, phi, rho, mgRho->nPostSmooth, m);
	gridBnd(c, phi, m);
	if(level > top) opProlong(c, devGrid(c, mgRes->grids[level-1]), phi, m);
}

// mgVRegular (src/multigrid.c:1559-1650), call for call - including its gSubFrom(phi, res) where mgVRecursive adds
static void opVRegular(Ctx *c, int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *m){
	for(int cur = level; cur < bottom; cur++){
		DevGrid *phi = devGrid(c, mgPhi->grids[cur]), *rho = devGrid(c, mgRho->grids[cur]), *res = devGrid(c, mgRes->grids[cur]);
		gridHalo(c, phi, m, 0, 0);
		gridBnd(c, phi, m);
		gridNeutralize(c, rho, m);
		opSmooth(c, (SmoothFn)mgRho->preSmooth, phi, rho, mgRho->nPreSmooth, m);
		gridHalo(c, rho, m, 0, 0);
		gridBnd(c, phi, m);
		gridZero(c, res);
		opResidual(c, res, rho, phi);
		gridHalo(c, res, m, 0, 0);
		opRestrict(c, res, devGrid(c, mgRho->grids[cur+1]));
	}
	{
		DevGrid *phi = devGrid(c, mgPhi->grids[bottom]), *rho = devGrid(c, mgRho->grids[bottom]);
		gridNeutralize(c, rho, m);
		gridHalo(c, rho, m, 0, 0);
		opSmooth(c, (SmoothFn)mgRho->coarseSolv, phi, rho, mgRho->nCoarseSolve, m);
		gridHalo(c, phi, m, 0, 0);
		gridBnd(c, phi, m);
		opProlong(c, devGrid(c, mgRes->grids[bottom-1]), phi, m);
	}
	for(int cur = bottom-1; cur >= top; cur--){
		DevGrid *phi = devGrid(c, mgPhi->grids[cur]), *rho = devGrid(c, mgRho->grids[cur]), *res = devGrid(c, mgRes->grids[cur]);
		gridSubFrom(c, phi, res);
		gridHalo(c, phi, m, 0, 0);
		gridBnd(c, phi, m);
		opSmooth(c, (SmoothFn)mgRho->postSmooth, phi, rho, mgRho->nPostSmooth, m);
		gridBnd(c, phi, m);
		if(cur > top) opProlong(c, devGrid(c, mgRes->grids[cur-1]), phi, m);
	}
}
// one pass of the cycle a solver names (getMgAlgo, src/multigrid.c:113-125)
} // namespace pinc
extern "C" {
void mgVRegular(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo);
void mgFMG(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo);
void mgW(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo);
void mgVRecursive(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo);
}
namespace pinc {
static void opCycle(Ctx *c, funPtr algo, int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *m){
	if(!algo || algo == (funPtr)mgVRecursive) opVCycle(c, level, bottom, top, mgRho, mgPhi, mgRes, m);
	else if(algo == (funPtr)mgVRegular) opVRegular(c, level, bottom, top, mgRho, mgPhi, mgRes, m);
	else if(algo == (funPtr)mgW){                        // src/multigrid.c:1675-1683
		int middle = bottom/2;
		opVCycle(c, 0, bottom, middle, mgRho, mgPhi, mgRes, m);
		opVCycle(c, middle, bottom, 0, mgRho, mgPhi, mgRes, m);
	} else if(algo == (funPtr)mgFMG)
		fatal("multigrid: cycle = mgFMG is not provided: the reference's mgFMG (src/multigrid.c:1652-1672) overwrites mgRho->grids[0] with the coarsest grid, after which mgSolveRaw reads out of bounds");
	else fatal("multigrid: unknown cycle function");
}

// =================================================================================================
// fused mode: the whole tolerance loop in one persistent cooperative kernel
// =================================================================================================
#define MG_MAXLEV 10
#define MG_BLOCK 512
#define MG_SMALL MC_SMALL    // levels with at most this many true nodes run inside CTA 0 (shared memory)
// Block-resident smoothing of a grid-wide level: CTA 1+b keeps block b (bx x by x bz true nodes + one halo layer) of phi
// in its shared memory for a whole mgGS3D call; the faces travel through per-CTA mailboxes in L2 (see bGS).
struct BLvl {
	int on, bx, by, bz, nbx, nby, nbz, nb, rows;
	int offPhi, offRho;      // shared-memory offsets in doubles; offRho < 0: rho is read from global memory
	uint4 *mail;             // nb x 2(by*bz + bx*bz + bx*by) slots
	double *rhoS;            // row smoother: colour-separated copy of rho, nb x bx*by*bz doubles (mgrows.cuh)
};
// Hybrid multi-rank solve (hybridSolve below): levels q < qDist are DISTRIBUTED (every rank smooths its own sub-domain
// block-resident, faces across a sub-domain boundary travel through tagged 16-byte slots in the NEIGHBOUR'S arena over
// NVLink), levels q >= qDist are REPLICATED (every rank holds and solves the global level).  All cross-GPU traffic is
// "data is its own flag" (st.relaxed.sys / ld.relaxed.sys of {lo, tag, hi, tag}): no fence, no flag word.
#define XD_MAXR 32
struct XDist {
	int on, qDist, R, me;
	int ns[3], sub[3];             // sub-domains per dimension, this rank's sub-domain
	int nbr[6];                    // rank across face f = 2*dim + side (side 0: lower)
	uint4 *peer[XD_MAXR];          // every rank's arena (peer[me] is this rank's own)
	unsigned *ctr;                 // persisted counters in my arena: [0] mailbox tags, [1] sums, [2] halo fills, [3] gathers
	unsigned offSum, offPlane, planeCap, offGath, gathCap;      // slot (16-byte) offsets into an arena
};
struct MgPlan {
	Lvl L[MG_MAXLEV];
	BLvl B[MG_MAXLEV];
	XDist X;
	unsigned *seqWord;       // running half-sweep number of the mailbox protocol (persists across launches)
	uint4 *mailAll; unsigned long long mailSlots;       // all mailbox slots of this context (cleared when the 32-bit tags are about to wrap)
	int nLevels, nPre, nPost, nCoarse, qSmall, maxCycles;
	double tol, totTrue;
	double *partial;         // 2*gridDim doubles
	unsigned *bar;
	double *hist;            // [0] cycles, [1..] barRes per V-cycle
	unsigned barBase;
	int exact;               // gBnd after every half-sweep (pending shifts) instead of once per smoother call
	int smemSmall;           // small levels live in CTA 0's shared memory (descriptors in C)
	int offZ;                // exchange buffer of the 16^3 smoother in CTA 0's shared memory (doubles), or < 0
	int pyramid;             // small levels are cubic 16/8/4: specialised routines (mgsmall.cuh)
	long long *prof;         // optional cycle accounting ($PINC_B200_MGPROF)
	CPlan C;
};

// ---- cross-GPU slots: one 16-byte store {value.lo, tag, value.hi, tag} at system scope; the reader polls until both tags match
__device__ __forceinline__ void llStoreSys(uint4 *p, double v, unsigned tag){
	unsigned lo = (unsigned)__double_as_longlong(v), hi = (unsigned)(__double_as_longlong(v) >> 32);
	asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __noinline__ void llTimeout(const uint4 *p, unsigned tag, unsigned a, unsigned b){
	printf("PINC-B200 ERROR: rank-to-rank slot %p never reached tag %u (holds %u/%u) - a neighbour rank did not arrive\n", (const void*)p, tag, a, b);
	__trap();
}
__device__ __forceinline__ double llWaitSys(const uint4 *p, unsigned tag){
	unsigned a, b, c, d, spins = 0;
	unsigned long long t0 = 0;
	for(;;){
		asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
		if(b == tag && d == tag) break;
		if((++spins & 0xfffu) == 0){                  // bounded: ~20 s of wall clock, then fail loudly instead of hanging the device
			unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
			if(!t0) t0 = t; else if(t - t0 > 20000000000ULL) llTimeout(p, tag, b, d);
		}
	}
	return __longlong_as_double(((long long)c << 32) | (long long)a);
}

struct Scope {
	const XDist *X;          // hybrid multi-rank solve: cross-GPU description (MgPlan::X), else unused
	unsigned xMail, xSum, xHalo, xGath;     // running tags of the cross-GPU operations (identical on all ranks)
	bool single;             // only this CTA takes part (block-level barriers)
	unsigned *bar; unsigned gen;
	double *partial; int flip;
	double *sh;              // 18 doubles of shared memory
	CK *K;                   // cycle accounting
	__device__ __forceinline__ long tid() const { return single ? threadIdx.x : blockIdx.x*(long)blockDim.x + threadIdx.x; }
	__device__ __forceinline__ long nthr() const { return single ? blockDim.x : (long)gridDim.x*blockDim.x; }
	// Grid barrier: one monotonically increasing arrival counter (no reset, no generation word).  The fence
	// before the arrival orders this CTA's stores; readers fetch other CTAs' data with ld.global.cg (L2), so no
	// L1 invalidation is needed afterwards.  Measured (tools/ubench_gridbar.cu): 2780 cycles against 4820 for the
	// fence + counter + generation + fence scheme.
	__device__ __noinline__ void sync(){
		__syncthreads();
		if(single) return;
		if(threadIdx.x == 0){
			__threadfence();
			atomicAdd(&bar[0], 1u);
			gen += gridDim.x;
			while((int)(*((volatile unsigned*)&bar[0]) - gen) < 0){ }
			unsigned seen;
			asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&bar[0]) : "memory");
			(void)seen;
		}
		__syncthreads();
	}
	// sum over all participating threads; the same bits in every thread; acts as a barrier
	__device__ __noinline__ double allSum(double v){
		int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
		for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		if(lane == 0) sh[w] = v;
		__syncthreads();
		if(w == 0){
			double t = lane < nw ? sh[lane] : 0.0;
			for(int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
			if(lane == 0) sh[17] = t;
		}
		__syncthreads();
		double tot = sh[17];
		if(single){ __syncthreads(); return tot; }
		if(threadIdx.x == 0) __stcg(&partial[flip*gridDim.x + blockIdx.x], tot);
		sync();
		if(w == 0){
			double a = 0;
			for(int i = lane; i < (int)gridDim.x; i += 32) a += __ldcg(&partial[flip*gridDim.x + i]);
			for(int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
			if(lane == 0) sh[17] = a;
		}
		__syncthreads();
		tot = sh[17];
		__syncthreads();
		flip ^= 1;
		return tot;
	}
	// the same over all ranks: rank r's CTA 0 stores its total into slot r of every rank, every CTA polls the R slots of
	// its own arena and adds them in rank order (identical bits everywhere); two slot sets alternate, a set is rewritten
	// only after every rank has sent its next value, i.e. after it has read this one
	__device__ __noinline__ double allSumX(double v){
		double tot = allSum(v);
		const XDist &x = *X;
		const unsigned tag = ++xSum, set = tag & 1u;
		if(blockIdx.x == 0 && (int)threadIdx.x < x.R) llStoreSys(x.peer[threadIdx.x] + x.offSum + set*XD_MAXR + x.me, tot, tag);
		if(threadIdx.x < 32){
			const int lane = threadIdx.x;
			double t = lane < x.R ? llWaitSys(x.peer[x.me] + x.offSum + set*XD_MAXR + lane, tag) : 0.0;
			double a = 0;
			for(int r = 0; r < x.R; r++) a += __shfl_sync(0xffffffffu, t, r);
			if(lane == 0) sh[17] = a;
		}
		__syncthreads();
		tot = sh[17];
		__syncthreads();
		return tot;
	}
};

// gNeutralizeGrid on the true nodes: returns after the subtraction is visible to everyone
template<bool X = false> __device__ __noinline__ void fNeutralize(double *v, int s0, int s1, int s2, Scope &S){
	ProfScope psn(*S.K, S.single ? 27 : 28);
	int t0 = s0-2, t1 = s1-2, t2 = s2-2; long nt = (long)t0*t1*t2;
	double acc = 0;
	for(long i = S.tid(); i < nt; i += S.nthr()){ int j,k,l; truePoint(i,t0,t1,j,k,l); acc += ldg2(v + ix(j,k,l,s0,s1)); }
	double avg = X ? S.allSumX(acc)/((double)nt*(double)S.X->R) : S.allSum(acc)/(double)nt;
	for(long i = S.tid(); i < nt; i += S.nthr()){ int j,k,l; truePoint(i,t0,t1,j,k,l); long g = ix(j,k,l,s0,s1); v[g] = ldg2(v + g) - avg; }
	S.sync();
}

// mgGS3D with gBnd's mean subtraction carried as pending shifts.  sIn: shift still pending on every value at entry.
__device__ __noinline__ void fGS(const Lvl &L, int nCycles, double sIn, int exact, Scope &S){
	ProfScope ps(*S.K, S.single ? PS_GS_SMALL : PS_GS_BIG);
	int s0 = L.s0, s1 = L.s1, s2 = L.s2;
	int t0 = s0-2, t1 = s1-2, t2 = s2-2; long nt = (long)t0*t1*t2;
	int half = t0/2; long items = (long)half*t1*t2;
	if(nCycles <= 0 || !exact){
		if(sIn != 0.0){
			for(long i = S.tid(); i < nt; i += S.nthr()){ int j,k,l; truePoint(i,t0,t1,j,k,l); long g = ix(j,k,l,s0,s1); L.phi[g] = ldg2(L.phi + g) - sIn; }
			S.sync();
		}
		if(nCycles <= 0) return;
		sIn = 0.0;
	}
	double sR = sIn, sPrev = 0;
	for(int h = 0; h < 2*nCycles; h++){
		int parity = (h & 1) ? 0 : 1;
		double acc = 0;
		// two nodes per trip: both nodes' loads are issued before either store (the compiler cannot prove that an
		// own-colour store does not alias the next node's other-colour loads), so their L2 latencies overlap
		const long step = S.nthr();
		for(long i = S.tid(); i < items; i += 2*step){
			double vn[2], vo[2]; long go[2]; bool ok[2];
			#pragma unroll
			for(int w = 0; w < 2; w++){
				long iw = i + w*step;
				ok[w] = iw < items;
				if(!ok[w]) continue;
				unsigned ui = (unsigned)iw, r = ui / (unsigned)half; int m = (int)(ui - r*(unsigned)half);
				unsigned lq = r / (unsigned)t1; int k = (int)(r - lq*(unsigned)t1) + 1; int l = (int)lq + 1;
				int ja = 2*m+1;
				int jOwn = (((ja+k+l)&1) == parity) ? ja : ja+1;
				int jOth = 2*ja+1 - jOwn;
				vn[w] = gsPoint<true>(L.phi, L.rho, jOwn, k, l, s0, s1, s2, sR);
				if(exact) vo[w] = ldg2(L.phi + ix(jOth,k,l,s0,s1)) - sR;
				go[w] = ix(jOwn,k,l,s0,s1);
			}
			#pragma unroll
			for(int w = 0; w < 2; w++){
				if(!ok[w]) continue;
				if(exact){ acc += vn[w]; acc += vo[w]; }
				L.phi[go[w]] = vn[w];
			}
		}
		if(exact){
			double avg = S.allSum(acc)/(double)nt;
			sPrev = sR; sR = avg;
		} else S.sync();
	}
	if(!exact){ fNeutralize(L.phi, s0, s1, s2, S); return; }     // the 2*nCycles gBnd calls, applied once
	// materialise: the colour written last (even nodes) carries sR, the other one sPrev then sR
	for(long i = S.tid(); i < nt; i += S.nthr()){
		int j,k,l; truePoint(i,t0,t1,j,k,l); long g = ix(j,k,l,s0,s1);
		double v = ldg2(L.phi + g);
		if((j+k+l)&1) v -= sPrev;
		v -= sR;
		L.phi[g] = v;
	}
	S.sync();
}


// division by a run-time divisor that is used many times: q = umulhi(n, ceil(2^32/d)), exact while n*d < 2^32
struct FD {
	unsigned m, d;
	__device__ __forceinline__ explicit FD(int dd) : m(dd > 1 ? 0xFFFFFFFFu/(unsigned)dd + 1u : 0u), d((unsigned)dd) {}
	__device__ __forceinline__ int div(int n) const { return d > 1 ? (int)__umulhi((unsigned)n, m) : n; }
	__device__ __forceinline__ void divmod(int n, int &q, int &r) const { q = div(n); r = n - q*(int)d; }
};

// ---- block-resident mgGS3D ---------------------------------------------------------------------------------------
// One mailbox slot = {value.lo, tag, value.hi, tag}: a 16-byte store whose two halves each carry the tag, so a reader
// that sees both tags has the whole value (8-byte single-copy atomicity is all this needs).  No fence, no barrier: the
// data is its own flag.  .cg accesses go to L2, the coherence point (measured, tools/ubench_gridbar.cu: 1500 cycles per
// exchange of 512 slots per CTA against 2780 for a grid barrier + 300 for the dependent L2 loads behind it; volatile
// accesses are 800 cycles slower).
__device__ __forceinline__ void llStore(uint4 *p, double v, unsigned tag){
	unsigned lo = (unsigned)__double_as_longlong(v), hi = (unsigned)(__double_as_longlong(v) >> 32);
	asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ bool llTry(const uint4 *p, unsigned tag, double &v){
	unsigned a, b, c, d;
	asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
	v = __longlong_as_double(((long long)c << 32) | (long long)a);
	return b == tag && d == tag;
}
__device__ __forceinline__ double llWait(const uint4 *p, unsigned tag){
	unsigned a, b, c, d, spins = 0;
	for(;;){
		asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
		if(b == tag && d == tag) break;
		if(++spins > (1u << 22)) __trap();             // ~2 s: a lost neighbour must not hang the device
	}
	return __longlong_as_double(((long long)c << 32) | (long long)a);
}
// ---- the inner loop of the block smoother for small blocks: everything a thread needs is in registers ------------
#define LL_NONE 0xffffffffu
struct FastNode { int idx; unsigned t0, t1, t2; double rho; };      // idx 0: no node; t*: mailbox slots (from B.mail) it is sent to
// EX, PL: the block's row and plane strides as compile-time constants (0: the run-time values ex, pl).  With constants the six
// neighbours are immediate offsets from one address register per node; with run-time strides they are four more registers per
// node, which ptxas spilled (20 local-memory loads per half-sweep, several of them on the chain poll -> barrier -> update -> send)
template<int EX, int PL> __device__ __forceinline__ double fastVal(const double *Ph, int idx, int ex, int pl, double rho){
	const double coeff = 1./6.;
	const int e_ = EX ? EX : ex, p_ = PL ? PL : pl;
	double a = Ph[idx+1], b = Ph[idx-1], c = Ph[idx+e_], d = Ph[idx-e_], e = Ph[idx+p_], f = Ph[idx-p_];
	return coeff*(a + b + c + d + e + f + rho);
}
// X (hybrid multi-rank solve, distributed level): a slot offset carries the face it crosses in its top three bits, the face's
// base is this rank's mailbox array or the neighbour rank's (xBase, shared memory), and stores/polls are system-scope
#define LL_FACE_SHIFT 29
#define LL_OFF_MASK 0x1fffffffu
template<bool X> __device__ __forceinline__ void fastSend(uint4 *mail, uint4 *const *xBase, const FastNode &n, double v, unsigned stag){
	if constexpr(X){
		if(n.t0 != LL_NONE) llStoreSys(xBase[n.t0 >> LL_FACE_SHIFT] + (n.t0 & LL_OFF_MASK), v, stag);
		if(n.t1 != LL_NONE) llStoreSys(xBase[n.t1 >> LL_FACE_SHIFT] + (n.t1 & LL_OFF_MASK), v, stag);
		if(n.t2 != LL_NONE) llStoreSys(xBase[n.t2 >> LL_FACE_SHIFT] + (n.t2 & LL_OFF_MASK), v, stag);
	} else {
		if(n.t0 != LL_NONE) llStore(mail + n.t0, v, stag);
		if(n.t1 != LL_NONE) llStore(mail + n.t1, v, stag);
		if(n.t2 != LL_NONE) llStore(mail + n.t2, v, stag);
	}
}
template<bool X, int EX, int PL> __device__ __forceinline__ void fastHalf(double *Ph, int ex, int pl, uint4 *mail, uint4 *const *xBase, const uint4 *pAddr, int pIdx, bool recv, unsigned tagIn,
		const FastNode &a, const FastNode &b, unsigned stag, bool send){
	if(recv && pAddr) Ph[pIdx] = X ? llWaitSys(pAddr, tagIn) : llWait(pAddr, tagIn);
	__syncthreads();
	double va = 0, vb = 0;
	if(a.idx) va = fastVal<EX,PL>(Ph, a.idx, ex, pl, a.rho);
	if(b.idx) vb = fastVal<EX,PL>(Ph, b.idx, ex, pl, b.rho);
	if(a.idx){ Ph[a.idx] = va; if(send) fastSend<X>(mail, xBase, a, va, stag); }
	if(b.idx){ Ph[b.idx] = vb; if(send) fastSend<X>(mail, xBase, b, vb, stag); }
}
// c0*: nodes of colour 0 and the halo node of colour 0 this thread receives; same for colour 1.  Half-sweep h updates
// colour 1 (h even) or 0 (h odd) and first receives the other colour's face nodes of half-sweep h-1.
// X: the halo loaded from this rank's memory is not the neighbour rank's data, so the call starts with an exchange of the
// colour-0 boundary nodes (tag seq+1) and the half-sweeps use tags seq+2 .. (the caller advances seq by 2*nCycles + 2).
template<bool X, int EX, int PL> __device__ __noinline__ void bSmoothFast(double *Ph, int ex, int pl, uint4 *mail, uint4 *const *xBase, const uint4 *p0Addr, int p0Idx, const uint4 *p1Addr, int p1Idx,
		FastNode a0, FastNode b0, FastNode a1, FastNode b1, int nCycles, unsigned seq, long long *pf, long long tEnter){
	long long tW = 0;
	if(pf){ pf[2*9] += clock64() - tEnter; pf[2*9+1] += 1; tW = clock64(); }
	if constexpr(X){
		seq += 1u;
		if(a0.idx) fastSend<X>(mail, xBase, a0, Ph[a0.idx], seq);
		if(b0.idx) fastSend<X>(mail, xBase, b0, Ph[b0.idx], seq);
	}
	for(int h2 = 0; h2 < nCycles; h2++){
		const unsigned t = seq + 2u*(unsigned)h2;
		fastHalf<X,EX,PL>(Ph, ex, pl, mail, xBase, p0Addr, p0Idx, X || h2 > 0, t, a1, b1, t + 1u, true);
		fastHalf<X,EX,PL>(Ph, ex, pl, mail, xBase, p1Addr, p1Idx, true, t + 1u, a0, b0, t + 2u, h2 + 1 < nCycles);
	}
	if(pf){ pf[2*10] += clock64() - tW; pf[2*10+1] += 2*nCycles; }
}
// mgGS3D (src/multigrid.c:683-767) followed by the batched gBnd, on a level that is split into blocks.  The update of a
// node reads the six neighbours of the other colour, so a half-sweep needs from the neighbouring blocks exactly the
// face nodes they updated in the previous half-sweep: each CTA sends the boundary nodes it updates straight into the
// mailboxes of the (up to) three neighbours that need them, tagged with the running half-sweep number, and before the
// next half-sweep copies the tagged values it was sent into its halo layer.  A slot is rewritten two half-sweeps later,
// which needs the value its reader produces in between: the protocol is its own back-pressure.  Same arithmetic per
// node as fGS, hence the same bits.
// pendOut (single-rank levels): the batched gBnd is NOT applied to the values written back; the mean goes to *pendOut and every
// later reader of this level's phi forms (stored - mean) itself (resPoint, prolPoint, the prolongation's add, the next call's sIn,
// the pass that ends the solve).  The same bits as subtracting here - and the barrier inside the sum publishes the write-back, so
// the smoother call ends with one grid barrier instead of two.
template<bool X> __device__ __noinline__ void bGS(const Lvl &L, const BLvl &B, int nCycles, double sIn, Scope &S, unsigned &seq, double *pendOut = nullptr){
	ProfScope ps(*S.K, PS_GS_BIG);
	__shared__ uint4 *xBase[6];
	const int t0 = L.s0-2, t1 = L.s1-2, t2 = L.s2-2;
	const int bid = (int)blockIdx.x - 1;
	const bool act = bid >= 0 && bid < B.nb;
	const int bx = B.bx, by = B.by, bz = B.bz, ex = bx+2, ey = by+2, pl = ex*ey;
	const int nA = by*bz, nB = bx*bz, nC = bx*by;
	const FD dEx(ex), dEy(ey), dBx(bx), dBy(by), dHx(bx/2), dHy(by/2);
	double *Ph = mgS + B.offPhi;
	const double *Rh = mgS + B.offRho;
	double bsum = 0;
	const long long tEnter = clock64();
	int ox = 0, oy = 0, oz = 0;
	if(act){
		const int cx = bid % B.nbx, cr = bid / B.nbx, cy = cr % B.nby, cz = cr / B.nby;
		ox = cx*bx; oy = cy*by; oz = cz*bz;
		const int slots = 2*(nA + nB + nC);
		const int fb[6] = {0, nA, 2*nA, 2*nA + nB, 2*nA + 2*nB, 2*nA + 2*nB + nC};
		const uint4 *mine = B.mail + (size_t)bid*slots;
		// where my boundary nodes go: the neighbour across face f receives them on its face f^1
		unsigned out[6];                               // slot offsets from B.mail
		{
			int xm = cx ? cx-1 : B.nbx-1, xp = cx+1 < B.nbx ? cx+1 : 0, ym = cy ? cy-1 : B.nby-1, yp = cy+1 < B.nby ? cy+1 : 0;
			int zm = cz ? cz-1 : B.nbz-1, zp = cz+1 < B.nbz ? cz+1 : 0;
			out[0] = (unsigned)(xm + B.nbx*(cy + B.nby*cz))*slots + fb[1];
			out[1] = (unsigned)(xp + B.nbx*(cy + B.nby*cz))*slots + fb[0];
			out[2] = (unsigned)(cx + B.nbx*(ym + B.nby*cz))*slots + fb[3];
			out[3] = (unsigned)(cx + B.nbx*(yp + B.nby*cz))*slots + fb[2];
			out[4] = (unsigned)(cx + B.nbx*(cy + B.nby*zm))*slots + fb[5];
			out[5] = (unsigned)(cx + B.nbx*(cy + B.nby*zp))*slots + fb[4];
			if constexpr(X){
				// a face on the sub-domain boundary of a decomposed dimension goes to the same slot of the neighbour RANK's array
				if(threadIdx.x < 6){
					const int f = threadIdx.x, d = f >> 1, side = f & 1;
					const int cc = d == 0 ? cx : (d == 1 ? cy : cz), nn = d == 0 ? B.nbx : (d == 1 ? B.nby : B.nbz);
					const bool edge = side ? cc == nn-1 : cc == 0;
					const XDist &x = *S.X;
					const size_t rel = (size_t)(B.mail - x.peer[x.me]);
					xBase[f] = (edge && x.ns[d] > 1) ? x.peer[x.nbr[f]] + rel : B.mail;
				}
			}
		}
		if constexpr(X) __syncthreads();
		// block + halo layer from global memory (periodic image), with the pending mean shift applied
		const int ne = pl*(bz+2);
		auto srcOf = [&](int i) -> const double* {
			int jl, r, kl, ll; dEx.divmod(i, r, jl); dEy.divmod(r, ll, kl);
			int gj = ox + jl, gk = oy + kl, gl = oz + ll;
			gj = gj == 0 ? t0 : (gj == t0+1 ? 1 : gj);
			gk = gk == 0 ? t1 : (gk == t1+1 ? 1 : gk);
			gl = gl == 0 ? t2 : (gl == t2+1 ? 1 : gl);
			return L.phi + ix(gj,gk,gl,L.s0,L.s1);
		};
		if(ne <= 8*(int)blockDim.x){
			// all loads in flight before the first use: one L2 latency instead of eight
			double v[8];
			#pragma unroll
			for(int u = 0; u < 8; u++){ int i = threadIdx.x + u*(int)blockDim.x; v[u] = i < ne ? ldg2(srcOf(i)) : 0.0; }
			#pragma unroll
			for(int u = 0; u < 8; u++){ int i = threadIdx.x + u*(int)blockDim.x; if(i < ne) Ph[i] = sIn != 0.0 ? v[u] - sIn : v[u]; }
		} else
			for(int i = threadIdx.x; i < ne; i += blockDim.x){
				double v = ldg2(srcOf(i));
				if(sIn != 0.0) v -= sIn;
				Ph[i] = v;
			}
		const int hx = bx/2, items = hx*by*bz;
		const int nHalo = nA + nB + nC;                 // halo nodes of one colour over the six faces
		const bool fast = B.on == 1 && items <= 2*(int)blockDim.x && nHalo <= (int)blockDim.x;
		if(B.offRho >= 0 && !fast)
			for(int i = threadIdx.x; i < bx*by*bz; i += blockDim.x){
				int jl, r, kl, ll; dBx.divmod(i, r, jl); dBy.divmod(r, ll, kl);
				mgS[B.offRho + i] = ldg2(L.rho + ix(ox+jl+1, oy+kl+1, oz+ll+1, L.s0, L.s1));
			}
		const double coeff = 1./6.;
		// halo node i (0 <= i < nHalo) of colour c: mailbox slot and index into the block array
		auto haloNode = [&](int i, int c, int &slot, int &hidx){
			int f, w, j, k, l;
			// (a face's slots are colour-separated: colour c's nodes of a row sit next to each other, so that the 16-byte stores
			// and polls of consecutive threads are consecutive in memory - half the L2 sectors of the interleaved layout)
			if(i < nA){
				f = i >= nA/2; w = i - f*(nA/2);
				int uu; dHy.divmod(w, l, uu); l += 1; j = f ? bx+1 : 0;
				k = 2*uu + 1; k += ((j + k + l) & 1) != c;
				slot = fb[f] + c*(nA/2) + uu + (by/2)*(l-1);
			} else if(i < nA + nB){
				w = i - nA; f = w >= nB/2; w -= f*(nB/2);
				int uu; dHx.divmod(w, l, uu); l += 1; k = f ? by+1 : 0;
				j = 2*uu + 1; j += ((j + k + l) & 1) != c;
				slot = fb[2+f] + c*(nB/2) + uu + (bx/2)*(l-1);
			} else {
				w = i - nA - nB; f = w >= nC/2; w -= f*(nC/2);
				int uu; dHx.divmod(w, k, uu); k += 1; l = f ? bz+1 : 0;
				j = 2*uu + 1; j += ((j + k + l) & 1) != c;
				slot = fb[4+f] + c*(nC/2) + uu + (bx/2)*(k-1);
			}
			hidx = j + ex*(k + ey*l);
		};
		// slot of a boundary node inside the face it is sent across (see haloNode): colour half, then (u/2, v)
		auto slotA = [&](int c, int k, int l){ return (unsigned)(c*(nA/2) + ((k-1) >> 1) + (by/2)*(l-1)); };
		auto slotB = [&](int c, int j, int l){ return (unsigned)(c*(nB/2) + ((j-1) >> 1) + (bx/2)*(l-1)); };
		auto slotC = [&](int c, int j, int k){ return (unsigned)(c*(nC/2) + ((j-1) >> 1) + (bx/2)*(k-1)); };
		auto sendNode = [&](int j, int k, int l, double v, unsigned stag){
			const int c = (j + k + l) & 1;
			if constexpr(X){
				if(j == 1)  llStoreSys(xBase[0] + out[0] + slotA(c,k,l), v, stag);
				if(j == bx) llStoreSys(xBase[1] + out[1] + slotA(c,k,l), v, stag);
				if(k == 1)  llStoreSys(xBase[2] + out[2] + slotB(c,j,l), v, stag);
				if(k == by) llStoreSys(xBase[3] + out[3] + slotB(c,j,l), v, stag);
				if(l == 1)  llStoreSys(xBase[4] + out[4] + slotC(c,j,k), v, stag);
				if(l == bz) llStoreSys(xBase[5] + out[5] + slotC(c,j,k), v, stag);
			} else {
				if(j == 1)  llStore(B.mail + out[0] + slotA(c,k,l), v, stag);
				if(j == bx) llStore(B.mail + out[1] + slotA(c,k,l), v, stag);
				if(k == 1)  llStore(B.mail + out[2] + slotB(c,j,l), v, stag);
				if(k == by) llStore(B.mail + out[3] + slotB(c,j,l), v, stag);
				if(l == 1)  llStore(B.mail + out[4] + slotC(c,j,k), v, stag);
				if(l == bz) llStore(B.mail + out[5] + slotC(c,j,k), v, stag);
			}
		};
		if(fast){
			// fast path (at most two nodes per thread and colour, one halo node per thread and colour): node addresses,
			// rho and mailbox slots are worked out once per call and handed to a lean routine, so that a half-sweep is
			// wait -> barrier -> 12 shared loads -> 14 additions -> stores
			FastNode nd[2][2]; const uint4 *pAddr[2]; int pIdx[2];
			#pragma unroll
			for(int c = 0; c < 2; c++){
				pAddr[c] = nullptr; pIdx[c] = 0;
				if((int)threadIdx.x < nHalo){ int slot; haloNode(threadIdx.x, c, slot, pIdx[c]); pAddr[c] = mine + slot; }
				#pragma unroll
				for(int w = 0; w < 2; w++){
					int iw = threadIdx.x + w*(int)blockDim.x;
					FastNode &n = nd[c][w];
					n.idx = 0; n.t0 = n.t1 = n.t2 = LL_NONE; n.rho = 0;
					if(iw < items){
						int m, r, k, l; dHx.divmod(iw, r, m); dBy.divmod(r, l, k); k += 1; l += 1;
						int j = ((((1+k+l)&1) == c) ? 1 : 2) + 2*m;
						n.idx = j + ex*(k + ey*l);
						n.rho = ldg2(L.rho + ix(ox+j, oy+k, oz+l, L.s0, L.s1));
						// a node lies on at most one face per dimension unless the block is two nodes wide, where the second
						// face of that dimension takes a slot of its own
						unsigned t[6]; int nt = 0;
						constexpr unsigned FS = X ? (1u << LL_FACE_SHIFT) : 0u;       // X: the face rides in the top bits
						if(j == 1)  t[nt++] = out[0] + slotA(c,k,l);
						if(j == bx) t[nt++] = out[1] + slotA(c,k,l) + 1u*FS;
						if(k == 1)  t[nt++] = out[2] + slotB(c,j,l) + 2u*FS;
						if(k == by) t[nt++] = out[3] + slotB(c,j,l) + 3u*FS;
						if(l == 1)  t[nt++] = out[4] + slotC(c,j,k) + 4u*FS;
						if(l == bz) t[nt++] = out[5] + slotC(c,j,k) + 5u*FS;
						if(nt > 0) n.t0 = t[0];
						if(nt > 1) n.t1 = t[1];
						if(nt > 2) n.t2 = t[2];
					}
				}
			}
			long long *pf = (S.K->prof && bid == 0 && threadIdx.x == 0) ? S.K->prof : nullptr;
			// the two block shapes of the benchmark pyramid (64^3 -> 16x16x8, 32^3 -> 8x8x4) with their strides as constants
			if(ex == 18 && pl == 324) bSmoothFast<X,18,324>(Ph, ex, pl, B.mail, xBase, pAddr[0], pIdx[0], pAddr[1], pIdx[1], nd[0][0], nd[0][1], nd[1][0], nd[1][1], nCycles, seq, pf, tEnter);
			else if(ex == 10 && pl == 100) bSmoothFast<X,10,100>(Ph, ex, pl, B.mail, xBase, pAddr[0], pIdx[0], pAddr[1], pIdx[1], nd[0][0], nd[0][1], nd[1][0], nd[1][1], nCycles, seq, pf, tEnter);
			else bSmoothFast<X,0,0>(Ph, ex, pl, B.mail, xBase, pAddr[0], pIdx[0], pAddr[1], pIdx[1], nd[0][0], nd[0][1], nd[1][0], nd[1][1], nCycles, seq, pf, tEnter);
		} else {
		const unsigned seqG = X ? seq + 1u : seq;
		if constexpr(X){
			// the call starts with an exchange of the colour-0 boundary nodes (see bSmoothFast)
			__syncthreads();
			for(int i = threadIdx.x; i < items; i += blockDim.x){
				int m, r, k, l; dHx.divmod(i, r, m); dBy.divmod(r, l, k); k += 1; l += 1;
				int j = ((((1+k+l)&1) == 0) ? 1 : 2) + 2*m;
				sendNode(j, k, l, Ph[j + ex*(k + ey*l)], seqG);
			}
		}
		for(int h = 0; h < 2*nCycles; h++){
			const int parity = (h & 1) ? 0 : 1;
			if(X || h > 0){
				// receive the other colour's face nodes of half-sweep h-1
				const unsigned tag = seqG + (unsigned)h;
				const int c = 1 - parity;
				for(int i = threadIdx.x; i < nHalo; i += blockDim.x){
					int slot, hidx;
					haloNode(i, c, slot, hidx);
					Ph[hidx] = X ? llWaitSys(mine + slot, tag) : llWait(mine + slot, tag);
				}
			}
			__syncthreads();
			const bool send = h + 1 < 2*nCycles;
			const unsigned stag = seqG + (unsigned)h + 1u;
			// two nodes per trip: all loads before the stores (own-colour stores never alias other-colour loads)
			for(int i = threadIdx.x; i < items; i += 2*blockDim.x){
				double vn[2]; int id[2], jj[2], kk[2], lq[2]; bool ok[2];
				#pragma unroll
				for(int w = 0; w < 2; w++){
					int iw = i + w*(int)blockDim.x;
					ok[w] = iw < items;
					if(!ok[w]) continue;
					int m, r, k, l; dHx.divmod(iw, r, m); dBy.divmod(r, l, k); k += 1; l += 1;
					int j = ((((1+k+l)&1) == parity) ? 1 : 2) + 2*m;
					int idx = j + ex*(k + ey*l);
					double rho = B.offRho >= 0 ? Rh[(j-1) + bx*((k-1) + by*(l-1))] : ldg2(L.rho + ix(ox+j, oy+k, oz+l, L.s0, L.s1));
					vn[w] = coeff*(Ph[idx+1] + Ph[idx-1] + Ph[idx+ex] + Ph[idx-ex] + Ph[idx+pl] + Ph[idx-pl] + rho);
					id[w] = idx; jj[w] = j; kk[w] = k; lq[w] = l;
				}
				#pragma unroll
				for(int w = 0; w < 2; w++){
					if(!ok[w]) continue;
					Ph[id[w]] = vn[w];
					if(send) sendNode(jj[w], kk[w], lq[w], vn[w], stag);
				}
			}
		}
		}
		__syncthreads();
		for(int i = threadIdx.x; i < bx*by*bz; i += blockDim.x){
			int jl, r, kl, ll; dBx.divmod(i, r, jl); dBy.divmod(r, ll, kl);
			bsum += Ph[(jl+1) + ex*((kl+1) + ey*(ll+1))];
		}
	}
	seq += 2u*(unsigned)nCycles + (X ? 2u : 0u);
	const long long tTail = clock64();
	if(!X && pendOut){
		if(act)
			for(int i = threadIdx.x; i < bx*by*bz; i += blockDim.x){
				int jl, r, kl, ll; dBx.divmod(i, r, jl); dBy.divmod(r, ll, kl);
				L.phi[ix(ox+jl+1, oy+kl+1, oz+ll+1, L.s0, L.s1)] = Ph[(jl+1) + ex*((kl+1) + ey*(ll+1))];
			}
		*pendOut = S.allSum(bsum)/((double)t0*t1*t2);
		if(S.K->prof && bid == 0 && threadIdx.x == 0){ S.K->prof[2*12] += clock64() - tTail; S.K->prof[2*12+1] += 1; }
		return;
	}
	// the 2*nCycles gBnd calls, applied once (as fGS does in batched mode), on the way back to global memory
	double avg = X ? S.allSumX(bsum)/((double)t0*t1*t2*(double)S.X->R) : S.allSum(bsum)/((double)t0*t1*t2);
	if(act)
		for(int i = threadIdx.x; i < bx*by*bz; i += blockDim.x){
			int jl, r, kl, ll; dBx.divmod(i, r, jl); dBy.divmod(r, ll, kl);
			L.phi[ix(ox+jl+1, oy+kl+1, oz+ll+1, L.s0, L.s1)] = Ph[(jl+1) + ex*((kl+1) + ey*(ll+1))] - avg;
		}
	S.sync();
	if(S.K->prof && bid == 0 && threadIdx.x == 0){ S.K->prof[2*12] += clock64() - tTail; S.K->prof[2*12+1] += 1; }
}

// gBnd(rho) = gNeutralizeGrid (src/grid.c:730-779) ahead of a block-resident smoother call: every block CTA handles the
// rho of its own nodes with the thread <-> node mapping of bGS's fast path, so the smoother reads what the same thread
// wrote and no grid barrier is needed after the subtraction (readers in other CTAs come after the smoother's barriers)
template<bool X> __device__ __noinline__ void bNeutRho(const Lvl &L, const BLvl &B, Scope &S){
	ProfScope psn(*S.K, 28);
	const int bid = (int)blockIdx.x - 1;
	const bool act = bid >= 0 && bid < B.nb;
	const int bx = B.bx, by = B.by, bz = B.bz, hx = bx/2, items = hx*by*bz;
	const FD dHx(hx), dBy(by);
	long g[4] = {-1, -1, -1, -1}; double r[4] = {0, 0, 0, 0};
	double acc = 0;
	if(act){
		const int cx = bid % B.nbx, cr = bid / B.nbx, cy = cr % B.nby, cz = cr / B.nby;
		const int ox = cx*bx, oy = cy*by, oz = cz*bz;
		#pragma unroll
		for(int u = 0; u < 4; u++){
			const int c = u >> 1, iw = threadIdx.x + (u & 1)*(int)blockDim.x;
			if(iw < items){
				int m, q, k, l; dHx.divmod(iw, q, m); dBy.divmod(q, l, k); k += 1; l += 1;
				int j = ((((1+k+l)&1) == c) ? 1 : 2) + 2*m;
				g[u] = ix(ox+j, oy+k, oz+l, L.s0, L.s1);
			}
		}
		#pragma unroll
		for(int u = 0; u < 4; u++) if(g[u] >= 0) r[u] = ldg2(L.rho + g[u]);
		#pragma unroll
		for(int u = 0; u < 4; u++) acc += r[u];
	}
	const double nt = (double)(L.s0-2)*(L.s1-2)*(L.s2-2);
	const double avg = X ? S.allSumX(acc)/(nt*(double)S.X->R) : S.allSum(acc)/nt;
	#pragma unroll
	for(int u = 0; u < 4; u++) if(g[u] >= 0) L.rho[g[u]] = r[u] - avg;
}
} // namespace pinc
#include "mgrows.cuh"
namespace pinc {

// ---- hybrid multi-rank solve: the distributed levels --------------------------------------------------------------------
// Ghost FACES of a rank-local array (all a 7-point stencil reads): a non-decomposed dimension wraps locally, a decomposed
// one travels through the plane slots of the neighbours' arenas.  Two plane sets alternate; a set is rewritten two calls
// later, after the neighbour has sent its next planes, i.e. after the grid barrier that ended its reading of this set.
// sides: bit 0 = fill the lower ghost layers, bit 1 = the upper ones (the restriction only reads lower ghosts).
__device__ __noinline__ void xHalo(double *v, int s0, int s1, int s2, Scope &S, int sides = 3){
	ProfScope psx(*S.K, PS_GS_BIG_SYNC);
	const XDist &x = *S.X;
	const unsigned tag = ++S.xHalo, set = tag & 1u;
	const int t[3] = {s0-2, s1-2, s2-2};
	const bool pr = S.K->prof && blockIdx.x == 0 && threadIdx.x == 0;
	long long tq = pr ? clock64() : 0;
	uint4 *mine = x.peer[x.me] + x.offPlane + (size_t)set*6*x.planeCap;
	const int n0 = t[1]*t[2], n1 = t[0]*t[2], n2 = t[0]*t[1];
	for(int pass = 0; pass < 2; pass++){
		// one flat index over the three face pairs: a thread has at most one node per pass on the benchmark grids, so a pass
		// is one L2 round trip, not three
		for(int ii = (int)S.tid(); ii < n0 + n1 + n2; ii += (int)S.nthr()){
			const int d = ii < n0 ? 0 : (ii < n0 + n1 ? 1 : 2), i = ii - (d == 0 ? 0 : (d == 1 ? n0 : n0 + n1));
			const bool dec = x.ns[d] > 1;
			if(pass == 1 && !dec) continue;
			const int ta = d == 0 ? t[1] : t[0];
			const int b0 = i / ta, a = i - b0*ta + 1, b = b0 + 1;
			long gLo, gHi, hLo, hHi;          // boundary layers, ghost layers
			if(d == 0){ gLo = ix(1,a,b,s0,s1); gHi = ix(t[0],a,b,s0,s1); hLo = ix(0,a,b,s0,s1); hHi = ix(t[0]+1,a,b,s0,s1); }
			else if(d == 1){ gLo = ix(a,1,b,s0,s1); gHi = ix(a,t[1],b,s0,s1); hLo = ix(a,0,b,s0,s1); hHi = ix(a,t[1]+1,b,s0,s1); }
			else { gLo = ix(a,b,1,s0,s1); gHi = ix(a,b,t[2],s0,s1); hLo = ix(a,b,0,s0,s1); hHi = ix(a,b,t[2]+1,s0,s1); }
			if(pass == 0){
				// my lower layer fills somebody's UPPER ghost, my upper layer somebody's LOWER ghost
				double vLo = 0, vHi = 0;
				if(sides & 2) vLo = ldg2(v + gLo);
				if(sides & 1) vHi = ldg2(v + gHi);
				if(dec){
					if(sides & 2) llStoreSys(x.peer[x.nbr[2*d]] + x.offPlane + (size_t)(set*6 + 2*d+1)*x.planeCap + i, vLo, tag);
					if(sides & 1) llStoreSys(x.peer[x.nbr[2*d+1]] + x.offPlane + (size_t)(set*6 + 2*d)*x.planeCap + i, vHi, tag);
				} else { if(sides & 2) v[hHi] = vLo; if(sides & 1) v[hLo] = vHi; }
			} else {
				if(sides & 1) v[hLo] = llWaitSys(mine + (size_t)(2*d)*x.planeCap + i, tag);
				if(sides & 2) v[hHi] = llWaitSys(mine + (size_t)(2*d+1)*x.planeCap + i, tag);
			}
		}
		if(pr){ long long tn = clock64(); S.K->prof[2*(13+pass)] += tn - tq; S.K->prof[2*(13+pass)+1] += 1; tq = tn; }
	}
	S.sync();
	if(pr){ S.K->prof[2*15] += clock64() - tq; S.K->prof[2*15+1] += 1; }
}

// X: level q is distributed (hybrid multi-rank solve); its coarse level q+1 is the first replicated one (qDist = q+1)
// ROWS: the plan has a row-smoothed level (mgrows.cuh); a compile-time switch because the persistent kernel is sensitive to
// its instruction footprint (the row smoother's code cost the 64^3 solve 7 % while it was part of the default kernel)
// pend[q]: mean shift still pending on the stored phi of level q (see bGS); X and row-smoothed levels always store final values
template<bool X = false, bool ROWS = false> __device__ __noinline__ void fDown(const MgPlan &P, int q, Scope &S, unsigned &seq, double *pend){
	const Lvl &L = P.L[q], &C = P.L[q+1];
	const bool blk = P.B[q].on && !S.single && P.nPre > 0;
	if constexpr(X){
		if(P.B[q].on == 1) bNeutRho<true>(L, P.B[q], S); else fNeutralize<true>(L.rho, L.s0, L.s1, L.s2, S);
		bGS<true>(L, P.B[q], P.nPre, 0.0, S, S.xMail);
		xHalo(L.phi, L.s0, L.s1, L.s2, S);
	} else {
	bool rows = false;
	if constexpr(ROWS) rows = blk && P.B[q].on == 3;
	if constexpr(ROWS){ if(rows){ if(P.B[q].bx == 32) rNeutRho<16>(L, P.B[q], S); else rNeutRho<8>(L, P.B[q], S); } }
	if(!rows){ if(blk && P.B[q].on == 1) bNeutRho<false>(L, P.B[q], S); else fNeutralize(L.rho, L.s0, L.s1, L.s2, S); }
	const double sIn = pend[q];            // left by the previous V-cycle's post-smoothing of this level
	if constexpr(ROWS){ if(rows){ if(P.B[q].bx == 32) rGS<16>(L, P.B[q], P.nPre, sIn, S, seq); else rGS<8>(L, P.B[q], P.nPre, sIn, S, seq); pend[q] = 0.0; } }
	if(!rows){ if(blk) bGS<false>(L, P.B[q], P.nPre, sIn, S, seq, &pend[q]); else { fGS(L, P.nPre, sIn, P.exact, S); pend[q] = 0.0; } }
	}
	{
		ProfScope psr(*S.K, S.single ? 27 : 29);
		int t0 = L.s0-2, t1 = L.s1-2, t2 = L.s2-2; long nt = (long)t0*t1*t2;
		for(long i = S.tid(); i < nt; i += S.nthr()){ int j,k,l; truePoint(i,t0,t1,j,k,l);
			L.res[ix(j,k,l,L.s0,L.s1)] = resPoint<!X>(L.phi, L.rho, j, k, l, L.s0, L.s1, L.s2, X ? 0.0 : pend[q]); }
		S.sync();
	}
	if constexpr(X){
		// restriction of my sub-domain, gathered into every rank's replicated rho(q+1): each coarse value goes into slot
		// (my rank, node) of every rank's gather buffer; then every rank unpacks all R segments into the global array
		xHalo(L.res, L.s0, L.s1, L.s2, S, 1);
		ProfScope psr(*S.K, 30);
		const XDist &x = *S.X;
		const unsigned tag = ++S.xGath, set = tag & 1u;
		const int c0 = (L.s0-2)/2, c1 = (L.s1-2)/2, c2 = (L.s2-2)/2; const long nc = (long)c0*c1*c2;
		for(long i = S.tid(); i < nc; i += S.nthr()){
			int J,K,Lz; truePoint(i,c0,c1,J,K,Lz);
			const double v = restrictPoint<false>(L.res, J, K, Lz, L.s0, L.s1, L.s2);
			for(int r = 0; r < x.R; r++) llStoreSys(x.peer[r] + x.offGath + (size_t)set*x.gathCap + (size_t)x.me*nc + i, v, tag);
		}
		const uint4 *in = x.peer[x.me] + x.offGath + (size_t)set*x.gathCap;
		for(long ii = S.tid(); ii < nc*x.R; ii += S.nthr()){
			const int r = (int)(ii / nc); const long i = ii - (long)r*nc;
			int J,K,Lz; truePoint(i,c0,c1,J,K,Lz);
			const int rx = r % x.ns[0], rr = r / x.ns[0], ry = rr % x.ns[1], rz = rr / x.ns[1];
			C.rho[ix(rx*c0 + J, ry*c1 + K, rz*c2 + Lz, C.s0, C.s1)] = llWaitSys(in + ii, tag);
		}
		S.sync();
	} else {
		ProfScope psr(*S.K, S.single ? 27 : 30);
		int t0 = C.s0-2, t1 = C.s1-2, t2 = C.s2-2; long nt = (long)t0*t1*t2;
		for(long i = S.tid(); i < nt; i += S.nthr()){ int J,K,Lz; truePoint(i,t0,t1,J,K,Lz);
			C.rho[ix(J,K,Lz,C.s0,C.s1)] = restrictPoint<true>(L.res, J, K, Lz, L.s0, L.s1, L.s2); }
		S.sync();
	}
}
__device__ __noinline__ void fBottom(const MgPlan &P, Scope &S){
	const Lvl &L = P.L[P.nLevels-1];
	fNeutralize(L.rho, L.s0, L.s1, L.s2, S);
	fGS(L, P.nCoarse, 0.0, P.exact, S);
	if(P.exact || P.nCoarse <= 0) fNeutralize(L.phi, L.s0, L.s1, L.s2, S);      // batched mode: fGS just ended with this gBnd
}
// res(q) := P(phi(q+1)); phi(q) += res(q); gBnd; post-smooth; gBnd
template<bool X = false, bool ROWS = false> __device__ __noinline__ void fUp(const MgPlan &P, int q, Scope &S, unsigned &seq, double *pend){
	const Lvl &L = P.L[q], &C = P.L[q+1];
	int t0 = L.s0-2, t1 = L.s1-2, t2 = L.s2-2; long nt = (long)t0*t1*t2;
	const long long tUp = clock64();
	double acc = 0;
	// X: the coarse level is the replicated global one; my fine node (j,k,l) is global node (ox+j, oy+k, oz+l)
	const int ox = X ? S.X->sub[0]*t0 : 0, oy = X ? S.X->sub[1]*t1 : 0, oz = X ? S.X->sub[2]*t2 : 0;
	const double sC = pend[q+1], sF = X ? 0.0 : pend[q];      // pending shifts of the coarse level's phi and of this level's
	for(long i = S.tid(); i < nt; i += S.nthr()){
		int j,k,l; truePoint(i,t0,t1,j,k,l); long g = ix(j,k,l,L.s0,L.s1);
		double p = prolPoint(C.phi, ox+j, oy+k, oz+l, C.s0, C.s1, C.s2, sC);
		L.res[g] = p;
		double v = ldg2(L.phi + g) - sF; v += p;
		L.phi[g] = v;
		acc += v;
	}
	if(!X) pend[q] = 0.0;
	if constexpr(X){
		double avg = S.allSumX(acc)/((double)nt*(double)S.X->R);
		bGS<true>(L, P.B[q], P.nPost, avg, S, S.xMail);
		return;
	}
	double avg = S.allSum(acc)/(double)nt;
	if(S.K->prof && blockIdx.x == 0 && threadIdx.x == 0 && !S.single){ S.K->prof[2*31] += clock64() - tUp; S.K->prof[2*31+1] += 1; }
	bool rows = false;
	if constexpr(ROWS){
		rows = P.B[q].on == 3 && !S.single && P.nPost > 0;
		if(rows){ if(P.B[q].bx == 32) rGS<16>(L, P.B[q], P.nPost, avg, S, seq); else rGS<8>(L, P.B[q], P.nPost, avg, S, seq); }
	}
	if(!rows){ if(P.B[q].on && !S.single && P.nPost > 0) bGS<false>(L, P.B[q], P.nPost, avg, S, seq, &pend[q]); else fGS(L, P.nPost, avg, P.exact, S); }
	if(P.exact || P.nPost <= 0) fNeutralize(L.phi, L.s0, L.s1, L.s2, S);
}
__device__ __noinline__ void fGhosts(double *v, int s0, int s1, int s2, Scope &S){
	long n = (long)s0*s1*s2;
	for(long i = S.tid(); i < n; i += S.nthr()){
		int j = (int)(i % s0); long r = i / s0; int k = (int)(r % s1); int l = (int)(r / s1);
		int jw = j == 0 ? s0-2 : (j == s0-1 ? 1 : j);
		int kw = k == 0 ? s1-2 : (k == s1-1 ? 1 : k);
		int lw = l == 0 ? s2-2 : (l == s2-1 ? 1 : l);
		if(jw != j || kw != k || lw != l) v[i] = ldg2(v + ix(jw,kw,lw,s0,s1));
	}
}

// small levels of the all-SM kernel: CTA 0, shared memory, the routines of mgsmem.cuh
template<bool EXACT> __device__ __noinline__ void smallSection(const MgPlan &P, CK &K){
	const CPlan &C = P.C;
	const int b = C.nLevels - 1, qs = P.qSmall;
	{	// rho(qs) was restricted into global memory by the grid-wide part
		const CLvl &L = C.L[qs];
		int n = L.nx*L.ny*L.nz;
		for(int i = threadIdx.x; i < n; i += blockDim.x){ int j,k,l; ownNode(L,1,i,j,k,l); mgS[L.offRho + i] = __ldcg(L.rho + gix(L,j,k,l)); }
		__syncthreads();
	}
	for(int q = qs; q < b; q++) cDown<true,EXACT>(C, q, K);
	cBottom<true,EXACT>(C, K);
	for(int q = b-1; q >= qs; q--) cUp<true,EXACT>(C, q, K);
	{	// phi(qs) back to global for the grid-wide prolongation
		const CLvl &L = C.L[qs];
		int n = L.nx*L.ny*L.nz;
		for(int i = threadIdx.x; i < n; i += blockDim.x){
			int j,k,l; ownNode(L,1,i,j,k,l);
			L.phiG[gix(L,j,k,l)] = mgS[L.offPhi + i];
			if(qs == 0) L.rho[gix(L,j,k,l)] = mgS[L.offRho + i];      // the residual norm reads the neutralised rho
		}
	}
}

// small levels of the benchmark pyramid (cubic, 16^3 / 8^3 / 4^3, gBnd batched): the routines of mgsmall.cuh
template<int N> __device__ __forceinline__ void pyDown(const CPlan &C, int q, CK &K, double *Z){
	const CLvl &L = C.L[q];
	{ ProfScope ps(K, PS_GS_SMALL); ProfScope psl(K, psLvl(L, 0)); sSmooth<N>(mgS + L.offPhi, mgS + L.offRho, Z, C.nPre, 0.0, K); }
	{ ProfScope ps(K, PS_RESTRICT); ProfScope psl(K, psLvl(L, 1)); sRestrict<N>(mgS + L.offPhi, mgS + L.offRho, mgS + C.L[q+1].offRho, K); }
}
template<int N> __device__ __forceinline__ void pyUp(const CPlan &C, int q, CK &K, double *Z){
	const CLvl &L = C.L[q];
	double avg;
	{ ProfScope ps(K, PS_PROLONG); ProfScope psl(K, psLvl(L, 2)); avg = sProlongAdd<N>(mgS + L.offPhi, mgS + C.L[q+1].offPhi, L.res, L.s0, L.s1, K); }
	{ ProfScope ps(K, PS_GS_SMALL); ProfScope psl(K, psLvl(L, 0)); sSmooth<N>(mgS + L.offPhi, mgS + L.offRho, Z, C.nPost, avg, K); }
}
__device__ __noinline__ void smallPyramid(const MgPlan &P, CK &K){
	const CPlan &C = P.C;
	const int b = C.nLevels - 1, qs = P.qSmall;
	double *Z = mgS + (P.offZ >= 0 ? P.offZ : 0);
	{	// rho(qs) was restricted into global memory by the grid-wide part (or is the solver's input); gBnd(rho)
		const CLvl &L = C.L[qs];
		const int N = L.nx, n = N*N*N;
		for(int i = threadIdx.x; i < n; i += blockDim.x){ int x = i % N, r = i / N, y = r % N, z = r / N; mgS[L.offRho + i] = __ldcg(L.rho + gix(L,x+1,y+1,z+1)); }
		__syncthreads();
		ProfScope ps(K, PS_NEUT_RHO);
		if(N == 16) sNeutralize<16>(mgS + L.offRho, K); else if(N == 8) sNeutralize<8>(mgS + L.offRho, K); else sNeutralize<4>(mgS + L.offRho, K);
	}
	for(int q = qs; q < b; q++){ if(C.L[q].nx == 16) pyDown<16>(C, q, K, Z); else pyDown<8>(C, q, K, Z); }
	{
		const CLvl &L = C.L[b];
		ProfScope ps(K, PS_GS_SMALL); ProfScope psl(K, psLvl(L, 0));
		if(L.nx == 16) sSmooth<16>(mgS + L.offPhi, mgS + L.offRho, Z, C.nCoarse, 0.0, K);
		else if(L.nx == 8) sSmooth<8>(mgS + L.offPhi, mgS + L.offRho, Z, C.nCoarse, 0.0, K);
		else sSmooth<4>(mgS + L.offPhi, mgS + L.offRho, Z, C.nCoarse, 0.0, K);
	}
	for(int q = b-1; q >= qs; q--){ if(C.L[q].nx == 16) pyUp<16>(C, q, K, Z); else pyUp<8>(C, q, K, Z); }
	{	// phi(qs) back to global for the grid-wide prolongation / the residual norm
		const CLvl &L = C.L[qs];
		const int N = L.nx, n = N*N*N;
		for(int i = threadIdx.x; i < n; i += blockDim.x){
			int x = i % N, r = i / N, y = r % N, z = r / N;
			L.phiG[gix(L,x+1,y+1,z+1)] = mgS[L.offPhi + i];
			if(qs == 0) L.rho[gix(L,x+1,y+1,z+1)] = mgS[L.offRho + i];
		}
	}
}

// EXACT (gBnd after every half-sweep, modes 1 and 3) is a compile-time parameter only to keep the code of the default
// kernel small: the persistent kernel is latency-bound and sensitive to its instruction footprint
template<bool EXACT, bool DIST, bool ROWS> __global__ void __launch_bounds__(MG_BLOCK, 1) k_mg_solve(const __grid_constant__ MgPlan P){
	__shared__ double sh[18];
	__shared__ double red[40];
	CK K{ cg::this_cluster(), (int)blockIdx.x, 1, mgS, red, 0, P.prof, P.offZ };
	if(P.smemSmall && blockIdx.x == 0){
		for(int q = P.qSmall; q < P.nLevels; q++){
			const CLvl &L = P.C.L[q];
			int n = L.nx*L.ny*L.nz;
			for(int i = threadIdx.x; i < n; i += blockDim.x){
				int j,k,l; ownNode(L,1,i,j,k,l);
				mgS[L.offPhi + i] = __ldcg(L.phiG + gix(L,j,k,l));
				mgS[L.offRho + i] = __ldcg(L.rho + gix(L,j,k,l));
			}
		}
		__syncthreads();
	}
	Scope Sg; Sg.single = false; Sg.bar = P.bar; Sg.partial = P.partial; Sg.flip = 0; Sg.sh = sh; Sg.gen = 0;
	Sg.X = &P.X; Sg.xMail = Sg.xSum = Sg.xHalo = Sg.xGath = 0;
	if constexpr(DIST){ Sg.xMail = P.X.ctr[0]; Sg.xSum = P.X.ctr[1]; Sg.xHalo = P.X.ctr[2]; Sg.xGath = P.X.ctr[3]; }      // rewritten after the last grid barrier of this launch
	if(threadIdx.x == 0) Sg.gen = P.barBase;            // arrival count at kernel start (host-tracked)
	Sg.K = &K;
	Scope S1 = Sg; S1.single = true;
	int b = P.nLevels - 1;
	int qs = P.qSmall < 0 ? 0 : P.qSmall;
	double barRes = 2.;
	int cycles = 0;
	double pend[MG_MAXLEV + 1];           // per level: mean shift still pending on the stored phi (bGS); identical in every thread
	for(int q = 0; q <= MG_MAXLEV; q++) pend[q] = 0.0;
	unsigned seq = *((volatile unsigned*)P.seqWord);       // rewritten only after the last grid barrier of this launch
	if(seq > 0xE0000000u){
		// the 32-bit tags are about to wrap (after ~10^6 time steps): back to the initial state - every slot zero,
		// sequence 0 - so that a stale tag can never match a fresh one
		for(unsigned long long i = (unsigned long long)Sg.tid(); i < P.mailSlots; i += (unsigned long long)Sg.nthr()) P.mailAll[i] = make_uint4(0u, 0u, 0u, 0u);
		__threadfence();
		Sg.sync();
		seq = 0;
	}
	while(barRes > P.tol && cycles < P.maxCycles){
		for(int q = 0; q <= b && q < qs; q++){
			if(DIST && q < P.X.qDist) fDown<DIST>(P, q, Sg, seq, pend);
			else if(q < b) fDown<false,ROWS>(P, q, Sg, seq, pend); else fBottom(P, Sg);
		}
		if(qs <= b){
			if(blockIdx.x == 0){
				ProfScope pss(K, PS_LEVEL0);
				if(P.pyramid) smallPyramid(P, K);
				else if(P.smemSmall){
					smallSection<EXACT>(P, K);
				} else {
					for(int q = qs; q < b; q++) fDown(P, q, S1, seq, pend);
					fBottom(P, S1);
					for(int q = b-1; q >= qs; q--) fUp(P, q, S1, seq, pend);
				}
			}
			Sg.sync();
		}
		for(int q = (qs <= b ? qs : b) - 1; q >= 0; q--){ if(DIST && q < P.X.qDist) fUp<DIST>(P, q, Sg, seq, pend); else fUp<false,ROWS>(P, q, Sg, seq, pend); }
		// mgSolveRaw :1700-1704: residual, square in place, true-grid sum, RMS
		const Lvl &L = P.L[0];
		ProfScope psn(K, PS_NORM);
		int t0 = L.s0-2, t1 = L.s1-2, t2 = L.s2-2; long nt = (long)t0*t1*t2;
		double acc = 0;
		if constexpr(DIST) xHalo(L.phi, L.s0, L.s1, L.s2, Sg);
		for(long i = Sg.tid(); i < nt; i += Sg.nthr()){
			int j,k,l; truePoint(i,t0,t1,j,k,l);
			double r = resPoint<!DIST>(L.phi, L.rho, j, k, l, L.s0, L.s1, L.s2, pend[0]);
			r = r*r;
			L.res[ix(j,k,l,L.s0,L.s1)] = r;
			acc += r;
		}
		barRes = DIST ? Sg.allSumX(acc) : Sg.allSum(acc);
		barRes /= P.totTrue;
		barRes = sqrt(barRes);
		if(blockIdx.x == 0 && threadIdx.x == 0 && cycles < 250) P.hist[1+cycles] = barRes;
		cycles++;
	}
	if(blockIdx.x == 0 && threadIdx.x == 0){ P.hist[0] = (double)cycles; P.hist[251] = barRes; }
	const unsigned seqEnd = seq;
	{	// the shifts that are still pending become part of the stored values: what the arrays hold after the solve is final
		bool any = false;
		for(int q = 0; q <= b; q++){
			if(pend[q] == 0.0) continue;
			any = true;
			const Lvl &L = P.L[q];
			int t0 = L.s0-2, t1 = L.s1-2, t2 = L.s2-2; long nt = (long)t0*t1*t2;
			for(long i = Sg.tid(); i < nt; i += Sg.nthr()){ int j,k,l; truePoint(i,t0,t1,j,k,l); long g = ix(j,k,l,L.s0,L.s1); L.phi[g] = ldg2(L.phi + g) - pend[q]; }
		}
		if(any) Sg.sync();
	}
	if(P.smemSmall){
		if(blockIdx.x == 0)
			for(int q = P.qSmall; q <= b; q++){
				const CLvl &L = P.C.L[q];
				int n = L.nx*L.ny*L.nz;
				for(int i = threadIdx.x; i < n; i += blockDim.x){
					int j,k,l; ownNode(L,1,i,j,k,l);
					L.phiG[gix(L,j,k,l)] = mgS[L.offPhi + i];
					L.rho[gix(L,j,k,l)] = mgS[L.offRho + i];
				}
			}
		Sg.sync();
	}
	for(int q = 0; q <= b; q++){
		if(DIST && q < P.X.qDist) continue;          // rank-local levels: the host fills their ghost layers from the neighbours (hybridSolve)
		fGhosts(P.L[q].phi, P.L[q].s0, P.L[q].s1, P.L[q].s2, Sg);
		fGhosts(P.L[q].rho, P.L[q].s0, P.L[q].s1, P.L[q].s2, Sg);
		fGhosts(P.L[q].res, P.L[q].s0, P.L[q].s1, P.L[q].s2, Sg);
	}
	if(blockIdx.x == 0 && threadIdx.x == 0){
		*P.seqWord = seqEnd;
		if constexpr(DIST){ P.X.ctr[0] = Sg.xMail; P.X.ctr[1] = Sg.xSum; P.X.ctr[2] = Sg.xHalo; P.X.ctr[3] = Sg.xGath; P.hist[252] = (double)Sg.xMail; }
	}
}

int g_mgMode = -1;
// -1: from $PINC_B200_MG at first use; 0 ops; 1 fused (grid-wide persistent kernel, gBnd per half-sweep);
// 2 cluster (DSMEM-resident, gBnd batched per smoother call; default); 3 cluster with gBnd per half-sweep
int g_mgForceCluster = 0;     // $PINC_B200_MG=cluster-always: use the cluster kernel whenever it fits
int g_mgNoCluster = 0;        // $PINC_B200_MG=allsm: use the all-SM kernel whatever the size
int g_mgRowMode = -1;         // row smoother (mgrows.cuh): 0 off, 1 for big blocks (default), 2 whenever the block shape allows; -1: from $PINC_B200_MG_ROWMODE
int g_mgReplica = -1;         // multi-rank solves replicated (1) or distributed (0); -1: from $PINC_B200_MG_REPLICA at first use
static void ensureHist(Ctx *c){
	if(c->d_mgHist) return;
	PINC_CUDA(cudaMalloc(&c->d_mgHist, 256*sizeof(double)));
	PINC_CUDA(cudaMallocHost(&c->h_mgHist, 256*sizeof(double)));
	c->h_mgHist[0] = 0;
}

static bool fusedEligible(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *m){
	if(!mgMode()) return false;
	if(m->mpiSize != 1) return false;
	int nL = mgRho->nLevels;
	if(nL < 2 || nL > MG_MAXLEV) return false;
	for(int q = 0; q < nL; q++){
		const Grid *g = mgRho->grids[q];
		if(g->rank != 4) return false;
		for(int d = 1; d < 4; d++){
			if(g->trueSize[d] < 2 || (g->trueSize[d] & 1)) return false;
			if(g->bnd[d] != PERIODIC || g->bnd[d+g->rank] != PERIODIC) return false;
		}
	}
	(void)c; (void)mgPhi; (void)mgRes;
	return true;
}

// Decomposition of a t0 x t1 x t2 level into at most maxBlocks equal blocks with even edges: as many blocks as
// possible, then the smallest surface, then the longest x edge.
static bool planBlocks(int t0, int t1, int t2, int maxBlocks, long smemCap, BLvl &B){
	B = BLvl{};
	long bestSurf = 0; int bestNb = 0;
	for(int nx = 1; nx <= t0; nx++){
		if(t0 % nx || ((t0/nx) & 1)) continue;
		for(int ny = 1; ny <= t1 && nx*ny <= maxBlocks; ny++){
			if(t1 % ny || ((t1/ny) & 1)) continue;
			for(int nz = 1; nz <= t2 && nx*ny*nz <= maxBlocks; nz++){
				if(t2 % nz || ((t2/nz) & 1)) continue;
				int bx = t0/nx, by = t1/ny, bz = t2/nz, nb = nx*ny*nz;
				long phiB = (long)(bx+2)*(by+2)*(bz+2)*8;
				if(phiB > smemCap) continue;
				long surf = (long)by*bz + (long)bx*bz + (long)bx*by;
				bool better = nb > bestNb || (nb == bestNb && (surf < bestSurf || (surf == bestSurf && bx > B.bx)));
				if(!better) continue;
				bestNb = nb; bestSurf = surf;
				B.bx = bx; B.by = by; B.bz = bz; B.nbx = nx; B.nby = ny; B.nbz = nz; B.nb = nb;
				B.offPhi = 0;
				B.offRho = (phiB + (long)bx*by*bz*8 <= smemCap) ? (bx+2)*(by+2)*(bz+2) : -1;
			}
		}
	}
	B.on = bestNb >= 8;
	return B.on != 0;
}

// arena of the hybrid multi-rank solve: one allocation per rank, mapped into every other rank (Transport::peerAlloc)
struct XArena { char *mine = nullptr; std::vector<char*> peers; size_t bytes = 0; int share = 1; };
static void freeXArena(Ctx *c){
	XArena *A = (XArena*)c->mgXArena;
	if(!A) return;
	if(A->mine && c->tp) c->tp->peerFree(c, A->mine, A->peers);
	delete A;
	c->mgXArena = nullptr;
}
// collective: every rank asks for the same size (same plan)
static XArena *ensureXArena(Ctx *c, size_t bytes){
	XArena *A = (XArena*)c->mgXArena;
	if(A && A->bytes >= bytes) return A;
	int share = A ? A->share : 0;
	if(A){ streamSync(c); c->tp->barrier(c); freeXArena(c); }
	A = new XArena();
	bytes += bytes/4;
	if(!c->tp->peerAlloc(c, bytes, &A->mine, A->peers)){ delete A; return nullptr; }
	A->bytes = bytes;
	if(!share){
		// ranks sharing this rank's device (thread ranks of the tests): their persistent kernels must be co-resident
		std::vector<long> dev(c->size);
		long mine = c->device;
		c->tp->allgatherLong(c, &mine, 1, dev.data());
		for(int r = 0; r < c->size; r++) if(dev[r] == mine && !strcmp(c->tp->name(), "threads")) share++;
		if(share < 1) share = 1;
	}
	A->share = share;
	c->mgXArena = A;
	c->mgXTagHigh = 0;
	return A;
}

// XH: hybrid multi-rank solve (levels < XH->qDist are this rank's sub-domain, the others the replicated global levels);
// returns false if the distributed levels cannot be smoothed block-resident (nothing launched, nothing changed)
static bool fusedSolve(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, double tol, int maxCycles, int exact, const XDist *XH = nullptr){
	MgPlan P{};
	int nL = mgRho->nLevels;
	const int qDist = XH ? XH->qDist : 0;
	int share = 1;
	if(XH){
		if(exact || nL < 2 || qDist != 1) return false;
		XArena *A0 = ensureXArena(c, 4096);
		if(!A0) return false;
		share = A0->share;
	}
	P.nLevels = nL; P.nPre = mgRho->nPreSmooth; P.nPost = mgRho->nPostSmooth; P.nCoarse = mgRho->nCoarseSolve;
	P.qSmall = nL; P.offZ = -1;
	double work = 0;
	for(int q = 0; q < nL; q++){
		DevGrid *r = devGrid(c, mgRho->grids[q]), *p = devGrid(c, mgPhi->grids[q]), *e = devGrid(c, mgRes->grids[q]);
		if(r->n != p->n || r->n != e->n || r->nv != 1) fatal("multigrid level %d: rho/phi/res differ in shape", q);
		P.L[q] = Lvl{ p->d, r->d, e->d, r->size[0], r->size[1], r->size[2] };
		long nt = trueCount(r);
		if(nt <= MG_SMALL && q < P.qSmall) P.qSmall = q;
		if(nt > MG_SMALL) P.qSmall = nL;         // small levels must be a suffix
		work += 24.0*nt*(q == nL-1 ? P.nCoarse : P.nPre + P.nPost) + 50.0*nt;
	}
	// recompute the suffix of small levels properly
	P.qSmall = nL;
	for(int q = nL-1; q >= 0; q--){ if(trueCount(devGrid(c, mgRho->grids[q])) <= MG_SMALL) P.qSmall = q; else break; }
	P.tol = tol; P.maxCycles = maxCycles; P.exact = exact;
	// small levels in CTA 0's shared memory (needs at least one grid-wide level above them)
	size_t smem = 0;
	P.smemSmall = 0;
	if(P.qSmall >= 0 && P.qSmall < nL){
		long off = 0;
		bool ok = true;
		P.C.nLevels = nL; P.C.nBig = P.qSmall; P.C.nPre = P.nPre; P.C.nPost = P.nPost; P.C.nCoarse = P.nCoarse; P.C.nc = 1;
		for(int q = P.qSmall; q < nL; q++){
			DevGrid *r = devGrid(c, mgRho->grids[q]), *p = devGrid(c, mgPhi->grids[q]), *e = devGrid(c, mgRes->grids[q]);
			CLvl &L = P.C.L[q];
			L.phiG = p->d; L.rho = r->d; L.res = e->d;
			L.nx = r->tsize[0]; L.ny = r->tsize[1]; L.nz = r->tsize[2]; L.s0 = r->size[0]; L.s1 = r->size[1];
			L.ppc = L.nz; L.small = 1;
			long nt = (long)L.nx*L.ny*L.nz;
			if((L.nx/2)*L.ny*L.nz > MC_U*MG_BLOCK) ok = false;
			L.offPhi = (int)off; off += nt;
			L.offRho = (int)off; off += nt;
		}
		P.offZ = -1;
		for(int q = P.qSmall; q < nL; q++) if(P.C.L[q].nx == 16 && P.C.L[q].ny == 16 && P.C.L[q].nz == 16){ P.offZ = (int)off; off += MS_ZBUF; break; }
		smem = (size_t)off*sizeof(double);
		P.smemSmall = ok ? 1 : 0;
		bool pyr = ok && !exact && P.nPre > 0 && P.nPost > 0 && P.nCoarse > 0;
		for(int q = P.qSmall; q < nL; q++){
			const CLvl &L = P.C.L[q];
			if(L.nx != L.ny || L.ny != L.nz || (L.nx != 16 && L.nx != 8 && L.nx != 4)) pyr = false;
			if(L.nx == 4 && q != nL-1) pyr = false;
			if(L.nx == 16 && P.offZ < 0) pyr = false;
		}
		static const bool noPyr = getenv("PINC_B200_MG_PYRAMID") && atoi(getenv("PINC_B200_MG_PYRAMID")) == 0;
		P.pyramid = pyr && !noPyr ? 1 : 0;
		if(!ok) smem = 0;
	}
	DevGrid *r0 = devGrid(c, mgRho->grids[0]);
	P.totTrue = (double)trueCount(r0)*(XH ? XH->R : 1);
	int grid = P.qSmall == 0 ? 1 : c->numSMs/share;
	if(XH && (P.qSmall < qDist || grid < 10)) return false;          // a distributed level must be a grid-wide one
	// block-resident smoothing of the grid-wide levels (batched gBnd only; CTA 0 is left to the small levels)
	P.seqWord = c->d_bar + 16;
	static const bool blocksOff = getenv("PINC_B200_MG_BLOCKS") && atoi(getenv("PINC_B200_MG_BLOCKS")) == 0;
	size_t xMailOff[MG_MAXLEV] = {0}, xSlots = 0;
	if(XH && (blocksOff || P.nPre <= 0 || P.nPost <= 0)) return false;
	if(!exact && grid > 8 && !blocksOff){
		int smemOptin = 0;
		PINC_CUDA(cudaDeviceGetAttribute(&smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
		const long smemCap = smemOptin - 4096;
		size_t mailSlots = 0, blockSmem = 0, mailOff[MG_MAXLEV] = {0};
		for(int q = 0; q < P.qSmall && q < nL; q++){
			DevGrid *r = devGrid(c, mgRho->grids[q]);
			BLvl &B = P.B[q];
			if(!planBlocks(r->tsize[0], r->tsize[1], r->tsize[2], grid-1, smemCap, B)){ if(q < qDist) return false; continue; }
			if(B.bx > 1000 || B.by > 1000 || B.bz > 1000){ B.on = 0; if(q < qDist) return false; continue; }        // packed node coordinates in bGS
			if(q < qDist){
				// distributed level of a hybrid solve: always block-resident (bGS<true>), mailboxes in the peer-mapped arena
				if((B.bx/2)*B.by*B.bz > 2*MG_BLOCK || B.by*B.bz + B.bx*B.bz + B.bx*B.by > MG_BLOCK) B.on = 2;
				B.rows = 1;
				xMailOff[q] = xSlots;
				xSlots += (size_t)B.nb*2*(B.by*B.bz + B.bx*B.bz + B.bx*B.by);
				size_t need = (size_t)(B.offRho >= 0 ? B.offRho + B.bx*B.by*B.bz : (B.bx+2)*(B.by+2)*(B.bz+2))*sizeof(double);
				if(need > blockSmem) blockSmem = need;
				continue;
			}
			{	// big blocks: one x-row per thread, colour-separated rows (mgrows.cuh); $PINC_B200_MG_ROWMODE=0 off, 2 also for small blocks
				if(g_mgRowMode < 0) g_mgRowMode = getenv("PINC_B200_MG_ROWMODE") ? atoi(getenv("PINC_B200_MG_ROWMODE")) : 1;
				const int rowMode = g_mgRowMode;
				const bool big = (B.bx/2)*B.by*B.bz > 2*MG_BLOCK || B.by*B.bz + B.bx*B.bz + B.bx*B.by > MG_BLOCK;
				if(rowMode && !XH && (big || rowMode == 2) && (B.bx == 16 || B.bx == 32) && B.by*B.bz <= MG_BLOCK){
					B.on = 3; B.offRho = -1;
					mailOff[q] = mailSlots;
					mailSlots += (size_t)B.nb*2*(B.by*B.bz + B.bx*B.bz + B.bx*B.by);
					size_t need = (size_t)(B.bx+2)*(B.by+2)*(B.bz+2)*sizeof(double);
					if(need > blockSmem) blockSmem = need;
					continue;
				}
			}
			{ static const int maxItems = getenv("PINC_B200_MG_MAXITEMS") ? atoi(getenv("PINC_B200_MG_MAXITEMS")) : 1024;     // bigger blocks are throughput-bound and the grid-wide sweep (fGS) is as fast or faster (us per V-cycle, fGS vs blocks: 64x64x128 532 vs 532, 64x128x128 736 vs 820, 128^3 1122 vs 1367)
			  if((B.bx/2)*B.by*B.bz > maxItems){ B.on = 0; continue; } }
			static const bool noRows = getenv("PINC_B200_MG_ROWS") && atoi(getenv("PINC_B200_MG_ROWS")) == 0;
			B.rows = noRows ? 0 : 1;
			static const bool noFast = getenv("PINC_B200_MG_FAST") && atoi(getenv("PINC_B200_MG_FAST")) == 0;
			if(noFast || (B.bx/2)*B.by*B.bz > 2*MG_BLOCK || B.by*B.bz + B.bx*B.bz + B.bx*B.by > MG_BLOCK) B.on = 2;      // 1: fast path of bGS
			mailOff[q] = mailSlots;
			mailSlots += (size_t)B.nb*2*(B.by*B.bz + B.bx*B.bz + B.bx*B.by);
			size_t need = (size_t)(B.offRho >= 0 ? B.offRho + B.bx*B.by*B.bz : (B.bx+2)*(B.by+2)*(B.bz+2))*sizeof(double);
			if(need > blockSmem) blockSmem = need;
		}
		if(mailSlots){
			if(mailSlots*sizeof(uint4) > c->mgMailBytes){
				if(c->d_mgMail) PINC_CUDA(cudaFree(c->d_mgMail));
				c->mgMailBytes = mailSlots*sizeof(uint4);
				PINC_CUDA(cudaMalloc(&c->d_mgMail, c->mgMailBytes));
				PINC_CUDA(cudaMemsetAsync(c->d_mgMail, 0, c->mgMailBytes, c->stream));
			}
			for(int q = qDist; q < P.qSmall && q < nL; q++) if(P.B[q].on) P.B[q].mail = (uint4*)c->d_mgMail + mailOff[q];
			{	// colour-separated rho copies of the row-smoothed levels
				size_t need = 0, off[MG_MAXLEV] = {0};
				for(int q = 0; q < P.qSmall && q < nL; q++) if(P.B[q].on == 3){ off[q] = need; need += (size_t)P.B[q].nb*P.B[q].bx*P.B[q].by*P.B[q].bz; }
				if(need*sizeof(double) > c->mgRhoSBytes){
					if(c->d_mgRhoS){ streamSync(c); PINC_CUDA(cudaFree(c->d_mgRhoS)); }
					c->mgRhoSBytes = need*sizeof(double);
					PINC_CUDA(cudaMalloc(&c->d_mgRhoS, c->mgRhoSBytes));
				}
				for(int q = 0; q < P.qSmall && q < nL; q++) if(P.B[q].on == 3) P.B[q].rhoS = (double*)c->d_mgRhoS + off[q];
			}
			P.mailAll = (uint4*)c->d_mgMail; P.mailSlots = c->mgMailBytes/sizeof(uint4);
		}
		if(blockSmem > smem) smem = blockSmem;
	}
	if(XH){
		// arena layout in 16-byte slots: [16 counters][2 x XD_MAXR sums][2 x 6 planes][2 x R gather segments][mailboxes of the distributed levels]
		for(int q = 0; q < qDist; q++) if(!P.B[q].on) return false;
		P.X = *XH;
		P.X.on = 1;
		size_t planeCap = 0;
		for(int q = 0; q < qDist; q++){
			DevGrid *r = devGrid(c, mgRho->grids[q]);
			const size_t t0 = r->tsize[0], t1 = r->tsize[1], t2 = r->tsize[2];
			planeCap = std::max(planeCap, std::max(t0*t1, std::max(t0*t2, t1*t2)));
		}
		DevGrid *rl = devGrid(c, mgRho->grids[qDist-1]);
		const size_t nc = (size_t)(rl->tsize[0]/2)*(rl->tsize[1]/2)*(rl->tsize[2]/2);
		P.X.offSum = 16; P.X.offPlane = P.X.offSum + 2*XD_MAXR; P.X.planeCap = (unsigned)planeCap;
		P.X.offGath = P.X.offPlane + 12*(unsigned)planeCap; P.X.gathCap = (unsigned)(nc*XH->R);
		const size_t offMail = (size_t)P.X.offGath + 2*(size_t)P.X.gathCap;
		const size_t slots = offMail + xSlots;
		if(slots >= (1u << LL_FACE_SHIFT)) return false;
		XArena *A = ensureXArena(c, slots*sizeof(uint4));
		if(!A) return false;
		for(int r = 0; r < XH->R; r++) P.X.peer[r] = (uint4*)A->peers[r];
		P.X.ctr = (unsigned*)A->mine;
		for(int q = 0; q < qDist; q++) P.B[q].mail = (uint4*)A->mine + offMail + xMailOff[q];
	}
	bool rows = false;
	for(int q = 0; q < nL; q++) if(P.B[q].on == 3) rows = true;
	if(rows && (XH || exact)) fatal("multigrid plan: a row-smoothed level in a hybrid or exact solve");
	const void *kern = XH ? (const void*)k_mg_solve<false,true,false> : exact ? (const void*)k_mg_solve<true,false,false>
		: rows ? (const void*)k_mg_solve<false,false,true> : (const void*)k_mg_solve<false,false,false>;
	{	// one high-water mark for the kernel's dynamic shared memory
		size_t &attrSet = c->mgAttrSmem[XH ? 2 : exact ? 1 : rows ? 3 : 0];       // per context: the attribute belongs to the device
		if(smem > attrSet){
			if(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) attrSet = smem;
			else { cudaGetLastError(); if(XH) return false; for(int q = 0; q < nL; q++) P.B[q].on = 0; P.smemSmall = 0; P.pyramid = 0; smem = 0; }
		}
	}
	P.partial = partialBuffer(c, 2L*grid);
	P.bar = c->d_bar;
	PINC_CUDA(cudaMemsetAsync(c->d_bar, 0, 2*sizeof(unsigned), c->stream));
	P.barBase = 0;
	ensureHist(c);
	P.hist = c->d_mgHist;
	P.prof = (long long*)mgProfBuffer(c);
	void *args[] = { &P };
	// rank threads sharing a device: every allocation above may synchronise the device, so all of them launch together
	if(XH && share > 1){ streamSync(c); c->tp->barrier(c); }
	{
		LaunchScope ls(c, K_MGFUSED, work);
		if(XH && share > 1){
			// (a cooperative launch only adds the co-residency check, which does not know about the other ranks' kernels)
			k_mg_solve<false,true,false><<<dim3(grid), dim3(MG_BLOCK), smem, c->stream>>>(P);
			PINC_CUDA(cudaGetLastError());
		} else
			PINC_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(MG_BLOCK), args, smem, c->stream));
	}
	PINC_CUDA(cudaMemcpyAsync(c->h_mgHist, c->d_mgHist, 256*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	c->mgHistPending = true;
	c->mgCheckPending = true; c->mgTol = tol; c->mgMaxCycles = maxCycles;
	c->mgXPending = XH != nullptr;
	if(XH && share > 1){ streamSync(c); c->tp->barrier(c); }
	return true;
}

// All-reduce of one double over peer memory (gNeutralizeGrid, residual norm): block-reduce the partial sums, store the
// total into every rank's value slot, publish, wait for every rank, add the slots in rank order (identical bits on
// all ranks).  The operation counter lives on the device and is bumped by the kernel itself: graph-replay safe.
struct AllSumArgs { char *peer[64]; char *mine; int size, rank; };
__global__ void k_allsum_p2p(const double *__restrict__ partial, int n, double *__restrict__ out, AllSumArgs A, int *flags){
	__shared__ double sh[8];
	__shared__ unsigned long long seqS;
	double acc = 0;
	for(int i = threadIdx.x; i < n; i += 256) acc += partial[i];
	for(int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
	if((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
	__syncthreads();
	if(threadIdx.x == 0){
		double t = 0;
		for(int w = 0; w < 8; w++) t += sh[w];
		unsigned long long *ctr = (unsigned long long*)A.mine + 11;
		unsigned long long seq = *ctr + 1; *ctr = seq; seqS = seq;
		sh[0] = t;
	}
	__syncthreads();
	const unsigned long long seq = seqS;
	const int par = (int)(seq & 1);
	if(threadIdx.x < A.size){
		double *slot = (double*)((unsigned long long*)A.peer[threadIdx.x] + 128) + par*64 + A.rank;
		*((volatile double*)slot) = sh[0];
		__threadfence_system();
		*((volatile unsigned long long*)((unsigned long long*)A.peer[threadIdx.x] + 32 + A.rank)) = seq;
	}
	__syncthreads();
	if(threadIdx.x < A.size){
		long long c0 = clock64();
		while(*((volatile unsigned long long*)((unsigned long long*)A.mine + 32 + threadIdx.x)) < seq)
			if(clock64() - c0 > 6000000000LL){ atomicOr(flags, ERR_P2P_TIMEOUT); break; }
		__threadfence_system();
	}
	__syncthreads();
	if(threadIdx.x == 0){
		const volatile double *v = (const volatile double*)((unsigned long long*)A.mine + 128) + par*64;
		double tot = v[0];
		for(int r = 1; r < A.size; r++) tot += v[r];
		out[0] = tot;
	}
}
// replaces k_final_sum + ncclAllReduce when the peer arena exists; false if unavailable
bool allSumP2P(Ctx *c, const double *partial, int n, double *out, const MpiInfo *m){
	P2P *p = c->tp->p2p();
	if(!p || m->mpiSize > 64 || mgMode() != 2) return false;
	AllSumArgs A{};
	for(int r = 0; r < m->mpiSize; r++) A.peer[r] = p->peerArena[r];
	A.mine = p->arena; A.size = m->mpiSize; A.rank = m->mpiRank;
	PINC_LAUNCH(c, K_REDUCE, 8.0*n, (k_allsum_p2p<<<1,256,0,c->stream>>>(partial, n, out, A, c->d_flags)));
	return true;
}

__global__ void k_seq_advance(unsigned long long *base, unsigned long long n){ *base += n; }

// One V-cycle + residual + norm of a multi-rank solve is ~230 small kernels and ~30 NCCL calls that never change from
// cycle to cycle: the second V-cycle of a solver is captured into a CUDA graph and every later one is a graph launch
// (host enqueue cost was the bottleneck: 28 000 launches per time step).  The sequence numbers of the peer-memory
// operations are offsets from a device word that the graph itself advances, so a replay continues the sequence.

static void oneCycle(Ctx *c, funPtr algo, int bottom, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *m, DevGrid *res, DevGrid *rho, DevGrid *phi){
	opCycle(c, algo, 0, bottom, 0, mgRho, mgPhi, mgRes, m);
	opResidual(c, res, rho, phi);
	gridHalo(c, res, m, 0, 0);
	gridSumTrueAll(c, res, 1, 3, m);
}

static void opsSolve(Ctx *c, funPtr mgAlgo, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *m, double tol, int maxCycles){
	int bottom = mgRho->nLevels - 1;
	c->mgHistory.clear();
	c->mgHistPending = false;
	if(mgRho->nLevels > 1){
		DevGrid *res = devGrid(c, mgRes->grids[0]), *rho = devGrid(c, mgRho->grids[0]), *phi = devGrid(c, mgPhi->grids[0]);
		double barRes = 2.;
		int cycles = 0;
		P2P *p = c->tp->p2p();
		static int noGraph = -1;
		if(noGraph < 0) noGraph = getenv("PINC_B200_NO_GRAPH") ? 1 : 0;
		const bool graphable = p && m->mpiSize > 1 && mgMode() == 2 && !noGraph && !c->profOn;
		CycleGraph &G = c->cycleGraphs[mgRho];         // per context: rank threads of one process never share it
		while(barRes > tol && cycles < maxCycles){
			if(graphable && G.state == 1){
				PINC_CUDA(cudaGraphLaunch(G.exec, c->stream));
				p->seq += G.nOps; p->base += G.nOps;
				c->launches += G.launches;
			} else if(graphable && G.state == 0 && (cycles >= 1 || !c->mgHistory.empty() || G.launches < 0)){
				// everything the cycle needs exists after the first (eager) cycle: capture this one
				unsigned long long seq0 = p->seq; long l0 = c->launches;
				cudaGraph_t graph = nullptr;
				cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
				if(e == cudaSuccess){
					oneCycle(c, mgAlgo, bottom, mgRho, mgPhi, mgRes, m, res, rho, phi);
					unsigned long long nOps = p->seq - seq0;
					k_seq_advance<<<1,1,0,c->stream>>>(P2P::flag(p->arena, 10), nOps);
					e = cudaStreamEndCapture(c->stream, &graph);
					if(e == cudaSuccess) e = cudaGraphInstantiate(&G.exec, graph, 0);
					if(graph) cudaGraphDestroy(graph);
					if(e == cudaSuccess){
						G.nOps = nOps; G.launches = c->launches - l0; G.state = 1;
						PINC_CUDA(cudaGraphLaunch(G.exec, c->stream));       // the captured cycle has not run yet
						p->base += nOps;                                      // p->seq already counts its operations
					}
				}
				if(e != cudaSuccess){
					cudaGetLastError();
					fprintf(stderr, "PINC-B200 WARNING: CUDA graph capture of the V-cycle failed (%s); staying with eager launches\n", cudaGetErrorString(e));
					G.state = -1;
					p->seq = seq0;                                        // nothing of the captured cycle ran
					oneCycle(c, mgAlgo, bottom, mgRho, mgPhi, mgRes, m, res, rho, phi);
				}
			} else {
				oneCycle(c, mgAlgo, bottom, mgRho, mgPhi, mgRes, m, res, rho, phi);
			}
			barRes = readScalar(c, 3);
			barRes /= (double)gTotTruesize(mgRho->grids[0], m);
			barRes = sqrt(barRes);
			c->mgHistory.push_back(barRes);
			cycles++;
		}
		c->mgLastBarRes = barRes; c->mgLastCycles = cycles;
		if(p && m->mpiSize > 1) checkDeviceFlags(c, "mgSolve");        // a peer-memory timeout must not reach the pusher
		if(!(barRes <= tol)) fatal("mgSolve: the multigrid solver did not reach barRes <= %g within %d V-cycles (barRes = %g); the reference would loop forever (src/multigrid.c:1697)", tol, cycles, barRes);
	} else {
		DevGrid *rho = devGrid(c, mgRho->grids[0]), *phi = devGrid(c, mgPhi->grids[0]);
		for(int cyc = 0; cyc < mgRho->nMGCycles; cyc++){
			gridHalo(c, rho, m, 0, 0);
			gridNeutralize(c, rho, m);
			opSmooth(c, (SmoothFn)mgRho->coarseSolv, phi, rho, mgRho->nCoarseSolve, m);
		}
	}
}

// =================================================================================================
// multi-rank solves by replication
// =================================================================================================
// The distributed solve is bound by ~180 cross-GPU round trips per V-cycle (6.9 us each over NVLink against 1.5 us for a
// half-sweep inside one GPU) and the grids are tiny, so for every multi-rank configuration that fits one GPU it is faster
// to NOT distribute it: every rank gathers rho and all levels of phi of all ranks (one grouped exchange of a few MB over
// NVLink), solves the GLOBAL problem redundantly with the single-GPU persistent kernel, and copies its own sub-domain of
// every level (ghost layers included) back into the rank-local grids, which are left exactly as the distributed
// algorithm leaves them (same arithmetic per node; only gBnd's and the norm's sums are taken in another order).
struct GlobalMg {
	Grid *rho = nullptr, *phi = nullptr;
	MultigridSolver *solver = nullptr;
	double *d_pack = nullptr;
	long seg = 0;                   // doubles per rank in d_pack: rho(0) then phi(0..L-1), true nodes, x fastest
	int nLevels = 0, t[3] = {0,0,0}, ns[3] = {0,0,0};
};
__global__ void k_rep_pack(const double *__restrict__ g, int s0, int s1, int t0, int t1, int t2, double *__restrict__ seg){
	long nt = (long)t0*t1*t2, st = (long)gridDim.x*blockDim.x;
	for(long i = blockIdx.x*(long)blockDim.x + threadIdx.x; i < nt; i += st){ int j,k,l; truePoint(i,t0,t1,j,k,l); seg[i] = g[ix(j,k,l,s0,s1)]; }
}
// global true node <- the owning rank's segment
__global__ void k_rep_unpack(double *__restrict__ G, int S0, int S1, const double *__restrict__ pack, long segStride, long segOff,
		int t0, int t1, int t2, int nsx, int nsy, int nsz){
	int T0 = t0*nsx, T1 = t1*nsy; long nT = (long)T0*T1*t2*nsz, st = (long)gridDim.x*blockDim.x;
	for(long i = blockIdx.x*(long)blockDim.x + threadIdx.x; i < nT; i += st){
		int J,K,Lz; truePoint(i,T0,T1,J,K,Lz);
		int bx = (J-1)/t0, by = (K-1)/t1, bz = (Lz-1)/t2;
		int j = (J-1) - bx*t0, k = (K-1) - by*t1, l = (Lz-1) - bz*t2;
		long r = bx + (long)nsx*(by + (long)nsy*bz);
		G[ix(J,K,Lz,S0,S1)] = pack[r*segStride + segOff + j + (long)t0*(k + (long)t1*l)];
	}
}
// rank-local grid, ghost layers included <- the global grid at the periodic image of (offset + local index)
__global__ void k_rep_extract(double *__restrict__ g, int s0, int s1, int s2, const double *__restrict__ G, int S0, int S1, int T0, int T1, int T2,
		int ox, int oy, int oz){
	long n = (long)s0*s1*s2, st = (long)gridDim.x*blockDim.x;
	for(long i = blockIdx.x*(long)blockDim.x + threadIdx.x; i < n; i += st){
		int j = (int)(i % s0); long r = i / s0; int k = (int)(r % s1); int l = (int)(r / s1);
		int J = ox + j, K = oy + k, Lz = oz + l;
		J = J < 1 ? J + T0 : (J > T0 ? J - T0 : J); K = K < 1 ? K + T1 : (K > T1 ? K - T1 : K); Lz = Lz < 1 ? Lz + T2 : (Lz > T2 ? Lz - T2 : Lz);
		g[i] = G[ix(J,K,Lz,S0,S1)];
	}
}
static void freeGlobalMg(GlobalMg *G){
	if(!G) return;
	if(G->d_pack) cudaFree(G->d_pack);
	if(G->solver) mgFreeSolver(G->solver);
	if(G->rho) pincGridFree(G->rho);
	if(G->phi) pincGridFree(G->phi);
	delete G;
}
static bool replicaEligible(Ctx *c, Multigrid *mgRho, const MpiInfo *m){
	if(g_mgReplica < 0) g_mgReplica = (getenv("PINC_B200_MG_REPLICA") && atoi(getenv("PINC_B200_MG_REPLICA")) == 0) ? 0 : 1;
	if(!g_mgReplica || mgMode() != 2 || m->mpiSize < 2 || !c->tp) return false;
	int nL = mgRho->nLevels;
	if(nL < 2 || nL > MG_MAXLEV) return false;
	long nGlobal = 1;
	for(int q = 0; q < nL; q++){
		const Grid *g = mgRho->grids[q];
		if(g->rank != 4 || g->size[0] != 1) return false;
		for(int d = 1; d < 4; d++){
			if(g->trueSize[d] < 1 || g->nGhostLayers[d] != 1 || g->nGhostLayers[d+g->rank] != 1) return false;
			if(g->bnd[d] != PERIODIC || g->bnd[d+g->rank] != PERIODIC) return false;
			if(((long)g->trueSize[d]*m->nSubdomains[d-1]) & 1) return false;
			if(q == 0) nGlobal *= (long)g->trueSize[d]*m->nSubdomains[d-1];
		}
	}
	return nGlobal <= (1L << 24);          // 16 M nodes = 128 MB per array: far below what one GPU holds, and still L2-friendly
}
static void solveSingle(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, double tol, int maxCycles);
// the replicated global hierarchy of a solver (created on first use, rebuilt when the shapes change)
static GlobalMg *globalMgFor(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, const MpiInfo *m){
	const int nL = mgRho->nLevels, R = m->mpiSize;
	const Grid *g0 = mgRho->grids[0];
	GlobalMg *G = nullptr;
	auto it = c->mgGlobal.find(mgRho);
	if(it != c->mgGlobal.end()){
		G = (GlobalMg*)it->second;
		bool same = G->nLevels == nL;
		for(int d = 0; d < 3; d++) same = same && G->t[d] == g0->trueSize[d+1] && G->ns[d] == m->nSubdomains[d];
		if(!same){ c->mgGlobal.erase(it); freeGlobalMg(G); G = nullptr; }
	}
	if(!G){
		G = new GlobalMg();
		G->nLevels = nL;
		int ts[3], gl[6] = {1,1,1,1,1,1}, bnd[6] = {PERIODIC,PERIODIC,PERIODIC,PERIODIC,PERIODIC,PERIODIC};
		for(int d = 0; d < 3; d++){ G->t[d] = g0->trueSize[d+1]; G->ns[d] = m->nSubdomains[d]; ts[d] = G->t[d]*G->ns[d]; }
		G->rho = pincGridAlloc(3, ts, gl, 1, bnd);
		G->phi = pincGridAlloc(3, ts, gl, 1, bnd);
		G->solver = pincMgAllocSolver(G->rho, G->phi, nL, mgRho->nMGCycles, mgRho->nPreSmooth, mgRho->nPostSmooth, mgRho->nCoarseSolve);
		long seg = trueCount(devGrid(c, mgRho->grids[0]));
		for(int q = 0; q < nL; q++) seg += trueCount(devGrid(c, mgPhi->grids[q]));
		G->seg = seg;
		PINC_CUDA(cudaMalloc(&G->d_pack, (size_t)seg*R*sizeof(double)));
		c->mgGlobal[mgRho] = G;
	}
	return G;
}
// every rank's true nodes of rho(0) (if withRho) and of phi(qFirst .. L-1) into the global hierarchy of every rank: pack,
// ONE grouped exchange, unpack
static void gatherGlobal(Ctx *c, GlobalMg *G, Multigrid *mgRho, Multigrid *mgPhi, const MpiInfo *m, bool withRho, int qFirst){
	const int nL = mgRho->nLevels, R = m->mpiSize, me = m->mpiRank;
	Multigrid *gRho = G->solver->mgRho, *gPhi = G->solver->mgPhi;
	double *mine = G->d_pack + (long)me*G->seg;
	long off = 0;
	std::vector<long> segOff(nL + 1, -1);
	for(int q = -1; q < nL; q++){
		if(q < 0 ? !withRho : q < qFirst) continue;
		DevGrid *g = devGrid(c, q < 0 ? mgRho->grids[0] : mgPhi->grids[q]);
		long nt = trueCount(g);
		segOff[q+1] = off;
		PINC_LAUNCH(c, K_GRIDOP, 16.0*nt, (k_rep_pack<<<tGrid(c,nt),256,0,c->stream>>>(g->d, g->size[0], g->size[1], g->tsize[0], g->tsize[1], g->tsize[2], mine + off)));
		off += nt;
	}
	if(!off) return;
	std::vector<Msg> sends, recvs;
	for(int r = 0; r < R; r++){
		if(r == me) continue;
		sends.push_back(Msg{ r, 7100, (void*)mine, (size_t)off*sizeof(double) });
		recvs.push_back(Msg{ r, 7100, (void*)(G->d_pack + (long)r*G->seg), (size_t)off*sizeof(double) });
	}
	c->tp->exchange(c, sends, recvs);
	for(int q = -1; q < nL; q++){
		if(segOff[q+1] < 0) continue;
		DevGrid *loc = devGrid(c, q < 0 ? mgRho->grids[0] : mgPhi->grids[q]);
		DevGrid *glo = devGrid(c, q < 0 ? gRho->grids[0] : gPhi->grids[q]);
		long nT = trueCount(glo);
		PINC_LAUNCH(c, K_GRIDOP, 16.0*nT, (k_rep_unpack<<<tGrid(c,nT),256,0,c->stream>>>(glo->d, glo->size[0], glo->size[1], G->d_pack, G->seg, segOff[q+1],
			loc->tsize[0], loc->tsize[1], loc->tsize[2], G->ns[0], G->ns[1], G->ns[2])));
	}
}
// my sub-domain of levels qFirst .. L-1 of the global phi, rho and res, ghost layers included
static void extractLocal(Ctx *c, GlobalMg *G, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *m, int qFirst){
	Multigrid *gRho = G->solver->mgRho, *gPhi = G->solver->mgPhi, *gRes = G->solver->mgRes;
	for(int q = qFirst; q < mgRho->nLevels; q++){
		Multigrid *locs[3] = {mgPhi, mgRho, mgRes}, *glos[3] = {gPhi, gRho, gRes};
		for(int a = 0; a < 3; a++){
			DevGrid *loc = devGrid(c, locs[a]->grids[q]), *glo = devGrid(c, glos[a]->grids[q]);
			PINC_LAUNCH(c, K_GRIDOP, 16.0*loc->n, (k_rep_extract<<<tGrid(c,loc->n),256,0,c->stream>>>(loc->d, loc->size[0], loc->size[1], loc->size[2],
				glo->d, glo->size[0], glo->size[1], glo->tsize[0], glo->tsize[1], glo->tsize[2],
				m->subdomain[0]*loc->tsize[0], m->subdomain[1]*loc->tsize[1], m->subdomain[2]*loc->tsize[2])));
		}
	}
}
static void replicaSolve(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *m, double tol, int maxCycles){
	GlobalMg *G = globalMgFor(c, mgRho, mgPhi, m);
	gatherGlobal(c, G, mgRho, mgPhi, m, true, 0);
	// the global problem, redundantly on every rank
	solveSingle(c, G->solver->mgRho, G->solver->mgPhi, G->solver->mgRes, tol, maxCycles);
	const int path = c->mgLastPath + 4;
	extractLocal(c, G, mgRho, mgPhi, mgRes, m, 0);
	c->mgLastPath = path;
}

// =================================================================================================
// multi-rank solves, hybrid: the finest level distributed, the coarser ones replicated
// =================================================================================================
// Replicating the whole solve makes every GPU sweep the whole global finest level, which no longer fits the SMs' shared
// memory (8 ranks: 128^3, 944 us per V-cycle).  Here every rank keeps what it owns anyway - its sub-domain of the finest
// level - and smooths it with the block-resident smoother of the single-GPU kernel; the faces of blocks on a sub-domain
// boundary go into the NEIGHBOUR RANK'S mailboxes over NVLink with the same tagged 16-byte stores as inside one GPU
// (st.relaxed.sys, the data is its own flag: no fence, no flag word, no NCCL call).  The restricted residual is gathered
// into every rank's copy of the next level (same slots), and from there down the hierarchy is the replicated global one:
// those levels are latency-bound, so solving them redundantly costs nothing, and the prolongation back reads the local
// copy.  gBnd's means and the residual norm are sums over all ranks, added in rank order (identical bits everywhere).
// One persistent kernel per rank and solve (k_mg_solve<false,true>); the ranks only meet in those slots.
int g_mgHybrid = -1;          // 1 (default): hybrid whenever it applies; 0: replicate the whole solve; from $PINC_B200_MG_HYBRID
static bool hybridSolve(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *m, double tol, int maxCycles){
	if(g_mgHybrid < 0) g_mgHybrid = (getenv("PINC_B200_MG_HYBRID") && atoi(getenv("PINC_B200_MG_HYBRID")) == 0) ? 0 : 1;
	if(!g_mgHybrid || m->mpiSize > XD_MAXR) return false;
	const int nL = mgRho->nLevels;
	const Grid *g0 = mgRho->grids[0];
	for(int d = 1; d < 4; d++) if(g0->trueSize[d] < 4 || (g0->trueSize[d] & 3)) return false;      // blocks with even edges on level 0, even level 1
	if(c->mgXPending){ streamSync(c); c->mgXPending = false; }
	if(c->mgXTagHigh > 3.0e9){
		// the 32-bit tags are about to wrap (identical on every rank): back to the initial state, collectively
		XArena *A = (XArena*)c->mgXArena;
		c->tp->barrier(c);
		if(A){ PINC_CUDA(cudaMemsetAsync(A->mine, 0, A->bytes, c->stream)); streamSync(c); }
		c->tp->barrier(c);
		c->mgXTagHigh = 0;
	}
	XDist X{};
	X.qDist = 1; X.R = m->mpiSize; X.me = m->mpiRank;
	for(int d = 0; d < 3; d++){ X.ns[d] = m->nSubdomains[d]; X.sub[d] = m->subdomain[d]; }
	for(int d = 0; d < 3; d++){ X.nbr[2*d] = dimNb(m, d, -1); X.nbr[2*d+1] = dimNb(m, d, +1); }
	GlobalMg *G = globalMgFor(c, mgRho, mgPhi, m);
	// level 0: this rank's own grids; levels >= 1: the global hierarchy
	std::vector<Grid*> gr(nL), gp(nL), ge(nL);
	for(int q = 0; q < nL; q++){
		gr[q] = q < X.qDist ? mgRho->grids[q] : G->solver->mgRho->grids[q];
		gp[q] = q < X.qDist ? mgPhi->grids[q] : G->solver->mgPhi->grids[q];
		ge[q] = q < X.qDist ? mgRes->grids[q] : G->solver->mgRes->grids[q];
	}
	Multigrid xr = *mgRho, xp = *mgPhi, xe = *mgRes;
	xr.grids = gr.data(); xp.grids = gp.data(); xe.grids = ge.data();
	gatherGlobal(c, G, mgRho, mgPhi, m, false, X.qDist);        // the coarse levels keep their phi from solve to solve (src/multigrid.c:1496-1548)
	if(!fusedSolve(c, &xr, &xp, &xe, tol, maxCycles, 0, &X)) return false;
	extractLocal(c, G, mgRho, mgPhi, mgRes, m, X.qDist);
	for(int q = 0; q < X.qDist; q++){
		gridHalo(c, devGrid(c, mgPhi->grids[q]), m, 0, 0);
		gridHalo(c, devGrid(c, mgRho->grids[q]), m, 0, 0);
		gridHalo(c, devGrid(c, mgRes->grids[q]), m, 0, 0);
	}
	c->mgLastPath = 1 + 8;
	return true;
}
void mgFreeArena(Ctx *c){ freeXArena(c); }

// which single-GPU kernel runs a periodic solve (the caller has checked fusedEligible's shape conditions)
static void solveSingle(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, double tol, int maxCycles){
	// the cluster kernel runs on 16 SMs: it wins while the finest level is small enough to be latency-bound
	// (measured: 32^3 304 vs 526 us per V-cycle, 64^3 908 vs 805); larger grids go to the all-SM kernel
	const Grid *g0 = mgRho->grids[0];
	long nt0 = (long)g0->trueSize[1]*g0->trueSize[2]*g0->trueSize[3];
	// ... unless its small levels are the cubic 16/8/4 pyramid, which the all-SM kernel runs through specialised
	// routines (measured us per V-cycle, all-SM vs cluster: 16^3 65 vs 123, 32^3 148 vs 288)
	bool pyramid = g_mgMode == 2;
	for(int q = 0; q < mgRho->nLevels; q++){
		const Grid *g = mgRho->grids[q];
		long nt = (long)g->trueSize[1]*g->trueSize[2]*g->trueSize[3];
		if(nt > MG_SMALL) continue;
		int n = g->trueSize[1];
		if(g->trueSize[2] != n || g->trueSize[3] != n || (n != 16 && n != 8 && n != 4) || (n == 4 && q != mgRho->nLevels-1)) pyramid = false;
	}
	bool preferCluster = (g_mgMode == 3 || g_mgForceCluster || (nt0 <= 65536 && !pyramid)) && !g_mgNoCluster;
	if(g_mgMode >= 2 && preferCluster && clusterSolve(c, mgRho, mgPhi, mgRes, tol, maxCycles, g_mgMode == 3)){ c->mgLastPath = 2; return; }
	fusedSolve(c, mgRho, mgPhi, mgRes, tol, maxCycles, g_mgMode == 1 || g_mgMode == 3);
	c->mgLastPath = 1;
}

// The reference's tolerance loop has no bound (src/multigrid.c:1697): a solve that stalls hangs the run.  Here the loop
// is bounded ($PINC_B200_MG_MAXCYCLES, default 10000 V-cycles) and a solve that ends above the tolerance is a fatal
// error on every path, raised at the first stream synchronisation after the solve (the persistent kernels report
// through the history buffer, which lands with the solve) - an unconverged E never reaches the pusher unnoticed.
int mgMaxCyclesDefault(){
	static int v = 0;
	if(!v){ const char *e = getenv("PINC_B200_MG_MAXCYCLES"); v = e && atoi(e) > 0 ? atoi(e) : 10000; }
	return v;
}
void mgConvergenceCheck(Ctx *c){
	c->mgCheckPending = false;
	if(!c->h_mgHist) return;
	c->mgLastCycles = (int)c->h_mgHist[0];
	c->mgLastBarRes = c->h_mgHist[251];
	if(c->mgXPending){ c->mgXTagHigh = c->h_mgHist[252]; c->mgXPending = false; }
	if(!(c->mgLastBarRes <= c->mgTol))
		fatal("mgSolve: the multigrid solver did not reach barRes <= %g within %d V-cycles (barRes = %g); the reference would loop forever (src/multigrid.c:1697)",
			c->mgTol, c->mgLastCycles, c->mgLastBarRes);
}

void mgForgetPlans(Ctx *c){
	for(auto &kv : c->cycleGraphs) if(kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
	c->cycleGraphs.clear();
	if(c->mgGlobalBusy) return;                      // freeing a replica's own grids comes back here
	c->mgGlobalBusy = true;
	std::unordered_map<const void*, void*> gone;
	gone.swap(c->mgGlobal);
	for(auto &kv : gone) freeGlobalMg((GlobalMg*)kv.second);
	c->mgGlobalBusy = false;
}

} // namespace pinc

using namespace pinc;

extern "C" {

void mgGS3D(Grid *phi, const Grid *rho, int nCycles, const MpiInfo *mpiInfo){
	Ctx *c = cur(); opGS(c, devGrid(c, phi), devGrid(c, rho), nCycles, mpiInfo);
}
void mgResidual(Grid *res, const Grid *rho, const Grid *phi, const MpiInfo *mpiInfo){
	(void)mpiInfo; Ctx *c = cur(); opResidual(c, devGrid(c, res), devGrid(c, rho), devGrid(c, phi));
}
void mgHalfRestrict3D(const Grid *fine, Grid *coarse){ Ctx *c = cur(); opRestrict(c, devGrid(c, fine), devGrid(c, coarse)); }
void mgBilinProl3D(Grid *fine, const Grid *coarse, const MpiInfo *mpiInfo){ Ctx *c = cur(); opProlong(c, devGrid(c, fine), devGrid(c, coarse), mpiInfo); }
double mgSumTrueSquared(Grid *error, const MpiInfo *mpiInfo){
	Ctx *c = cur();
	gridSumTrue(c, devGrid(c, error), 1, nullptr, 3);
	if(mpiInfo->mpiSize > 1) c->tp->allreduceSum(c, c->d_scal + 3, 1);
	return readScalar(c, 3);
}

static void checkPlugins(const Multigrid *mg){
	if(mg->nLevels < 1) fatal("multigrid: nLevels < 1");
	auto okS = [](const void *f){ return !f || f == (const void*)mgGS3D || f == (const void*)mgJacob3D; };
	if(!okS((const void*)mg->coarseSolv) || !okS((const void*)mg->preSmooth) || !okS((const void*)mg->postSmooth))
		fatal("multigrid: the smoothers provided are gaussSeidelRB (mgGS3D) and jacobian (mgJacob3D)");
	if((mg->restrictor && mg->restrictor != mgHalfRestrict3D) || (mg->prolongator && mg->prolongator != mgBilinProl3D))
		fatal("multigrid: only halfWeight restriction and bilinear prolongation are implemented");
}
// the persistent kernels implement exactly one configuration: mgVRecursive with red-black Gauss-Seidel everywhere
static bool defaultPlugins(funPtr algo, const Multigrid *mg){
	auto gs = [](const void *f){ return !f || f == (const void*)mgGS3D; };
	return (!algo || algo == (funPtr)mgVRecursive) && gs((const void*)mg->coarseSolv) && gs((const void*)mg->preSmooth) && gs((const void*)mg->postSmooth);
}

// src/multigrid.c:1314-1379: boundary values of every coarser level := every second value of the finer level's slices
// (host arrays; the reference defines this function and never calls it - its coarse bndSlice arrays stay uninitialised,
// so a host that wants non-periodic multigrid calls it after gSetBndSlices on the finest level)
void mgRestrictBnd(Multigrid *mg){
	Ctx *c = curOrNull();                        // host arrays only; the device mirrors follow where they exist
	for(int lvl = 0; lvl < mg->nLevels-1; lvl++){
		Grid *f = mg->grids[lvl], *g = mg->grids[lvl+1];
		const int rank = f->rank;
		if(!f->bndSlice || !g->bndSlice) fatal("mgRestrictBnd: level %d has no bndSlice", lvl);
		long nF = 0, nC = 0;
		for(int d = 0; d < rank; d++){
			long a = 1, b = 1;
			for(int dd = 0; dd < rank; dd++) if(dd != d){ a *= f->size[dd]; b *= g->size[dd]; }
			if(a > nF) nF = a;
			if(b > nC) nC = b;
		}
		for(int d = 1; d < 2*rank; d++){
			if(d == rank) continue;
			for(long s = 0; s < nC; s++) g->bndSlice[s + nC*d] = f->bndSlice[2*s + nF*d];
		}
		if(!c) continue;
		auto it = c->grids.find(g);
		if(it != c->grids.end() && it->second->nonPeriodic) gridUploadBnd(c, it->second);
	}
}

void mgJacob3D(Grid *phi, const Grid *rho, const int nCycles, const MpiInfo *mpiInfo){
	Ctx *c = cur(); opJacobi(c, devGrid(c, phi), devGrid(c, rho), nCycles, mpiInfo);
}
void mgVRecursive(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo){
	checkPlugins(mgRho);
	opVCycle(cur(), level, bottom, top, mgRho, mgPhi, mgRes, mpiInfo);
}
void mgVRegular(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo){
	checkPlugins(mgRho);
	opVRegular(cur(), level, bottom, top, mgRho, mgPhi, mgRes, mpiInfo);
}
void mgW(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo){
	(void)level; (void)top;
	checkPlugins(mgRho);
	opCycle(cur(), (funPtr)mgW, 0, bottom, 0, mgRho, mgPhi, mgRes, mpiInfo);
}
void mgFMG(int level, int bottom, int top, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo){
	opCycle(cur(), (funPtr)mgFMG, level, bottom, top, mgRho, mgPhi, mgRes, mpiInfo);
}

void mgSolveRaw(funPtr mgAlgo, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, const MpiInfo *mpiInfo){
	checkPlugins(mgRho);
	Ctx *c = cur();
	const double tol = 1.E-10;                       // src/multigrid.c:1695
	const int maxCycles = pinc::mgMaxCyclesDefault();    // the reference has no bound (it would hang); this one fails loudly, see mgConvergenceCheck
	const bool dflt = defaultPlugins(mgAlgo, mgRho);
	if(dflt && fusedEligible(c, mgRho, mgPhi, mgRes, mpiInfo)) pinc::solveSingle(c, mgRho, mgPhi, mgRes, tol, maxCycles);
	else if(dflt && pinc::replicaEligible(c, mgRho, mpiInfo)){
		if(!pinc::hybridSolve(c, mgRho, mgPhi, mgRes, mpiInfo, tol, maxCycles)) pinc::replicaSolve(c, mgRho, mgPhi, mgRes, mpiInfo, tol, maxCycles);
	}
	else { opsSolve(c, mgAlgo, mgRho, mgPhi, mgRes, mpiInfo, tol, maxCycles); c->mgLastPath = 0; }
}

void mgSolve(const MultigridSolver *solver, const Grid *rho, const Grid *phi, const MpiInfo *mpiInfo){
	(void)rho; (void)phi;                            // as the reference: the grids captured at allocation are used
	mgSolveRaw(solver->mgAlgo, solver->mgRho, solver->mgPhi, solver->mgRes, mpiInfo);
}

void mgSolver(void (**solve)(), MultigridSolver *(**solverAlloc)(), void (**solverFree)()){
	*solve = (void(*)())mgSolve;
	*solverAlloc = (MultigridSolver*(*)())mgAllocSolver;          // the reference's signature (ini, rho, phi): inihost.cpp
	*solverFree = (void(*)())mgFreeSolver;
}

void pincMgSetMode(int mode){ pinc::g_mgForceCluster = mode == 4; pinc::g_mgNoCluster = mode == 5; if(mode == 4 || mode == 5) mode = 2; pinc::g_mgMode = mode < 0 ? 0 : (mode > 3 ? 3 : mode); }

int pincMgLastPath(void){ return cur()->mgLastPath; }
int pincMgLastHistory(double *barRes, int cap){
	Ctx *c = cur();
	if(c->mgHistPending){
		streamSync(c);
		int n = (int)c->h_mgHist[0];
		c->mgHistory.assign(c->h_mgHist + 1, c->h_mgHist + 1 + (n < 250 ? n : 250));
		c->mgHistPending = false;
		c->mgLastCycles = n;
	}
	int n = (int)c->mgHistory.size();
	for(int i = 0; i < n && i < cap; i++) barRes[i] = c->mgHistory[i];
	return c->mgLastCycles > n ? c->mgLastCycles : n;      // more than 250 V-cycles: the count is exact, the history holds the first 250
}
double pincMgLastBarRes(void){ Ctx *c = cur(); if(c->mgHistPending || c->mgCheckPending) streamSync(c); return c->mgLastBarRes; }
void pincMgSetReplica(int on){ pinc::g_mgReplica = on ? 1 : 0; }
void pincMgSetHybrid(int on){ pinc::g_mgHybrid = on ? 1 : 0; }
void pincMgSetRowMode(int mode){ pinc::g_mgRowMode = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }

} // extern "C"
