// initial.cu — initial conditions on the device: pPosLattice, pPosUniform, pPosPerturb, pVelMaxwell,
// pVelZero of src/population.c:110-276, 367-428 with plain arguments (the reference reads them from the ini).
//
// As in the reference every rank walks ALL global particles and keeps the ones of its own sub-domain
// (population.c:134-151, 196-215), so the particle set does not depend on the decomposition.  The random numbers come
// from Philox4x32-10 keyed by (seed, species) with the GLOBAL particle index as counter (GSL's MT19937 + ziggurat of
// the reference cannot be reproduced without GSL, and a counter-based generator needs no sequential state);
// velocities are keyed by (seed, rank, species) and the local index, as the reference's stream is rank-local too.
// Host arrays are stale afterwards (pincSyncPopToHost), like after any other entry point.
#include "common.h"

namespace pinc {

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4]){
	#pragma unroll
	for(int r = 0; r < 10; r++){
		unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u*c0;
		unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u*c2;
		unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// uniform in (0,1) from 53 random bits (gsl_rng_uniform_pos excludes 0)
__device__ __forceinline__ double u01(unsigned hi, unsigned lo){
	return ((double)(hi >> 5)*67108864.0 + (double)(lo >> 6) + 0.5)*(1.0/9007199254740992.0);
}

struct IcGeom { double L[3]; double invTs[3]; int sub[3]; int off[3]; };

// mode 0: lattice (population.c:191-207), mode 1: uniform (:134-141).  Keeps the particles of this sub-domain.
__global__ void k_ic_positions(double *__restrict__ P, long cap, long a, long capS, long n, int mode, double l, IcGeom G,
		unsigned long long seed, int species, unsigned long long *count){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st){
		double pos[3];
		if(mode == 0){
			double lin = l*(double)i;
			for(int d = 0; d < 3; d++){ pos[d] = fmod(lin, G.L[d]); lin /= G.L[d]; }
		} else {
			unsigned r[4], q[4];
			philox4x32_10((unsigned)i, (unsigned)((unsigned long long)i >> 32), 0u, 0u, (unsigned)seed ^ (0x9E37u*(unsigned)species), (unsigned)(seed >> 32) + 1u, r);
			philox4x32_10((unsigned)i, (unsigned)((unsigned long long)i >> 32), 1u, 0u, (unsigned)seed ^ (0x9E37u*(unsigned)species), (unsigned)(seed >> 32) + 1u, q);
			pos[0] = G.L[0]*u01(r[0], r[1]); pos[1] = G.L[1]*u01(r[2], r[3]); pos[2] = G.L[2]*u01(q[0], q[1]);
		}
		bool mine = true;
		for(int d = 0; d < 3; d++) mine = mine && (G.sub[d] == (int)(G.invTs[d]*pos[d]));
		if(!mine) continue;
		unsigned long long slot = atomicAdd(count, 1ULL);
		if((long)slot >= capS) continue;                                   // reported by the host from the final count
		long q = a + (long)slot;
		for(int d = 0; d < 3; d++) P[q + d*cap] = pos[d] - (double)G.off[d];      // pToLocalFrame (population.c:727)
	}
}
// pPosPerturb (population.c:242-276): global frame, x += A cos(2 pi m x / L), local frame
__global__ void k_ic_perturb(double *__restrict__ P, long cap, long a, long n, double A0, double A1, double A2, double m0, double m1, double m2, IcGeom G){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	const double A[3] = {A0, A1, A2}, M[3] = {m0, m1, m2};
	const double pi = 3.14159265358979323846;
	for(; i < n; i += st)
		#pragma unroll
		for(int d = 0; d < 3; d++){
			double x = P[a + i + d*cap];
			x += (double)G.off[d];
			double theta = 2.0*pi*M[d]*x/G.L[d];
			x += A[d]*cos(theta);
			x -= (double)G.off[d];
			P[a + i + d*cap] = x;
		}
}
// pVelMaxwell (population.c:367-392): drift + thermal * N(0,1) per component (Box-Muller on Philox)
__global__ void k_ic_maxwell(double *__restrict__ P, long cap, long a, long n, double drift, double vth, unsigned long long seed, int rank, int species){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	const double twoPi = 6.28318530717958647692;
	for(; i < n; i += st){
		unsigned r[4], q[4];
		unsigned k0 = (unsigned)seed ^ (0x85EBu*(unsigned)species), k1 = (unsigned)(seed >> 32) ^ (0xC2B2u*(unsigned)(rank + 1));
		philox4x32_10((unsigned)i, (unsigned)((unsigned long long)i >> 32), 2u, 0u, k0, k1, r);
		philox4x32_10((unsigned)i, (unsigned)((unsigned long long)i >> 32), 3u, 0u, k0, k1, q);
		double ra = sqrt(-2.0*log(u01(r[0], r[1]))), ta = twoPi*u01(r[2], r[3]);
		double rb = sqrt(-2.0*log(u01(q[0], q[1]))), tb = twoPi*u01(q[2], q[3]);
		P[a + i + 3*cap] = drift + vth*(ra*cos(ta));
		P[a + i + 4*cap] = drift + vth*(ra*sin(ta));
		P[a + i + 5*cap] = drift + vth*(rb*cos(tb));
	}
}

static IcGeom geomOf(const MpiInfo *m, const int *trueSize){
	IcGeom G;
	for(int d = 0; d < 3; d++){
		G.L[d] = (double)(m->nSubdomains[d]*trueSize[d]);
		G.invTs[d] = m->posToSubdomain[d];
		G.sub[d] = m->subdomain[d];
		G.off[d] = m->offset[d];
	}
	return G;
}
static void resetOrder(DevPop *dp){
	for(int s = 0; s < dp->nS; s++) dp->sortedN[s] = 0;
	dp->keysValid = false; dp->extracted = false;
	if(dp->predep){ dp->predep->fixDirty = true; dp->predep = nullptr; }
}
static void generatePositions(Population *pop, const MpiInfo *m, const long int *nParticles, const int *trueSize, int mode, unsigned long long seed){
	Ctx *c = cur();
	DevPop *dp = devPop(c, pop, false);
	IcGeom G = geomOf(m, trueSize);
	double V = G.L[0]*G.L[1]*G.L[2];
	unsigned long long *count = (unsigned long long*)c->d_long;
	PINC_CUDA(cudaMemsetAsync(count, 0, 8*sizeof(unsigned long long), c->stream));
	for(int s = 0; s < dp->nS; s++){
		long n = nParticles[s];
		double l = mode == 0 ? pow(V/(double)n, 1.0/3.0) : 0.0;
		if(n > 0) PINC_LAUNCH(c, K_LAYOUT, 24.0*n, (k_ic_positions<<<gridFor(n,256,c->numSMs*8),256,0,c->stream>>>(dp->base, dp->cap, pop->iStart[s],
			pop->iStart[s+1]-pop->iStart[s], n, mode, l, G, seed, s, count + s)));
	}
	PINC_CUDA(cudaMemcpyAsync(c->h_long, count, 8*sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
	streamSync(c);
	for(int s = 0; s < dp->nS; s++){
		long made = c->h_long[s], room = pop->iStart[s+1] - pop->iStart[s];
		if(made > room) fatal("allocated only %ld particles of specie %i per node but %ld generated", room, s, made);   // population.c:155-160
		pop->iStop[s] = pop->iStart[s] + made;
	}
	resetOrder(dp);
}

} // namespace pinc

using namespace pinc;

extern "C" {

void pincPosLattice(Population *pop, const MpiInfo *mpiInfo, const long int *nParticles, const int *trueSize){
	generatePositions(pop, mpiInfo, nParticles, trueSize, 0, 0);
}
void pincPosUniform(Population *pop, const MpiInfo *mpiInfo, const long int *nParticles, const int *trueSize, unsigned long long seed){
	generatePositions(pop, mpiInfo, nParticles, trueSize, 1, seed);
}
void pincPosPerturb(Population *pop, const MpiInfo *mpiInfo, const double *amplitude, const double *mode, const int *trueSize){
	Ctx *c = cur(); DevPop *dp = devPop(c, pop);
	IcGeom G = geomOf(mpiInfo, trueSize);
	for(int s = 0; s < dp->nS; s++){
		long a = pop->iStart[s], n = pop->iStop[s] - a;
		if(n > 0) PINC_LAUNCH(c, K_LAYOUT, 48.0*n, (k_ic_perturb<<<gridFor(n,256,c->numSMs*8),256,0,c->stream>>>(dp->base, dp->cap, a, n,
			amplitude[3*s], amplitude[3*s+1], amplitude[3*s+2], mode[3*s], mode[3*s+1], mode[3*s+2], G)));
	}
	resetOrder(dp);
}
void pincVelMaxwell(Population *pop, const MpiInfo *mpiInfo, const double *drift, const double *thermalVelocity, unsigned long long seed){
	Ctx *c = cur(); DevPop *dp = devPop(c, pop);
	for(int s = 0; s < dp->nS; s++){
		long a = pop->iStart[s], n = pop->iStop[s] - a;
		if(n > 0) PINC_LAUNCH(c, K_LAYOUT, 24.0*n, (k_ic_maxwell<<<gridFor(n,256,c->numSMs*8),256,0,c->stream>>>(dp->base, dp->cap, a, n,
			drift[s], thermalVelocity[s], seed, mpiInfo->mpiRank, s)));
	}
}
void pincVelZero(Population *pop){
	Ctx *c = cur(); DevPop *dp = devPop(c, pop);
	for(int s = 0; s < dp->nS; s++){
		long a = pop->iStart[s], n = pop->iStop[s] - a;
		for(int d = 3; d < 6 && n > 0; d++) PINC_CUDA(cudaMemsetAsync(dp->base + (size_t)d*dp->cap + a, 0, (size_t)n*sizeof(double), c->stream));
	}
}

} // extern "C"
