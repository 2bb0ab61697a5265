// mgsmall.cuh — mgGS3D (src/multigrid.c:683-767) for the cubic small levels of the benchmark pyramid (16^3, 8^3, 4^3)
// inside one CTA of 512 threads, gBnd batched (mode 2).  Included by mgsmem.cuh; same arithmetic per node as
// cGSSmall (x+ + x- + y+ + y- + z+ + z- + rho, times 1/6), hence the same bits.
//
//  16^3  every thread keeps a pencil of eight x-consecutive nodes (and their rho) in registers for the whole call:
//        lane = 2*row + half, warp = z-plane.  x-neighbours are the thread's own registers (the pencil's end comes
//        from lane^1), y-neighbours are lanes +-2 of the same warp (16 rows = the whole periodic extent, so the lane
//        rotation IS the periodic wrap), z-neighbours go through a lane-contiguous exchange buffer in shared memory.
//        The generic routine is bound by shared-memory wavefronts (7 stride-2 loads per node, ~2600 cycles per
//        half-sweep); this one needs 2 loads + 1 store per node, all conflict-free.
//  8^3, 4^3  one node per thread, neighbour addresses and rho in registers.
#pragma once

namespace pinc {

#define MS_ZBUF 4096         // doubles: 2 colours x 16 planes x 4 nodes x 32 lanes

static __device__ __forceinline__ double shflD(double v, int src){
	int lo = __double2loint(v), hi = __double2hiint(v);
	lo = __shfl_sync(0xffffffffu, lo, src); hi = __shfl_sync(0xffffffffu, hi, src);
	return __hiloint2double(hi, lo);
}

// P, R: the level's phi and rho in shared memory, x fastest; Z: MS_ZBUF doubles of scratch.  Returns with gBnd applied.
static __device__ __noinline__ void sGS16(double *P, const double *R, double *Z, int nCycles, double sIn, CK &K){
	const int lane = threadIdx.x & 31, l = threadIdx.x >> 5, k = lane >> 1, hf = lane & 1;
	const int rho_ = (k + l) & 1;                       // colour-1 nodes of this row sit at pencil positions 2i + rho_
	const int base = 8*hf + 16*(k + 16*l);
	double a[4], b[4], ra[4], rb[4];                   // a: colour 1 (updated first), b: colour 0
	#pragma unroll
	for(int i = 0; i < 4; i++){
		int pa = base + 2*i + rho_, pb = base + 2*i + 1 - rho_;
		a[i] = P[pa]; b[i] = P[pb]; ra[i] = R[pa]; rb[i] = R[pb];
		if(sIn != 0.0){ a[i] -= sIn; b[i] -= sIn; }
	}
	double *Z0 = Z, *Z1 = Z + 2048;                     // [plane][i][lane] of colour 0 / colour 1
	const int zme = l*128 + lane, zup = ((l+1)&15)*128 + lane, zdn = ((l+15)&15)*128 + lane;
	#pragma unroll
	for(int i = 0; i < 4; i++) Z0[zme + 32*i] = b[i];
	__syncthreads();
	const int up = (lane + 2) & 31, dn = (lane + 30) & 31;
	const double coeff = 1./6.;
	for(int c = 0; c < nCycles; c++){
		{	// colour 1: a from b
			double edge = shflD(rho_ ? b[0] : b[3], lane ^ 1);
			double n[4];
			#pragma unroll
			for(int i = 0; i < 4; i++){
				double xp = rho_ ? (i < 3 ? b[i < 3 ? i+1 : 3] : edge) : b[i];
				double xm = rho_ ? b[i] : (i > 0 ? b[i > 0 ? i-1 : 0] : edge);
				double yp = shflD(b[i], up), ym = shflD(b[i], dn);
				double zp = Z0[zup + 32*i], zm = Z0[zdn + 32*i];
				n[i] = coeff*(xp + xm + yp + ym + zp + zm + ra[i]);
			}
			#pragma unroll
			for(int i = 0; i < 4; i++){ a[i] = n[i]; Z1[zme + 32*i] = n[i]; }
		}
		__syncthreads();
		{	// colour 0: b from a
			double edge = shflD(rho_ ? a[3] : a[0], lane ^ 1);
			double n[4];
			#pragma unroll
			for(int i = 0; i < 4; i++){
				double xp = rho_ ? a[i] : (i < 3 ? a[i < 3 ? i+1 : 3] : edge);
				double xm = rho_ ? (i > 0 ? a[i > 0 ? i-1 : 0] : edge) : a[i];
				double yp = shflD(a[i], up), ym = shflD(a[i], dn);
				double zp = Z1[zup + 32*i], zm = Z1[zdn + 32*i];
				n[i] = coeff*(xp + xm + yp + ym + zp + zm + rb[i]);
			}
			#pragma unroll
			for(int i = 0; i < 4; i++){ b[i] = n[i]; Z0[zme + 32*i] = n[i]; }
		}
		__syncthreads();
	}
	// the 2*nCycles gBnd calls, applied once
	double acc = 0;
	#pragma unroll
	for(int i = 0; i < 4; i++){ acc += a[i]; acc += b[i]; }
	const double avg = blockSumC(K, acc)/4096.0;
	#pragma unroll
	for(int i = 0; i < 4; i++){
		P[base + 2*i + rho_] = a[i] - avg;
		P[base + 2*i + 1 - rho_] = b[i] - avg;
	}
	__syncthreads();
}

// N = 8 or 4: node t = x + N*(y + N*z) belongs to thread t
template<int N> static __device__ __noinline__ void sGSOne(double *P, const double *R, int nCycles, double sIn, CK &K){
	constexpr int NN = N*N*N;
	const int t = threadIdx.x;
	const bool act = t < NN;
	const int x = t & (N-1), y = (t / N) & (N-1), z = (t / (N*N)) & (N-1);
	const int colour = (x + y + z + 1) & 1;             // (j+k+l)&1 of the 1-based true node
	const int row = N*(y + N*z);
	const int ixp = row + ((x+1) & (N-1)), ixm = row + ((x+N-1) & (N-1));
	const int iyp = x + N*(((y+1) & (N-1)) + N*z), iym = x + N*(((y+N-1) & (N-1)) + N*z);
	const int izp = x + N*(y + N*((z+1) & (N-1))), izm = x + N*(y + N*((z+N-1) & (N-1)));
	double v = 0, r = 0;
	if(act){ v = P[t]; r = R[t]; if(sIn != 0.0){ v -= sIn; P[t] = v; } }
	__syncthreads();
	const double coeff = 1./6.;
	for(int h = 0; h < 2*nCycles; h++){
		const int parity = (h & 1) ? 0 : 1;
		if(act && colour == parity){
			v = coeff*(P[ixp] + P[ixm] + P[iyp] + P[iym] + P[izp] + P[izm] + r);
			P[t] = v;
		}
		__syncthreads();
	}
	const double avg = blockSumC(K, act ? v : 0.0)/(double)NN;
	if(act) P[t] = v - avg;
	__syncthreads();
}

// ---- the rest of the V-cycle on the cubic levels N = 16, 8, 4: compile-time sizes, masks instead of divisions ----------
template<int N> static __device__ __forceinline__ int sIdx(int x, int y, int z){ return (x & (N-1)) + N*((y & (N-1)) + N*(z & (N-1))); }
// gNeutralizeGrid (src/grid.c:730-779)
template<int N> static __device__ __forceinline__ void sNeutralize(double *A, CK &K){
	constexpr int NN = N*N*N;
	double acc = 0;
	for(int i = threadIdx.x; i < NN; i += blockDim.x) acc += A[i];
	const double avg = blockSumC(K, acc)/(double)NN;
	for(int i = threadIdx.x; i < NN; i += blockDim.x) A[i] -= avg;
	__syncthreads();
}
// residual at (x,y,z): -6 phi; += six neighbours; += rho (src/grid.c:318-322, src/multigrid.c:1400)
template<int N> static __device__ __forceinline__ double sResAt(const double *P, const double *R, int x, int y, int z){
	const int c = sIdx<N>(x,y,z);
	double r = -6.*P[c];
	r += P[sIdx<N>(x+1,y,z)] + P[sIdx<N>(x-1,y,z)] + P[sIdx<N>(x,y+1,z)] + P[sIdx<N>(x,y-1,z)] + P[sIdx<N>(x,y,z+1)] + P[sIdx<N>(x,y,z-1)];
	r += R[c];
	return r;
}
// mgResidual + mgHalfRestrict3D (src/multigrid.c:844-911) into the coarse rho, followed by the coarse level's gBnd(rho)
template<int N> static __device__ __noinline__ void sRestrict(const double *P, const double *R, double *Rc, CK &K){
	constexpr int H = N/2, HH = H*H*H, U = (HH + 511)/512;
	double mine[U];
	double acc = 0;
	#pragma unroll
	for(int u = 0; u < U; u++){
		const int i = threadIdx.x + u*512;
		mine[u] = 0;
		if(i < HH){
			const int X = i & (H-1), Y = (i / H) & (H-1), Z = i / (H*H);
			const int x = 2*X, y = 2*Y, z = 2*Z;
			const double coeff = 1./12.;
			double v = coeff*(6*sResAt<N>(P,R,x,y,z)
				+ sResAt<N>(P,R,x+1,y,z) + sResAt<N>(P,R,x-1,y,z)
				+ sResAt<N>(P,R,x,y+1,z) + sResAt<N>(P,R,x,y-1,z)
				+ sResAt<N>(P,R,x,y,z+1) + sResAt<N>(P,R,x,y,z-1));
			mine[u] = v; acc += v;
		}
	}
	const double avg = blockSumC(K, acc)/(double)HH;
	#pragma unroll
	for(int u = 0; u < U; u++){ const int i = threadIdx.x + u*512; if(i < HH) Rc[i] = mine[u] - avg; }
	__syncthreads();
}
// trilinear prolongation in the nesting of the reference's three passes (z, then y, then x; multigrid.c:1127-1238);
// H = coarse size, (x,y,z) = 0-based fine node
// (branch-free: an even fine node's two coarse neighbours are the same node and 0.5*(a + a) == a exactly, see prolPoint)
template<int H> static __device__ __forceinline__ double sProl(const double *C, int x, int y, int z){
	const int xa = x >> 1, xb = (x+1) >> 1, ya = y >> 1, yb = (y+1) >> 1, za = z >> 1, zb = (z+1) >> 1;
	const double v000 = C[sIdx<H>(xa,ya,za)], v001 = C[sIdx<H>(xa,ya,zb)], v010 = C[sIdx<H>(xa,yb,za)], v011 = C[sIdx<H>(xa,yb,zb)];
	const double v100 = C[sIdx<H>(xb,ya,za)], v101 = C[sIdx<H>(xb,ya,zb)], v110 = C[sIdx<H>(xb,yb,za)], v111 = C[sIdx<H>(xb,yb,zb)];
	const double pa = 0.5*(0.5*(v000 + v001) + 0.5*(v010 + v011));
	const double pb = 0.5*(0.5*(v100 + v101) + 0.5*(v110 + v111));
	return 0.5*(pa + pb);
}
// res := P(phi coarse); phi += res; returns the mean of the new phi (the gBnd that follows is applied by the smoother)
template<int N> static __device__ __noinline__ double sProlongAdd(double *P, const double *Pc, double *resG, int s0, int s1, CK &K){
	constexpr int NN = N*N*N, U = (NN + 511)/512;
	double acc = 0;
	#pragma unroll
	for(int u = 0; u < U; u++){
		const int i = threadIdx.x + u*512;
		if(i < NN){
			const int x = i & (N-1), y = (i / N) & (N-1), z = i / (N*N);
			const double p = sProl<N/2>(Pc, x, y, z);
			resG[(x+1) + (long)s0*((y+1) + (long)s1*(z+1))] = p;
			double v = P[i]; v += p;
			P[i] = v;
			acc += v;
		}
	}
	return blockSumC(K, acc)/(double)NN;
}
template<int N> static __device__ __forceinline__ void sSmooth(double *P, const double *R, double *Z, int nCycles, double sIn, CK &K){
	if(N == 16) sGS16(P, R, Z, nCycles, sIn, K); else sGSOne<N>(P, R, nCycles, sIn, K);
}

} // namespace pinc
