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

} // namespace pinc
