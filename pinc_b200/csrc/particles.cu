// particles.cu — particle kernels and the PINC pusher entry points (src/pusher.h).
//
// Device layout: structure of arrays, six planes x,y,z,vx,vy,vz of `cap` doubles (the host keeps the
// reference's AoS pos[3i+d] / vel[3i+d]; conversion happens in pincSyncPop*).  Species s lives in
// [iStart[s], iStop[s]) of every plane, exactly as in src/core.h:72-86.
//
// Cell binning.  puExtractEmigrants3D is a counting sort of every species by the key
//     key = cell index  j + nc0*(k + nc1*l)        for particles that stay,
//     key = nCells + ne                            for emigrants to neighbour ne (src/pusher.c:819-826),
// so that afterwards (i) the stayers are contiguous and ordered by cell, (ii) the emigrants sit behind
// iStop[s] packed neighbour by neighbour (the in-GPU compaction for migrants), and (iii) cellStart[]
// gives every cell's particle range.  The accelerator then reads E with warp-uniform addresses and the
// deposition runs one warp per cell with no per-particle atomics.
//
// Deposition is deterministic: each trilinear weight w in [0,1] is accumulated as round(w*2^46) in a
// 64-bit integer (warp butterfly per cell, then one integer RED per node), so the sum is exact and
// independent of particle order; the integer grid is converted to fp64 in the reference's order of
// operations (src/pusher.c:514,522,568: rho = (rho*(1/q) + sum w)*q per species).
//
// All arithmetic that reaches particle state follows the reference's expression order and the library is
// compiled with -fmad=false (the reference's x86 build has no FMA contraction).
#include "common.h"
#include <atomic>
#include <mutex>

namespace pinc {

struct Thr { double lo[3], up[3]; };
struct CellSpace { int nc0, nc1, nc2; long nCells; };

__device__ __forceinline__ unsigned classify(double x, double y, double z, const Thr &T, const CellSpace &C, int *flags){
	int nx = -(x < T.lo[0]) + (x >= T.up[0]);
	int ny = -(y < T.lo[1]) + (y >= T.up[1]);
	int nz = -(z < T.lo[2]) + (z >= T.up[2]);
	int ne = 13 + nx + 3*ny + 9*nz;
	if(ne != 13) return (unsigned)(C.nCells + ne);
	int j = (int)x, k = (int)y, l = (int)z;
	if(j < 0 || j >= C.nc0 || k < 0 || k >= C.nc1 || l < 0 || l >= C.nc2 || !(x >= 0) || !(y >= 0) || !(z >= 0)){
		atomicOr(flags, ERR_POS_RANGE);
		j = min(max(j,0),C.nc0-1); k = min(max(k,0),C.nc1-1); l = min(max(l,0),C.nc2-1);
	}
	return (unsigned)(j + C.nc0*(k + (long)C.nc1*l));
}

// all 32 lanes call; lanes with the same key add once
__device__ __forceinline__ void histAdd(unsigned *hist, unsigned key, bool act){
	unsigned k = act ? key : 0xffffffffu;
	unsigned peers = __match_any_sync(0xffffffffu, k);
	int lane = threadIdx.x & 31;
	if(act && lane == __ffs(peers)-1) atomicAdd(&hist[key], (unsigned)__popc(peers));
}

// round-to-nearest-even of w*2^46 as an integer, by the 1.5*2^52 magic constant: one FMA (w*2^46 is an exact
// scaling, so the single rounding of the FMA is the rounding to integer) and one integer subtract, instead of a
// multiply and a (quarter-rate) F2I.S64; identical values to __double2ll_rn(w*2^46) for 0 <= w <= 1.
__device__ __forceinline__ long long fixw(double w){
	const double M = 6755399441055744.0;               // 1.5 * 2^52
	return __double_as_longlong(__fma_rn(w, (double)(1LL<<PINC_FIX_BITS), M)) - __double_as_longlong(M);
}
// fixw without the subtraction: a sum of n of these is the sum of the n integers plus n*FIXW_BIAS (modulo 2^64)
__device__ __forceinline__ long long fixwRaw(double w){ return __double_as_longlong(__fma_rn(w, (double)(1LL<<PINC_FIX_BITS), 6755399441055744.0)); }
#define FIXW_BIAS 0x4338000000000000ULL            // bits of 1.5*2^52
// the eight trilinear weights of a particle at fractional position (xf,yf,zf), products in the reference's order
// (src/pusher.c:556-563), index = dx + 2*dy + 4*dz
__device__ __forceinline__ void cornerWeights(double xf, double yf, double zf, long long a[8]){
	double xc = 1-xf, yc = 1-yf, zc = 1-zf;
	double cc = xc*yc, fc = xf*yc, cf = xc*yf, ff = xf*yf;
	a[0] = fixw(cc*zc); a[1] = fixw(fc*zc); a[2] = fixw(cf*zc); a[3] = fixw(ff*zc);
	a[4] = fixw(cc*zf); a[5] = fixw(fc*zf); a[6] = fixw(cf*zf); a[7] = fixw(ff*zf);
}
// transposing butterfly: 8 values x 32 lanes -> the warp total of corner (lane>>2)&7 in every lane, 9 64-bit shuffles
__device__ __forceinline__ long long warpCornerTotal(long long a[8]){
	const int lane = threadIdx.x & 31;
	const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
	#pragma unroll
	for(int q = 0; q < 4; q++){
		long long send = u16 ? a[q] : a[q+4], keep = u16 ? a[q+4] : a[q];
		a[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
	}
	#pragma unroll
	for(int q = 0; q < 2; q++){
		long long send = u8 ? a[q] : a[q+2], keep = u8 ? a[q+2] : a[q];
		a[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
	}
	{
		long long send = u4 ? a[0] : a[1], keep = u4 ? a[1] : a[0];
		a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
	}
	a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
	a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
	return a[0];
}
// node offset of corner (lane>>2)&7 = dx + 2*dy + 4*dz ... held by the lanes with bits 4 (dx), 8 (dy), 16 (dz)
__device__ __forceinline__ long cornerOffset(long sx, long sxy){
	const int lane = threadIdx.x & 31;
	return ((lane & 4) ? 1 : 0) + ((lane & 8) ? sx : 0) + ((lane & 16) ? sxy : 0);
}

template<int BLOCK> __device__ __forceinline__ double blockSumP(double v){
	__shared__ double sh[BLOCK/32];
	for(int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if(lane == 0) sh[w] = v;
	__syncthreads();
	if(w == 0){
		v = lane < BLOCK/32 ? sh[lane] : 0.0;
		for(int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	}
	return v;
}
__global__ void k_final_sum_p(const double *__restrict__ partial, int n, double *__restrict__ out){
	double acc = 0;
	for(int i = threadIdx.x; i < n; i += 256) acc += partial[i];
	acc = blockSumP<256>(acc);
	if(threadIdx.x == 0) out[0] = acc;
}

// ---- move (src/pusher.c:86-119, quirk Q4: pos += vel) -------------------------------------------
__global__ void k_move(double *__restrict__ p, const double *__restrict__ v, long cap, long a, long n){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st){
		long q = a + i;
		p[q] += v[q]; p[q+cap] += v[q+cap]; p[q+2*cap] += v[q+2*cap];
	}
}

// ---- accelerate (src/pusher.c:147-214, 394-483) with the interpolation of :1089-1122 embedded -----
// KIND 0: leapfrog kick; 1: Boris (half kick, rotation by T then S of THIS particle's velocity, half kick).
// FUSE 1: also pos += vel (the next step's puMove) and classification of the new position into the sort key.
// FUSE 2: as 1, plus the deposition of the moved particle (the next step's puDistr3D1) for the particles that stay on
//         this rank: the lanes of a warp that share a cell (particles are in cell order, ~95 % keep their cell) are
//         summed with the 9-shuffle butterfly, two groups per warp, stragglers add their eight weights directly.
struct BorisPar { double T[3], S[3]; };
template<int KIND, int KE, int FUSE>
__global__ void __launch_bounds__(256, 4) k_acc(double *__restrict__ P, long cap, long a, long n,
		const double *__restrict__ E, long sx3, long sxy3, int gs0, int gs1, int gs2, BorisPar B, double *__restrict__ partial,
		Thr T, CellSpace C, unsigned *__restrict__ keys, unsigned *__restrict__ hist, int *flags,
		long long *__restrict__ fix, long fsx, long fsxy){
	double acc = 0;
	long stride = (long)gridDim.x*blockDim.x;
	for(long base = blockIdx.x*(long)blockDim.x; base < n; base += stride){
		long i = base + threadIdx.x;
		bool act = i < n;
		unsigned key = 0;
		double nx_ = 0, ny_ = 0, nz_ = 0;          // moved position (FUSE 2)
		if(act){
			long q = a + i;
			double x = P[q], y = P[q+cap], z = P[q+2*cap];
			double vx = P[q+3*cap], vy = P[q+4*cap], vz = P[q+5*cap];
			int j = (int)x, k = (int)y, l = (int)z;
			if(!(x >= 0) || !(y >= 0) || !(z >= 0) || j > gs0-2 || k > gs1-2 || l > gs2-2){
				atomicOr(flags, ERR_POS_RANGE);          // not migrated: the reference would read out of bounds
				j = min(max(j,0),gs0-2); k = min(max(k,0),gs1-2); l = min(max(l,0),gs2-2);
			}
			double xf = x-j, yf = y-k, zf = z-l;
			double xc = 1-xf, yc = 1-yf, zc = 1-zf;
			long p000 = 3L*j + k*sx3 + l*sxy3;
			const double *e0 = E + p000, *e1 = e0 + sx3, *e2 = e0 + sxy3, *e3 = e2 + sx3;
			double dv[3];
			#pragma unroll
			for(int v = 0; v < 3; v++)
				dv[v] = zc*( yc*(xc*__ldg(e0+v)+xf*__ldg(e0+3+v)) + yf*(xc*__ldg(e1+v)+xf*__ldg(e1+3+v)) )
				      + zf*( yc*(xc*__ldg(e2+v)+xf*__ldg(e2+3+v)) + yf*(xc*__ldg(e3+v)+xf*__ldg(e3+3+v)) );
			if(KIND == 0){
				if(KE){
					double v2 = 0;
					v2 += vx*(vx+dv[0]); v2 += vy*(vy+dv[1]); v2 += vz*(vz+dv[2]);
					acc += v2;
				}
				vx += dv[0]; vy += dv[1]; vz += dv[2];
			} else {
				vx += 0.5*dv[0]; vy += 0.5*dv[1]; vz += 0.5*dv[2];
				double px = vx, py = vy, pz = vz;
				px +=  (vy*B.T[2]-vz*B.T[1]); py += -(vx*B.T[2]-vz*B.T[0]); pz +=  (vx*B.T[1]-vy*B.T[0]);
				vx +=  (py*B.S[2]-pz*B.S[1]); vy += -(px*B.S[2]-pz*B.S[0]); vz +=  (px*B.S[1]-py*B.S[0]);
				if(KE){ double v2 = 0; v2 += vx*vx; v2 += vy*vy; v2 += vz*vz; acc += v2; }
				vx += 0.5*dv[0]; vy += 0.5*dv[1]; vz += 0.5*dv[2];
			}
			P[q+3*cap] = vx; P[q+4*cap] = vy; P[q+5*cap] = vz;
			if(FUSE){
				x += vx; y += vy; z += vz;
				P[q] = x; P[q+cap] = y; P[q+2*cap] = z;
				key = classify(x, y, z, T, C, flags);
				keys[q] = key;
				nx_ = x; ny_ = y; nz_ = z;
			}
		}
		if(FUSE) histAdd(hist, key, act);
		if(FUSE == 2){
			const int lane = threadIdx.x & 31;
			bool stay = act && key < (unsigned)C.nCells;
			long long w[8];
			#pragma unroll
			for(int qq = 0; qq < 8; qq++) w[qq] = 0;
			if(stay) cornerWeights(nx_ - (double)(int)nx_, ny_ - (double)(int)ny_, nz_ - (double)(int)nz_, w);
			unsigned remaining = __ballot_sync(0xffffffffu, stay);
			#pragma unroll 1
			for(int round = 0; round < 2 && remaining; round++){
				int leader = __ffs(remaining) - 1;
				unsigned gkey = __shfl_sync(0xffffffffu, key, leader);
				bool member = stay && key == gkey;
				long long v[8];
				#pragma unroll
				for(int qq = 0; qq < 8; qq++) v[qq] = member ? w[qq] : 0;
				long long tot = warpCornerTotal(v);
				if((lane & 3) == 0 && tot != 0){
					long cj = gkey % (unsigned)C.nc0, cr = gkey / (unsigned)C.nc0, ck = cr % C.nc1, cl = cr / C.nc1;
					atomicAdd((unsigned long long*)&fix[cj + fsx*ck + fsxy*cl + cornerOffset(fsx, fsxy)], (unsigned long long)tot);
				}
				unsigned done = __ballot_sync(0xffffffffu, member);
				remaining &= ~done;
				if(member) stay = false;
			}
			if(stay){                                   // a third cell in this warp (rare): eight direct integer REDs
				long cj = key % (unsigned)C.nc0, cr = key / (unsigned)C.nc0, ck = cr % C.nc1, cl = cr / C.nc1;
				unsigned long long *f = (unsigned long long*)fix + (cj + fsx*ck + fsxy*cl);
				atomicAdd(f, (unsigned long long)w[0]); atomicAdd(f+1, (unsigned long long)w[1]);
				atomicAdd(f+fsx, (unsigned long long)w[2]); atomicAdd(f+fsx+1, (unsigned long long)w[3]);
				atomicAdd(f+fsxy, (unsigned long long)w[4]); atomicAdd(f+fsxy+1, (unsigned long long)w[5]);
				atomicAdd(f+fsxy+fsx, (unsigned long long)w[6]); atomicAdd(f+fsxy+fsx+1, (unsigned long long)w[7]);
			}
		}
	}
	if(KE){
		acc = blockSumP<256>(acc);
		if(threadIdx.x == 0) partial[blockIdx.x] = acc;
	}
}

// ---- classification + histogram of the current positions ------------------------------------------
__global__ void __launch_bounds__(256) k_keys(const double *__restrict__ P, long cap, long a, long n, Thr T, CellSpace C,
		unsigned *__restrict__ keys, unsigned *__restrict__ hist, int *flags){
	long stride = (long)gridDim.x*blockDim.x;
	for(long base = blockIdx.x*(long)blockDim.x; base < n; base += stride){
		long i = base + threadIdx.x;
		bool act = i < n;
		unsigned key = 0;
		if(act){
			long q = a + i;
			key = classify(P[q], P[q+cap], P[q+2*cap], T, C, flags);
			keys[q] = key;
		}
		histAdd(hist, key, act);
	}
}

// ---- exclusive scan of the histogram (in place; element n receives the total) -----------------------
#define SCAN_CH 4096
__global__ void __launch_bounds__(256) k_scan_block(unsigned *__restrict__ h, long n, unsigned *__restrict__ blockSums){
	__shared__ unsigned sh[256];
	long base = (long)blockIdx.x*SCAN_CH + threadIdx.x*16;
	unsigned v[16], sum = 0;
	#pragma unroll
	for(int i = 0; i < 16; i++){ long g = base + i; v[i] = g < n ? h[g] : 0u; sum += v[i]; }
	sh[threadIdx.x] = sum;
	__syncthreads();
	for(int o = 1; o < 256; o <<= 1){
		unsigned t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
		__syncthreads();
		sh[threadIdx.x] += t;
		__syncthreads();
	}
	unsigned run = sh[threadIdx.x] - sum;            // exclusive prefix of this thread inside the chunk
	#pragma unroll
	for(int i = 0; i < 16; i++){ long g = base + i; if(g < n) h[g] = run; run += v[i]; }
	if(threadIdx.x == 255) blockSums[blockIdx.x] = sh[255];
}
__global__ void k_scan_sums(unsigned *__restrict__ s, int nb){
	__shared__ unsigned sh[1024];
	__shared__ unsigned carry;
	if(threadIdx.x == 0) carry = 0;
	__syncthreads();
	for(int b0 = 0; b0 < nb; b0 += 1024){
		int i = b0 + threadIdx.x;
		unsigned v = i < nb ? s[i] : 0u;
		sh[threadIdx.x] = v;
		__syncthreads();
		for(int o = 1; o < 1024; o <<= 1){
			unsigned t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
			__syncthreads();
			sh[threadIdx.x] += t;
			__syncthreads();
		}
		if(i < nb) s[i] = carry + sh[threadIdx.x] - v;
		__syncthreads();
		if(threadIdx.x == 1023) carry += sh[1023];
		__syncthreads();
	}
}
__global__ void k_scan_add(unsigned *__restrict__ h, long n, const unsigned *__restrict__ blockSums){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x;
	if(i < n) h[i] += blockSums[i / SCAN_CH];
}

// ---- scatter into cell order ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(const double *__restrict__ src, double *__restrict__ dst, long cap, long a, long n,
		const unsigned *__restrict__ keys, const unsigned *__restrict__ off, unsigned *__restrict__ cursor){
	long stride = (long)gridDim.x*blockDim.x;
	int lane = threadIdx.x & 31;
	for(long base = blockIdx.x*(long)blockDim.x; base < n; base += stride){
		long i = base + threadIdx.x;
		bool act = i < n;
		unsigned key = act ? keys[a+i] : 0xffffffffu;
		unsigned peers = __match_any_sync(0xffffffffu, key);
		int leader = __ffs(peers)-1;
		unsigned b = 0;
		if(act && lane == leader) b = atomicAdd(&cursor[key], (unsigned)__popc(peers));
		b = __shfl_sync(0xffffffffu, b, leader);
		if(act){
			long d = a + off[key] + b + __popc(peers & ((1u<<lane)-1u));
			long q = a + i;
			#pragma unroll
			for(int w = 0; w < 6; w++) dst[d + w*cap] = src[q + w*cap];
		}
	}
}

// ---- deposition (src/pusher.c:512-572) -----------------------------------------------------------------

// sorted prefix: one warp per cell
__global__ void __launch_bounds__(256, 4) k_distr_cells(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
		const unsigned *__restrict__ cs, CellSpace C, long sx, long sxy, long long *__restrict__ fix){
	int lane = threadIdx.x & 31;
	long warp = (blockIdx.x*(long)blockDim.x + threadIdx.x) >> 5;
	long nWarps = ((long)gridDim.x*blockDim.x) >> 5;
	for(long c = warp; c < C.nCells; c += nWarps){
		unsigned b = cs[c], e = cs[c+1];
		if(b == e) continue;
		long long a[8];
		#pragma unroll
		for(int q = 0; q < 8; q++) a[q] = 0;
		// every particle of the bin has (int)x == cj etc. (that is how the key was formed): convert once per cell
		const int cj = (int)(c % C.nc0); const long cr = c / C.nc0; const int ck = (int)(cr % C.nc1); const int cl = (int)(cr / C.nc1);
		const double dj = (double)cj, dk = (double)ck, dl = (double)cl;
		// up to 96 particles of the cell per trip, all loads issued before the arithmetic (memory-level parallelism)
		for(unsigned i0 = b; i0 < e; i0 += 96){
			double x[3], y[3], z[3];
			#pragma unroll
			for(int u = 0; u < 3; u++){
				unsigned i = i0 + lane + 32*u;
				bool ok = i < e;
				x[u] = ok ? X[i] : 0.0; y[u] = ok ? Y[i] : 0.0; z[u] = ok ? Z[i] : 0.0;
			}
			#pragma unroll
			for(int u = 0; u < 3; u++){
				if(i0 + lane + 32*u >= e) continue;
				// raw bits of fma(w, 2^46, 1.5*2^52): the bias comes off the warp total (e x bits(1.5*2^52), modulo 2^64), which
				// saves the 64-bit subtraction per weight in a kernel that is bound by instruction issue
				const double xf = x[u]-dj, yf = y[u]-dk, zf = z[u]-dl;
				const double xc = 1-xf, yc = 1-yf, zc = 1-zf;
				const double cc = xc*yc, fc = xf*yc, cf = xc*yf, ff = xf*yf;
				a[0] += fixwRaw(cc*zc); a[1] += fixwRaw(fc*zc); a[2] += fixwRaw(cf*zc); a[3] += fixwRaw(ff*zc);
				a[4] += fixwRaw(cc*zf); a[5] += fixwRaw(fc*zf); a[6] += fixwRaw(cf*zf); a[7] += fixwRaw(ff*zf);
			}
		}
		long long tot = warpCornerTotal(a);
		tot -= (long long)((unsigned long long)(e - b)*FIXW_BIAS);
		if((lane & 3) == 0 && tot != 0)
			atomicAdd((unsigned long long*)&fix[cj + sx*ck + sxy*cl + cornerOffset(sx, sxy)], (unsigned long long)tot);
	}
}
// unsorted tail (immigrants appended by puMigrate, or a population that was never binned)
__global__ void __launch_bounds__(256) k_distr_tail(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
		long n, int s0, int s1, int s2, long long *__restrict__ fix, int *flags){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	long sx = s0, sxy = (long)s0*s1;
	for(; i < n; i += st){
		double x = X[i], y = Y[i], z = Z[i];
		int j = (int)x, k = (int)y, l = (int)z;
		if(!(x >= 0) || !(y >= 0) || !(z >= 0) || j > s0-2 || k > s1-2 || l > s2-2){ atomicOr(flags, ERR_POS_RANGE); continue; }
		long long w[8];
		cornerWeights(x-j, y-k, z-l, w);
		unsigned long long *f = (unsigned long long*)fix + (j + sx*k + sxy*l);
		atomicAdd(f,          (unsigned long long)w[0]);
		atomicAdd(f+1,        (unsigned long long)w[1]);
		atomicAdd(f+sx,       (unsigned long long)w[2]);
		atomicAdd(f+sx+1,     (unsigned long long)w[3]);
		atomicAdd(f+sxy,      (unsigned long long)w[4]);
		atomicAdd(f+sxy+1,    (unsigned long long)w[5]);
		atomicAdd(f+sxy+sx,   (unsigned long long)w[6]);
		atomicAdd(f+sxy+sx+1, (unsigned long long)w[7]);
	}
}
// rho = (rho*(1/q) + sum w)*q, and the integer grid is cleared for the next species
__global__ void k_distr_finalize(double *__restrict__ rho, long long *__restrict__ fix, long n, double invq, double q, int *flags){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	const double inv = 1.0/(double)(1LL<<PINC_FIX_BITS);
	for(; i < n; i += st){
		double r = rho[i];
		const long long f = fix[i];
		// the accumulator wraps at 2^63 = 2^17 full-weight particles on one node: report at half of that instead of
		// handing a wrapped sum to the solver
		if(f > (1LL<<62) || f < -(1LL<<62)) atomicOr(flags, ERR_FIX_OVERFLOW);
		r *= invq;
		r += (double)f*inv;
		r *= q;
		rho[i] = r;
		fix[i] = 0;
	}
}

// ---- migrants --------------------------------------------------------------------------------------------
struct PackPar { long srcOff[28]; long dstOff[27]; };     // per neighbour: first emigrant (relative), first record
__global__ void k_pack_emigrants(const double *__restrict__ P, long cap, long a, long n, PackPar pp, double *__restrict__ out){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st){
		long r = pp.srcOff[0] + i;
		int ne = 0;
		while(ne < 26 && r >= pp.srcOff[ne+1]) ne++;
		long d = pp.dstOff[ne] + (r - pp.srcOff[ne]);
		long q = a + r;
		#pragma unroll
		for(int w = 0; w < 6; w++) out[6*d + w] = P[q + w*cap];
	}
}
struct ImportPar { long srcOff[27]; long cnt[27]; double shift[27][3]; };
__global__ void k_import(double *__restrict__ P, long cap, long dst0, long n, ImportPar ip, const double *__restrict__ in){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st){
		long r = i; int ne = 0;
		while(ne < 26 && r >= ip.cnt[ne]){ r -= ip.cnt[ne]; ne++; }
		const double *rec = in + 6*(ip.srcOff[ne] + r);
		long q = dst0 + i;
		#pragma unroll
		for(int w = 0; w < 3; w++){ double v = rec[w]; v += ip.shift[ne][w]; P[q + w*cap] = v; }
		#pragma unroll
		for(int w = 3; w < 6; w++) P[q + w*cap] = rec[w];
	}
}

// =================================================================================================================
// Cell-slotted storage ("slotted mode"): the steady state of pincAccMove3D1KE -> puExtractEmigrants3D -> puMigrate ->
// puDistr3D1.  The counting sort above rewrites all 48 B of every particle every step although ~95 % of them keep
// their cell.  Here cell c of species s owns `cap` slots of six planes; one warp per cell pushes its particles, keeps the
// ones that stay in place (compacted towards the front of the cell's run) and appends the others - to another cell of
// this rank, or emigrants - to the species' mover list (the second population buffer, which the sort no longer needs);
// puExtractEmigrants3D then drops the local movers into their new cells (one atomic per mover) and counts the emigrants,
// puMigrate packs those from the list and imports the immigrants straight into their cells, puDistr3D1 runs its
// warp-per-cell deposition over the slots.  Per particle and step the path moves 96 B (push) + 24 B (deposit) instead
// of 96 + 4 + 100 + 24.  A cell that outgrows its slots raises ERR_SLOT_OVERFLOW; the host then goes back to the
// contiguous layout (popLeaveSlotted) and re-enters with a larger capacity.  Same arithmetic per particle as k_acc, same
// integer deposition: fields are bit-identical to the sorted path, the particle ORDER differs (as it already does
// between two runs of the scatter).
// =================================================================================================================
#define SLOT_OVER_KEY 0xfffffffeu          // a mover that did not fit into its cell: waits in the list for popLeaveSlotted
#define SLOT_EMPTY_KEY 0xffffffffu         // an unused entry of the mover list (the list is handed out in chunks)
#define MV_CHUNK 128u                      // >= the 96 particles of one trip of k_cell_push
struct SlotPar { double *S; long plane; long off; int cap; unsigned *cnt; };

// MODE 0: pincAccMove3D1KE (kick + move + re-binning); 1: puAcc3D1(KE) alone (velocities in place, nothing moves);
// 2: puMove alone (pos += vel, re-binning).  1 followed by 2 is the reference's call order and leaves the same bits as 0.
template<int KE, int MODE, int DEP> __global__ void __launch_bounds__(256, 2) k_cell_push(SlotPar Q, const double *__restrict__ E, long sx3, long sxy3,
		CellSpace C, Thr T, double *__restrict__ partial, double *__restrict__ M, long mPlane, unsigned *__restrict__ mKey, unsigned *mCount, unsigned mCap, int *flags,
		long long *__restrict__ fix, long sx, long sxy){
	const int lane = threadIdx.x & 31;
	const unsigned lt = (1u << lane) - 1u;
	long warp = (blockIdx.x*(long)blockDim.x + threadIdx.x) >> 5;
	const long nWarps = ((long)gridDim.x*blockDim.x) >> 5;
	__shared__ double evS[8][24];
	double acc = 0;
	unsigned chBase = 0, chUsed = MV_CHUNK;          // this warp's chunk of the mover list: none yet
	unsigned nNext = warp < C.nCells ? Q.cnt[warp] : 0u;
	for(long c = warp; c < C.nCells; c += nWarps){
		const unsigned n = nNext;
		if(c + nWarps < C.nCells) nNext = Q.cnt[c + nWarps];          // the next cell's count is on its way while this cell is pushed
		if(n == 0) continue;
		const int cj = (int)(c % C.nc0); const long cr = c / C.nc0; const int ck = (int)(cr % C.nc1); const int cl = (int)(cr / C.nc1);
		const double dj = (double)cj, dk = (double)ck, dl = (double)cl;
		const double *e0 = E + 3L*cj + ck*sx3 + cl*sxy3, *e1 = e0 + sx3, *e2 = e0 + sxy3, *e3 = e2 + sx3;
		double *P = Q.S + Q.off + c*(long)Q.cap;
		unsigned wr = 0;
		long long a[8];
		if(DEP){
			#pragma unroll
			for(int q = 0; q < 8; q++) a[q] = 0;
		}
		for(unsigned i0 = 0; i0 < n; i0 += 96){
			// up to 96 particles per trip (nearly always the whole cell), every load in flight before the first use
			double x[3], y[3], z[3], vx[3], vy[3], vz[3];
			#pragma unroll
			for(int u = 0; u < 3; u++){
				const unsigned i = i0 + lane + 32*u;
				const bool ok = i < n;
				x[u] = ok ? P[i] : 0.0; y[u] = ok ? P[i + Q.plane] : 0.0; z[u] = ok ? P[i + 2*Q.plane] : 0.0;
				vx[u] = ok ? P[i + 3*Q.plane] : 0.0; vy[u] = ok ? P[i + 4*Q.plane] : 0.0; vz[u] = ok ? P[i + 5*Q.plane] : 0.0;
			}
			if(MODE != 2 && i0 == 0){
				// the cell's eight corner fields once per cell: lane q*3+v fetches component v of corner q into the warp's shared row
				__syncwarp();
				if(lane < 24){
					const int q = lane/3, v = lane - 3*q;
					const double *ep = (q & 4 ? ((q & 2) ? e3 : e2) : ((q & 2) ? e1 : e0)) + ((q & 1) ? 3 : 0) + v;
					evS[threadIdx.x >> 5][lane] = __ldg(ep);
				}
				__syncwarp();
			}
			const double *ev = evS[threadIdx.x >> 5];
			unsigned key[3];
			#pragma unroll
			for(int u = 0; u < 3; u++){
				key[u] = 0xffffffffu;
				if(i0 + lane + 32*u >= n) continue;
				if(MODE != 2){
					const double xf = x[u]-dj, yf = y[u]-dk, zf = z[u]-dl;
					const double xc = 1-xf, yc = 1-yf, zc = 1-zf;
					double dv[3];
					#pragma unroll
					for(int v = 0; v < 3; v++)
						dv[v] = zc*( yc*(xc*ev[v]+xf*ev[3+v]) + yf*(xc*ev[6+v]+xf*ev[9+v]) )
						      + zf*( yc*(xc*ev[12+v]+xf*ev[15+v]) + yf*(xc*ev[18+v]+xf*ev[21+v]) );
					if(KE){
						double v2 = 0;
						v2 += vx[u]*(vx[u]+dv[0]); v2 += vy[u]*(vy[u]+dv[1]); v2 += vz[u]*(vz[u]+dv[2]);
						acc += v2;
					}
					vx[u] += dv[0]; vy[u] += dv[1]; vz[u] += dv[2];
				}
				if(MODE == 1){
					const unsigned i = i0 + lane + 32*u;
					P[i + 3*Q.plane] = vx[u]; P[i + 4*Q.plane] = vy[u]; P[i + 5*Q.plane] = vz[u];
					continue;
				}
				x[u] += vx[u]; y[u] += vy[u]; z[u] += vz[u];
				key[u] = classify(x[u], y[u], z[u], T, C, flags);
			}
			if(MODE == 1) continue;
			// (every load of the trip has landed - the keys need them - before the first store: the compaction may overwrite
			// slots that other lanes read in this trip, never slots of a later trip, since wr <= i0)
			unsigned ms[3], mm[3], nm = 0;
			#pragma unroll
			for(int u = 0; u < 3; u++){
				const bool ok = i0 + lane + 32*u < n;
				ms[u] = __ballot_sync(0xffffffffu, ok && key[u] == (unsigned)c);
				mm[u] = __ballot_sync(0xffffffffu, ok && key[u] != (unsigned)c);
				nm += __popc(mm[u]);
			}
			if(nm){
				// room for the trip's movers in the species' list: a warp takes MV_CHUNK entries at a time (one atomic on the
				// list's counter per ~30 cells instead of one per cell: every warp of the grid hits that one address)
				if(chUsed + nm > MV_CHUNK){
					for(unsigned k = chUsed + lane; k < MV_CHUNK; k += 32) mKey[chBase + k] = SLOT_EMPTY_KEY;
					unsigned b = 0;
					if(lane == 0) b = atomicAdd(mCount, (unsigned)MV_CHUNK);
					chBase = __shfl_sync(0xffffffffu, b, 0);
					chUsed = 0;
					if(chBase + MV_CHUNK > mCap){ atomicOr(flags, ERR_CAPACITY); chBase = 0; }       // (the host sizes the grid so that this cannot happen)
				}
			}
			#pragma unroll
			for(int u = 0; u < 3; u++){
				const bool ok = i0 + lane + 32*u < n;
				if(ok && key[u] == (unsigned)c){
					const unsigned d = wr + __popc(ms[u] & lt);
					P[d] = x[u]; P[d + Q.plane] = y[u]; P[d + 2*Q.plane] = z[u];
					if(MODE == 0 || d != i0 + lane + 32*u){ P[d + 3*Q.plane] = vx[u]; P[d + 4*Q.plane] = vy[u]; P[d + 5*Q.plane] = vz[u]; }      // (puMove leaves velocities alone)
					if(DEP){
						// the next step's puDistr3D1 for the particles that keep their cell (see k_distr_cells for the raw-bits sum)
						const double xf = x[u]-dj, yf = y[u]-dk, zf = z[u]-dl;
						const double xc = 1-xf, yc = 1-yf, zc = 1-zf;
						const double cc = xc*yc, fc = xf*yc, cf = xc*yf, ff = xf*yf;
						a[0] += fixwRaw(cc*zc); a[1] += fixwRaw(fc*zc); a[2] += fixwRaw(cf*zc); a[3] += fixwRaw(ff*zc);
						a[4] += fixwRaw(cc*zf); a[5] += fixwRaw(fc*zf); a[6] += fixwRaw(cf*zf); a[7] += fixwRaw(ff*zf);
					}
				} else if(ok){
					const unsigned d = chBase + chUsed + __popc(mm[u] & lt);
					M[d] = x[u]; M[d + mPlane] = y[u]; M[d + 2*mPlane] = z[u];
					M[d + 3*mPlane] = vx[u]; M[d + 4*mPlane] = vy[u]; M[d + 5*mPlane] = vz[u];
					mKey[d] = key[u];
				}
				wr += __popc(ms[u]);
				chUsed += __popc(mm[u]);
			}
		}
		if(MODE != 1 && lane == 0) Q.cnt[c] = wr;
		if(DEP && wr){
			long long tot = warpCornerTotal(a);
			tot -= (long long)((unsigned long long)wr*FIXW_BIAS);
			if((lane & 3) == 0 && tot != 0)
				atomicAdd((unsigned long long*)&fix[cj + sx*ck + sxy*cl + cornerOffset(sx, sxy)], (unsigned long long)tot);
		}
	}
	for(unsigned k = chUsed + lane; k < MV_CHUNK; k += 32) mKey[chBase + k] = SLOT_EMPTY_KEY;       // the unused rest of the last chunk
	if(KE){
		acc = blockSumP<256>(acc);
		if(threadIdx.x == 0) partial[blockIdx.x] = acc;
	}
}
// the eight weights of one particle straight into the accumulators (movers and immigrants: ~5 % of the population)
__device__ __forceinline__ void depositOne(long long *__restrict__ fix, long sx, long sxy, double x, double y, double z, int j, int k, int l){
	long long w[8];
	cornerWeights(x - (double)j, y - (double)k, z - (double)l, w);
	unsigned long long *f = (unsigned long long*)fix + (j + sx*k + sxy*l);
	atomicAdd(f, (unsigned long long)w[0]); atomicAdd(f + 1, (unsigned long long)w[1]);
	atomicAdd(f + sx, (unsigned long long)w[2]); atomicAdd(f + sx + 1, (unsigned long long)w[3]);
	atomicAdd(f + sxy, (unsigned long long)w[4]); atomicAdd(f + sxy + 1, (unsigned long long)w[5]);
	atomicAdd(f + sxy + sx, (unsigned long long)w[6]); atomicAdd(f + sxy + sx + 1, (unsigned long long)w[7]);
}
// movers of this rank into their new cells; emigrants are counted per neighbour (they stay in the list for puMigrate)
// fix != nullptr (pincAccMoveDistr3D1KE): the movers that stay on this rank are deposited here, whether or not their cell has room
__global__ void k_mv_insert(SlotPar Q, const double *__restrict__ M, long mPlane, unsigned *__restrict__ mKey, const unsigned *mCount,
		CellSpace C, unsigned *__restrict__ emHist, int *flags, long long *__restrict__ fix, long sx, long sxy){
	const long n = *mCount;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st){
		const unsigned key = mKey[i];
		if(key >= (unsigned)C.nCells){ if(key < SLOT_OVER_KEY) atomicAdd(&emHist[key - (unsigned)C.nCells], 1u); continue; }
		if(fix){
			const int cj = (int)(key % (unsigned)C.nc0); const unsigned cr = key / (unsigned)C.nc0; const int ck = (int)(cr % (unsigned)C.nc1), cl = (int)(cr / (unsigned)C.nc1);
			depositOne(fix, sx, sxy, M[i], M[i + mPlane], M[i + 2*mPlane], cj, ck, cl);
		}
		const unsigned slot = atomicAdd(&Q.cnt[key], 1u);
		if(slot >= (unsigned)Q.cap){ atomicOr(flags, ERR_SLOT_OVERFLOW); mKey[i] = SLOT_OVER_KEY; continue; }
		double *P = Q.S + Q.off + (long)key*Q.cap + slot;
		#pragma unroll
		for(int w = 0; w < 6; w++) P[w*Q.plane] = M[i + w*mPlane];
	}
}
// emigrants from the mover list into the message buffer, neighbour by neighbour (order inside a neighbour: arrival)
struct MvPackPar { long dstOff[27]; };
__global__ void k_mv_pack(const double *__restrict__ M, long mPlane, const unsigned *__restrict__ mKey, const unsigned *mCount, CellSpace C,
		MvPackPar pp, unsigned *__restrict__ cursor, double *__restrict__ out){
	const long n = *mCount;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st){
		const unsigned key = mKey[i];
		if(key < (unsigned)C.nCells || key >= SLOT_OVER_KEY) continue;
		const int ne = (int)(key - (unsigned)C.nCells);
		const long d = pp.dstOff[ne] + atomicAdd(&cursor[ne], 1u);
		#pragma unroll
		for(int w = 0; w < 6; w++) out[6*d + w] = M[i + w*mPlane];
	}
}
// immigrants (records, shifted into the local frame as k_import does) straight into their cells
__global__ void k_import_cells(SlotPar Q, long n, ImportPar ip, const double *__restrict__ in, CellSpace C,
		double *__restrict__ M, long mPlane, unsigned *__restrict__ mKey, unsigned *mCount, int *flags, long long *__restrict__ fix, long sx, long sxy){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st){
		long r = i; int ne = 0;
		while(ne < 26 && r >= ip.cnt[ne]){ r -= ip.cnt[ne]; ne++; }
		const double *rec = in + 6*(ip.srcOff[ne] + r);
		double v[6];
		#pragma unroll
		for(int w = 0; w < 3; w++){ v[w] = rec[w]; v[w] += ip.shift[ne][w]; }
		#pragma unroll
		for(int w = 3; w < 6; w++) v[w] = rec[w];
		int j = (int)v[0], k = (int)v[1], l = (int)v[2];
		if(!(v[0] >= 0) || !(v[1] >= 0) || !(v[2] >= 0) || j >= C.nc0 || k >= C.nc1 || l >= C.nc2){
			atomicOr(flags, ERR_POS_RANGE);
			j = min(max(j,0),C.nc0-1); k = min(max(k,0),C.nc1-1); l = min(max(l,0),C.nc2-1);
		}
		const unsigned key = (unsigned)(j + C.nc0*(k + (long)C.nc1*l));
		if(fix) depositOne(fix, sx, sxy, v[0], v[1], v[2], j, k, l);
		const unsigned slot = atomicAdd(&Q.cnt[key], 1u);
		if(slot >= (unsigned)Q.cap){
			atomicOr(flags, ERR_SLOT_OVERFLOW);
			const unsigned d = atomicAdd(mCount, 1u);
			#pragma unroll
			for(int w = 0; w < 6; w++) M[d + w*mPlane] = v[w];
			mKey[d] = SLOT_OVER_KEY;
			continue;
		}
		double *P = Q.S + Q.off + (long)key*Q.cap + slot;
		#pragma unroll
		for(int w = 0; w < 6; w++) P[w*Q.plane] = v[w];
	}
}
// contiguous cell-ordered planes <-> slots: one warp per cell copies the cell's run
__global__ void k_cells_copy(SlotPar Q, double *__restrict__ B, long bPlane, const unsigned *__restrict__ cellStart, long nCells, int toSlots){
	const int lane = threadIdx.x & 31;
	long warp = (blockIdx.x*(long)blockDim.x + threadIdx.x) >> 5;
	const long nWarps = ((long)gridDim.x*blockDim.x) >> 5;
	for(long c = warp; c < nCells; c += nWarps){
		const unsigned b = cellStart[c], n = cellStart[c+1] - b;
		double *P = Q.S + Q.off + c*(long)Q.cap;
		for(unsigned i = lane; i < n; i += 32){
			#pragma unroll
			for(int w = 0; w < 6; w++){ if(toSlots) P[i + w*Q.plane] = B[b + i + w*bPlane]; else B[b + i + w*bPlane] = P[i + w*Q.plane]; }
		}
		if(toSlots && lane == 0) Q.cnt[c] = n;
	}
}
// particles that are not in cell order (the unsorted tail behind sortedN) into their cells
__global__ void k_tail_insert(SlotPar Q, const double *__restrict__ B, long bPlane, long n, CellSpace C, int *flags){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st){
		double x = B[i], y = B[i + bPlane], z = B[i + 2*bPlane];
		int j = (int)x, k = (int)y, l = (int)z;
		if(!(x >= 0) || !(y >= 0) || !(z >= 0) || j >= C.nc0 || k >= C.nc1 || l >= C.nc2){ atomicOr(flags, ERR_SLOT_OVERFLOW); continue; }
		const unsigned key = (unsigned)(j + C.nc0*(k + (long)C.nc1*l));
		const unsigned slot = atomicAdd(&Q.cnt[key], 1u);
		if(slot >= (unsigned)Q.cap){ atomicOr(flags, ERR_SLOT_OVERFLOW); continue; }
		double *P = Q.S + Q.off + (long)key*Q.cap + slot;
		#pragma unroll
		for(int w = 0; w < 6; w++) P[w*Q.plane] = B[i + w*bPlane];
	}
}
// cell counts (clipped to the capacity: an overflowing cell's extra particles wait in the mover list) -> histogram for the scan
__global__ void k_cnt_to_hist(const unsigned *__restrict__ cnt, unsigned *__restrict__ hist, long nCells, long nKeys1, int cap){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < nKeys1; i += st) hist[i] = i < nCells ? min(cnt[i], (unsigned)cap) : 0u;
}
// movers of the list that are still waiting (all of them before puExtractEmigrants3D, the overflowed ones after) -> behind the cells' particles
__global__ void k_mv_append(const double *__restrict__ M, long mPlane, const unsigned *__restrict__ mKey, const unsigned *mCount, int onlyOver,
		double *__restrict__ B, long bPlane, unsigned *cursor){
	const long n = *mCount;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st){
		if(mKey[i] == SLOT_EMPTY_KEY || (onlyOver && mKey[i] != SLOT_OVER_KEY)) continue;
		const unsigned d = atomicAdd(cursor, 1u);
		#pragma unroll
		for(int w = 0; w < 6; w++) B[d + w*bPlane] = M[i + w*mPlane];
	}
}
__global__ void k_max_run(const unsigned *__restrict__ cellStart, long nCells, unsigned *out){
	unsigned m = 0;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < nCells; i += st) m = max(m, cellStart[i+1] - cellStart[i]);
	for(int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
	if((threadIdx.x & 31) == 0) atomicMax(out, m);
}
// deposition over the slots: k_distr_cells with the cell's run taken from the slot array
__global__ void __launch_bounds__(256, 4) k_distr_slots(SlotPar Q, CellSpace C, long sx, long sxy, long long *__restrict__ fix){
	int lane = threadIdx.x & 31;
	long warp = (blockIdx.x*(long)blockDim.x + threadIdx.x) >> 5;
	long nWarps = ((long)gridDim.x*blockDim.x) >> 5;
	for(long c = warp; c < C.nCells; c += nWarps){
		const unsigned e = Q.cnt[c];
		if(e == 0) continue;
		const double *X = Q.S + Q.off + c*(long)Q.cap, *Y = X + Q.plane, *Z = Y + Q.plane;
		long long a[8];
		#pragma unroll
		for(int q = 0; q < 8; q++) a[q] = 0;
		const int cj = (int)(c % C.nc0); const long cr = c / C.nc0; const int ck = (int)(cr % C.nc1); const int cl = (int)(cr / C.nc1);
		const double dj = (double)cj, dk = (double)ck, dl = (double)cl;
		for(unsigned i0 = 0; i0 < e; i0 += 96){
			double x[3], y[3], z[3];
			#pragma unroll
			for(int u = 0; u < 3; u++){
				unsigned i = i0 + lane + 32*u;
				bool ok = i < e;
				x[u] = ok ? X[i] : 0.0; y[u] = ok ? Y[i] : 0.0; z[u] = ok ? Z[i] : 0.0;
			}
			#pragma unroll
			for(int u = 0; u < 3; u++){
				if(i0 + lane + 32*u >= e) continue;
				const double xf = x[u]-dj, yf = y[u]-dk, zf = z[u]-dl;
				const double xc = 1-xf, yc = 1-yf, zc = 1-zf;
				const double cc = xc*yc, fc = xf*yc, cf = xc*yf, ff = xf*yf;
				a[0] += fixwRaw(cc*zc); a[1] += fixwRaw(fc*zc); a[2] += fixwRaw(cf*zc); a[3] += fixwRaw(ff*zc);
				a[4] += fixwRaw(cc*zf); a[5] += fixwRaw(fc*zf); a[6] += fixwRaw(cf*zf); a[7] += fixwRaw(ff*zf);
			}
		}
		long long tot = warpCornerTotal(a);
		tot -= (long long)((unsigned long long)e*FIXW_BIAS);          // (see k_distr_cells)
		if((lane & 3) == 0 && tot != 0)
			atomicAdd((unsigned long long*)&fix[cj + sx*ck + sxy*cl + cornerOffset(sx, sxy)], (unsigned long long)tot);
	}
}

// ---- the N-dimensional select() targets of the accelerator and the distributor, for nDims = 3 (src/pusher.c:215-391, 574-678):
// puAccND1[KE] / puDistrND1 are first order like the 3D1 forms but evaluate the trilinear weights in the order of the
// reference's recursion (puInterpND1Inner :1147, puDistrND1Inner :626: factor = (y weight)*(z weight), then (x weight*factor)*E,
// corners visited x fastest, z slowest, each term added to the running sum); puAccND0[KE] / puDistrND0 are zeroth order
// (nearest grid point, (int)(pos + 0.5), :1164, :644).  One thread per particle, per-particle integer REDs for the deposit:
// these are the slow-path variants, the 3D1 forms are the ones the time loop is built around.
template<int ORDER, int KE> __global__ void __launch_bounds__(256) k_acc_nd(double *__restrict__ P, long cap, long a, long n,
		const double *__restrict__ E, long sx3, long sxy3, int gs0, int gs1, int gs2, double *__restrict__ partial, int *flags){
	double acc = 0;
	long stride = (long)gridDim.x*blockDim.x;
	for(long i = blockIdx.x*(long)blockDim.x + threadIdx.x; i < n; i += stride){
		const long q = a + i;
		const double x = P[q], y = P[q+cap], z = P[q+2*cap];
		double vx = P[q+3*cap], vy = P[q+4*cap], vz = P[q+5*cap];
		double dv[3] = {0, 0, 0};
		if(ORDER == 0){
			int j = (int)(x+0.5), k = (int)(y+0.5), l = (int)(z+0.5);
			if(!(x >= -0.5) || !(y >= -0.5) || !(z >= -0.5) || j > gs0-1 || k > gs1-1 || l > gs2-1){ atomicOr(flags, ERR_POS_RANGE); j = min(max(j,0),gs0-1); k = min(max(k,0),gs1-1); l = min(max(l,0),gs2-1); }
			const double *e = E + 3L*j + k*sx3 + l*sxy3;
			dv[0] = __ldg(e); dv[1] = __ldg(e+1); dv[2] = __ldg(e+2);
		} else {
			int j = (int)x, k = (int)y, l = (int)z;
			if(!(x >= 0) || !(y >= 0) || !(z >= 0) || j > gs0-2 || k > gs1-2 || l > gs2-2){ atomicOr(flags, ERR_POS_RANGE); j = min(max(j,0),gs0-2); k = min(max(k,0),gs1-2); l = min(max(l,0),gs2-2); }
			const double xf = x-j, yf = y-k, zf = z-l, xc = 1-xf, yc = 1-yf, zc = 1-zf;
			const double *e = E + 3L*j + k*sx3 + l*sxy3;
			#pragma unroll
			for(int dz = 0; dz < 2; dz++){
				const double fz = (dz ? zf : zc)*1.0;
				#pragma unroll
				for(int dy = 0; dy < 2; dy++){
					const double f = (dy ? yf : yc)*fz;
					const double *ep = e + dy*sx3 + dz*sxy3;
					#pragma unroll
					for(int d = 0; d < 3; d++){
						dv[d] += xc*f*__ldg(ep+d);
						dv[d] += xf*f*__ldg(ep+3+d);
					}
				}
			}
		}
		if(KE){
			double v2 = 0;
			v2 += vx*(vx+dv[0]); v2 += vy*(vy+dv[1]); v2 += vz*(vz+dv[2]);
			acc += v2;
		}
		vx += dv[0]; vy += dv[1]; vz += dv[2];
		P[q+3*cap] = vx; P[q+4*cap] = vy; P[q+5*cap] = vz;
	}
	if(KE){
		acc = blockSumP<256>(acc);
		if(threadIdx.x == 0) partial[blockIdx.x] = acc;
	}
}
template<int ORDER> __global__ void __launch_bounds__(256) k_distr_nd(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
		long n, int s0, int s1, int s2, long long *__restrict__ fix, int *flags){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	const long sx = s0, sxy = (long)s0*s1;
	for(; i < n; i += st){
		const double x = X[i], y = Y[i], z = Z[i];
		if(ORDER == 0){
			int j = (int)(x+0.5), k = (int)(y+0.5), l = (int)(z+0.5);
			if(!(x >= -0.5) || !(y >= -0.5) || !(z >= -0.5) || j > s0-1 || k > s1-1 || l > s2-1){ atomicOr(flags, ERR_POS_RANGE); continue; }
			atomicAdd((unsigned long long*)fix + (j + sx*k + sxy*l), (unsigned long long)fixw(1.0));
		} else {
			int j = (int)x, k = (int)y, l = (int)z;
			if(!(x >= 0) || !(y >= 0) || !(z >= 0) || j > s0-2 || k > s1-2 || l > s2-2){ atomicOr(flags, ERR_POS_RANGE); continue; }
			const double xf = x-j, yf = y-k, zf = z-l, xc = 1-xf, yc = 1-yf, zc = 1-zf;
			unsigned long long *f = (unsigned long long*)fix + (j + sx*k + sxy*l);
			#pragma unroll
			for(int dz = 0; dz < 2; dz++){
				const double fz = (dz ? zf : zc)*1.0;
				#pragma unroll
				for(int dy = 0; dy < 2; dy++){
					const double fy = (dy ? yf : yc)*fz;
					atomicAdd(f + dy*sx + dz*sxy,     (unsigned long long)fixw(xc*fy));
					atomicAdd(f + dy*sx + dz*sxy + 1, (unsigned long long)fixw(xf*fy));
				}
			}
		}
	}
}

// ---- debug scans of the driver loop (src/population.c:316-365) ------------------------------------------------
// three planes starting at `P` are compared with lo <= v <= hi[d]; the first offender is recorded (species-local index,
// dimension) by an atomicMin on index*4+dimension
__global__ void k_assert_range(const double *__restrict__ P, long cap, long n, double lo0, double lo1, double lo2,
		double hi0, double hi1, double hi2, int useLo, unsigned long long *first){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	const double lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
	for(; i < n; i += st)
		#pragma unroll
		for(int d = 0; d < 3; d++){
			double v = P[i + d*cap];
			if(v > hi[d] || (useLo && v < lo[d])) atomicMin(first, (unsigned long long)i*4 + d);
		}
}

// ---- host side ----------------------------------------------------------------------------------------------
static inline int pGrid(Ctx *c, long n){ return gridFor(n, 256, c->numSMs*8); }

static Thr thrOf(const double *thr6){
	Thr T;
	for(int d = 0; d < 3; d++){ T.lo[d] = thr6[d]; T.up[d] = thr6[3+d]; }
	return T;
}
static Thr thrOf(const MpiInfo *m){ return thrOf(m->thresholds); }
static void setupCells(Ctx *c, DevPop *dp, const double *thr6);
static void setupCells(Ctx *c, DevPop *dp, const MpiInfo *m){ setupCells(c, dp, m->thresholds); }
// the cell space every non-emigrant position falls into: 0 <= lower threshold, x < upper threshold <= nc
static void setupCells(Ctx *c, DevPop *dp, const double *thr6){
	int nc[3];
	for(int d = 0; d < 3; d++){
		if(thr6[d] < 0) fatal("negative lower migration threshold");
		nc[d] = (int)ceil(thr6[3+d]);
		if(nc[d] < 1) nc[d] = 1;
		dp->thr6[d] = thr6[d]; dp->thr6[3+d] = thr6[3+d];
	}
	dp->haveThr = true;
	long nCells = (long)nc[0]*nc[1]*nc[2];
	if(nCells + 28 >= 0xffffffffL) fatal("cell space too large for 32-bit keys");
	if(dp->nCells == nCells && dp->nc[0] == nc[0] && dp->nc[1] == nc[1] && dp->d_hist[0]) return;
	streamSync(c);
	if(dp->slotted) fatal("setupCells: the cell space changed under a slotted population");
	for(int s = 0; s < dp->nS; s++){
		if(dp->d_hist[s]) cudaFree(dp->d_hist[s]);
		if(dp->d_cursor[s]) cudaFree(dp->d_cursor[s]);
		if(dp->d_cnt[s]){ cudaFree(dp->d_cnt[s]); dp->d_cnt[s] = nullptr; }
		PINC_CUDA(cudaMalloc(&dp->d_hist[s], (size_t)(nCells+28)*sizeof(unsigned)));
		PINC_CUDA(cudaMalloc(&dp->d_cursor[s], (size_t)(nCells+28)*sizeof(unsigned)));
		dp->sortedN[s] = 0;
	}
	for(int d = 0; d < 3; d++) dp->nc[d] = nc[d];
	dp->nCells = nCells;
	dp->keysValid = false;
}
// pNew / pCut (src/population.c:430-466): one particle in or out of the SoA planes
__global__ void k_particle_put(double *__restrict__ P, long cap, long i, double x, double y, double z, double vx, double vy, double vz){
	if(threadIdx.x == 0){ P[i] = x; P[i + cap] = y; P[i + 2*cap] = z; P[i + 3*cap] = vx; P[i + 4*cap] = vy; P[i + 5*cap] = vz; }
}
__global__ void k_particle_cut(double *__restrict__ P, long cap, long i, long last, double *__restrict__ out){
	if(threadIdx.x < 6){ long o = threadIdx.x*cap; out[threadIdx.x] = P[i + o]; P[i + o] = P[last + o]; }
}
static CellSpace cellsOf(const DevPop *dp){ return CellSpace{ dp->nc[0], dp->nc[1], dp->nc[2], dp->nCells }; }

static void dropPredeposit(DevPop *dp);
static void invalidateOrder(DevPop *dp){
	dropPredeposit(dp);
	for(int s = 0; s < dp->nS; s++) dp->sortedN[s] = 0;
	dp->keysValid = false;
}

static void scanHist(Ctx *c, unsigned *h, long n);
// ---- slotted mode, host side --------------------------------------------------------------------------------
static int g_slotted = -1;              // $PINC_B200_SLOTTED=0 keeps the counting sort every step
static int g_slotHeadroom = 25;         // per cent of the fullest cell, plus g_slotExtra slots, kept free in every cell
static int g_slotExtra = 16;
static std::atomic<long> g_slotOverflows{0};   // how often a full cell sent a population back to the contiguous layout (rank threads share it)
static bool slottedEnabled(){
	if(g_slotted < 0){
		g_slotted = (getenv("PINC_B200_SLOTTED") && atoi(getenv("PINC_B200_SLOTTED")) == 0) ? 0 : 1;
		int h, e;
		if(getenv("PINC_B200_SLOT_HEADROOM") && sscanf(getenv("PINC_B200_SLOT_HEADROOM"), "%d,%d", &h, &e) == 2 && h >= 0 && e >= 0){ g_slotHeadroom = h; g_slotExtra = e; }
	}
	return g_slotted != 0;
}
static SlotPar slotPar(const DevPop *dp, int s){ return SlotPar{ dp->slot, dp->slotPlane, dp->slotOff[s], dp->slotCapS[s], dp->d_cnt[s] }; }
static double *mvBase(const DevPop *dp, int s){ return dp->alt + dp->host->iStart[s]; }         // six planes of stride dp->cap
static unsigned *mvKeys(const DevPop *dp, int s){ return dp->d_keys + dp->host->iStart[s]; }
enum { MV_COUNT = 0, MV_EMHIST = 8, MV_SCRATCH = 8 + 8*27, MV_WORDS = 8 + 8*27 + 16 };
static int cellBlocks(Ctx *c, long nCells){ return gridFor(nCells*32, 256, c->numSMs*8); }
// did a cell run out of slots since the last call?  (other device errors stay fatal)
static bool takeSlotOverflow(Ctx *c, const char *where){
	PINC_CUDA(cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	streamSync(c);
	int f = c->h_flags[0];
	if(!(f & ERR_SLOT_OVERFLOW)){ if(f) checkDeviceFlags(c, where); return false; }
	if(f & ~ERR_SLOT_OVERFLOW) checkDeviceFlags(c, where);
	PINC_CUDA(cudaMemsetAsync(c->d_flags, 0, sizeof(int), c->stream));
	return true;
}
// slots -> contiguous cell-ordered planes (+ whatever waits in the mover lists as an unsorted tail): the state the sorted
// path leaves, so every other entry point works unchanged
void popLeaveSlotted(Ctx *c, DevPop *dp){
	if(!dp->slotted) return;
	dropPredeposit(dp);           // what pincAccMoveDistr3D1KE deposited from the slots is redone by the next puDistr3D1 (accumulators cleared first)
	const Population *pop = dp->host;
	const long nKeys = dp->nCells + 27;
	for(int s = 0; s < dp->nS; s++){
		const long a = pop->iStart[s];
		PINC_LAUNCH(c, K_SORT, 8.0*nKeys, (k_cnt_to_hist<<<gridFor(nKeys+1,256,c->numSMs*8),256,0,c->stream>>>(dp->d_cnt[s], dp->d_hist[s], dp->nCells, nKeys+1, dp->slotCapS[s])));
		scanHist(c, dp->d_hist[s], nKeys+1);
		long nIn = pop->iStop[s] - a;
		if(nIn > 0) PINC_LAUNCH(c, K_SORT, 96.0*nIn, (k_cells_copy<<<cellBlocks(c,dp->nCells),256,0,c->stream>>>(slotPar(dp,s), dp->base + a, dp->cap, dp->d_hist[s], dp->nCells, 0)));
		unsigned *cursor = dp->d_mvCount + MV_SCRATCH + s;
		PINC_CUDA(cudaMemcpyAsync(cursor, dp->d_hist[s] + dp->nCells, sizeof(unsigned), cudaMemcpyDeviceToDevice, c->stream));
		PINC_LAUNCH(c, K_SORT, 96.0, (k_mv_append<<<pGrid(c, nIn/8 + 1024),256,0,c->stream>>>(mvBase(dp,s), dp->cap, mvKeys(dp,s), dp->d_mvCount + MV_COUNT + s,
			dp->mvPending ? 0 : 1, dp->base + a, dp->cap, cursor)));
		PINC_CUDA(cudaMemcpyAsync(c->h_long + 2*s, dp->d_hist[s] + dp->nCells, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
		PINC_CUDA(cudaMemcpyAsync((unsigned*)(c->h_long + 2*s) + 1, cursor, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
	}
	streamSync(c);
	for(int s = 0; s < dp->nS; s++){
		const unsigned *h = (const unsigned*)(c->h_long + 2*s);
		long n = pop->iStop[s] - pop->iStart[s];
		if((long)h[1] != n) fatal("slotted population of species %d holds %u particles, the host counts %ld", s, h[1], n);
		dp->sortedN[s] = h[0];
	}
	dp->slotted = false;
	if(dp->mvPending){ dp->keysValid = false; dp->mvPending = false; }
}
// contiguous cell-ordered planes -> slots; false (population unchanged) if it is not binned yet or a cell would not fit
static bool enterSlotted(Ctx *c, DevPop *dp, const double *thr6){
	const Population *pop = dp->host;
	if(!slottedEnabled() || dp->extracted || dp->predep || dp->slotKicks >= 3) return false;       // (a host whose loop keeps calling entry points without a slotted form gains nothing)
	{ double t[6]; for(int d = 0; d < 6; d++) t[d] = thr6[d]; setupCells(c, dp, t); }
	for(int s = 0; s < dp->nS; s++) if(pop->iStop[s] > pop->iStart[s] && dp->sortedN[s] == 0) return false;       // the sort has to run once
	if(!dp->alt) PINC_CUDA(cudaMalloc(&dp->alt, (size_t)6*(dp->cap > 0 ? dp->cap : 1)*sizeof(double)));
	if(!dp->d_mvCount) PINC_CUDA(cudaMalloc(&dp->d_mvCount, MV_WORDS*sizeof(unsigned)));
	PINC_CUDA(cudaMemsetAsync(dp->d_mvCount, 0, MV_WORDS*sizeof(unsigned), c->stream));
	for(int s = 0; s < dp->nS; s++)
		if(dp->sortedN[s] > 0) PINC_LAUNCH(c, K_SORT, 4.0*dp->nCells, (k_max_run<<<gridFor(dp->nCells,256,c->numSMs*8),256,0,c->stream>>>(dp->d_hist[s], dp->nCells, dp->d_mvCount + MV_SCRATCH + s)));
	PINC_CUDA(cudaMemcpyAsync(c->h_long, dp->d_mvCount + MV_SCRATCH, 8*sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
	streamSync(c);
	long total = 0;
	for(int s = 0; s < dp->nS; s++){
		long mx = ((const unsigned*)c->h_long)[s];
		long cap = mx + mx*g_slotHeadroom/100 + g_slotExtra;       // head room for fluctuations: a cell that outgrows it sends the population back to the sort
		cap = (cap + 3) & ~3L;
		dp->slotCapS[s] = (int)cap;
		dp->slotOff[s] = total;
		total += dp->nCells*cap;
	}
	dp->slotOff[dp->nS] = total;
	if(total > dp->slotPlane){
		size_t freeB = 0, totB = 0;
		cudaMemGetInfo(&freeB, &totB);
		size_t have = dp->slot ? (size_t)6*dp->slotPlane*sizeof(double) : 0;
		if((size_t)6*total*sizeof(double) > (freeB + have)/2) return false;        // not worth squeezing the device for
		if(dp->slot){ PINC_CUDA(cudaFree(dp->slot)); dp->slot = nullptr; dp->slotPlane = 0; }
		long want = (total + total/16 + 3) & ~3L;
		if(cudaMalloc(&dp->slot, (size_t)6*want*sizeof(double)) != cudaSuccess){ cudaGetLastError(); return false; }
		dp->slotPlane = want;
	}
	PINC_CUDA(cudaMemsetAsync(dp->d_mvCount, 0, MV_WORDS*sizeof(unsigned), c->stream));
	for(int s = 0; s < dp->nS; s++){
		if(!dp->d_cnt[s]) PINC_CUDA(cudaMalloc(&dp->d_cnt[s], (size_t)(dp->nCells+28)*sizeof(unsigned)));
		PINC_CUDA(cudaMemsetAsync(dp->d_cnt[s], 0, (size_t)(dp->nCells+28)*sizeof(unsigned), c->stream));
		const long a = pop->iStart[s], n = pop->iStop[s] - a, ns = dp->sortedN[s] < n ? dp->sortedN[s] : n;
		if(ns > 0) PINC_LAUNCH(c, K_SORT, 96.0*ns, (k_cells_copy<<<cellBlocks(c,dp->nCells),256,0,c->stream>>>(slotPar(dp,s), dp->base + a, dp->cap, dp->d_hist[s], dp->nCells, 1)));
		if(n - ns > 0) PINC_LAUNCH(c, K_SORT, 96.0*(n-ns), (k_tail_insert<<<pGrid(c,n-ns),256,0,c->stream>>>(slotPar(dp,s), dp->base + a + ns, dp->cap, n - ns, cellsOf(dp), c->d_flags)));
	}
	if(takeSlotOverflow(c, "entering slotted mode")){ g_slotOverflows++; return false; }
	dp->slotted = true; dp->mvPending = false; dp->emigInMovers = false;
	return true;
}
// warps k_cell_push may run for species s: the mover list (the species' share of the second buffer) must hold every particle
// plus one unused chunk per warp
static long slotWarpsFor(const Population *pop, int s){
	long room = (pop->iStart[s+1] - pop->iStart[s]) - (pop->iStop[s] - pop->iStart[s]);
	return room/(long)MV_CHUNK;
}
static bool slotRoom(const DevPop *dp){
	const Population *pop = dp->host;
	for(int s = 0; s < dp->nS; s++) if(pop->iStop[s] > pop->iStart[s] && slotWarpsFor(pop, s) < 64) return false;
	return pop->iStart[dp->nS] < 0xfffffff0L;
}
// pincAccMove3D1KE on the slots: kick + move + re-binning of every species, one warp per cell
// mode 0: kick + move + re-binning (pincAccMove3D1KE); 1: kick only (puAcc3D1[KE]); 2: move + re-binning only (puMove)
// rho != nullptr (mode 0 only): the stayers are deposited into rho's accumulators by the push (pincAccMoveDistr3D1KE)
static void cellPush(Ctx *c, DevPop *dp, Population *pop, DevGrid *E, int ke, const double *thr6, int mode, DevGrid *rho = nullptr){
	const long sx3 = E ? 3L*E->size[0] : 0, sxy3 = E ? sx3*E->size[1] : 0;
	if(E && (dp->nc[0] > E->size[0]-1 || dp->nc[1] > E->size[1]-1 || dp->nc[2] > E->size[2]-1)) fatal("accelerator: E is smaller than the migration thresholds allow");
	const Thr thr = thrOf(thr6); const CellSpace C = cellsOf(dp);
	if(mode == 2) ke = 0;
	if(rho && mode != 0) fatal("cellPush: the fused deposition belongs to the fused push");
	const int maxBlocks = cellBlocks(c, dp->nCells);
	double *partial = ke ? partialBuffer(c, (long)maxBlocks*dp->nS) : nullptr;
	if(mode != 1) PINC_CUDA(cudaMemsetAsync(dp->d_mvCount, 0, MV_WORDS*sizeof(unsigned), c->stream));
	for(int s = 0; s < dp->nS; s++){
		const long n = pop->iStop[s] - pop->iStart[s];
		if(E) gridScale(c, E, pop->charge[s]/pop->mass[s]);          // quirk Q2, as accelerate()
		double *part = ke ? partial + (long)s*maxBlocks : nullptr;
		// every warp may leave up to one chunk of the mover list unused: as many warps as the species' allocation has room for
		const long mCap = pop->iStart[s+1] - pop->iStart[s];
		int blocks = (int)std::min<long>(maxBlocks, slotWarpsFor(pop, s)/8);
		if(n > 0){
#define CELL_LAUNCH(KEE,MODE,DEP,CLS,BYTES) PINC_LAUNCH(c, CLS, (BYTES)*n, (k_cell_push<KEE,MODE,DEP><<<blocks,256,0,c->stream>>>(slotPar(dp,s), E ? E->d : nullptr, sx3, sxy3, C, thr, part, mvBase(dp,s), dp->cap, mvKeys(dp,s), dp->d_mvCount + MV_COUNT + s, (unsigned)mCap, c->d_flags, \
				rho ? rho->d_fixS[s] : nullptr, rho ? (long)rho->size[0] : 0, rho ? (long)rho->size[0]*rho->size[1] : 0)))
			if(mode == 0 && rho){ if(ke) CELL_LAUNCH(1,0,1,K_PUSH,120.0); else CELL_LAUNCH(0,0,1,K_PUSH,120.0); }
			else if(mode == 0){ if(ke) CELL_LAUNCH(1,0,0,K_PUSH,96.0); else CELL_LAUNCH(0,0,0,K_PUSH,96.0); }
			else if(mode == 1){ if(ke) CELL_LAUNCH(1,1,0,K_PUSH,72.0); else CELL_LAUNCH(0,1,0,K_PUSH,72.0); }
			else CELL_LAUNCH(0,2,0,K_MOVE,72.0);
#undef CELL_LAUNCH
		}
		if(ke) PINC_LAUNCH(c, K_REDUCE, 8.0*blocks, (k_final_sum_p<<<1,256,0,c->stream>>>(part, n > 0 ? blocks : 0, c->d_scal + 16 + s)));
		if(E) gridScale(c, E, pop->mass[s]/pop->charge[s]);
	}
	if(mode != 1){
		dp->mvPending = true;
		for(int d = 0; d < 6; d++) dp->keyThr[d] = thr6[d];
	}
	if(ke){
		PINC_CUDA(cudaMemcpyAsync(c->h_scal + 16, c->d_scal + 16, dp->nS*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
		for(int s = 0; s < dp->nS; s++){
			pop->kinEnergy[s] = c->h_scal[16+s];
			pop->kinEnergy[s] *= 0.5*pop->mass[s];
		}
	}
}
// puExtractEmigrants3D on the slots: the local movers drop into their cells, the emigrants are counted and wait in the list
static void cellExtract(Ctx *c, DevPop *dp, Population *pop, MpiInfo *m){
	const int nS = dp->nS;
	const CellSpace C = cellsOf(dp);
	for(int s = 0; s < nS; s++){
		const long n = pop->iStop[s] - pop->iStart[s];
		DevGrid *rho = dp->predep;          // pincAccMoveDistr3D1KE deposited the stayers: the local movers follow here
		if(n > 0) PINC_LAUNCH(c, K_EXTRACT, 100.0*(n/16), (k_mv_insert<<<pGrid(c, n/8 + 1024),256,0,c->stream>>>(slotPar(dp,s), mvBase(dp,s), dp->cap, mvKeys(dp,s), dp->d_mvCount + MV_COUNT + s,
			C, dp->d_mvCount + MV_EMHIST + 27*s, c->d_flags, rho ? rho->d_fixS[s] : nullptr, rho ? (long)rho->size[0] : 0, rho ? (long)rho->size[0]*rho->size[1] : 0)));
	}
	PINC_CUDA(cudaMemcpyAsync(c->h_long, dp->d_mvCount, MV_SCRATCH*sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
	const bool over = takeSlotOverflow(c, "puExtractEmigrants3D");          // (synchronises)
	const unsigned *h = (const unsigned*)c->h_long;
	dp->mvPending = false;
	dp->emigInMovers = true;
	dp->extracted = true;
	for(int s = 0; s < nS; s++){
		long em = 0;
		for(int ne = 0; ne < 27; ne++){ m->nEmigrants[ne*nS+s] = h[MV_EMHIST + 27*s + ne]; em += h[MV_EMHIST + 27*s + ne]; }
		pop->iStop[s] -= em;
	}
	if(over){ g_slotOverflows++; popLeaveSlotted(c, dp); }       // the movers that found their cell full become the unsorted tail of the contiguous planes
}

enum AccKind { ACC_LEAP = 0, ACC_BORIS = 1 };
static void ensureFix(Ctx *c, DevGrid *rho, int nS){
	for(int s = 0; s < nS; s++) if(!rho->d_fixS[s]){
		PINC_CUDA(cudaMalloc(&rho->d_fixS[s], (size_t)rho->n*sizeof(long long)));
		PINC_CUDA(cudaMemsetAsync(rho->d_fixS[s], 0, (size_t)rho->n*sizeof(long long), c->stream));
	}
	if(rho->fixDirty){
		for(int s = 0; s < nS; s++) PINC_CUDA(cudaMemsetAsync(rho->d_fixS[s], 0, (size_t)rho->n*sizeof(long long), c->stream));
		rho->fixDirty = false;
	}
}
static void dropPredeposit(DevPop *dp){ if(dp->predep){ dp->predep->fixDirty = true; dp->predep = nullptr; } }

static void accelerate(Ctx *c, Population *pop, Grid *Egrid, int kind, int ke, const double *T, const double *S, const MpiInfo *fuse, Grid *rhoGrid = nullptr){
	DevPop *dp = devPopRaw(c, pop);
	DevGrid *E = devGrid(c, Egrid);
	if(E->nv != 3) fatal("accelerator needs a 3-vector field grid");
	{	// slotted mode: the steady state of pincAccMove3D1KE (leapfrog kick + move + re-binning, no fused deposition)
		// (with `fuse` also the move and the re-binning; without, the kick alone - the reference's call order)
		// (with rho - pincAccMoveDistr3D1KE - the stayers are deposited by the push as well)
		bool eligible = kind == ACC_LEAP && (!rhoGrid || fuse) && slottedEnabled() && (fuse || dp->haveThr);
		if(dp->slotted && (!eligible || dp->mvPending || dp->extracted || !slotRoom(dp))) popLeaveSlotted(c, dp);
		if(eligible && !dp->slotted && slotRoom(dp)) enterSlotted(c, dp, fuse ? fuse->thresholds : dp->thr6);
		if(dp->slotted){
			double t[6]; for(int d = 0; d < 6; d++) t[d] = fuse ? fuse->thresholds[d] : dp->thr6[d];
			DevGrid *rho = nullptr;
			dropPredeposit(dp);
			if(rhoGrid){
				rho = devGrid(c, rhoGrid);
				if(rho->nv != 1) fatal("pincAccMoveDistr3D1KE needs a scalar grid for rho");
				if(dp->nc[0] > rho->size[0]-1 || dp->nc[1] > rho->size[1]-1 || dp->nc[2] > rho->size[2]-1) fatal("pincAccMoveDistr3D1KE: rho is smaller than the migration thresholds allow");
				ensureFix(c, rho, dp->nS);
			}
			cellPush(c, dp, pop, E, ke, t, fuse ? 0 : 1, rho);
			dp->predep = rho;
			return;
		}
	}
	long sx3 = 3L*E->size[0], sxy3 = sx3*E->size[1];
	Thr thr{}; CellSpace C{1,1,1,1};
	DevGrid *rho = nullptr;
	if(fuse) dropPredeposit(dp);
	if(fuse && rhoGrid){
		rho = devGrid(c, rhoGrid);
		if(rho->nv != 1) fatal("pincAccMoveDistr3D1KE needs a scalar grid for rho");
		ensureFix(c, rho, dp->nS);
	}
	if(fuse){
		setupCells(c, dp, fuse);
		thr = thrOf(fuse); C = cellsOf(dp);
		for(int s = 0; s < dp->nS; s++)
			PINC_CUDA(cudaMemsetAsync(dp->d_hist[s], 0, (size_t)(dp->nCells+28)*sizeof(unsigned), c->stream));
	}
	int maxBlocks = 0;
	for(int s = 0; s < dp->nS; s++){ int b = pGrid(c, pop->iStop[s]-pop->iStart[s]); if(b > maxBlocks) maxBlocks = b; }
	double *partial = ke ? partialBuffer(c, (long)maxBlocks*dp->nS) : nullptr;
	for(int s = 0; s < dp->nS; s++){
		long a = pop->iStart[s], n = pop->iStop[s] - a;
		// the reference rescales the whole field by q/m and back for every species (quirk Q2); same here
		gridScale(c, E, pop->charge[s]/pop->mass[s]);
		BorisPar B{};
		if(kind == ACC_BORIS) for(int d = 0; d < 3; d++){ B.T[d] = T[3*s+d]; B.S[d] = S[3*s+d]; }
		int blocks = pGrid(c, n);
		double *part = ke ? partial + (long)s*maxBlocks : nullptr;
		if(n > 0){
			double bytes = (fuse ? 100.0 : 72.0)*n;
#define ACC_LAUNCH(K,KEE,F) PINC_LAUNCH(c, K_PUSH, bytes, (k_acc<K,KEE,F><<<blocks,256,0,c->stream>>>(dp->base, dp->cap, a, n, E->d, sx3, sxy3, E->size[0], E->size[1], E->size[2], B, part, thr, C, dp->d_keys, fuse ? dp->d_hist[s] : nullptr, c->d_flags, rho ? rho->d_fixS[s] : nullptr, rho ? (long)rho->size[0] : 0, rho ? (long)rho->size[0]*rho->size[1] : 0)))
			if(kind == ACC_LEAP && rho){
				if(dp->nc[0] > rho->size[0]-1 || dp->nc[1] > rho->size[1]-1 || dp->nc[2] > rho->size[2]-1) fatal("pincAccMoveDistr3D1KE: rho is smaller than the migration thresholds allow");
				bytes = 124.0*n;
				ACC_LAUNCH(0,1,2);
			} else if(kind == ACC_LEAP){
				if(ke){ if(fuse) ACC_LAUNCH(0,1,1); else ACC_LAUNCH(0,1,0); }
				else  { if(fuse) ACC_LAUNCH(0,0,1); else ACC_LAUNCH(0,0,0); }
			} else {
				if(ke){ if(fuse) ACC_LAUNCH(1,1,1); else ACC_LAUNCH(1,1,0); }
				else  { if(fuse) ACC_LAUNCH(1,0,1); else ACC_LAUNCH(1,0,0); }
			}
#undef ACC_LAUNCH
		}
		if(ke) PINC_LAUNCH(c, K_REDUCE, 8.0*blocks, (k_final_sum_p<<<1,256,0,c->stream>>>(part, n > 0 ? blocks : 0, c->d_scal + 16 + s)));
		gridScale(c, E, pop->mass[s]/pop->charge[s]);
	}
	if(fuse){
		for(int s = 0; s < dp->nS; s++) dp->sortedN[s] = 0;
		dp->keysValid = true;
		for(int d = 0; d < 6; d++) dp->keyThr[d] = fuse->thresholds[d];
		dp->predep = rho;
	}
	if(ke){
		PINC_CUDA(cudaMemcpyAsync(c->h_scal + 16, c->d_scal + 16, dp->nS*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
		for(int s = 0; s < dp->nS; s++){
			pop->kinEnergy[s] = c->h_scal[16+s];
			pop->kinEnergy[s] *= 0.5*pop->mass[s];
		}
	}
}

// puAccND1[KE], puAccND0[KE] (src/pusher.c:215-391) for nDims = 3
static void accelerateND(Ctx *c, Population *pop, Grid *Egrid, int order, int ke){
	DevPop *dp = devPop(c, pop);
	DevGrid *E = devGrid(c, Egrid);
	if(E->nv != 3) fatal("accelerator needs a 3-vector field grid");
	const long sx3 = 3L*E->size[0], sxy3 = sx3*E->size[1];
	int maxBlocks = 0;
	for(int s = 0; s < dp->nS; s++){ int b = pGrid(c, pop->iStop[s]-pop->iStart[s]); if(b > maxBlocks) maxBlocks = b; }
	double *partial = ke ? partialBuffer(c, (long)maxBlocks*dp->nS) : nullptr;
	for(int s = 0; s < dp->nS; s++){
		long a = pop->iStart[s], n = pop->iStop[s] - a;
		gridScale(c, E, pop->charge[s]/pop->mass[s]);
		int blocks = pGrid(c, n);
		double *part = ke ? partial + (long)s*maxBlocks : nullptr;
		if(n > 0){
#define ND_LAUNCH(O,K) PINC_LAUNCH(c, K_PUSH, 72.0*n, (k_acc_nd<O,K><<<blocks,256,0,c->stream>>>(dp->base, dp->cap, a, n, E->d, sx3, sxy3, E->size[0], E->size[1], E->size[2], part, c->d_flags)))
			if(order == 1){ if(ke) ND_LAUNCH(1,1); else ND_LAUNCH(1,0); } else { if(ke) ND_LAUNCH(0,1); else ND_LAUNCH(0,0); }
#undef ND_LAUNCH
		}
		if(ke) PINC_LAUNCH(c, K_REDUCE, 8.0*blocks, (k_final_sum_p<<<1,256,0,c->stream>>>(part, n > 0 ? blocks : 0, c->d_scal + 16 + s)));
		gridScale(c, E, pop->mass[s]/pop->charge[s]);
	}
	if(ke){
		PINC_CUDA(cudaMemcpyAsync(c->h_scal + 16, c->d_scal + 16, dp->nS*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
		for(int s = 0; s < dp->nS; s++){ pop->kinEnergy[s] = c->h_scal[16+s]; pop->kinEnergy[s] *= 0.5*pop->mass[s]; }
	}
	checkDeviceFlags(c, "puAccND");
}
// puDistrND1, puDistrND0 (src/pusher.c:578-678) for nDims = 3
static void distributeND(Ctx *c, const Population *pop, Grid *rhoGrid, int order){
	DevPop *dp = devPop(c, pop); DevGrid *rho = devGrid(c, rhoGrid);
	if(rho->nv != 1) fatal("puDistrND needs a scalar grid");
	dropPredeposit(dp);
	ensureFix(c, rho, dp->nS);
	gridZero(c, rho);
	for(int s = 0; s < dp->nS; s++){
		long a = pop->iStart[s], n = pop->iStop[s] - a;
		const double *X = dp->base + a, *Y = X + dp->cap, *Z = Y + dp->cap;
		if(n > 0){
			if(order == 1) PINC_LAUNCH(c, K_DEPOSIT, 24.0*n, (k_distr_nd<1><<<pGrid(c,n),256,0,c->stream>>>(X, Y, Z, n, rho->size[0], rho->size[1], rho->size[2], rho->d_fixS[s], c->d_flags)));
			else PINC_LAUNCH(c, K_DEPOSIT, 24.0*n, (k_distr_nd<0><<<pGrid(c,n),256,0,c->stream>>>(X, Y, Z, n, rho->size[0], rho->size[1], rho->size[2], rho->d_fixS[s], c->d_flags)));
		}
		PINC_LAUNCH(c, K_DEPOSIT, 32.0*rho->n, (k_distr_finalize<<<gridFor(rho->n,256,c->numSMs*8),256,0,c->stream>>>(rho->d, rho->d_fixS[s], rho->n, 1.0/pop->charge[s], pop->charge[s], c->d_flags)));
	}
}

static void scanHist(Ctx *c, unsigned *h, long n /* entries incl. the total slot */){
	int nb = (int)((n + SCAN_CH - 1)/SCAN_CH);
	unsigned *sums = (unsigned*)tmpBuffer(c, (size_t)nb*sizeof(unsigned));
	PINC_LAUNCH(c, K_SORT, 8.0*n, (k_scan_block<<<nb,256,0,c->stream>>>(h, n, sums)));
	PINC_LAUNCH(c, K_SORT, 8.0*nb, (k_scan_sums<<<1,1024,0,c->stream>>>(sums, nb)));
	PINC_LAUNCH(c, K_SORT, 8.0*n, (k_scan_add<<<(unsigned)((n+255)/256),256,0,c->stream>>>(h, n, sums)));
}

} // namespace pinc

using namespace pinc;

extern "C" {

void puMove(Population *pop, Object *obj){
	(void)obj;
	Ctx *c = cur(); DevPop *dp = devPopRaw(c, pop);
	if(dp->slotted && !dp->mvPending && !dp->extracted && dp->haveThr && slotRoom(dp)){
		double t[6]; for(int d = 0; d < 6; d++) t[d] = dp->thr6[d];
		cellPush(c, dp, pop, nullptr, 0, t, 2);          // pos += vel and the re-binning of the next puExtractEmigrants3D in one pass
		return;
	}
	if(dp->slotted) popLeaveSlotted(c, dp);
	for(int s = 0; s < dp->nS; s++){
		long a = pop->iStart[s], n = pop->iStop[s] - a;
		if(n > 0) PINC_LAUNCH(c, K_MOVE, 72.0*n, (k_move<<<pGrid(c,n),256,0,c->stream>>>(dp->base, dp->base + 3*dp->cap, dp->cap, a, n)));
	}
	invalidateOrder(dp);
}

void puAccND1(Population *pop, Grid *E){ accelerateND(cur(), pop, E, 1, 0); }            /* pusher.c:269 */
void puAccND1KE(Population *pop, Grid *E){ accelerateND(cur(), pop, E, 1, 1); }          /* pusher.c:219 */
void puAccND0(Population *pop, Grid *E){ accelerateND(cur(), pop, E, 0, 0); }            /* pusher.c:357 */
void puAccND0KE(Population *pop, Grid *E){ accelerateND(cur(), pop, E, 0, 1); }          /* pusher.c:311 */
void puDistrND1(const Population *pop, Grid *rho){ distributeND(cur(), pop, rho, 1); }   /* pusher.c:578 */
void puDistrND0(const Population *pop, Grid *rho){ distributeND(cur(), pop, rho, 0); }   /* pusher.c:644 */
void puAcc3D1(Population *pop, Grid *E){ accelerate(cur(), pop, E, ACC_LEAP, 0, nullptr, nullptr, nullptr); }
void puAcc3D1KE(Population *pop, Grid *E){ accelerate(cur(), pop, E, ACC_LEAP, 1, nullptr, nullptr, nullptr); }
// Quirk Q3: the reference rotates particle 0's velocity for every particle; this implements the Boris
// rotation of each particle's own velocity (identical for BExt = 0, which every BASELINE config has).
void puBoris3D1(Population *pop, Grid *E, const double *T, const double *S){ accelerate(cur(), pop, E, ACC_BORIS, 0, T, S, nullptr); }
void puBoris3D1KE(Population *pop, Grid *E, const double *T, const double *S){ accelerate(cur(), pop, E, ACC_BORIS, 1, T, S, nullptr); }
void pincAccMove3D1KE(Population *pop, Grid *E, MpiInfo *mpiInfo){ accelerate(cur(), pop, E, ACC_LEAP, 1, nullptr, nullptr, mpiInfo); }
void pincAccMoveDistr3D1KE(Population *pop, Grid *E, Grid *rho, MpiInfo *mpiInfo){ accelerate(cur(), pop, E, ACC_LEAP, 1, nullptr, nullptr, mpiInfo, rho); }

// src/pusher.c:485-505
void pincGet3DRotationParameters(int nSpecies, const double *BExt, const double *charge, const double *mass, double *T, double *S){
	for(int s = 0; s < nSpecies; s++){
		double factor = 0.5*charge[s]/mass[s];
		double denom = 1;
		for(int p = 0; p < 3; p++){ T[3*s+p] = factor*BExt[p]; denom += T[3*s+p]*T[3*s+p]; }
		double mul = 2.0/denom;
		for(int p = 0; p < 3; p++) S[3*s+p] = mul*T[3*s+p];
	}
}

void puExtractEmigrants3D(Population *pop, MpiInfo *mpiInfo);
// src/pusher.c:864-910: the N-dimensional form classifies by the same thresholds into the same 27 neighbours (only the order
// in which the reference back-fills the holes differs, and that order is not reproduced by either)
void puExtractEmigrantsND(Population *pop, MpiInfo *mpiInfo){ puExtractEmigrants3D(pop, mpiInfo); }
// src/pusher.c:782-855 as a counting sort (see the header comment)
void puExtractEmigrants3D(Population *pop, MpiInfo *mpiInfo){
	Ctx *c = cur(); DevPop *dp = devPopRaw(c, pop);
	if(dp->extracted) fatal("puExtractEmigrants3D called twice without puMigrate");
	if(dp->slotted){
		bool sameThr = true;
		for(int d = 0; d < 6; d++) if(dp->keyThr[d] != mpiInfo->thresholds[d]) sameThr = false;
		if(dp->mvPending && sameThr){ cellExtract(c, dp, pop, mpiInfo); return; }
		popLeaveSlotted(c, dp);
	}
	setupCells(c, dp, mpiInfo);
	if(dp->keysValid) for(int d = 0; d < 6; d++) if(dp->keyThr[d] != mpiInfo->thresholds[d]) dp->keysValid = false;
	if(!dp->alt) PINC_CUDA(cudaMalloc(&dp->alt, (size_t)6*(dp->cap > 0 ? dp->cap : 1)*sizeof(double)));
	Thr thr = thrOf(mpiInfo); CellSpace C = cellsOf(dp);
	long nKeys = dp->nCells + 27;
	int nS = dp->nS;
	for(int s = 0; s < nS; s++){
		long a = pop->iStart[s], n = pop->iStop[s] - a;
		if(!dp->keysValid){
			PINC_CUDA(cudaMemsetAsync(dp->d_hist[s], 0, (size_t)(nKeys+1)*sizeof(unsigned), c->stream));
			if(n > 0) PINC_LAUNCH(c, K_EXTRACT, 28.0*n, (k_keys<<<pGrid(c,n),256,0,c->stream>>>(dp->base, dp->cap, a, n, thr, C, dp->d_keys, dp->d_hist[s], c->d_flags)));
		}
		scanHist(c, dp->d_hist[s], nKeys+1);
		PINC_CUDA(cudaMemsetAsync(dp->d_cursor[s], 0, (size_t)nKeys*sizeof(unsigned), c->stream));
		if(n > 0) PINC_LAUNCH(c, K_SORT, 100.0*n, (k_scatter<<<pGrid(c,n),256,0,c->stream>>>(dp->base, dp->alt, dp->cap, a, n, dp->d_keys, dp->d_hist[s], dp->d_cursor[s])));
		// offsets of the 27 emigrant bins + total: 28 values
		PINC_CUDA(cudaMemcpyAsync(c->h_long + 32*s, dp->d_hist[s] + dp->nCells, 28*sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
	}
	std::swap(dp->base, dp->alt);
	streamSync(c);
	checkDeviceFlags(c, "puExtractEmigrants3D");
	for(int s = 0; s < nS; s++){
		const unsigned *o = (const unsigned*)(c->h_long + 32*s);
		long n = pop->iStop[s] - pop->iStart[s];
		if((long)o[27] != n) fatal("puExtractEmigrants3D: histogram total %u != %ld particles of species %d", o[27], n, s);
		for(int ne = 0; ne < 27; ne++) mpiInfo->nEmigrants[ne*nS+s] = (long)o[ne+1] - (long)o[ne];
		dp->sortedN[s] = o[0];
		pop->iStop[s] = pop->iStart[s] + o[0];
	}
	dp->keysValid = false;
	dp->extracted = true;
}

// src/pusher.c:914-1035: counts first, then the (x,y,z,vx,vy,vz) records; immigrants are shifted by
// (direction they came from)*trueSize and appended species by species in ascending neighbour index.
void puMigrate(Population *pop, MpiInfo *mpiInfo, Grid *grid){
	Ctx *c = cur(); DevPop *dp = devPopRaw(c, pop);
	if(!dp->extracted) fatal("puMigrate: call puExtractEmigrants3D first");
	int nS = dp->nS;
	if(27*nS > 1024/2) fatal("too many species");
	long *nEm = mpiInfo->nEmigrants, *nIm = mpiInfo->nImmigrants;
	// --- counts (exchangeNMigrants :914) ---
	if(mpiInfo->mpiSize == 1){
		for(int ne = 0; ne < 27; ne++) for(int s = 0; s < nS; s++)
			nIm[ne*nS+s] = ne == 13 ? 0 : nEm[neighborToReciprocal(ne,3)*nS+s];
	} else {
		std::vector<long> all((size_t)27*nS*mpiInfo->mpiSize);
		c->tp->allgatherLong(c, nEm, 27*nS, all.data());
		for(int ne = 0; ne < 27; ne++){
			int src = neighborToRank(mpiInfo, ne), rec = neighborToReciprocal(ne, 3);
			for(int s = 0; s < nS; s++) nIm[ne*nS+s] = ne == 13 ? 0 : all[(size_t)src*27*nS + rec*nS + s];
		}
	}
	// --- pack ---
	long emOff[28], imOff[28];
	emOff[0] = imOff[0] = 0;
	for(int ne = 0; ne < 27; ne++){
		long e = 0, i = 0;
		for(int s = 0; s < nS; s++){ e += nEm[ne*nS+s]; i += nIm[ne*nS+s]; }
		emOff[ne+1] = emOff[ne] + e; imOff[ne+1] = imOff[ne] + i;
	}
	if(emOff[27] > dp->emigCap){
		if(dp->d_emig){ streamSync(c); cudaFree(dp->d_emig); }
		dp->emigCap = emOff[27] + emOff[27]/2 + 1024;
		PINC_CUDA(cudaMalloc(&dp->d_emig, (size_t)dp->emigCap*6*sizeof(double)));
	}
	if(imOff[27] > dp->immigCap){
		if(dp->d_immig){ streamSync(c); cudaFree(dp->d_immig); }
		dp->immigCap = imOff[27] + imOff[27]/2 + 1024;
		PINC_CUDA(cudaMalloc(&dp->d_immig, (size_t)dp->immigCap*6*sizeof(double)));
	}
	if(dp->emigInMovers){
		// slotted mode: the emigrants wait in the species' mover lists (cellExtract)
		for(int s = 0; s < nS; s++){
			MvPackPar pp; long tot = 0;
			for(int ne = 0; ne < 27; ne++){
				long inMsg = 0;
				for(int s2 = 0; s2 < s; s2++) inMsg += nEm[ne*nS+s2];
				pp.dstOff[ne] = emOff[ne] + inMsg; tot += nEm[ne*nS+s];
			}
			if(tot <= 0) continue;
			PINC_CUDA(cudaMemsetAsync(dp->d_cursor[s], 0, 27*sizeof(unsigned), c->stream));
			long cap = pop->iStart[s+1] - pop->iStart[s];
			PINC_LAUNCH(c, K_EXTRACT, 96.0*tot, (k_mv_pack<<<pGrid(c, cap/8 + 1024),256,0,c->stream>>>(mvBase(dp,s), dp->cap, mvKeys(dp,s), dp->d_mvCount + MV_COUNT + s, cellsOf(dp), pp, dp->d_cursor[s], dp->d_emig)));
		}
	} else
	for(int s = 0; s < nS; s++){
		PackPar pp;
		long run = pop->iStop[s] - pop->iStart[s];          // emigrants start right behind the stayers
		long tot = 0;
		for(int ne = 0; ne < 27; ne++){
			pp.srcOff[ne] = run;
			long inMsg = 0;
			for(int s2 = 0; s2 < s; s2++) inMsg += nEm[ne*nS+s2];
			pp.dstOff[ne] = emOff[ne] + inMsg;
			run += nEm[ne*nS+s]; tot += nEm[ne*nS+s];
		}
		pp.srcOff[27] = run;
		if(tot > 0) PINC_LAUNCH(c, K_EXTRACT, 96.0*tot, (k_pack_emigrants<<<pGrid(c,tot),256,0,c->stream>>>(dp->base, dp->cap, pop->iStart[s], tot, pp, dp->d_emig)));
	}
	// --- exchange (exchangeMigrants :988) ---
	std::vector<Msg> sends, recvs;
	for(int ne = 0; ne < 27; ne++){
		if(ne == 13) continue;
		int peer = neighborToRank(mpiInfo, ne);
		sends.push_back({peer, neighborToReciprocal(ne,3), dp->d_emig + 6*emOff[ne], (size_t)(emOff[ne+1]-emOff[ne])*48});
		recvs.push_back({peer, ne, dp->d_immig + 6*imOff[ne], (size_t)(imOff[ne+1]-imOff[ne])*48});
	}
	c->tp->exchange(c, sends, recvs);
	// --- import (shiftImmigrants :941, importParticles :967) ---
	for(int s = 0; s < nS; s++){
		ImportPar ip;
		long tot = 0;
		for(int ne = 0; ne < 27; ne++){
			long inMsg = 0;
			for(int s2 = 0; s2 < s; s2++) inMsg += nIm[ne*nS+s2];
			ip.srcOff[ne] = imOff[ne] + inMsg;
			ip.cnt[ne] = nIm[ne*nS+s];
			tot += ip.cnt[ne];
			int q = ne;
			for(int d = 0; d < 3; d++){ int n = q%3 - 1; q /= 3; ip.shift[ne][d] = (double)(n*grid->trueSize[d+1]); }
		}
		if(pop->iStop[s] + tot > pop->iStart[s+1])
			fatal("puMigrate: species %d overflows its allocation (%ld + %ld immigrants > %ld)", s, pop->iStop[s]-pop->iStart[s], tot, pop->iStart[s+1]-pop->iStart[s]);
		DevGrid *pre = dp->slotted ? dp->predep : nullptr;          // pincAccMoveDistr3D1KE: the immigrants are deposited as they arrive
		if(tot > 0 && dp->slotted) PINC_LAUNCH(c, K_IMPORT, 96.0*tot, (k_import_cells<<<pGrid(c,tot),256,0,c->stream>>>(slotPar(dp,s), tot, ip, dp->d_immig, cellsOf(dp),
			mvBase(dp,s), dp->cap, mvKeys(dp,s), dp->d_mvCount + MV_COUNT + s, c->d_flags, pre ? pre->d_fixS[s] : nullptr, pre ? (long)pre->size[0] : 0, pre ? (long)pre->size[0]*pre->size[1] : 0)));
		else if(tot > 0) PINC_LAUNCH(c, K_IMPORT, 96.0*tot, (k_import<<<pGrid(c,tot),256,0,c->stream>>>(dp->base, dp->cap, pop->iStop[s], tot, ip, dp->d_immig)));
		pop->iStop[s] += tot;
	}
	dp->extracted = false;
	dp->emigInMovers = false;
	if(dp->slotted && takeSlotOverflow(c, "puMigrate")){ g_slotOverflows++; popLeaveSlotted(c, dp); }      // immigrants that found their cell full are the tail now
}

void puDistr3D1(const Population *pop, Grid *rhoGrid){
	Ctx *c = cur(); DevPop *dp = devPopRaw(c, pop); DevGrid *rho = devGrid(c, rhoGrid);
	if(rho->nv != 1) fatal("puDistr3D1 needs a scalar grid");
	if(dp->slotted && (dp->mvPending || dp->nc[0] > rho->size[0]-1 || dp->nc[1] > rho->size[1]-1 || dp->nc[2] > rho->size[2]-1)) popLeaveSlotted(c, dp);
	if(dp->slotted){
		const bool pre = dp->predep == rho;           // stayers by the push, movers by k_mv_insert, immigrants by k_import_cells: only the conversion is left
		if(!pre) dropPredeposit(dp);
		ensureFix(c, rho, dp->nS);
		gridZero(c, rho);
		const long sx = rho->size[0], sxy = sx*rho->size[1];
		for(int s = 0; s < dp->nS; s++){
			const long n = pop->iStop[s] - pop->iStart[s];
			if(n > 0 && !pre) PINC_LAUNCH(c, K_DEPOSIT, 24.0*n, (k_distr_slots<<<cellBlocks(c,dp->nCells),256,0,c->stream>>>(slotPar(dp,s), cellsOf(dp), sx, sxy, rho->d_fixS[s])));
			PINC_LAUNCH(c, K_DEPOSIT, 32.0*rho->n, (k_distr_finalize<<<gridFor(rho->n,256,c->numSMs*8),256,0,c->stream>>>(rho->d, rho->d_fixS[s], rho->n, 1.0/pop->charge[s], pop->charge[s], c->d_flags)));
		}
		dp->predep = nullptr;
		return;
	}
	bool pre = dp->predep == rho;                 // the stayers were deposited by pincAccMoveDistr3D1KE
	if(!pre) dropPredeposit(dp);
	ensureFix(c, rho, dp->nS);
	gridZero(c, rho);
	long sx = rho->size[0], sxy = sx*rho->size[1];
	for(int s = 0; s < dp->nS; s++){
		long a = pop->iStart[s], n = pop->iStop[s] - a;
		long ns = dp->sortedN[s] < n ? dp->sortedN[s] : n;
		if(ns > 0 && (dp->nc[0] > rho->size[0]-1 || dp->nc[1] > rho->size[1]-1 || dp->nc[2] > rho->size[2]-1)) ns = 0;
		if(pre && ns != dp->sortedN[s]) fatal("puDistr3D1: pre-deposited population changed size");
		const double *X = dp->base + a, *Y = X + dp->cap, *Z = Y + dp->cap;
		if(ns > 0 && !pre){
			long warps = dp->nCells;
			int blocks = gridFor(warps*32, 256, c->numSMs*8);
			PINC_LAUNCH(c, K_DEPOSIT, 24.0*ns, (k_distr_cells<<<blocks,256,0,c->stream>>>(X, Y, Z, dp->d_hist[s], cellsOf(dp), sx, sxy, rho->d_fixS[s])));
		}
		if(n - ns > 0)
			PINC_LAUNCH(c, K_DEPOSIT, 24.0*(n-ns), (k_distr_tail<<<pGrid(c,n-ns),256,0,c->stream>>>(X+ns, Y+ns, Z+ns, n-ns, rho->size[0], rho->size[1], rho->size[2], rho->d_fixS[s], c->d_flags)));
		PINC_LAUNCH(c, K_DEPOSIT, 32.0*rho->n, (k_distr_finalize<<<gridFor(rho->n,256,c->numSMs*8),256,0,c->stream>>>(rho->d, rho->d_fixS[s], rho->n, 1.0/pop->charge[s], pop->charge[s], c->d_flags)));
	}
	dp->predep = nullptr;
}

// the same over the occupied slots of a slotted population (the offender is reported by its slot index)
__global__ void k_assert_slots(SlotPar Q, int which, long nCells, double lo0, double lo1, double lo2,
		double hi0, double hi1, double hi2, int useLo, unsigned long long *first){
	const long n = nCells*Q.cap;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	const double lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
	for(; i < n; i += st){
		const long c = i / Q.cap;
		if((unsigned)(i - c*Q.cap) >= Q.cnt[c]) continue;
		#pragma unroll
		for(int d = 0; d < 3; d++){
			double v = Q.S[Q.off + i + (3*which + d)*Q.plane];
			if(v > hi[d] || (useLo && v < lo[d])) atomicMin(first, (unsigned long long)i*4 + d);
		}
	}
}
static void assertScan(Ctx *c, const Population *pop, int which, const double *lo, const double *hi, int useLo, const char *what){
	DevPop *dp = devPopRaw(c, pop);
	unsigned long long *first = (unsigned long long*)c->d_long;
	if(dp->slotted && dp->mvPending) popLeaveSlotted(c, dp);
	if(dp->slotted){
		// the reference's main loop runs these scans every step (src/main.c:207,221): they must not cost the slotted layout
		for(int s = 0; s < dp->nS; s++){
			long n = pop->iStop[s] - pop->iStart[s];
			if(n <= 0) continue;
			PINC_CUDA(cudaMemsetAsync(first, 0xff, sizeof(unsigned long long), c->stream));
			PINC_LAUNCH(c, K_MOVE, 24.0*n, (k_assert_slots<<<pGrid(c,dp->nCells*dp->slotCapS[s]),256,0,c->stream>>>(slotPar(dp,s), which, dp->nCells,
				lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], useLo, first)));
			PINC_CUDA(cudaMemcpyAsync(c->h_long, first, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
			streamSync(c);
			unsigned long long f = (unsigned long long)c->h_long[0];
			if(f != ~0ULL) fatal("Particle in slot %llu (of specie %i) %s in dimension %i", (unsigned long long)(f/4), s, what, (int)(f%4));
		}
		return;
	}
	for(int s = 0; s < dp->nS; s++){
		long a = pop->iStart[s], n = pop->iStop[s] - a;
		if(n <= 0) continue;
		PINC_CUDA(cudaMemsetAsync(first, 0xff, sizeof(unsigned long long), c->stream));
		PINC_LAUNCH(c, K_MOVE, 24.0*n, (k_assert_range<<<pGrid(c,n),256,0,c->stream>>>(dp->base + (size_t)(3*which)*dp->cap + a, dp->cap, n,
			lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], useLo, first)));
		PINC_CUDA(cudaMemcpyAsync(c->h_long, first, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
		unsigned long long f = (unsigned long long)c->h_long[0];
		if(f != ~0ULL) fatal("Particle i=%llu (of specie %i) %s in dimension %i", (unsigned long long)(a + (long)(f/4)), s, what, (int)(f%4));
	}
}

// src/population.c:316-341: every live particle inside [0, size-1] in every dimension, else msg(ERROR)
/* population.c:430-450: append one particle to species s (ignored with a warning when the species is full) */
void pNew(Population *pop, int s, const double *pos, const double *vel){
	Ctx *c = cur(); DevPop *dp = devPop(c, pop);
	if(pop->nDims != 3) fatal("pNew: only nDims=3 is implemented");
	if(pop->iStop[s] >= pop->iStart[s+1]){
		fprintf(stderr, "PINC-B200 WARNING: Not enough allocated memory to add new particle to specie %i. New particle ignored.\n", s);
		return;
	}
	invalidateOrder(dp);
	PINC_LAUNCH(c, K_MOVE, 48.0, (k_particle_put<<<1,32,0,c->stream>>>(dp->base, dp->cap, pop->iStop[s], pos[0], pos[1], pos[2], vel[0], vel[1], vel[2])));
	pop->iStop[s]++;
}
/* population.c:452-466: take the particle whose first coordinate sits at flat index p (= particle index * nDims) out of
 * species s, return its position and velocity and fill the hole with the species' last particle */
void pCut(Population *pop, int s, long int p, double *pos, double *vel){
	Ctx *c = cur(); DevPop *dp = devPop(c, pop);
	if(pop->nDims != 3) fatal("pCut: only nDims=3 is implemented");
	long i = p/3, last = pop->iStop[s] - 1;
	if(p % 3 || i < pop->iStart[s] || i > last) fatal("pCut: flat index %ld is not a live particle of specie %i", p, s);
	invalidateOrder(dp);
	double *d_out = (double*)c->d_long;
	PINC_LAUNCH(c, K_MOVE, 96.0, (k_particle_cut<<<1,32,0,c->stream>>>(dp->base, dp->cap, i, last, d_out)));
	double h[6];
	PINC_CUDA(cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, c->stream));
	streamSync(c);
	for(int d = 0; d < 3; d++){ pos[d] = h[d]; vel[d] = h[3+d]; }
	pop->iStop[s]--;
}
void pPosAssertInLocalFrame(const Population *pop, const Grid *grid){
	double lo[3] = {0, 0, 0}, hi[3];
	for(int d = 0; d < 3; d++) hi[d] = (double)(grid->size[d+1] - 1);
	assertScan(cur(), pop, 0, lo, hi, 1, "is out of bounds");
}
// src/population.c:343-365: no velocity component above max, else msg(ERROR)
void pVelAssertMax(const Population *pop, double max){
	double lo[3] = {0, 0, 0}, hi[3] = {max, max, max};
	assertScan(cur(), pop, 1, lo, hi, 0, "travels too fast");
}

// src/population.c:700-710 (host arithmetic on the small per-species scalars)
void pincSetSlotted(int on, int headroomPercent, int extraSlots){
	g_slotted = on ? 1 : 0;
	if(headroomPercent >= 0) g_slotHeadroom = headroomPercent;
	if(extraSlots >= 0) g_slotExtra = extraSlots;
}
int pincPopLayout(const Population *pop){ Ctx *c = cur(); auto it = c->pops.find(pop); return it != c->pops.end() && it->second->slotted ? 1 : 0; }
long pincSlottedOverflows(void){ return g_slotOverflows.load(); }

void pSumKinEnergy(Population *pop){
	int nS = pop->nSpecies;
	pop->kinEnergy[nS] = 0;
	for(int s = 0; s < nS; s++) pop->kinEnergy[nS] += pop->kinEnergy[s];
}

} // extern "C"
