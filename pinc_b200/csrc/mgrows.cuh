// mgrows.cuh — block-resident mgGS3D (src/multigrid.c:683-767) for BIG blocks: the throughput-bound counterpart of
// bGS's register-descriptor path.  Included by multigrid.cu (needs Lvl, BLvl, Scope, llStore/llWait, ix, ldg2, mgS).
//
// Levels of >= 1 M nodes (128^3: warm128, and the replicated global solve of 4 and 8 ranks) were smoothed as grid-wide
// sweeps through L2 (fGS: seven L2 loads per update, half of every sector unused, 16.6 us per half-sweep on 128^3).
// Here CTA 1+b keeps its block (bx x by x bz = 32 x 32 x 16 on 128^3) plus one halo layer in shared memory for the whole
// smoother call, as bGS does, but organised for bandwidth instead of latency:
//
//  * one x-ROW of the block per thread (bx = 2*HB nodes, HB = 8 or 16 per colour); the row's rho lives in registers;
//  * COLOUR-SEPARATED rows: row r = (y, z) stores its colour-0 and colour-1 nodes (colour = (x+y+z)&1, block-local) as two
//    arrays of HB+1 doubles, slot(x) = (x+1)>>1 for x = -1 .. bx, so the one x-halo node an array needs is its first
//    (odd x) or last (even x) slot.  For a node x = 2i+o of the colour being updated (o = (colour + y + z)&1) the
//    x-neighbours are slots i and i+1 of the row's other-colour array and the y/z-neighbours are slot i+o of the four
//    adjacent rows' other-colour arrays: every access of the half-sweep is a unit-stride run of HB doubles;
//  * row pitch HB+1 doubles (odd): the 16 lanes of a half-warp, which own consecutive rows, hit 16 different bank pairs
//    with every 64-bit access - no bank conflicts, against 2-way conflicts of the stride-2 red-black layout;
//  * faces travel through the same tagged 16-byte mailboxes as bGS's (the data is its own flag); a thread sends the
//    face nodes of its own row (contiguous slots: the mailbox faces are colour-separated too), receives are spread over
//    all threads.
//
// Same arithmetic per node as everywhere else: 1/6*(x+ + x- + y+ + y- + z+ + z- + rho) summed left to right, hence the
// same bits as fGS/bGS; gBnd's mean subtraction is applied once per smoother call as in bGS.
#pragma once

namespace pinc {

template<int HB> struct RowMap {
	static constexpr int BX = 2*HB, W = HB + 1;
	int by, bz, RY, CB;                        // rows per plane incl. halo rows; doubles per colour plane
	__device__ __forceinline__ RowMap(const BLvl &B) : by(B.by), bz(B.bz), RY(B.by + 2), CB((B.by + 2)*(B.bz + 2)*W) {}
	__device__ __forceinline__ int row(int y, int z) const { return (y + 1) + RY*(z + 1); }        // y in -1..by, z in -1..bz
	// shared-memory index of block-local node (x, y, z), x in -1..BX
	__device__ __forceinline__ int at(int x, int y, int z) const { return ((x + y + z) & 1)*CB + row(y, z)*W + ((x + 1) >> 1); }
};

__device__ __forceinline__ int wrapG(int g, int t){ return g == 0 ? t : (g == t + 1 ? 1 : g); }      // periodic image of a ghost index

// Colour-separated copy of the block's rho in global memory (L2-resident): [colour][i][row], so that for a fixed i the lanes
// of a warp (consecutive rows) read consecutive doubles.  Keeping rho in registers (32-64 of the 128 a thread may have)
// made ptxas spill it to local memory, i.e. to L2 anyway, with extra traffic; this way the reads are explicit and coalesced.
template<int HB> __device__ __forceinline__ long rhoSIdx(int c, int i, int row, int nRows){ return ((long)(c*HB + i))*nRows + row; }

// gBnd(rho) = gNeutralizeGrid (src/grid.c:730-779) ahead of rGS with rGS's thread <-> row mapping: the smoother's
// thread reads what it wrote itself, so no grid barrier is needed after the subtraction
template<int HB> __device__ __noinline__ void rNeutRho(const Lvl &L, const BLvl &B, Scope &S){
	ProfScope psn(*S.K, 28);
	constexpr int BX = 2*HB;
	const int bid = (int)blockIdx.x - 1;
	const int nRows = B.by*B.bz;
	const bool own = bid >= 0 && bid < B.nb && (int)threadIdx.x < nRows;
	double r[BX];
	double acc = 0;
	double *g = nullptr;
	int p = 0;
	if(own){
		const int cx = bid % B.nbx, cr = bid / B.nbx, cy = cr % B.nby, cz = cr / B.nby;
		const int y = threadIdx.x % B.by, z = threadIdx.x / B.by;
		p = (y + z) & 1;
		g = L.rho + ix(cx*BX + 1, cy*B.by + y + 1, cz*B.bz + z + 1, L.s0, L.s1);
		#pragma unroll
		for(int x = 0; x < BX; x++) r[x] = ldg2(g + x);
		#pragma unroll
		for(int x = 0; x < BX; x++) acc += r[x];
	}
	const double avg = S.allSum(acc)/((double)(L.s0-2)*(L.s1-2)*(L.s2-2));
	if(own){
		double *rs = B.rhoS + (size_t)bid*BX*nRows;
		#pragma unroll
		for(int x = 0; x < BX; x++){
			const double v = r[x] - avg;
			g[x] = v;
			rs[rhoSIdx<HB>((x + p) & 1, x >> 1, threadIdx.x, nRows)] = v;       // x = 2i + o with o = (colour + p)&1
		}
	}
}

// one half-sweep of colour C over the thread's row: E = the row's other-colour array, o = (C + y + z)&1, O = its own-colour
// array, rho = the row's colour-C values in the colour-separated copy (stride rstride between consecutive i).  The rho
// loads of a chunk are issued first and consumed last, so their L2 latency hides behind the chunk's shared-memory work.
template<int HB, int CH> __device__ __forceinline__ void rowUpdate(const double *E, int o, int RYW, const double *rho, int rstride, double *O){
	constexpr int W = HB + 1;
	const double coeff = 1./6.;
	const double *Np = E + W + o, *Sp = E - W + o, *Up = E + RYW + o, *Dp = E - RYW + o;
	#pragma unroll
	for(int i0 = 0; i0 < HB; i0 += CH){
		// all loads of the chunk first (CH independent chains in flight), then the CH sums side by side
		double e[CH+1], n[CH], s[CH], u[CH], d[CH], rr[CH], v[CH];
		#pragma unroll
		for(int i = 0; i < CH; i++) rr[i] = __ldcg(rho + (long)(i0+i)*rstride);
		#pragma unroll
		for(int i = 0; i <= CH; i++) e[i] = E[i0+i];
		#pragma unroll
		for(int i = 0; i < CH; i++){ n[i] = Np[i0+i]; s[i] = Sp[i0+i]; u[i] = Up[i0+i]; d[i] = Dp[i0+i]; }
		#pragma unroll
		for(int i = 0; i < CH; i++) v[i] = e[i+1] + e[i];
		#pragma unroll
		for(int i = 0; i < CH; i++) v[i] += n[i];
		#pragma unroll
		for(int i = 0; i < CH; i++) v[i] += s[i];
		#pragma unroll
		for(int i = 0; i < CH; i++) v[i] += u[i];
		#pragma unroll
		for(int i = 0; i < CH; i++) v[i] += d[i];
		#pragma unroll
		for(int i = 0; i < CH; i++) v[i] += rr[i];
		#pragma unroll
		for(int i = 0; i < CH; i++) v[i] *= coeff;
		#pragma unroll
		for(int i = 0; i < CH; i++) O[o+i0+i] = v[i];
	}
}

// what every thread of the CTA needs to address the mailboxes: kept in shared memory, not in registers
struct RowComm { unsigned out[6]; int fb[6]; };

__device__ __forceinline__ uint4 llLoad(const uint4 *p){
	uint4 v;
	asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
	return v;
}

template<int HB> __device__ __noinline__ void rGS(const Lvl &L, const BLvl &B, int nCycles, double sIn, Scope &S, unsigned &seq){
	ProfScope ps(*S.K, PS_GS_BIG);
	constexpr int BX = 2*HB, W = HB + 1, NP = 4;
	__shared__ RowComm RC;
	const RowMap<HB> M(B);
	const int t0 = L.s0-2, t1 = L.s1-2, t2 = L.s2-2;
	const int bid = (int)blockIdx.x - 1;
	const bool act = bid >= 0 && bid < B.nb;
	const int by = B.by, bz = B.bz, nRows = by*bz;
	const int nA = nRows, nBh = HB*bz, nCh = HB*by;                 // per colour: x-halo nodes, nodes of ONE y-face, of ONE z-face
	const int nHalo = nA + 2*nBh + 2*nCh;
	const int tid = threadIdx.x;
	const bool own = act && tid < nRows;
	double *A0 = mgS + B.offPhi;
	const long long tEnter = clock64();
	int ox = 0, oy = 0, oz = 0;
	double bsum = 0;
	if(act){
		const int cx = bid % B.nbx, cr = bid / B.nbx, cy = cr % B.nby, cz = cr / B.nby;
		ox = cx*BX; oy = cy*by; oz = cz*bz;
		const int fA = by*bz, fB = BX*bz, fC = BX*by, slots = 2*(fA + fB + fC);
		const uint4 *mine = B.mail + (size_t)bid*slots;
		if(tid == 0){
			int *fb = RC.fb; unsigned *out = RC.out;
			fb[0] = 0; fb[1] = fA; fb[2] = 2*fA; fb[3] = 2*fA + fB; fb[4] = 2*fA + 2*fB; fb[5] = 2*fA + 2*fB + fC;
			// my boundary nodes go to the neighbour across face f, which receives them on its face f^1
			int xm = cx ? cx-1 : B.nbx-1, xp = cx+1 < B.nbx ? cx+1 : 0, ym = cy ? cy-1 : B.nby-1, yp = cy+1 < B.nby ? cy+1 : 0;
			int zm = cz ? cz-1 : B.nbz-1, zp = cz+1 < B.nbz ? cz+1 : 0;
			out[0] = (unsigned)(xm + B.nbx*(cy + B.nby*cz))*slots + fb[1];
			out[1] = (unsigned)(xp + B.nbx*(cy + B.nby*cz))*slots + fb[0];
			out[2] = (unsigned)(cx + B.nbx*(ym + B.nby*cz))*slots + fb[3];
			out[3] = (unsigned)(cx + B.nbx*(yp + B.nby*cz))*slots + fb[2];
			out[4] = (unsigned)(cx + B.nbx*(cy + B.nby*zm))*slots + fb[5];
			out[5] = (unsigned)(cx + B.nbx*(cy + B.nby*zp))*slots + fb[4];
		}
		// ---- block + halo layer from global memory (periodic image), pending mean shift applied; warp = row, lane = x ----
		for(int i = tid; i < nRows*(BX+2); i += blockDim.x){
			const int x = i % (BX+2) - 1, rr = i / (BX+2), yy = rr % by, zz = rr / by;
			double v = ldg2(L.phi + ix(wrapG(ox + x + 1, t0), oy + yy + 1, oz + zz + 1, L.s0, L.s1));
			if(sIn != 0.0) v -= sIn;
			A0[M.at(x, yy, zz)] = v;
		}
		for(int i = tid; i < 2*BX*bz; i += blockDim.x){              // y-halo rows
			const int f = i >= BX*bz, w = i - f*BX*bz, x = w % BX, zz = w / BX, yy = f ? by : -1;
			double v = ldg2(L.phi + ix(ox + x + 1, wrapG(oy + yy + 1, t1), oz + zz + 1, L.s0, L.s1));
			if(sIn != 0.0) v -= sIn;
			A0[M.at(x, yy, zz)] = v;
		}
		for(int i = tid; i < 2*BX*by; i += blockDim.x){              // z-halo rows
			const int f = i >= BX*by, w = i - f*BX*by, x = w % BX, yy = w / BX, zz = f ? bz : -1;
			double v = ldg2(L.phi + ix(ox + x + 1, oy + yy + 1, wrapG(oz + zz + 1, t2), L.s0, L.s1));
			if(sIn != 0.0) v -= sIn;
			A0[M.at(x, yy, zz)] = v;
		}
		const int y = own ? tid % by : 0, z = own ? tid / by : 0;
		const int r = M.row(y, z), p = (y + z) & 1, RYW = M.RY*W, q = y + by*z;
		const double *rs = B.rhoS + (size_t)bid*BX*nRows;
		long long *pf = (S.K->prof && bid == 0 && tid == 0) ? S.K->prof : nullptr;
		long long tW = 0, tA = 0, tB = 0;
		if(pf){ pf[2*9] += clock64() - tEnter; pf[2*9+1] += 1; tW = clock64(); }
		// ---- 2*nCycles half-sweeps: colour 0 = (x+y+z) even = the reference's first pass ((j+k+l) odd, 1-based) ---------
		for(int h = 0; h < 2*nCycles; h++){
			const int c = h & 1, cc = 1 - c;
			if(pf) tA = clock64();
			if(h > 0){
				// the other colour's face nodes of half-sweep h-1: NP polls in flight per thread
				const unsigned tag = seq + (unsigned)h;
				for(int i0 = tid; i0 < nHalo; i0 += NP*(int)blockDim.x){
					const uint4 *src[NP]; int dst[NP]; uint4 val[NP];
					#pragma unroll
					for(int u = 0; u < NP; u++){
						const int i = i0 + u*(int)blockDim.x;
						src[u] = nullptr; dst[u] = 0;
						if(i >= nHalo) continue;
						int slot;
						if(i < nA){
							const int yy = i % by, zz = i / by, f = (cc == ((yy + zz) & 1));   // colour (yy+zz)&1 sits at even x: its halo is x = BX
							slot = RC.fb[f] + i;
							dst[u] = cc*M.CB + M.row(yy, zz)*W + (f ? HB : 0);
						} else if(i < nA + 2*nBh){
							const int uu = i - nA, f = uu >= nBh, w = uu - f*nBh, m = w % HB, zz = w / HB, yy = f ? by : -1;
							slot = RC.fb[2+f] + m + HB*(2*zz + cc);
							dst[u] = cc*M.CB + M.row(yy, zz)*W + m + ((cc + yy + zz) & 1);
						} else {
							const int uu = i - nA - 2*nBh, f = uu >= nCh, w = uu - f*nCh, m = w % HB, yy = w / HB, zz = f ? bz : -1;
							slot = RC.fb[4+f] + m + HB*(2*yy + cc);
							dst[u] = cc*M.CB + M.row(yy, zz)*W + m + ((cc + yy + zz) & 1);
						}
						src[u] = mine + slot;
					}
					#pragma unroll
					for(int u = 0; u < NP; u++) if(src[u]) val[u] = llLoad(src[u]);
					#pragma unroll
					for(int u = 0; u < NP; u++){
						if(!src[u]) continue;
						unsigned spins = 0;
						while(val[u].y != tag || val[u].w != tag){
							if(++spins > (1u << 22)) __trap();
							val[u] = llLoad(src[u]);
						}
						A0[dst[u]] = __longlong_as_double(((long long)val[u].z << 32) | (long long)val[u].x);
					}
				}
			}
			if(pf){ tB = clock64(); pf[2*13] += tB - tA; pf[2*13+1] += 1; }
			__syncthreads();
			if(pf){ tA = clock64(); pf[2*11] += tA - tB; pf[2*11+1] += 1; }
			if(own){
				const int o = (c + p) & 1;
				const double *E = A0 + cc*M.CB + r*W;
				double *O = A0 + c*M.CB + r*W;
				rowUpdate<HB, 4>(E, o, RYW, rs + rhoSIdx<HB>(c, 0, tid, nRows), nRows, O);
				if(pf){ tB = clock64(); pf[2*14] += tB - tA; pf[2*14+1] += 1; }
				if(h + 1 < 2*nCycles){
					// the row's face nodes, read back from the thread's own stores: x = 0 (o = 0) or x = BX-1 (o = 1), and the
					// whole row if it lies on a y- or z-face
					const unsigned stag = seq + (unsigned)h + 1u;
					if(o == 0) llStore(B.mail + RC.out[0] + q, O[0], stag); else llStore(B.mail + RC.out[1] + q, O[HB], stag);
					if(y == 0 || y == by-1){
						uint4 *d = B.mail + RC.out[y == 0 ? 2 : 3] + HB*(2*z + c);
						#pragma unroll
						for(int i = 0; i < HB; i++) llStore(d + i, O[o+i], stag);
					}
					if(z == 0 || z == bz-1){
						uint4 *d = B.mail + RC.out[z == 0 ? 4 : 5] + HB*(2*y + c);
						#pragma unroll
						for(int i = 0; i < HB; i++) llStore(d + i, O[o+i], stag);
					}
				}
				if(pf){ tA = clock64(); pf[2*15] += tA - tB; pf[2*15+1] += 1; }
			}
		}
		if(pf){ pf[2*10] += clock64() - tW; pf[2*10+1] += 2*nCycles; }
		__syncthreads();
		for(int i = tid; i < nRows*BX; i += blockDim.x){
			const int x = i % BX, rr = i / BX;
			bsum += A0[M.at(x, rr % by, rr / by)];
		}
	}
	seq += 2u*(unsigned)nCycles;
	const long long tTail = clock64();
	// the 2*nCycles gBnd calls, applied once, on the way back to global memory
	const double avg = S.allSum(bsum)/((double)t0*t1*t2);
	if(act)
		for(int i = tid; i < nRows*BX; i += blockDim.x){
			const int x = i % BX, rr = i / BX, yy = rr % by, zz = rr / by;
			L.phi[ix(ox + x + 1, oy + yy + 1, oz + zz + 1, L.s0, L.s1)] = A0[M.at(x, yy, zz)] - avg;
		}
	S.sync();
	if(S.K->prof && bid == 0 && tid == 0){ S.K->prof[2*12] += clock64() - tTail; S.K->prof[2*12+1] += 1; }
}

} // namespace pinc
