// grid.cu — field-grid kernels and the PINC grid entry points (src/grid.h).
//
// Layout is the reference's (src/core.h:261-277): val[c + nv*(j + sx*(k + sy*l))], one ghost layer per side.
// Every kernel here is a streaming pass, HBM/L2-bound; the grids of the BASELINE configs (<= 130^3 points,
// 17.6 MB per scalar) live in the 126 MB L2 between kernels.
#include "common.h"

namespace pinc {

// ---- elementwise -------------------------------------------------------------------------------
__global__ void k_scale(double *__restrict__ v, long n, double num){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st) v[i] *= num;
}
__global__ void k_add_scalar(double *__restrict__ v, long n, double num){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st) v[i] += num;
}
__global__ void k_sub_scalar(double *__restrict__ v, long n, double num){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st) v[i] -= num;
}
__global__ void k_square(double *__restrict__ v, long n){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st){ double x = v[i]; v[i] = x*x; }
}
__global__ void k_addto(double *__restrict__ r, const double *__restrict__ a, long n){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st) r[i] += a[i];
}
__global__ void k_subfrom(double *__restrict__ r, const double *__restrict__ a, long n){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st) r[i] -= a[i];
}
// val -= sum/denom with the sum taken from a device scalar (gNeutralizeGrid, src/grid.c:730-779)
__global__ void k_sub_mean(double *__restrict__ v, long n, const double *__restrict__ sum, double denom){
	double avg = sum[0]/denom;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < n; i += st) v[i] -= avg;
}

static inline int ewGrid(Ctx *c, long n){ return gridFor(n, 256, c->numSMs*8); }

void gridScale(Ctx *c, DevGrid *g, double num){
	PINC_LAUNCH(c, K_GRIDOP, 16.0*g->n, (k_scale<<<ewGrid(c,g->n),256,0,c->stream>>>(g->d, g->n, num)));
}
void gridAddTo(Ctx *c, DevGrid *r, const DevGrid *a){
	if(a->n != r->n) fatal("gAddTo: grids differ in size");
	PINC_LAUNCH(c, K_GRIDOP, 24.0*r->n, (k_addto<<<ewGrid(c,r->n),256,0,c->stream>>>(r->d, a->d, r->n)));
}
void gridSubFrom(Ctx *c, DevGrid *r, const DevGrid *a){
	if(a->n != r->n) fatal("gSubFrom: grids differ in size");
	PINC_LAUNCH(c, K_GRIDOP, 24.0*r->n, (k_subfrom<<<ewGrid(c,r->n),256,0,c->stream>>>(r->d, a->d, r->n)));
}
void gridZero(Ctx *c, DevGrid *g){
	PINC_CUDA(cudaMemsetAsync(g->d, 0, (size_t)g->n*sizeof(double), c->stream));
}

// ---- slices (src/grid.c:72-147): all elements whose coordinate along dimension dd equals o ----------
struct Dims { int s0, s1, s2, nv; };
__device__ __forceinline__ long sliceElem(const Dims D, int dd, int o, long e){
	int cc = (int)(e % D.nv); long t = e / D.nv;
	long j, k, l;
	if(dd == 0){ k = t % D.s1; l = t / D.s1; j = o; }
	else if(dd == 1){ j = t % D.s0; l = t / D.s0; k = o; }
	else { j = t % D.s0; k = t / D.s0; l = o; }
	return D.nv*(j + D.s0*(k + (long)D.s1*l)) + cc;
}
__global__ void k_slice_pack(const double *__restrict__ v, double *__restrict__ buf, Dims D, int dd, int o, long ns){
	long e = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; e < ns; e += st) buf[e] = v[sliceElem(D, dd, o, e)];
}
__global__ void k_slice_unpack(double *__restrict__ v, const double *__restrict__ buf, Dims D, int dd, int o, long ns, int add){
	long e = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; e < ns; e += st){
		long g = sliceElem(D, dd, o, e);
		if(add) v[g] += buf[e]; else v[g] = buf[e];
	}
}
// both directions of one dimension when the neighbour is this rank itself (periodic wrap, no buffers)
__global__ void k_halo_self(double *__restrict__ v, Dims D, int dd, int upTake, int loPlace, int loTake, int upPlace, long ns, int add){
	long e = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; e < 2*ns; e += st){
		long ee = e < ns ? e : e - ns;
		long src = sliceElem(D, dd, e < ns ? upTake : loTake, ee);
		long dst = sliceElem(D, dd, e < ns ? loPlace : upPlace, ee);
		if(add) v[dst] += v[src]; else v[dst] = v[src];
	}
}

static Dims dimsOf(const DevGrid *g){ return Dims{ g->size[0], g->size[1], g->size[2], g->nv }; }

static int dimNeighbor(const MpiInfo *m, int dd, int dir){
	int sub[3], nb[3];
	for(int d = 0; d < 3; d++) sub[d] = nb[d] = m->subdomain[d];
	nb[dd] = (sub[dd] + dir + m->nSubdomains[dd]) % m->nSubdomains[dd];
	return nb[0] + m->nSubdomains[0]*(nb[1] + m->nSubdomains[1]*nb[2]);
}

// src/grid.c:349-406.  d in 1..3.  The upper take-layer travels up into the receiver's lower place, the
// lower take-layer travels down.  TOHALO: take size-2 / 1, place 0 / size-1.  FROMHALO: take size-1 / 0,
// place 1 / size-2.
void gridHaloDim(Ctx *c, DevGrid *g, const MpiInfo *m, int d, int add, int dir){
	int dd = d - 1;
	int sz = g->size[dd];
	int upTake = sz-2+dir, upPlace = sz-1-dir, loTake = 1-dir, loPlace = dir;
	long ns = g->n / sz;
	Dims D = dimsOf(g);
	int blocks = gridFor(2*ns, 256, c->numSMs*4);
	if(m->nSubdomains[dd] == 1){
		PINC_LAUNCH(c, K_HALO, 32.0*ns, (k_halo_self<<<blocks,256,0,c->stream>>>(g->d, D, dd, upTake, loPlace, loTake, upPlace, ns, add)));
		return;
	}
	PINC_LAUNCH(c, K_HALO, 16.0*ns, (k_slice_pack<<<blocks,256,0,c->stream>>>(g->d, g->d_send, D, dd, upTake, ns)));
	PINC_LAUNCH(c, K_HALO, 16.0*ns, (k_slice_pack<<<blocks,256,0,c->stream>>>(g->d, g->d_send + ns, D, dd, loTake, ns)));
	int upper = dimNeighbor(m, dd, +1), lower = dimNeighbor(m, dd, -1);
	size_t bytes = (size_t)ns*sizeof(double);
	std::vector<Msg> sends = { {upper, 0, g->d_send, bytes}, {lower, 1, g->d_send + ns, bytes} };
	std::vector<Msg> recvs = { {lower, 0, g->d_recv, bytes}, {upper, 1, g->d_recv + ns, bytes} };
	c->tp->exchange(c, sends, recvs);
	PINC_LAUNCH(c, K_HALO, 16.0*ns, (k_slice_unpack<<<blocks,256,0,c->stream>>>(g->d, g->d_recv, D, dd, loPlace, ns, add)));
	PINC_LAUNCH(c, K_HALO, 16.0*ns, (k_slice_unpack<<<blocks,256,0,c->stream>>>(g->d, g->d_recv + ns, D, dd, upPlace, ns, add)));
}
// Face-only ghost fill (setSlice, TOHALO) of the decomposed dimensions in ONE exchange: what a 7-point stencil
// needs between two half-sweeps.  Unlike gHaloOp's dimension-by-dimension sequence it does not propagate edge and
// corner ghosts (the rims of a face may be stale); callers that need those use gridHalo.
struct FacePar { int dd[6], take[6], place[6]; long ns[6], off[7]; int n; };
__global__ void k_faces_pack(const double *__restrict__ v, double *__restrict__ buf, Dims D, FacePar F){
	long e = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; e < F.off[F.n]; e += st){
		int f = 0; while(f+1 < F.n && e >= F.off[f+1]) f++;
		buf[e] = v[sliceElem(D, F.dd[f], F.take[f], e - F.off[f])];
	}
}
__global__ void k_faces_unpack(double *__restrict__ v, const double *__restrict__ buf, Dims D, FacePar F){
	long e = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; e < F.off[F.n]; e += st){
		int f = 0; while(f+1 < F.n && e >= F.off[f+1]) f++;
		v[sliceElem(D, F.dd[f], F.place[f], e - F.off[f])] = buf[e];
	}
}
void gridHaloFaces(Ctx *c, DevGrid *g, const MpiInfo *m){
	FacePar F{}; F.n = 0; F.off[0] = 0;
	std::vector<Msg> sends, recvs;
	for(int dd = 0; dd < 3; dd++){
		if(m->nSubdomains[dd] == 1) continue;
		int sz = g->size[dd];
		long ns = g->n / sz;
		int upper = dimNeighbor(m, dd, +1), lower = dimNeighbor(m, dd, -1);
		// slab 2*i: my upper true layer -> upper neighbour's lower ghost; slab 2*i+1: my lower true layer -> lower neighbour's upper ghost
		int f = F.n;
		F.dd[f] = dd; F.take[f] = sz-2; F.place[f] = 0;    F.ns[f] = ns; F.off[f+1] = F.off[f] + ns;
		F.dd[f+1] = dd; F.take[f+1] = 1; F.place[f+1] = sz-1; F.ns[f+1] = ns; F.off[f+2] = F.off[f+1] + ns;
		size_t bytes = (size_t)ns*sizeof(double);
		sends.push_back({upper, 2*dd,   g->d_send + F.off[f],   bytes});
		sends.push_back({lower, 2*dd+1, g->d_send + F.off[f+1], bytes});
		recvs.push_back({lower, 2*dd,   g->d_recv + F.off[f],   bytes});
		recvs.push_back({upper, 2*dd+1, g->d_recv + F.off[f+1], bytes});
		F.n += 2;
	}
	if(F.n == 0) return;
	if(F.off[F.n] > 6*g->maxSlice) fatal("gridHaloFaces: exchange buffer too small");
	int blocks = gridFor(F.off[F.n], 256, c->numSMs*4);
	PINC_LAUNCH(c, K_HALO, 16.0*F.off[F.n], (k_faces_pack<<<blocks,256,0,c->stream>>>(g->d, g->d_send, dimsOf(g), F)));
	c->tp->exchange(c, sends, recvs);
	PINC_LAUNCH(c, K_HALO, 16.0*F.off[F.n], (k_faces_unpack<<<blocks,256,0,c->stream>>>(g->d, g->d_recv, dimsOf(g), F)));
}
void gridHalo(Ctx *c, DevGrid *g, const MpiInfo *m, int add, int dir){
	if(!add && dir == 0 && gridHaloP2P(c, g, m)) return;          // ghost fill of a scalar grid over peer memory
	for(int d = 1; d <= 3; d++) gridHaloDim(c, g, m, d, add, dir);
}

// ---- reductions over the true grid ---------------------------------------------------------------
template<int BLOCK> __device__ __forceinline__ double blockSum(double v){
	__shared__ double sh[BLOCK/32];
	for(int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	__syncthreads();
	if(lane == 0) sh[w] = v;
	__syncthreads();
	if(w == 0){
		v = lane < BLOCK/32 ? sh[lane] : 0.0;
		for(int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	}
	return v;            // valid in thread 0
}
// mode 0: sum v; mode 2: sum v*w.  Scalar grids.
__global__ void k_sum_true(const double *__restrict__ v, const double *__restrict__ w, Dims D, int mode, double *__restrict__ partial){
	int t0 = D.s0-2, t1 = D.s1-2, t2 = D.s2-2;
	long nt = (long)t0*t1*t2;
	double acc = 0;
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; i < nt; i += st){
		int j = (int)(i % t0) + 1; long r = i / t0; int k = (int)(r % t1) + 1; int l = (int)(r / t1) + 1;
		long g = j + D.s0*(k + (long)D.s1*l);
		acc += mode == 2 ? v[g]*w[g] : v[g];
	}
	acc = blockSum<256>(acc);
	if(threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void k_final_sum(const double *__restrict__ partial, int n, double *__restrict__ out){
	double acc = 0;
	for(int i = threadIdx.x; i < n; i += 256) acc += partial[i];
	acc = blockSum<256>(acc);
	if(threadIdx.x == 0) out[0] = acc;
}

// true-grid sum of val (or val^2 after squaring in place) over ALL ranks into d_scal[slot]
void gridSumTrueAll(Ctx *c, DevGrid *g, int mode, int slot, const MpiInfo *m){
	if(m->mpiSize > 1 && c->tp->p2p()){
		if(mode == 1){
			PINC_LAUNCH(c, K_GRIDOP, 16.0*g->n, (k_square<<<ewGrid(c,g->n),256,0,c->stream>>>(g->d, g->n)));
			mode = 0;
		}
		long nt = (long)g->tsize[0]*g->tsize[1]*g->tsize[2];
		int blocks = gridFor(nt, 256, c->numSMs*4);
		double *partial = partialBuffer(c, blocks);
		PINC_LAUNCH(c, K_REDUCE, 8.0*nt, (k_sum_true<<<blocks,256,0,c->stream>>>(g->d, nullptr, dimsOf(g), mode, partial)));
		if(allSumP2P(c, partial, blocks, c->d_scal + slot, m)) return;
		PINC_LAUNCH(c, K_REDUCE, 8.0*blocks, (k_final_sum<<<1,256,0,c->stream>>>(partial, blocks, c->d_scal + slot)));
		c->tp->allreduceSum(c, c->d_scal + slot, 1);
		return;
	}
	gridSumTrue(c, g, mode, nullptr, slot);
	if(m->mpiSize > 1) c->tp->allreduceSum(c, c->d_scal + slot, 1);
}
void gridSumTrue(Ctx *c, DevGrid *g, int mode, const DevGrid *other, int slot){
	if(g->nv != 1) fatal("true-grid sums are implemented for scalar grids");
	if(mode == 1){
		PINC_LAUNCH(c, K_GRIDOP, 16.0*g->n, (k_square<<<ewGrid(c,g->n),256,0,c->stream>>>(g->d, g->n)));
		mode = 0;
	}
	long nt = (long)g->tsize[0]*g->tsize[1]*g->tsize[2];
	int blocks = gridFor(nt, 256, c->numSMs*4);
	double *partial = partialBuffer(c, blocks);
	PINC_LAUNCH(c, K_REDUCE, 8.0*nt*(mode == 2 ? 2 : 1), (k_sum_true<<<blocks,256,0,c->stream>>>(g->d, other ? other->d : nullptr, dimsOf(g), mode, partial)));
	PINC_LAUNCH(c, K_REDUCE, 8.0*blocks, (k_final_sum<<<1,256,0,c->stream>>>(partial, blocks, c->d_scal + slot)));
}

// src/grid.c:730-779: mean over the global true grid subtracted from every element, ghosts included
void gridNeutralize(Ctx *c, DevGrid *g, const MpiInfo *m){
	gridSumTrueAll(c, g, 0, 0, m);
	double denom = (double)((long)g->tsize[0]*g->tsize[1]*g->tsize[2])*m->mpiSize;
	PINC_LAUNCH(c, K_GRIDOP, 16.0*g->n, (k_sub_mean<<<ewGrid(c,g->n),256,0,c->stream>>>(g->d, g->n, c->d_scal, denom)));
}

// ---- finite differences ------------------------------------------------------------------------------
// src/grid.c:226-261 over the flat interior range (quirk Q9), all three components in one pass:
// E[3g+d] = 0.5*(phi[g+sp_d] - phi[g-sp_d]).  32 B per point algorithmic (8 read, 24 written).
__global__ void k_findiff1st(const double *__restrict__ phi, double *__restrict__ E, long start, long end, long sx, long sxy){
	long g = start + blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; g < end; g += st){
		E[3*g]   = 0.5*(phi[g+1]   - phi[g-1]);
		E[3*g+1] = 0.5*(phi[g+sx]  - phi[g-sx]);
		E[3*g+2] = 0.5*(phi[g+sxy] - phi[g-sxy]);
	}
}
// src/grid.c:296-334: result = -6*obj + sum of the six neighbours (this order), flat interior range
__global__ void k_findiff2nd(double *__restrict__ res, const double *__restrict__ obj, long start, long end, long sx, long sxy){
	long g = start + blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; g < end; g += st){
		double r = -6.*obj[g];
		r += obj[g+1] + obj[g-1] + obj[g+sx] + obj[g-sx] + obj[g+sxy] + obj[g-sxy];
		res[g] = r;
	}
}

// ---- boundary conditions (src/grid.c:921-1023) ---------------------------------------------------------------------
void gridUploadBnd(Ctx *c, DevGrid *g){
	if(!g->d_bnd) return;
	if(!g->host->bndSlice) fatal("a grid with a DIRICHLET/NEUMANN edge needs Grid::bndSlice (src/grid.c:467)");
	PINC_CUDA(cudaMemcpyAsync(g->d_bnd, g->host->bndSlice, (size_t)2*g->host->rank*g->bndStride*sizeof(double), cudaMemcpyHostToDevice, c->stream));
	streamSync(c);              // the host array is pageable and may change right after
}
// ghost := value two layers further in - 2*A (src/grid.c:958-990)
__global__ void k_neumann(double *__restrict__ v, const double *__restrict__ bnd, Dims D, int dd, int take, int place, long ns){
	long e = blockIdx.x*(long)blockDim.x + threadIdx.x, st = (long)gridDim.x*blockDim.x;
	for(; e < ns; e += st){
		double a = v[sliceElem(D, dd, take, e)];
		a -= 2*bnd[e];
		v[sliceElem(D, dd, place, e)] = a;
	}
}
// one edge: gDirichlet (slice 1 on a lower, size-1 on an upper edge := bndSlice; the asymmetry is the reference's,
// src/grid.c:940) or gNeumann (ghost slice 0 / size-1 := slice 2 / size-3 minus twice bndSlice)
void gridEdge(Ctx *c, DevGrid *g, int boundary, int kind){
	const int rank = g->host->rank, d = boundary % rank, upper = boundary > rank;
	if(d < 1 || !g->d_bnd) fatal("gDirichlet/gNeumann: boundary %d of a grid without non-periodic edges", boundary);
	const int dd = d - 1, sz = g->size[dd];
	const long ns = g->n / sz;
	const double *b = g->d_bnd + (long)boundary*g->bndStride;
	const int blocks = gridFor(ns, 256, c->numSMs*4);
	if(kind == DIRICHLET){
		const int offset = 1 + upper*(sz - 2);
		PINC_LAUNCH(c, K_HALO, 16.0*ns, (k_slice_unpack<<<blocks,256,0,c->stream>>>(g->d, b, dimsOf(g), dd, offset, ns, 0)));
	} else {
		const int offset = upper*(sz - 1);
		PINC_LAUNCH(c, K_HALO, 24.0*ns, (k_neumann<<<blocks,256,0,c->stream>>>(g->d, b, dimsOf(g), dd, offset + 2 - 4*upper, offset, ns)));
	}
}
void gridBnd(Ctx *c, DevGrid *g, const MpiInfo *m){
	const Grid *h = g->host;
	const int rank = h->rank;
	bool periodic = false;
	for(int d = 1; d < rank; d++) if(h->bnd[d] == PERIODIC) periodic = true;         // (the reference looks at the lower edges only)
	if(periodic) gridNeutralize(c, g, m);
	if(!g->nonPeriodic) return;
	for(int d = 1; d < rank; d++)
		if(m->subdomain[d-1] == 0){
			if(h->bnd[d] == DIRICHLET) gridEdge(c, g, d, DIRICHLET);
			else if(h->bnd[d] == NEUMANN) gridEdge(c, g, d, NEUMANN);
		}
	for(int d = rank+1; d < 2*rank; d++)
		if(m->subdomain[d-rank-1] == m->nSubdomains[d-rank-1]-1){
			if(h->bnd[d] == DIRICHLET) gridEdge(c, g, d, DIRICHLET);
			if(h->bnd[d] == NEUMANN) gridEdge(c, g, d, NEUMANN);
		}
}

} // namespace pinc

using namespace pinc;

extern "C" {

void gZero(Grid *grid){ Ctx *c = cur(); gridZero(c, devGrid(c, grid)); }
void gMul(Grid *grid, double num){ Ctx *c = cur(); gridScale(c, devGrid(c, grid), num); }
void gAdd(Grid *grid, double num){
	Ctx *c = cur(); DevGrid *g = devGrid(c, grid);
	PINC_LAUNCH(c, K_GRIDOP, 16.0*g->n, (k_add_scalar<<<ewGrid(c,g->n),256,0,c->stream>>>(g->d, g->n, num)));
}
void gSub(Grid *grid, double num){
	Ctx *c = cur(); DevGrid *g = devGrid(c, grid);
	PINC_LAUNCH(c, K_GRIDOP, 16.0*g->n, (k_sub_scalar<<<ewGrid(c,g->n),256,0,c->stream>>>(g->d, g->n, num)));
}
void gSquare(Grid *grid){
	Ctx *c = cur(); DevGrid *g = devGrid(c, grid);
	PINC_LAUNCH(c, K_GRIDOP, 16.0*g->n, (k_square<<<ewGrid(c,g->n),256,0,c->stream>>>(g->d, g->n)));
}
void gCopy(const Grid *original, Grid *copy){
	Ctx *c = cur(); DevGrid *a = devGrid(c, original), *b = devGrid(c, copy);
	if(a->n != b->n) fatal("gCopy: grids differ in size");
	PINC_CUDA(cudaMemcpyAsync(b->d, a->d, (size_t)a->n*sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
}
void gAddTo(Grid *result, Grid *addition){
	Ctx *c = cur();
	gridAddTo(c, devGrid(c, result), devGrid(c, addition));
}
void gSubFrom(Grid *result, const Grid *subtraction){ Ctx *c = cur(); gridSubFrom(c, devGrid(c, result), devGrid(c, subtraction)); }
double gSumTruegrid(const Grid *grid){
	Ctx *c = cur();
	gridSumTrue(c, devGrid(c, grid), 0, nullptr, 1);
	return readScalar(c, 1);
}
long int gTotTruesize(const Grid *grid, const MpiInfo *mpiInfo){
	long int tot = 1;
	for(int r = 1; r < grid->rank; r++) tot *= (long)mpiInfo->nSubdomains[r-1]*grid->trueSize[r];
	return tot;
}
void gNeutralizeGrid(Grid *grid, const MpiInfo *mpiInfo){ Ctx *c = cur(); gridNeutralize(c, devGrid(c, grid), mpiInfo); }

// src/grid.c:992-1023
void gBnd(Grid *grid, const MpiInfo *mpiInfo){ Ctx *c = cur(); gridBnd(c, devGrid(c, grid), mpiInfo); }
// src/grid.c:929-956 / :958-990; boundary = d (lower edge) or rank + d (upper edge), d in 1..3
void gDirichlet(Grid *grid, const int boundary, const MpiInfo *mpiInfo){ (void)mpiInfo; Ctx *c = cur(); gridEdge(c, devGrid(c, grid), boundary, DIRICHLET); }
void gNeumann(Grid *grid, const int boundary, const MpiInfo *mpiInfo){ (void)mpiInfo; Ctx *c = cur(); gridEdge(c, devGrid(c, grid), boundary, NEUMANN); }
// src/grid.c:608-662: constant boundary values (1 on Dirichlet, 2 on Neumann edges of the global domain) into grid->bndSlice
void gSetBndSlices(Grid *grid, MpiInfo *mpiInfo){
	const int rank = grid->rank;
	if(!grid->bndSlice) fatal("gSetBndSlices: the grid has no bndSlice");
	long nMax = 0;
	for(int d = 0; d < rank; d++){ long n = 1; for(int dd = 0; dd < rank; dd++) if(dd != d) n *= grid->size[dd]; if(n > nMax) nMax = n; }
	for(int d = 1; d < rank; d++){
		if(mpiInfo->subdomain[d-1] == 0 && (grid->bnd[d] == DIRICHLET || grid->bnd[d] == NEUMANN))
			for(long s = 0; s < nMax; s++) grid->bndSlice[s + nMax*d] = grid->bnd[d] == DIRICHLET ? 1. : 2.;
		if(mpiInfo->subdomain[d-1] == mpiInfo->nSubdomains[d-1]-1 && (grid->bnd[d+rank] == DIRICHLET || grid->bnd[d+rank] == NEUMANN))
			for(long s = 0; s < nMax; s++) grid->bndSlice[s + nMax*(d+rank)] = grid->bnd[d+rank] == DIRICHLET ? 1. : 2.;
	}
	Ctx *c = curOrNull();                        // host arrays only; the device mirror follows if the grid already has one
	if(!c) return;
	auto it = c->grids.find(grid);
	if(it != c->grids.end() && it->second->nonPeriodic) gridUploadBnd(c, it->second);
}

void gHaloOpDim(funPtr sliceOp, Grid *grid, const MpiInfo *mpiInfo, int d, opDirection dir){
	int add;
	if(sliceOp == (funPtr)setSlice) add = 0;
	else if(sliceOp == (funPtr)addSlice) add = 1;
	else fatal("gHaloOp: sliceOp must be setSlice or addSlice of libpinc_b200");
	Ctx *c = cur();
	gridHaloDim(c, devGrid(c, grid), mpiInfo, d, add, dir == FROMHALO ? 1 : 0);
}
void gHaloOp(funPtr sliceOp, Grid *grid, const MpiInfo *mpiInfo, opDirection dir){
	for(int d = 1; d < grid->rank; d++) gHaloOpDim(sliceOp, grid, mpiInfo, d, dir);
}

// host-buffer slice access (src/grid.c:85-147); d in 0..rank-1 as in the reference (0 = component axis)
static void sliceHost(double *slice, const Grid *grid, int d, int offset, int mode){
	if(d < 1 || d > 3) fatal("slices along the component axis are not supported");
	Ctx *c = cur(); DevGrid *g = devGrid(c, grid);
	long ns = g->n / g->size[d-1];
	double *buf = (double*)tmpBuffer(c, ns*sizeof(double));
	int blocks = gridFor(ns, 256, c->numSMs*4);
	if(mode == 0){
		PINC_LAUNCH(c, K_HALO, 16.0*ns, (k_slice_pack<<<blocks,256,0,c->stream>>>(g->d, buf, dimsOf(g), d-1, offset, ns)));
		PINC_CUDA(cudaMemcpyAsync(slice, buf, ns*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
	} else {
		PINC_CUDA(cudaMemcpyAsync(buf, slice, ns*sizeof(double), cudaMemcpyHostToDevice, c->stream));
		PINC_LAUNCH(c, K_HALO, 16.0*ns, (k_slice_unpack<<<blocks,256,0,c->stream>>>(g->d, buf, dimsOf(g), d-1, offset, ns, mode == 2)));
		streamSync(c);
	}
}
void getSlice(double *slice, const Grid *grid, int d, int offset){ sliceHost(slice, grid, d, offset, 0); }
void setSlice(const double *slice, Grid *grid, int d, int offset){ sliceHost(const_cast<double*>(slice), grid, d, offset, 1); }
void addSlice(const double *slice, Grid *grid, int d, int offset){ sliceHost(const_cast<double*>(slice), grid, d, offset, 2); }

void gFinDiff1st(const Grid *scalar, Grid *field){
	Ctx *c = cur(); DevGrid *p = devGrid(c, scalar), *e = devGrid(c, field);
	if(p->nv != 1 || e->nv != 3 || e->n != 3*p->n) fatal("gFinDiff1st: need a scalar and a 3-vector grid of equal extent");
	long sx = p->size[0], sxy = (long)p->size[0]*p->size[1];
	long start = 1 + sx + sxy, end = p->n - start;
	PINC_LAUNCH(c, K_FINDIFF, 32.0*(end-start), (k_findiff1st<<<ewGrid(c,end-start),256,0,c->stream>>>(p->d, e->d, start, end, sx, sxy)));
}
void gFinDiff2nd3D(Grid *result, const Grid *object){
	Ctx *c = cur(); DevGrid *r = devGrid(c, result), *o = devGrid(c, object);
	if(r->nv != 1 || o->nv != 1 || r->n != o->n) fatal("gFinDiff2nd3D: need two scalar grids of equal extent");
	long sx = o->size[0], sxy = (long)o->size[0]*o->size[1];
	long start = 1 + sx + sxy, end = o->n - start;
	PINC_LAUNCH(c, K_RESIDUAL, 16.0*(end-start), (k_findiff2nd<<<ewGrid(c,end-start),256,0,c->stream>>>(r->d, o->d, start, end, sx, sxy)));
}

// src/grid.c:1276-1321: potEnergy[nSpecies] = 0.5 * sum over the true grid of rho*phi (this rank)
void gPotEnergy(const Grid *rho, const Grid *phi, Population *pop){
	Ctx *c = cur();
	gridSumTrue(c, devGrid(c, rho), 2, devGrid(c, phi), 2);
	double e = readScalar(c, 2);
	e *= 0.5;
	pop->potEnergy[pop->nSpecies] = e;
}

} // extern "C"
