// context.cu — per-rank device context, host-struct mirrors, coherence, timing and accounting.
#include "common.h"
#include <cstdarg>
#include <mutex>
#include <execinfo.h>
#include <csignal>
#include <unistd.h>

namespace pinc {

const char *kclassName[K_NCLASS] = { "push", "move", "deposit", "extract", "import", "sort", "gridop", "halo",
	"reduce", "gs", "residual", "restrict", "prolong", "mgfused", "findiff", "layout" };

static thread_local Ctx *t_ctx = nullptr;
static std::mutex g_mu;
static std::string g_lastError;

// Errors follow the reference's msg(ERROR,...) (src/io.c:170-217): print and exit(EXIT_FAILURE).
[[noreturn]] void fatal(const char *fmt, ...){
	char buf[1024];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
	{ std::lock_guard<std::mutex> lk(g_mu); g_lastError = buf; }
	fprintf(stderr, "PINC-B200 ERROR: %s\n", buf);
	fflush(stderr);
	exit(EXIT_FAILURE);
}

static Ctx *createCtx(int device, int rank, int size){
	int nDev = 0;
	cudaError_t e = cudaGetDeviceCount(&nDev);
	if(e != cudaSuccess || nDev < 1)
		fatal("no usable CUDA device (%s); libpinc_b200 has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
	if(device < 0 || device >= nDev) fatal("device %d out of range (have %d)", device, nDev);
	PINC_CUDA(cudaSetDevice(device));
	Ctx *c = new Ctx();
	c->device = device; c->rank = rank; c->size = size;
	PINC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	PINC_CUDA(cudaMalloc(&c->d_scal, 256*sizeof(double)));
	PINC_CUDA(cudaMemset(c->d_scal, 0, 256*sizeof(double)));
	PINC_CUDA(cudaMallocHost(&c->h_scal, 256*sizeof(double)));
	PINC_CUDA(cudaMalloc(&c->d_long, 1024*sizeof(long)));
	PINC_CUDA(cudaMallocHost(&c->h_long, 1024*sizeof(long)));
	PINC_CUDA(cudaMalloc(&c->d_flags, 16*sizeof(int)));
	PINC_CUDA(cudaMemset(c->d_flags, 0, 16*sizeof(int)));
	PINC_CUDA(cudaMallocHost(&c->h_flags, 16*sizeof(int)));
	PINC_CUDA(cudaMalloc(&c->d_bar, 64*sizeof(unsigned)));
	PINC_CUDA(cudaMemset(c->d_bar, 0, 64*sizeof(unsigned)));
	cudaDeviceProp prop;
	PINC_CUDA(cudaGetDeviceProperties(&prop, device));
	c->numSMs = prop.multiProcessorCount;
	PINC_CUDA(cudaEventCreate(&c->tStart));
	PINC_CUDA(cudaEventCreate(&c->tStop));
	c->tp = makeSelfTransport();
	return c;
}

Ctx *curOrNull(){ return t_ctx; }
Ctx *cur(){
	if(t_ctx){ return t_ctx; }
	int device = 0;
	const char *e = getenv("PINC_B200_DEVICE");
	if(!e) e = getenv("LOCAL_RANK");
	if(e) device = atoi(e);
	t_ctx = createCtx(device, 0, 1);
	return t_ctx;
}

void streamSync(Ctx *c){
	PINC_CUDA(cudaStreamSynchronize(c->stream));
	if(c->mgCheckPending) mgConvergenceCheck(c);
}

void checkDeviceFlags(Ctx *c, const char *where){
	PINC_CUDA(cudaMemcpyAsync(c->h_flags, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	streamSync(c);
	int f = c->h_flags[0];
	if(!f) return;
	PINC_CUDA(cudaMemsetAsync(c->d_flags, 0, sizeof(int), c->stream));
	if(f & ERR_POS_RANGE) fatal("%s: particle outside the local grid (not migrated, or |v| >= 1 cell per step)", where);
	if(f & ERR_CAPACITY)  fatal("%s: particle buffer capacity exceeded", where);
	if(f & ERR_P2P_TIMEOUT) fatal("%s: a neighbour rank did not arrive within 3 s (peer-memory smoother)", where);
	if(f & ERR_FIX_OVERFLOW) fatal("%s: deposition accumulator overflow (more than 2^16 full-weight particles of one species on one grid node)", where);
	fatal("%s: device error flags 0x%x", where, f);
}

void *tmpBuffer(Ctx *c, size_t bytes){
	if(bytes > c->tmpBytes){
		if(c->d_tmp){ streamSync(c); PINC_CUDA(cudaFree(c->d_tmp)); }
		size_t want = bytes + bytes/4 + 4096;
		PINC_CUDA(cudaMalloc(&c->d_tmp, want));
		c->tmpBytes = want;
	}
	return c->d_tmp;
}

double *partialBuffer(Ctx *c, long n){
	if(n > c->partialCap){
		if(c->d_partial){ streamSync(c); PINC_CUDA(cudaFree(c->d_partial)); }
		PINC_CUDA(cudaMalloc(&c->d_partial, (size_t)(n + 1024)*sizeof(double)));
		c->partialCap = n + 1024;
	}
	return c->d_partial;
}

double readScalar(Ctx *c, int slot){
	PINC_CUDA(cudaMemcpyAsync(&c->h_scal[slot], &c->d_scal[slot], sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	streamSync(c);
	return c->h_scal[slot];
}

// ---- launch accounting -----------------------------------------------------------------------
static cudaEvent_t getEvent(Ctx *c){
	if(!c->evPool.empty()){ cudaEvent_t e = c->evPool.back(); c->evPool.pop_back(); return e; }
	cudaEvent_t e; PINC_CUDA(cudaEventCreate(&e)); return e;
}
LaunchScope::LaunchScope(Ctx *c_, int cls_, double bytes) : c(c_), cls(cls_) {
	c->launches++;
	c->profCount[cls]++;
	c->profBytes[cls] += bytes;
	if(c->profOn){ a = getEvent(c); cudaEventRecord(a, c->stream); }
}
LaunchScope::~LaunchScope(){
	if(a){ cudaEvent_t b = getEvent(c); cudaEventRecord(b, c->stream); c->profEvents.push_back({a, b, cls}); }
}
static void profResolve(Ctx *c){
	if(c->profEvents.empty()) return;
	streamSync(c);
	for(auto &p : c->profEvents){
		float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b);
		c->profMs[p.cls] += ms;
		c->evPool.push_back(p.a); c->evPool.push_back(p.b);
	}
	c->profEvents.clear();
}

// ---- layout kernels: host AoS (pos[3i+d]) <-> device SoA planes ---------------------------------
__global__ void k_aos_to_soa(const double *__restrict__ aos, double *__restrict__ p0, double *__restrict__ p1,
                             double *__restrict__ p2, long n){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x;
	long stride = (long)gridDim.x*blockDim.x;
	for(; i < n; i += stride){ p0[i] = aos[3*i]; p1[i] = aos[3*i+1]; p2[i] = aos[3*i+2]; }
}
__global__ void k_soa_to_aos(double *__restrict__ aos, const double *__restrict__ p0, const double *__restrict__ p1,
                             const double *__restrict__ p2, long n){
	long i = blockIdx.x*(long)blockDim.x + threadIdx.x;
	long stride = (long)gridDim.x*blockDim.x;
	for(; i < n; i += stride){ aos[3*i] = p0[i]; aos[3*i+1] = p1[i]; aos[3*i+2] = p2[i]; }
}

static void popUpload(Ctx *c, DevPop *dp){
	const Population *p = dp->host;
	for(int s = 0; s < dp->nS; s++){
		long a = p->iStart[s], n = p->iStop[s] - a;
		if(n <= 0) continue;
		double *tmp = (double*)tmpBuffer(c, (size_t)3*n*sizeof(double));
		for(int w = 0; w < 2; w++){
			const double *src = (w ? p->vel : p->pos) + 3*a;
			PINC_CUDA(cudaMemcpyAsync(tmp, src, (size_t)3*n*sizeof(double), cudaMemcpyHostToDevice, c->stream));
			double *q = dp->base + (size_t)(3*w)*dp->cap + a;
			PINC_LAUNCH(c, K_LAYOUT, 48.0*n, (k_aos_to_soa<<<gridFor(n,256,c->numSMs*8),256,0,c->stream>>>(tmp, q, q+dp->cap, q+2*dp->cap, n)));
		}
	}
	for(int s = 0; s < dp->nS; s++) dp->sortedN[s] = 0;
	dp->keysValid = false;
	dp->extracted = false;
	dp->slotted = false; dp->mvPending = false; dp->emigInMovers = false;      // the host arrays are the truth now
	if(dp->predep){ dp->predep->fixDirty = true; dp->predep = nullptr; }
}

static void popDownload(Ctx *c, DevPop *dp){
	Population *p = dp->host;
	for(int s = 0; s < dp->nS; s++){
		long a = p->iStart[s], n = p->iStop[s] - a;
		if(n <= 0) continue;
		double *tmp = (double*)tmpBuffer(c, (size_t)3*n*sizeof(double));
		for(int w = 0; w < 2; w++){
			double *dst = (w ? p->vel : p->pos) + 3*a;
			const double *q = dp->base + (size_t)(3*w)*dp->cap + a;
			PINC_LAUNCH(c, K_LAYOUT, 48.0*n, (k_soa_to_aos<<<gridFor(n,256,c->numSMs*8),256,0,c->stream>>>(tmp, q, q+dp->cap, q+2*dp->cap, n)));
			PINC_CUDA(cudaMemcpyAsync(dst, tmp, (size_t)3*n*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		}
	}
	streamSync(c);
}

DevGrid *devGrid(Ctx *c, const Grid *g, bool upload){
	auto it = c->grids.find(g);
	if(it != c->grids.end()) return it->second;
	if(g->rank != 4) fatal("only 3-D grids are supported (grid rank %d)", g->rank);
	DevGrid *dg = new DevGrid();
	dg->host = const_cast<Grid*>(g);
	dg->nv = g->size[0];
	for(int d = 0; d < 3; d++){
		dg->size[d] = g->size[d+1]; dg->tsize[d] = g->trueSize[d+1];
		if(g->nGhostLayers[d+1] != 1 || g->nGhostLayers[d+1+g->rank] != 1)
			fatal("exactly one ghost layer per side is required (src/pusher.c:1047 puSanity)");
	}
	dg->n = g->sizeProd[g->rank];
	PINC_CUDA(cudaMalloc(&dg->d, (size_t)dg->n*sizeof(double)));
	long ms = 0;
	for(int d = 0; d < 3; d++){ long sl = dg->n/dg->size[d]; if(sl > ms) ms = sl; }
	dg->maxSlice = ms;
	PINC_CUDA(cudaMalloc(&dg->d_send, (size_t)6*ms*sizeof(double)));     // up to 3 dimensions x 2 directions in one exchange
	PINC_CUDA(cudaMalloc(&dg->d_recv, (size_t)6*ms*sizeof(double)));
	if(upload) PINC_CUDA(cudaMemcpyAsync(dg->d, g->val, (size_t)dg->n*sizeof(double), cudaMemcpyHostToDevice, c->stream));
	else PINC_CUDA(cudaMemsetAsync(dg->d, 0, (size_t)dg->n*sizeof(double), c->stream));
	for(int r = 1; r < g->rank; r++) if(g->bnd && (g->bnd[r] != PERIODIC || g->bnd[r + g->rank] != PERIODIC)) dg->nonPeriodic = true;
	if(dg->nonPeriodic){
		// nSliceMax as src/grid.c:455-465 (the component axis counts as a dimension)
		long nMax = 0;
		for(int d = 0; d < g->rank; d++){ long n = 1; for(int dd = 0; dd < g->rank; dd++) if(dd != d) n *= g->size[dd]; if(n > nMax) nMax = n; }
		dg->bndStride = nMax;
		PINC_CUDA(cudaMalloc(&dg->d_bnd, (size_t)2*g->rank*nMax*sizeof(double)));
		gridUploadBnd(c, dg);
	}
	c->grids[g] = dg;
	return dg;
}

DevPop *devPop(Ctx *c, const Population *p, bool upload){
	DevPop *dp = devPopRaw(c, p, upload);
	if(dp->slotted){ dp->slotKicks++; popLeaveSlotted(c, dp); }
	return dp;
}
DevPop *devPopRaw(Ctx *c, const Population *p, bool upload){
	auto it = c->pops.find(p);
	if(it != c->pops.end()) return it->second;
	if(p->nDims != 3) fatal("only 3-D populations are supported");
	if(p->nSpecies > 8) fatal("at most 8 species are supported");
	DevPop *dp = new DevPop();
	dp->host = const_cast<Population*>(p);
	dp->nS = p->nSpecies;
	dp->cap = p->iStart[p->nSpecies];
	for(int s = 0; s <= dp->nS; s++) dp->iStart[s] = p->iStart[s];
	size_t bytes = (size_t)6*(dp->cap > 0 ? dp->cap : 1)*sizeof(double);
	PINC_CUDA(cudaMalloc(&dp->base, bytes));
	PINC_CUDA(cudaMalloc(&dp->d_keys, (size_t)(dp->cap > 0 ? dp->cap : 1)*sizeof(unsigned)));
	c->pops[p] = dp;
	if(upload) popUpload(c, dp);
	return dp;
}

static void freeDevGrid(DevGrid *g){
	cudaFree(g->d); cudaFree(g->d_send); cudaFree(g->d_recv); for(int s = 0; s < 8; s++) if(g->d_fixS[s]) cudaFree(g->d_fixS[s]);
	if(g->d_bnd) cudaFree(g->d_bnd);
	delete g;
}
static void freeDevPop(DevPop *p){
	cudaFree(p->base); if(p->alt) cudaFree(p->alt); cudaFree(p->d_keys);
	for(int s = 0; s < 8; s++){ if(p->d_hist[s]) cudaFree(p->d_hist[s]); if(p->d_cursor[s]) cudaFree(p->d_cursor[s]); }
	if(p->d_emig) cudaFree(p->d_emig); if(p->d_immig) cudaFree(p->d_immig);
	if(p->slot) cudaFree(p->slot); if(p->d_mvCount) cudaFree(p->d_mvCount);
	for(int s = 0; s < 8; s++) if(p->d_cnt[s]) cudaFree(p->d_cnt[s]);
	delete p;
}

} // namespace pinc

using namespace pinc;

extern "C" {

PincCtx *pincCtxCreate(int device, int rank, int size){
	Ctx *c = createCtx(device, rank, size);
	t_ctx = c;
	return (PincCtx*)c;
}
void pincCtxMakeCurrent(PincCtx *ctx){
	t_ctx = (Ctx*)ctx;
	if(t_ctx) PINC_CUDA(cudaSetDevice(t_ctx->device));
}
void pincCtxDestroy(PincCtx *ctx){
	Ctx *c = (Ctx*)ctx;
	if(!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	mgForgetPlans(c);
	for(auto &kv : c->grids) freeDevGrid(kv.second);
	for(auto &kv : c->pops) freeDevPop(kv.second);
	cudaFree(c->d_scal); cudaFreeHost(c->h_scal); cudaFree(c->d_long); cudaFreeHost(c->h_long);
	cudaFree(c->d_flags); cudaFreeHost(c->h_flags); cudaFree(c->d_bar);
	if(c->d_partial) cudaFree(c->d_partial);
	if(c->d_mgProf) cudaFree(c->d_mgProf);
	if(c->d_mgMail) cudaFree(c->d_mgMail);
	if(c->d_mgRhoS) cudaFree(c->d_mgRhoS);
	if(c->d_mgHist){ cudaFree(c->d_mgHist); cudaFreeHost(c->h_mgHist); }
	if(c->d_tmp) cudaFree(c->d_tmp);
	for(auto e : c->evPool) cudaEventDestroy(e);
	cudaEventDestroy(c->tStart); cudaEventDestroy(c->tStop);
	cudaStreamDestroy(c->stream);
	mgFreeArena(c);
	delete c->tp;
	if(t_ctx == c) t_ctx = nullptr;
	delete c;
}

void pincSyncGridToDevice(Grid *grid){
	Ctx *c = cur();
	DevGrid *g = devGrid(c, grid, false);
	PINC_CUDA(cudaMemcpyAsync(g->d, grid->val, (size_t)g->n*sizeof(double), cudaMemcpyHostToDevice, c->stream));
	if(g->nonPeriodic) gridUploadBnd(c, g);
	streamSync(c);
}
void pincSyncGridToHost(Grid *grid){
	Ctx *c = cur();
	DevGrid *g = devGrid(c, grid, true);
	PINC_CUDA(cudaMemcpyAsync(grid->val, g->d, (size_t)g->n*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	streamSync(c);
}
void pincSyncPopToDevice(Population *pop){
	Ctx *c = cur();
	auto it = c->pops.find(pop);
	if(it == c->pops.end()){ devPop(c, pop, true); streamSync(c); return; }
	popUpload(c, it->second);
	streamSync(c);
}
void pincSyncPopToHost(Population *pop){
	Ctx *c = cur();
	DevPop *dp = devPop(c, pop, true);
	popDownload(c, dp);
}
void pincForget(void *hostStruct){
	Ctx *c = t_ctx;                                  // nothing to forget if this thread never used the device
	if(!c) return;
	streamSync(c);
	auto g = c->grids.find(hostStruct);
	if(g != c->grids.end()){ mgForgetPlans(c); freeDevGrid(g->second); c->grids.erase(g); return; }
	auto p = c->pops.find(hostStruct);
	if(p != c->pops.end()){ freeDevPop(p->second); c->pops.erase(p); }
}
// page-lock caller-owned host arrays (pop->pos, grid->val, ...) so that pincSync* copies run at full PCIe rate
int pincHostRegister(void *ptr, size_t bytes){
	cur();
	return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? 0 : (cudaGetLastError(), 1);
}
int pincHostUnregister(void *ptr){ return cudaHostUnregister(ptr) == cudaSuccess ? 0 : (cudaGetLastError(), 1); }
void pincDeviceSynchronize(void){
	Ctx *c = cur();
	streamSync(c);
	checkDeviceFlags(c, "pincDeviceSynchronize");
}

void pincTimerStart(void){ Ctx *c = cur(); PINC_CUDA(cudaEventRecord(c->tStart, c->stream)); }
double pincTimerStopMs(void){
	Ctx *c = cur();
	PINC_CUDA(cudaEventRecord(c->tStop, c->stream));
	PINC_CUDA(cudaEventSynchronize(c->tStop));
	float ms = 0; PINC_CUDA(cudaEventElapsedTime(&ms, c->tStart, c->tStop));
	return ms;
}
void pincProfEnable(int on){ Ctx *c = cur(); profResolve(c); c->profOn = on; }
void pincProfReset(void){
	Ctx *c = cur(); profResolve(c);
	for(int i = 0; i < K_NCLASS; i++){ c->profMs[i] = 0; c->profCount[i] = 0; c->profBytes[i] = 0; }
}
int pincProfGet(int idx, char *name, int namelen, double *ms, long int *launches, double *algBytes){
	Ctx *c = cur(); profResolve(c);
	if(idx < 0 || idx >= K_NCLASS) return 0;
	if(name && namelen > 0){ strncpy(name, kclassName[idx], namelen-1); name[namelen-1] = 0; }
	if(ms) *ms = c->profMs[idx];
	if(launches) *launches = c->profCount[idx];
	if(algBytes) *algBytes = c->profBytes[idx];
	return 1;
}
long int pincLaunchCount(void){ return cur()->launches; }
// cycle accounting of the cluster multigrid kernel ($PINC_B200_MGPROF=1): copies 32 (cycles, calls) pairs and clears
int pincMgProfRead(long long *out64){
	Ctx *c = cur();
	long long *d = (long long*)mgProfBuffer(c);
	if(!d) return 0;
	streamSync(c);
	PINC_CUDA(cudaMemcpy(out64, d, 64*sizeof(long long), cudaMemcpyDeviceToHost));
	PINC_CUDA(cudaMemset(d, 0, 64*sizeof(long long)));
	return 1;
}
const char *pincVersion(void){ return "pinc-b200 0.1 (sm_100a)"; }
int pincLastError(char *buf, int len){
	std::lock_guard<std::mutex> lk(g_mu);
	if(buf && len > 0){ strncpy(buf, g_lastError.c_str(), len-1); buf[len-1] = 0; }
	return (int)g_lastError.size();
}

} // extern "C"

// $PINC_B200_SEGV_TRACE=1: a native back trace on SIGSEGV/SIGABRT (debugging aid for hosts that only report "Segmentation fault")
namespace {
void segvTrace(int sig){
	void *frames[64];
	int n = backtrace(frames, 64);
	const char msg[] = "PINC-B200: fatal signal, native back trace:\n";
	if(write(2, msg, sizeof msg - 1) < 0){}
	backtrace_symbols_fd(frames, n, 2);
	signal(sig, SIG_DFL);
	raise(sig);
}
struct SegvInstall { SegvInstall(){ const char *e = getenv("PINC_B200_SEGV_TRACE"); if(e && atoi(e)){ signal(SIGSEGV, segvTrace); signal(SIGABRT, segvTrace); } } } g_segvInstall;
}
