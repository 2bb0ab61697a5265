// common.h — internal declarations of libpinc_b200 (device mirrors, context, launch helpers).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/pinc_b200.h"

namespace pinc {

[[noreturn]] void fatal(const char *fmt, ...);

#define PINC_CUDA(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) \
	::pinc::fatal("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); } while(0)

// ---- kernel classes for the per-class device-time accounting (pincProfGet) ----
enum KClass { K_PUSH = 0, K_MOVE, K_DEPOSIT, K_EXTRACT, K_IMPORT, K_SORT, K_GRIDOP, K_HALO, K_REDUCE,
              K_GS, K_RESIDUAL, K_RESTRICT, K_PROLONG, K_MGFUSED, K_FINDIFF, K_LAYOUT, K_NCLASS };
extern const char *kclassName[K_NCLASS];

// device-side error bits (Ctx::d_flags[0])
enum DevErr { ERR_POS_RANGE = 1, ERR_CAPACITY = 2, ERR_VEL_MAX = 4, ERR_P2P_TIMEOUT = 8, ERR_FIX_OVERFLOW = 16, ERR_SLOT_OVERFLOW = 32 };

// fixed-point scale of the deposition accumulators: weights in [0,1] are summed as
// round(w * 2^46) in 64-bit integers, so a node can take 2^17 particles per species before the
// accumulator overflows and the sum does not depend on the order of the additions.
#define PINC_FIX_BITS 46

struct DevGrid {
	Grid *host = nullptr;
	double *d = nullptr;        // values, same layout as host (core.h:261-277)
	long n = 0;                 // sizeProd[rank]
	int nv = 1;                 // size[0]
	int size[3] = {1,1,1};      // ghost-inclusive nx,ny,nz
	int tsize[3] = {1,1,1};
	double *d_send = nullptr;   // 2 slabs (up, down) of the largest face
	double *d_recv = nullptr;
	long maxSlice = 0;
	long long *d_fixS[8] = {nullptr};   // per-species fixed-point accumulators of the deposition (scalar grids, lazily)
	bool fixDirty = false;              // accumulators hold deposits nobody consumed (cleared before the next use)
	// Dirichlet/Neumann boundary values (Grid::bndSlice, src/grid.c:467: 2*rank slices of nSliceMax doubles, slice index =
	// boundary = d for a lower and rank+d for an upper edge); mirrored when the grid has a non-periodic edge
	double *d_bnd = nullptr;
	long bndStride = 0;                 // nSliceMax
	bool nonPeriodic = false;
};

struct DevPop {
	Population *host = nullptr;
	int nS = 0;
	long cap = 0;               // iStart[nS]
	double *base = nullptr;     // 6*cap doubles: x,y,z,vx,vy,vz planes (SoA)
	double *alt = nullptr;      // second buffer of the same size (target of the cell sort), lazily
	long iStart[9] = {0};
	// cell binning (per species): particles [iStart[s], iStart[s]+sortedN[s]) are ordered by cell and
	// cellStart[s][c] is the offset (relative to iStart[s]) of the first particle of cell c.
	int nc[3] = {0,0,0};        // cells per dimension (= size-1) the binning was built for
	long nCells = 0;            // nc[0]*nc[1]*nc[2]; keys nCells..nCells+26 are the emigrant bins
	unsigned *d_keys = nullptr; // cap entries
	unsigned *d_hist[8] = {nullptr};      // nCells+27+1 counts, then (after the scan) offsets
	unsigned *d_cursor[8] = {nullptr};    // nCells+27 running cursors of the scatter
	long sortedN[8] = {0};
	bool keysValid = false;     // d_keys/d_hist hold the classification of the current positions
	double keyThr[6] = {0};     // thresholds the keys were computed with
	// migrants
	double *d_emig = nullptr;   // packed (x,y,z,vx,vy,vz) records, neighbour by neighbour, species inside
	long emigCap = 0;           // records
	long emigOff[28] = {0};     // record offset of neighbour ne in d_emig
	double *d_immig = nullptr;
	long immigCap = 0;
	bool extracted = false;     // emigrants sit behind iStop[s], binned by neighbour, not yet packed
	DevGrid *predep = nullptr;  // the stayers of the current positions are already deposited into this grid's accumulators
	// Cell-slotted storage (particles.cu, "slotted mode"): while pincAccMove3D1KE runs step after step, the particles
	// of cell c of species s live in slots [slotOff[s] + c*slotCapS[s], ... + d_cnt[s][c]) of six planes of `slotPlane`
	// doubles and only the ~5 % that change cell move; the contiguous planes above are stale until somebody needs them.
	bool slotted = false;
	double *slot = nullptr; long slotPlane = 0;
	long slotOff[9] = {0}; int slotCapS[8] = {0};
	unsigned *d_cnt[8] = {nullptr};       // particles per cell
	unsigned *d_mvCount = nullptr;        // [8] movers per species (the mover list of species s is alt[iStart[s]..], keys in d_keys) + [8..8+8*27) emigrants per neighbour
	bool mvPending = false;               // the last push left movers (other cell / other rank) in the mover lists
	bool emigInMovers = false;            // puMigrate packs the emigrants from the mover lists
	double thr6[6] = {0}; bool haveThr = false;     // migration thresholds of the most recent extraction (puMove has no MpiInfo)
	int slotKicks = 0;                    // how often an entry point without a slotted form forced the contiguous layout back
};

struct ProfEvent { cudaEvent_t a, b; int cls; };

struct Transport;

// a captured V-cycle of a multi-rank solve (multigrid.cu): 0 not captured yet, 1 ready, -1 capture failed
struct CycleGraph { cudaGraphExec_t exec = nullptr; unsigned long long nOps = 0; long launches = 0; int state = 0; };

struct Ctx {
	int device = 0, rank = 0, size = 1;
	cudaStream_t stream = nullptr;
	std::unordered_map<const void*, DevGrid*> grids;
	std::unordered_map<const void*, DevPop*> pops;
	// scratch
	double *d_scal = nullptr;       // 256 small device scalars (reductions, means)
	double *h_scal = nullptr;       // pinned mirror
	double *d_partial = nullptr;    // per-block partial sums
	long partialCap = 0;
	void *d_tmp = nullptr;          // general scratch (grows)
	size_t tmpBytes = 0;
	long *h_long = nullptr;         // 1024 pinned host longs
	long *d_long = nullptr;
	int *d_flags = nullptr;         // device error bits
	int *h_flags = nullptr;
	unsigned *d_bar = nullptr;      // grid barrier words of the persistent multigrid kernel
	int numSMs = 148;
	// accounting
	long launches = 0;
	int profOn = 0;
	std::vector<ProfEvent> profEvents;
	std::vector<cudaEvent_t> evPool;
	double profMs[K_NCLASS] = {0};
	long profCount[K_NCLASS] = {0};
	double profBytes[K_NCLASS] = {0};
	cudaEvent_t tStart = nullptr, tStop = nullptr;
	// multigrid residual history of the most recent solve (device + pinned host: [0] = cycles, [1..] = barRes)
	double *d_mgHist = nullptr, *h_mgHist = nullptr;
	void *d_mgRhoS = nullptr; size_t mgRhoSBytes = 0;       // colour-separated rho of the row smoother (mgrows.cuh)
	void *d_mgMail = nullptr; size_t mgMailBytes = 0;       // halo mailboxes of the block-resident smoother (multigrid.cu)
	long long *d_mgProf = nullptr;                          // optional cycle accounting ($PINC_B200_MGPROF)
	bool mgHistPending = false;
	std::vector<double> mgHistory;
	// convergence report of the most recent solve (checked at the next stream synchronisation, see streamSync)
	bool mgCheckPending = false;
	double mgTol = 1e-10, mgLastBarRes = 0;
	int mgMaxCycles = 0, mgLastCycles = 0;
	// per-device launch attributes of the persistent kernels (cudaFuncSetAttribute applies to the current device only)
	size_t mgAttrSmem[4] = {0, 0, 0, 0};
	int clNc = -1; size_t clSmem = 0;
	std::unordered_map<const void*, CycleGraph> cycleGraphs;     // keyed by the solver's mgRho
	std::unordered_map<const void*, void*> mgGlobal;             // replicated global hierarchies of multi-rank solves (multigrid.cu)
	bool mgGlobalBusy = false;
	void *mgXArena = nullptr;                                    // peer-mapped arena of the hybrid multi-rank solve (multigrid.cu)
	double mgXTagHigh = 0; bool mgXPending = false;              // its running mailbox tag after the most recent solve (reset before the 32-bit tags wrap)
	int mgLastPath = 0;                                          // pincMgLastPath: 0 ops, 1 all-SM kernel, 2 cluster kernel, +4 replicated
	Transport *tp = nullptr;
	std::string lastError;
};

Ctx *cur();                                    // current context of this host thread (created lazily)
Ctx *curOrNull();                              // ... or nullptr if this thread never touched the device (host-array entry points)
DevGrid *devGrid(Ctx *c, const Grid *g, bool upload = true);
DevPop *devPop(Ctx *c, const Population *p, bool upload = true);       // contiguous planes current (leaves slotted mode)
DevPop *devPopRaw(Ctx *c, const Population *p, bool upload = true);    // whatever mode the population is in
void popLeaveSlotted(Ctx *c, DevPop *dp);                               // particles.cu: slots -> contiguous planes
void *tmpBuffer(Ctx *c, size_t bytes);
double *partialBuffer(Ctx *c, long n);
void checkDeviceFlags(Ctx *c, const char *where);      // after a stream sync
void streamSync(Ctx *c);

// launch bookkeeping: PINC_LAUNCH(ctx, class, algorithmic bytes, kernel<<<...>>>(...))
struct LaunchScope {
	Ctx *c; int cls; cudaEvent_t a = nullptr;
	LaunchScope(Ctx *c, int cls, double bytes);
	~LaunchScope();
};
#define PINC_LAUNCH(ctx, cls, bytes, ...) do { { ::pinc::LaunchScope ls_((ctx), (cls), (double)(bytes)); __VA_ARGS__; } \
	cudaError_t e_ = cudaGetLastError(); if(e_ != cudaSuccess) ::pinc::fatal("launch failed (%s) at %s:%d: %s", \
	::pinc::kclassName[cls], __FILE__, __LINE__, cudaGetErrorString(e_)); } while(0)

inline int gridFor(long n, int block, int maxBlocks){
	long b = (n + block - 1)/block;
	if(b < 1) b = 1;
	if(b > maxBlocks) b = maxBlocks;
	return (int)b;
}

// ---- transport (exchange steps between ranks) ----
struct Msg { int peer; int tag; void *ptr; size_t bytes; };
// Peer-to-peer arena of a rank (NVLink): six mailbox planes (lower/upper side of x, y, z) that the neighbours' smoother
// kernels store their boundary layers into, and one arrival counter per plane.  Every rank's arena is mapped into
// every other rank (CUDA IPC), so a kernel addresses a neighbour's plane as peerArena[rank] + the same offset.
#define P2P_PLANE_CAP (1L<<18)            // doubles per mailbox plane (faces up to 512 x 512)
struct P2P {
	char *arena = nullptr;                // local arena: [256 x u64 flags/scalars][6 planes x P2P_PLANE_CAP doubles]
	std::vector<char*> peerArena;         // indexed by rank; [own rank] = arena
	unsigned long long seq = 0;           // exchanges issued so far (identical on all ranks: same call sequence)
	unsigned long long base = 0;          // host mirror of the device-side sequence base (flag slot 10)
	static size_t flagBytes(){ return 256*sizeof(unsigned long long); }
	// flag words: 0-5 arrival counters of the mailbox planes, 8 block ticket, 9 local generation, 10 sequence base,
	// 11 all-reduce counter, 32-95 all-reduce arrival per rank, 128-255 all-reduce values [2][64]
	static size_t bytes(){ return flagBytes() + 6*P2P_PLANE_CAP*sizeof(double); }
	static double *plane(char *a, int i){ return (double*)(a + flagBytes()) + (size_t)i*P2P_PLANE_CAP; }
	static unsigned long long *flag(char *a, int i){ return (unsigned long long*)a + i; }
};
struct Transport {
	virtual ~Transport() {}
	// all sends and receives of one exchange step; device pointers.  A message is identified by
	// (sender, receiver, tag); returns when the receives are ordered on ctx->stream.
	virtual void exchange(Ctx *c, std::vector<Msg> sends, std::vector<Msg> recvs) = 0;
	virtual void allreduceSum(Ctx *c, double *d_vals, int n) = 0;           // in place, device, stream ordered
	virtual void allgatherLong(Ctx *c, const long *h_in, int n, long *h_out) = 0;  // host values, blocking
	virtual void barrier(Ctx *c) = 0;
	virtual const char *name() const = 0;
	virtual P2P *p2p() { return nullptr; }     // peer-memory arena, if this transport has one
	// collective: a zeroed device allocation per rank that every rank can address (peers[r] = rank r's, peers[own] = *mine);
	// false (on every rank) if this transport cannot do that
	virtual bool peerAlloc(Ctx *, size_t, char **, std::vector<char*> &) { return false; }
	virtual void peerFree(Ctx *, char *, std::vector<char*> &) {}
};
Transport *makeSelfTransport();
void localCopies(Ctx *c, std::vector<Msg> &sends, std::vector<Msg> &recvs);   // matches and removes self messages

// ---- topology (src/grid.c:166-171, src/pusher.c:1181-1231) ----
int neighborToRank(const MpiInfo *m, int ne);
int neighborToReciprocal(int ne, int nDims);
int rankToNeighbor(const MpiInfo *m, int rank);

// ---- grid ops (grid.cu) ----
void gridScale(Ctx *c, DevGrid *g, double num);
void gridZero(Ctx *c, DevGrid *g);
void gridHaloDim(Ctx *c, DevGrid *g, const MpiInfo *m, int d /*1..3*/, int add, int dir);
void gridHalo(Ctx *c, DevGrid *g, const MpiInfo *m, int add, int dir);
bool gridHaloP2P(Ctx *c, DevGrid *g, const MpiInfo *m);
bool allSumP2P(Ctx *c, const double *partial, int n, double *out, const MpiInfo *m);   // multigrid.cu: one-double all-reduce over peer memory       // multigrid.cu: ghost fill over peer memory, false if unavailable
void gridHaloFaces(Ctx *c, DevGrid *g, const MpiInfo *m);      // faces of the decomposed dimensions only, one exchange
void gridNeutralize(Ctx *c, DevGrid *g, const MpiInfo *m);
void gridBnd(Ctx *c, DevGrid *g, const MpiInfo *m);             // gBnd (src/grid.c:992-1023): periodic, Dirichlet and Neumann edges
void gridEdge(Ctx *c, DevGrid *g, int boundary, int kind);     // gDirichlet / gNeumann of one edge
void gridUploadBnd(Ctx *c, DevGrid *g);                          // host bndSlice -> device (after the host changed it)
void gridAddTo(Ctx *c, DevGrid *r, const DevGrid *a);
void gridSubFrom(Ctx *c, DevGrid *r, const DevGrid *a);
// sum over the true grid of val (mode 0), val^2 after squaring in place (mode 1) or val*other (mode 2);
// the result lands in c->d_scal[slot] (this rank only, no all-reduce)
void gridSumTrue(Ctx *c, DevGrid *g, int mode, const DevGrid *other, int slot);
void gridSumTrueAll(Ctx *c, DevGrid *g, int mode, int slot, const MpiInfo *m);       // the same over all ranks
double readScalar(Ctx *c, int slot);           // D2H of d_scal[slot] + sync

// ---- multigrid (multigrid.cu) ----
void mgForgetPlans(Ctx *c);
void mgFreeArena(Ctx *c);                        // before the transport goes away
void mgConvergenceCheck(Ctx *c);                 // called by streamSync when a solve's history has landed
int mgMaxCyclesDefault();
void *mgProfBuffer(Ctx *c);
bool clusterSolve(Ctx *c, Multigrid *mgRho, Multigrid *mgPhi, Multigrid *mgRes, double tol, int maxCycles, int exact);

} // namespace pinc
