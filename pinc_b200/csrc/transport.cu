// transport.cu — the exchange steps between sub-domains (what MPI_Sendrecv / MPI_Isend / MPI_Allreduce
// do in the reference: src/grid.c:390-404, 744-745; src/pusher.c:914-1025; src/multigrid.c:1478).
//
//   self     one rank: every neighbour is this rank (periodic wrap), copies stay on the device;
//   threads  ranks are host threads of one process sharing one or more GPUs (tests): device-to-device
//            copies ordered by host barriers; no kernel ever waits on another rank;
//   nccl     one process per GPU: grouped ncclSend/ncclRecv + ncclAllReduce over NVLink.
#include "common.h"
#include <algorithm>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include <pthread.h>

namespace pinc {

// match self-addressed messages by tag, copy on the stream, and drop them from both lists
void localCopies(Ctx *c, std::vector<Msg> &sends, std::vector<Msg> &recvs){
	std::vector<Msg> s2, r2;
	std::vector<char> used(sends.size(), 0);
	for(auto &r : recvs){
		if(r.peer != c->rank){ r2.push_back(r); continue; }
		bool found = false;
		for(size_t i = 0; i < sends.size(); i++){
			if(used[i] || sends[i].peer != c->rank || sends[i].tag != r.tag) continue;
			if(sends[i].bytes != r.bytes) fatal("self message size mismatch (tag %d: %zu vs %zu)", r.tag, sends[i].bytes, r.bytes);
			if(r.bytes) PINC_CUDA(cudaMemcpyAsync(r.ptr, sends[i].ptr, r.bytes, cudaMemcpyDeviceToDevice, c->stream));
			used[i] = 1; found = true; break;
		}
		if(!found) fatal("self message with tag %d has no matching send", r.tag);
	}
	for(size_t i = 0; i < sends.size(); i++) if(!used[i]){
		if(sends[i].peer == c->rank) fatal("self send with tag %d has no matching receive", sends[i].tag);
		s2.push_back(sends[i]);
	}
	sends.swap(s2); recvs.swap(r2);
}

// ---------------------------------------------------------------------------------------------
struct SelfTransport : Transport {
	void exchange(Ctx *c, std::vector<Msg> sends, std::vector<Msg> recvs) override {
		localCopies(c, sends, recvs);
		if(!sends.empty() || !recvs.empty()) fatal("message to another rank but no transport was initialised (pincCommInit*)");
	}
	void allreduceSum(Ctx *, double *, int) override {}
	void allgatherLong(Ctx *, const long *h_in, int n, long *h_out) override { memcpy(h_out, h_in, n*sizeof(long)); }
	void barrier(Ctx *) override {}
	const char *name() const override { return "self"; }
};
Transport *makeSelfTransport(){ return new SelfTransport(); }

// ---------------------------------------------------------------------------------------------
struct Posted { int src, tag; void *ptr; size_t bytes; };
struct ThreadWorld {
	int n;
	pthread_barrier_t bar;
	std::mutex mu;
	std::vector<std::vector<Posted>> box;
	std::vector<std::vector<double>> dcoll;
	std::vector<std::vector<long>> lcoll;
	int refs;
};
struct ThreadTransport : Transport {
	ThreadWorld *w;
	explicit ThreadTransport(ThreadWorld *w_) : w(w_) {}
	~ThreadTransport() override {
		bool last;
		{ std::lock_guard<std::mutex> lk(w->mu); last = (--w->refs == 0); }
		if(last){ pthread_barrier_destroy(&w->bar); delete w; }
	}
	void exchange(Ctx *c, std::vector<Msg> sends, std::vector<Msg> recvs) override {
		localCopies(c, sends, recvs);
		streamSync(c);                                     // my send buffers are complete
		{
			std::lock_guard<std::mutex> lk(w->mu);
			for(auto &s : sends) w->box[s.peer].push_back({c->rank, s.tag, s.ptr, s.bytes});
		}
		pthread_barrier_wait(&w->bar);
		std::vector<Posted> mine;
		{ std::lock_guard<std::mutex> lk(w->mu); mine.swap(w->box[c->rank]); }
		std::vector<char> used(mine.size(), 0);
		for(auto &r : recvs){
			bool found = false;
			for(size_t i = 0; i < mine.size(); i++){
				if(used[i] || mine[i].src != r.peer || mine[i].tag != r.tag) continue;
				if(mine[i].bytes != r.bytes) fatal("message size mismatch from rank %d tag %d", r.peer, r.tag);
				if(r.bytes) PINC_CUDA(cudaMemcpyAsync(r.ptr, mine[i].ptr, r.bytes, cudaMemcpyDefault, c->stream));
				used[i] = 1; found = true; break;
			}
			if(!found) fatal("rank %d: no message from rank %d with tag %d", c->rank, r.peer, r.tag);
		}
		streamSync(c);                                     // copies done before the senders reuse their buffers
		pthread_barrier_wait(&w->bar);
	}
	void allreduceSum(Ctx *c, double *d_vals, int n) override {
		std::vector<double> &mine = w->dcoll[c->rank];
		mine.resize(n);
		PINC_CUDA(cudaMemcpyAsync(mine.data(), d_vals, n*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
		pthread_barrier_wait(&w->bar);
		std::vector<double> tot(n);
		for(int i = 0; i < n; i++){                        // rank order 0..n-1 on every rank: identical results
			double s = w->dcoll[0][i];
			for(int r = 1; r < w->n; r++) s += w->dcoll[r][i];
			tot[i] = s;
		}
		pthread_barrier_wait(&w->bar);
		PINC_CUDA(cudaMemcpyAsync(d_vals, tot.data(), n*sizeof(double), cudaMemcpyHostToDevice, c->stream));
		streamSync(c);
	}
	void allgatherLong(Ctx *c, const long *h_in, int n, long *h_out) override {
		w->lcoll[c->rank].assign(h_in, h_in + n);
		pthread_barrier_wait(&w->bar);
		for(int r = 0; r < w->n; r++) memcpy(h_out + (size_t)r*n, w->lcoll[r].data(), n*sizeof(long));
		pthread_barrier_wait(&w->bar);
	}
	void barrier(Ctx *c) override { streamSync(c); pthread_barrier_wait(&w->bar); }
	const char *name() const override { return "threads"; }
	// one process: the pointers are valid everywhere on the same device; other devices need peer access
	bool peerAlloc(Ctx *c, size_t bytes, char **mine, std::vector<char*> &peers) override {
		char *p = nullptr;
		long ok = cudaMalloc(&p, bytes) == cudaSuccess && cudaMemset(p, 0, bytes) == cudaSuccess;
		if(!ok) cudaGetLastError();
		std::vector<long> all(3*(size_t)w->n);
		long v[3] = { (long)(uintptr_t)p, (long)c->device, ok };
		allgatherLong(c, v, 3, all.data());
		for(int r = 0; r < w->n && ok; r++){
			if(!all[3*r+2]) ok = 0;
			else if(all[3*r+1] != c->device){
				int can = 0;
				cudaDeviceCanAccessPeer(&can, c->device, (int)all[3*r+1]);
				if(!can) ok = 0;
				else { cudaError_t e = cudaDeviceEnablePeerAccess((int)all[3*r+1], 0); if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0; cudaGetLastError(); }
			}
		}
		long okAll[1] = { ok };
		std::vector<long> oks(w->n);
		allgatherLong(c, okAll, 1, oks.data());
		for(int r = 0; r < w->n; r++) if(!oks[r]) ok = 0;
		if(!ok){ if(p) cudaFree(p); return false; }
		peers.assign(w->n, nullptr);
		for(int r = 0; r < w->n; r++) peers[r] = (char*)(uintptr_t)all[3*r];
		*mine = p;
		return true;
	}
	void peerFree(Ctx *, char *mine, std::vector<char*> &peers) override { if(mine) cudaFree(mine); peers.clear(); }
};

// ---------------------------------------------------------------------------------------------
struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*);
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*GroupStart)();
	ncclResult_t (*GroupEnd)();
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi *ncclApi(){
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [](){
		const char *names[] = { "libnccl.so.2", "libnccl.so" };
		for(const char *n : names){ api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if(api.h) break; }
		if(!api.h) fatal("cannot load libnccl.so.2: %s", dlerror());
#define L(f) *(void**)(&api.f) = dlsym(api.h, "nccl" #f); if(!api.f) fatal("libnccl lacks nccl" #f)
		L(GetUniqueId); L(CommInitRank); L(CommDestroy); L(GroupStart); L(GroupEnd); L(Send); L(Recv);
		L(AllReduce); L(AllGather); L(GetErrorString);
#undef L
	});
	return &api;
}
#define PINC_NCCL(call) do { ncclResult_t r_ = (call); if(r_ != ncclSuccess) \
	::pinc::fatal("NCCL error at %s:%d: %s", __FILE__, __LINE__, ncclApi()->GetErrorString(r_)); } while(0)

struct NcclTransport : Transport {
	ncclComm_t comm = nullptr;
	P2P peer; bool peerOk = false;
	P2P *p2p() override { return peerOk ? &peer : nullptr; }
	~NcclTransport() override {
		for(size_t r = 0; r < peer.peerArena.size(); r++) if(peer.peerArena[r] && peer.peerArena[r] != peer.arena) cudaIpcCloseMemHandle(peer.peerArena[r]);
		if(peer.arena) cudaFree(peer.arena);
		if(comm) ncclApi()->CommDestroy(comm);
	}
	// map every rank's arena into this process (all ranks are processes on one NVLink/NVSwitch node)
	void setupPeer(Ctx *c){
		if(getenv("PINC_B200_NO_P2P")) return;
		NcclApi *n = ncclApi();
		int ok = cudaMalloc(&peer.arena, P2P::bytes()) == cudaSuccess;
		cudaIpcMemHandle_t mine; memset(&mine, 0, sizeof mine);
		if(ok){ cudaMemset(peer.arena, 0, P2P::bytes()); ok = cudaIpcGetMemHandle(&mine, peer.arena) == cudaSuccess; }
		char *d = (char*)tmpBuffer(c, sizeof(mine)*(c->size + 1));
		PINC_CUDA(cudaMemcpyAsync(d, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
		PINC_NCCL(n->AllGather(d, d + sizeof mine, sizeof mine, ncclChar, comm, c->stream));
		std::vector<cudaIpcMemHandle_t> all(c->size);
		PINC_CUDA(cudaMemcpyAsync(all.data(), d + sizeof mine, sizeof(mine)*c->size, cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
		peer.peerArena.assign(c->size, nullptr);
		for(int r = 0; r < c->size && ok; r++){
			if(r == c->rank){ peer.peerArena[r] = peer.arena; continue; }
			void *p = nullptr;
			if(cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess){ ok = 0; cudaGetLastError(); }
			peer.peerArena[r] = (char*)p;
		}
		// everybody or nobody
		double v = ok ? 0.0 : 1.0;
		PINC_CUDA(cudaMemcpyAsync(c->d_scal + 254, &v, sizeof v, cudaMemcpyHostToDevice, c->stream));
		PINC_NCCL(n->AllReduce(c->d_scal + 254, c->d_scal + 254, 1, ncclDouble, ncclSum, comm, c->stream));
		PINC_CUDA(cudaMemcpyAsync(&v, c->d_scal + 254, sizeof v, cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
		peerOk = (v == 0.0);
		if(!peerOk && c->rank == 0) fprintf(stderr, "PINC-B200 WARNING: peer-to-peer arena unavailable, halo exchange stays on NCCL send/recv\n");
	}
	// one process per GPU on one NVLink/NVSwitch node: CUDA IPC handles, all-gathered through NCCL
	bool peerAlloc(Ctx *c, size_t bytes, char **mine, std::vector<char*> &peers) override {
		if(getenv("PINC_B200_NO_P2P")) return false;
		NcclApi *n = ncclApi();
		char *p = nullptr;
		int ok = cudaMalloc(&p, bytes) == cudaSuccess;
		cudaIpcMemHandle_t h; memset(&h, 0, sizeof h);
		if(ok) ok = cudaMemset(p, 0, bytes) == cudaSuccess && cudaIpcGetMemHandle(&h, p) == cudaSuccess;
		if(!ok) cudaGetLastError();
		char *d = (char*)tmpBuffer(c, sizeof(h)*(c->size + 1));
		PINC_CUDA(cudaMemcpyAsync(d, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
		PINC_NCCL(n->AllGather(d, d + sizeof h, sizeof h, ncclChar, comm, c->stream));
		std::vector<cudaIpcMemHandle_t> all(c->size);
		PINC_CUDA(cudaMemcpyAsync(all.data(), d + sizeof h, sizeof(h)*c->size, cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
		peers.assign(c->size, nullptr);
		for(int r = 0; r < c->size && ok; r++){
			if(r == c->rank){ peers[r] = p; continue; }
			void *q = nullptr;
			if(cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess){ ok = 0; cudaGetLastError(); }
			peers[r] = (char*)q;
		}
		double v = ok ? 0.0 : 1.0;                       // everybody or nobody
		PINC_CUDA(cudaMemcpyAsync(c->d_scal + 254, &v, sizeof v, cudaMemcpyHostToDevice, c->stream));
		PINC_NCCL(n->AllReduce(c->d_scal + 254, c->d_scal + 254, 1, ncclDouble, ncclSum, comm, c->stream));
		PINC_CUDA(cudaMemcpyAsync(&v, c->d_scal + 254, sizeof v, cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
		if(v != 0.0){
			for(int r = 0; r < c->size; r++) if(peers[r] && peers[r] != p) cudaIpcCloseMemHandle(peers[r]);
			if(p) cudaFree(p);
			peers.clear();
			return false;
		}
		*mine = p;
		return true;
	}
	void peerFree(Ctx *, char *mine, std::vector<char*> &peers) override {
		for(char *q : peers) if(q && q != mine) cudaIpcCloseMemHandle(q);
		if(mine) cudaFree(mine);
		peers.clear();
	}
	void exchange(Ctx *c, std::vector<Msg> sends, std::vector<Msg> recvs) override {
		localCopies(c, sends, recvs);
		// NCCL pairs the sends and receives of two ranks in posting order: order both sides by tag
		auto byPeerTag = [](const Msg &a, const Msg &b){ return a.peer != b.peer ? a.peer < b.peer : a.tag < b.tag; };
		std::sort(sends.begin(), sends.end(), byPeerTag);
		std::sort(recvs.begin(), recvs.end(), byPeerTag);
		NcclApi *n = ncclApi();
		bool any = false;
		for(auto &s : sends) if(s.bytes) any = true;
		for(auto &r : recvs) if(r.bytes) any = true;
		if(!any) return;
		PINC_NCCL(n->GroupStart());
		for(auto &s : sends) if(s.bytes) PINC_NCCL(n->Send(s.ptr, s.bytes, ncclChar, s.peer, comm, c->stream));
		for(auto &r : recvs) if(r.bytes) PINC_NCCL(n->Recv(r.ptr, r.bytes, ncclChar, r.peer, comm, c->stream));
		PINC_NCCL(n->GroupEnd());
	}
	void allreduceSum(Ctx *c, double *d_vals, int n) override {
		PINC_NCCL(ncclApi()->AllReduce(d_vals, d_vals, n, ncclDouble, ncclSum, comm, c->stream));
	}
	void allgatherLong(Ctx *c, const long *h_in, int n, long *h_out) override {
		long *d = (long*)tmpBuffer(c, (size_t)n*(c->size + 1)*sizeof(long));
		PINC_CUDA(cudaMemcpyAsync(d, h_in, n*sizeof(long), cudaMemcpyHostToDevice, c->stream));
		PINC_NCCL(ncclApi()->AllGather(d, d + n, n, ncclInt64, comm, c->stream));
		PINC_CUDA(cudaMemcpyAsync(h_out, d + n, (size_t)n*c->size*sizeof(long), cudaMemcpyDeviceToHost, c->stream));
		streamSync(c);
	}
	void barrier(Ctx *c) override {
		PINC_NCCL(ncclApi()->AllReduce(c->d_scal + 255, c->d_scal + 255, 1, ncclDouble, ncclSum, comm, c->stream));
		streamSync(c);
	}
	const char *name() const override { return "nccl"; }
};

} // namespace pinc

using namespace pinc;

extern "C" {

void pincCommInitThreads(PincCtx **ctxs, int n){
	ThreadWorld *w = new ThreadWorld();
	w->n = n; w->refs = n;
	pthread_barrier_init(&w->bar, nullptr, n);
	w->box.resize(n); w->dcoll.resize(n); w->lcoll.resize(n);
	for(int r = 0; r < n; r++){
		Ctx *c = (Ctx*)ctxs[r];
		if(c->rank != r || c->size != n) fatal("pincCommInitThreads: context %d has rank %d of %d", r, c->rank, c->size);
		mgFreeArena(c);
		delete c->tp;
		c->tp = new ThreadTransport(w);
	}
}

int pincNcclUniqueId(char *out128){
	ncclUniqueId id;
	PINC_NCCL(ncclApi()->GetUniqueId(&id));
	static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
	memcpy(out128, &id, 128);
	return 128;
}

void pincCommInitNccl(PincCtx *ctx, const char *uniqueId128){
	Ctx *c = (Ctx*)ctx;
	PINC_CUDA(cudaSetDevice(c->device));
	ncclUniqueId id;
	memcpy(&id, uniqueId128, 128);
	NcclTransport *t = new NcclTransport();
	PINC_NCCL(ncclApi()->CommInitRank(&t->comm, c->size, id, c->rank));
	mgFreeArena(c);
	delete c->tp;
	c->tp = t;
	t->setupPeer(c);
}

const char *pincTransportName(void){ return cur()->tp->name(); }

} // extern "C"
