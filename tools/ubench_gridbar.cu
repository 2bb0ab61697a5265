// Micro-benchmark: grid-wide barrier variants for the persistent multigrid kernel (148 CTAs x 512 threads).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda/atomic>

// A: fence + arrive counter + generation word + fence (what cooperative_groups' grid.sync does)
__device__ __forceinline__ void barA(unsigned *bar, unsigned &gen){
	__syncthreads();
	if(threadIdx.x == 0){
		__threadfence();
		unsigned old = atomicAdd(&bar[0], 1u);
		if(old == gridDim.x-1){ atomicExch(&bar[0], 0u); __threadfence(); atomicAdd(&bar[1], 1u); }
		else while(*((volatile unsigned*)&bar[1]) == gen){}
		__threadfence();
		gen++;
	}
	__syncthreads();
}
// B: monotonic counter, fence before arrive only; data is read with ld.cg after the barrier
__device__ __forceinline__ void barB(unsigned *bar, unsigned &gen){
	__syncthreads();
	if(threadIdx.x == 0){
		__threadfence();
		atomicAdd(&bar[0], 1u);
		unsigned target = (gen+1)*gridDim.x;
		while(*((volatile unsigned*)&bar[0]) < target){}
		gen++;
	}
	__syncthreads();
}
// C: release RMW / acquire poll (PTX memory model, no stand-alone fences)
__device__ __forceinline__ void barC(unsigned *bar, unsigned &gen){
	__syncthreads();
	if(threadIdx.x == 0){
		cuda::atomic_ref<unsigned, cuda::thread_scope_device> a(bar[0]);
		a.fetch_add(1u, cuda::memory_order_release);
		unsigned target = (gen+1)*gridDim.x;
		while(a.load(cuda::memory_order_acquire) < target){}
		gen++;
	}
	__syncthreads();
}
// D: no ordering at all (floor: atomic round trip + poll)
__device__ __forceinline__ void barD(unsigned *bar, unsigned &gen){
	__syncthreads();
	if(threadIdx.x == 0){
		atomicAdd(&bar[0], 1u);
		unsigned target = (gen+1)*gridDim.x;
		while(*((volatile unsigned*)&bar[0]) < target){}
		gen++;
	}
	__syncthreads();
}
// E: release RMW, relaxed polls, one acquire fence at the end
__device__ __forceinline__ void barE(unsigned *bar, unsigned &gen){
	__syncthreads();
	if(threadIdx.x == 0){
		cuda::atomic_ref<unsigned, cuda::thread_scope_device> a(bar[0]);
		a.fetch_add(1u, cuda::memory_order_release);
		unsigned target = (gen+1)*gridDim.x;
		while(a.load(cuda::memory_order_relaxed) < target){}
		gen++;
	}
	__syncthreads();
}
template<int V> __global__ void k(int n, unsigned *bar, long long *out, double *data){
	unsigned gen = 0;
	long long t0 = clock64();
	double acc = 0;
	for(int i = 0; i < n; i++){
		// a little work with global stores so that the fences have something to order
		data[blockIdx.x*blockDim.x + threadIdx.x] = acc + i;
		if(V == 0) barA(bar, gen); else if(V == 1) barB(bar, gen); else if(V == 2) barC(bar, gen); else if(V == 3) barD(bar, gen); else barE(bar, gen);
		acc += __ldcg(&data[((blockIdx.x+1)%gridDim.x)*blockDim.x + threadIdx.x]);
	}
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0){ out[0] = t1 - t0; }
	if(acc == 1.2345) out[1] = 1;
}
// F: no barrier at all: every thread stores {value, tag} as one 16-byte line into the mailbox of 6 other CTAs' threads
// and polls its own 6 slots until the tag matches (NCCL-LL style: data and flag travel together)
__device__ __forceinline__ void llStore(uint4 *p, double v, unsigned tag){
	unsigned lo = (unsigned)__double_as_longlong(v), hi = (unsigned)(__double_as_longlong(v) >> 32);
	asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ double llPoll(const uint4 *p, unsigned tag){
	unsigned a, b, c, d;
	do { asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory"); } while((int)(b - tag) < 0 || (int)(d - tag) < 0);      // >=: a one-directional ring has no back-pressure
	return __longlong_as_double(((long long)c << 32) | a);
}
template<int NB, int ACTIVE> __global__ void kLL(int n, uint4 *mail, long long *out){
	long long t0 = clock64();
	double acc = 1.0;
	__shared__ double sh[1024];
	const int offs[6] = {1, -1, 4, -4, 16, -16};
	for(int i = 0; i < n; i++){
		if(threadIdx.x < ACTIVE){
			for(int f = 0; f < NB; f++){
				int dst = ((int)blockIdx.x + offs[f] + 4*(int)gridDim.x) % (int)gridDim.x;
				llStore(mail + ((size_t)dst*6 + f)*512 + threadIdx.x, acc + f, (unsigned)(i+1));
			}
			double s = 0;
			for(int f = 0; f < NB; f++) s += llPoll(mail + ((size_t)blockIdx.x*6 + f)*512 + threadIdx.x, (unsigned)(i+1));
			sh[threadIdx.x] = s;
		}
		__syncthreads();
		acc = sh[(threadIdx.x*7) % ACTIVE]*0.125;
		__syncthreads();
	}
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0){ out[0] = t1 - t0; }
	if(acc == 1.2345) out[1] = 1;
}
// G/H/I: the shape the block smoother needs: every thread owns ONE slot (512 per CTA and phase, ~85 per face), sends it to
// the matching neighbour and receives one.  MODE 0: spin on the own slot; 1: six lanes spin on one sentinel slot per face,
// __syncthreads, then everybody reads its slot; 2: spin with __nanosleep back-off; 3: fence + flag per face + ldcg
__device__ __forceinline__ bool llTry(const uint4 *p, unsigned tag, double &v){
	unsigned a, b, c, d;
	asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
	v = __longlong_as_double(((long long)c << 32) | a);
	return (int)(b - tag) >= 0 && (int)(d - tag) >= 0;
}
__device__ __forceinline__ void llStoreCg(uint4 *p, double v, unsigned tag){
	unsigned lo = (unsigned)__double_as_longlong(v), hi = (unsigned)(__double_as_longlong(v) >> 32);
	asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ bool llTryCg(const uint4 *p, unsigned tag, double &v){
	unsigned a, b, c, d;
	asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
	v = __longlong_as_double(((long long)c << 32) | a);
	return (int)(b - tag) >= 0 && (int)(d - tag) >= 0;
}
__device__ int g_perFace = 85;
template<int MODE> __global__ void kLL2(int n, uint4 *mail, long long *out, unsigned *flags, double *plain){
	long long t0 = clock64();
	double acc = 1.0;
	__shared__ double sh[512];
	const int offs[6] = {1, -1, 4, -4, 16, -16};
	const int PF = g_perFace;
	const bool act = threadIdx.x < 6*PF;
	const int f = act ? threadIdx.x / PF : 0, idx = threadIdx.x % PF;          // face-contiguous slots, as in the real smoother
	const int dst = ((int)blockIdx.x + offs[f] + 4*(int)gridDim.x) % (int)gridDim.x;
	const int sendSlot = (f^1)*85 + idx, mySlot = f*85 + idx;
	if(MODE == 4){ for(int i = 0; i < n; i++){ unsigned tag = (unsigned)(i+1); double v = 0; if(act){ llStoreCg(mail + (size_t)dst*512 + sendSlot, acc, tag); while(!llTryCg(mail + (size_t)blockIdx.x*512 + mySlot, tag, v)){ } } sh[threadIdx.x] = v; __syncthreads(); acc = sh[(threadIdx.x*7) & 511]*0.125 + 1.0; }
		if(threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0; if(acc == 1.2345) out[1] = 1; return; }
	for(int i = 0; i < n; i++){
		unsigned tag = (unsigned)(i+1);
		double v = 0;
		if(MODE == 3){
			if(act) plain[(size_t)dst*512 + sendSlot] = acc;
			__syncthreads();
			if(threadIdx.x == 0) __threadfence();
			if(threadIdx.x < 6){
				int d2 = ((int)blockIdx.x + offs[threadIdx.x] + 4*(int)gridDim.x) % (int)gridDim.x;
				if(threadIdx.x == 0){ for(int q = 0; q < 6; q++){ int d3 = ((int)blockIdx.x + offs[q] + 4*(int)gridDim.x) % (int)gridDim.x; *((volatile unsigned*)&flags[d3*8 + (q^1)]) = tag; } }
				(void)d2;
				while((int)(*((volatile unsigned*)&flags[blockIdx.x*8 + threadIdx.x]) - tag) < 0){ }
			}
			__syncthreads();
			if(act) v = __ldcg(&plain[(size_t)blockIdx.x*512 + mySlot]);
		} else {
			if(act) llStore(mail + (size_t)dst*512 + sendSlot, acc, tag);
			if(MODE == 0 && act){ while(!llTry(mail + (size_t)blockIdx.x*512 + mySlot, tag, v)){ } }
			if(MODE == 2 && act){ while(!llTry(mail + (size_t)blockIdx.x*512 + mySlot, tag, v)){ __nanosleep(40); } }
			if(MODE == 1){
				if(threadIdx.x < 6){ double w; while(!llTry(mail + (size_t)blockIdx.x*512 + threadIdx.x*85 + 84, tag, w)){ } }
				__syncthreads();
				if(act) while(!llTry(mail + (size_t)blockIdx.x*512 + mySlot, tag, v)){ }
			}
		}
		sh[threadIdx.x] = v;
		__syncthreads();
		acc = sh[(threadIdx.x*7) & 511]*0.125 + 1.0;
	}
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0){ out[0] = t1 - t0; }
	if(acc == 1.2345) out[1] = 1;
}
template<int MODE> void runLL2(int grid, long long *out, int perFace = 85){
	cudaMemcpyToSymbol(g_perFace, &perFace, sizeof(int));
	uint4 *mail; cudaMalloc(&mail, (size_t)148*512*16); unsigned *flags; cudaMalloc(&flags, 148*8*4); double *plain; cudaMalloc(&plain, 148*512*8);
	const int n = 2000; long long h = 0;
	for(int rep = 0; rep < 2; rep++){
		cudaMemset(mail, 0, (size_t)148*512*16); cudaMemset(flags, 0, 148*8*4);
		void *args[] = {(void*)&n, (void*)&mail, (void*)&out, (void*)&flags, (void*)&plain};
		cudaLaunchCooperativeKernel((void*)kLL2<MODE>, dim3(grid), dim3(512), args, 0, 0);
		cudaDeviceSynchronize();
	}
	cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
	const char *nm[] = {"G spin on own slot", "H sentinel per face, then own slot", "I spin with nanosleep(40)", "J plain stores + fence + flag per face + ldcg", "K spin on own slot, st/ld.relaxed.gpu"};
	printf("%-46s grid=%3d slots/face=%2d : %7.1f cycles per phase  %s\n", nm[MODE], grid, perFace, (double)h/n, cudaGetErrorString(cudaGetLastError()));
	cudaFree(mail); cudaFree(flags); cudaFree(plain);
}
// S: how long does the ISSUE of three 16-byte stores to three different lines take (clock before/after, no wait)?
__device__ int g_act = 512;
template<int KIND> __global__ void kStoreIssue(uint4 *buf, long long *out, int n){
	const bool on = (int)threadIdx.x < g_act;
	long long tot = 0;
	for(int i = 0; i < n; i++){
		__syncthreads();
		long long t0 = clock64();
		#pragma unroll
		for(int q = 0; q < 3; q++){
			if(!on) continue;
			uint4 *p = buf + ((size_t)((blockIdx.x*7 + q*41 + i) % 148)*4096 + threadIdx.x + q*1024);
			unsigned a = i, b = i + 1;
			if(KIND == 0) asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(a), "r"(b) : "memory");
			if(KIND == 1) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(a), "r"(b) : "memory");
			if(KIND == 2) asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(a), "r"(b) : "memory");
			if(KIND == 3) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(a), "r"(b) : "memory");
		}
		tot += clock64() - t0;
	}
	if(threadIdx.x == 0 && blockIdx.x == 0) out[0] = tot;
}
template<int KIND> void runStoreIssue(long long *out, int act = 512){
	cudaMemcpyToSymbol(g_act, &act, sizeof(int));
	uint4 *buf; cudaMalloc(&buf, (size_t)148*4096*16);
	int n = 1000; long long h = 0;
	for(int rep = 0; rep < 2; rep++){ kStoreIssue<KIND><<<128, 512>>>(buf, out, n); cudaDeviceSynchronize(); }
	cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
	const char *nm[] = {"st.relaxed.gpu", "st.global.cg", "st.volatile", "st.global (default)"};
	printf("S issue of 3 x 16-byte %-20s active threads %3d : %7.1f cycles  %s\n", nm[KIND], act, (double)h/n, cudaGetErrorString(cudaGetLastError()));
	cudaFree(buf);
}
template<int NB, int ACTIVE> void runLL(int grid, long long *out){
	uint4 *mail; cudaMalloc(&mail, (size_t)148*6*512*16);
	const int n = 2000; long long h = 0;
	for(int rep = 0; rep < 2; rep++){
		cudaMemset(mail, 0, (size_t)148*6*512*16);
		void *args[] = {(void*)&n, (void*)&mail, (void*)&out};
		cudaLaunchCooperativeKernel((void*)kLL<NB,ACTIVE>, dim3(grid), dim3(512), args, 0, 0);
		cudaDeviceSynchronize();
	}
	cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
	printf("F LL mailbox, %d neighbours, %3d polling threads, grid=%3d : %7.1f cycles per (store + poll + 2 syncthreads)  %s\n", NB, ACTIVE, grid, (double)h/n, cudaGetErrorString(cudaGetLastError()));
	cudaFree(mail);
}
int main(){
	{ long long *o; cudaMalloc(&o, 16);
	  runStoreIssue<0>(o); runStoreIssue<1>(o); runStoreIssue<0>(o, 1); runStoreIssue<0>(o, 32); runStoreIssue<0>(o, 128); runStoreIssue<3>(o, 32);
	  runLL<2,32>(16, o); runLL<2,32>(128, o); runLL<6,32>(128, o); runLL<6,128>(128, o); runLL<6,512>(128, o); runLL<6,512>(148, o); runLL<2,512>(128, o);
	  runLL2<0>(128, o); runLL2<1>(128, o); runLL2<2>(128, o); runLL2<3>(128, o); runLL2<4>(128, o);
	  for(int pf : {1, 5, 21, 42}){ runLL2<0>(128, o, pf); runLL2<4>(128, o, pf); } runLL2<0>(16, o, 85); runLL2<0>(64, o, 85); cudaFree(o); }
	long long *out; cudaMalloc(&out, 16);
	unsigned *bar; cudaMalloc(&bar, 8);
	double *data; cudaMalloc(&data, 148*1024*8);
	const int n = 2000;
	const char *names[] = {"A fence+arrive+gen+fence", "B fence+monotonic counter", "C release add/acquire poll", "D no ordering (floor)", "E release add/relaxed poll"};
	void *funcs[] = {(void*)k<0>, (void*)k<1>, (void*)k<2>, (void*)k<3>, (void*)k<4>};
	for(int threads : {256, 512}) for(int v = 0; v < 5; v++) for(int grid : {16, 148}){
		long long h = 0;
		for(int rep = 0; rep < 2; rep++){
			cudaMemset(bar, 0, 8);
			void *args[] = {(void*)&n, (void*)&bar, (void*)&out, (void*)&data};
			cudaLaunchCooperativeKernel(funcs[v], dim3(grid), dim3(threads), args, 0, 0);
			cudaDeviceSynchronize();
		}
		cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
		printf("%-28s threads=%d grid=%3d : %7.1f cycles per (store + barrier + ldcg)  %s\n", names[v], threads, grid, (double)h/n, cudaGetErrorString(cudaGetLastError()));
	}
	return 0;
}
