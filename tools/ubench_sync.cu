// Micro-benchmark behind the multigrid design: cost of one dependent phase under different barriers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_sync ubench_sync.cu && ./ubench_sync
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void k_cluster_sync(int n, long long *out){
	cg::cluster_group cl = cg::this_cluster();
	long long t0 = clock64();
	for(int i = 0; i < n; i++) cl.sync();
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
__global__ void k_block_sync(int n, long long *out){
	long long t0 = clock64();
	for(int i = 0; i < n; i++) __syncthreads();
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
// one phase = read a neighbour CTA's shared value, add, write own, cluster barrier
__global__ void k_cluster_phase(int n, long long *out, double *sink){
	__shared__ double v[1024];
	cg::cluster_group cl = cg::this_cluster();
	int r = cl.block_rank(), nc = cl.num_blocks();
	v[threadIdx.x] = threadIdx.x;
	cl.sync();
	double *nb = cl.map_shared_rank(v, (r+1)%nc);
	long long t0 = clock64();
	for(int i = 0; i < n; i++){
		double x = nb[threadIdx.x] + v[(threadIdx.x+1)&511];
		cl.sync();
		v[threadIdx.x] = x*0.5;
		cl.sync();
	}
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0){ out[0] = t1 - t0; sink[0] = v[0]; }
}
__global__ void k_l2_chain(int n, const int *p, long long *out, int *sink){
	int idx = 0;
	long long t0 = clock64();
	for(int i = 0; i < n; i++) idx = __ldcg(p + idx);
	long long t1 = clock64();
	out[0] = t1 - t0; sink[0] = idx;
}
__global__ void k_grid_sync(int n, unsigned *bar, long long *out){
	unsigned gen = 0;
	long long t0 = clock64();
	for(int i = 0; i < n; i++){
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
	long long t1 = clock64();
	if(threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main(){
	long long *out; double *sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 8);
	int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	const int n = 2000;
	long long h;
	cudaFuncSetAttribute((const void*)k_cluster_sync, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
	cudaFuncSetAttribute((const void*)k_cluster_phase, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
	for(int threads : {128, 512, 1024}) for(int nc : {1, 2, 4, 8, 16}){
		cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(nc); cfg.blockDim = dim3(threads);
		cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
		cfg.attrs = at; cfg.numAttrs = 1;
		for(int rep = 0; rep < 2; rep++){ cudaLaunchKernelEx(&cfg, k_cluster_sync, n, out); cudaDeviceSynchronize(); }
		cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
		printf("cluster.sync     threads=%4d cluster=%2d : %7.1f cycles  (%s)\n", threads, nc, (double)h/n, cudaGetErrorString(cudaGetLastError()));
		if(threads == 512){
			for(int rep = 0; rep < 2; rep++){ cudaLaunchKernelEx(&cfg, k_cluster_phase, n, out, sink); cudaDeviceSynchronize(); }
			cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
			printf("dsmem phase(2 syncs) threads=%4d cluster=%2d : %7.1f cycles per phase\n", threads, nc, (double)h/n);
		}
	}
	for(int threads : {128, 512, 1024}){
		for(int rep = 0; rep < 2; rep++){ k_block_sync<<<1,threads>>>(n, out); cudaDeviceSynchronize(); }
		cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
		printf("__syncthreads    threads=%4d            : %7.1f cycles\n", threads, (double)h/n);
	}
	{
		int N = 1<<20; int *p, *hs = new int[N]; for(int i = 0; i < N; i++) hs[i] = (i*7919 + 12345) % N;
		cudaMalloc(&p, N*4); cudaMemcpy(p, hs, N*4, cudaMemcpyHostToDevice);
		int *s2; cudaMalloc(&s2, 4);
		for(int rep = 0; rep < 2; rep++){ k_l2_chain<<<1,1>>>(n, p, out, s2); cudaDeviceSynchronize(); }
		cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
		printf("dependent __ldcg (L2 hit, 4 MB set)      : %7.1f cycles\n", (double)h/n);
	}
	unsigned *bar; cudaMalloc(&bar, 8); cudaMemset(bar, 0, 8);
	for(int grid : {16, 32, 74, 148}){
		for(int rep = 0; rep < 2; rep++){
			cudaMemset(bar, 0, 8);
			void *args[] = { (void*)&n, (void*)&bar, (void*)&out };
			cudaLaunchCooperativeKernel((void*)k_grid_sync, dim3(grid), dim3(512), args, 0, 0); cudaDeviceSynchronize();
		}
		cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
		printf("grid barrier (atomics through L2) grid=%3d: %7.1f cycles\n", grid, (double)h/n);
	}
	printf("SM clock %d kHz\n", clk);
	return 0;
}
