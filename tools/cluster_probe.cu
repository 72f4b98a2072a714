// Probe: which (cluster shape, dynamic shared memory) launches this GPU accepts.
// nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_probe cluster_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
extern __shared__ unsigned char sm[];
__global__ void __launch_bounds__(192, 1) k_probe(int *out)
{
	unsigned r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	if (out == nullptr) sm[threadIdx.x] = (unsigned char) r;
	if (threadIdx.x == 0) atomicAdd(out, (int) r + 1);
}
int main()
{
	int v = 0;
	cudaDeviceGetAttribute(&v, cudaDevAttrClusterLaunch, 0);
	printf("cluster launch supported: %d\n", v);
	int *d; cudaMalloc(&d, 4);
	const size_t sizes[] = {0, 100 << 10, 197888, 220 << 10};
	for (size_t s : sizes)
		for (int shape = 0; shape < 2; shape++)
		{
			cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s);
			cudaLaunchConfig_t cfg = {};
			cfg.gridDim = shape ? dim3(4, 2) : dim3(2, 4); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = s;
			cudaLaunchAttribute a[1];
			a[0].id = cudaLaunchAttributeClusterDimension;
			a[0].val.clusterDim.x = shape ? 1 : 2; a[0].val.clusterDim.y = shape ? 2 : 1; a[0].val.clusterDim.z = 1;
			cfg.attrs = a; cfg.numAttrs = 1;
			int nc = -1;
			cudaError_t eo = cudaOccupancyMaxActiveClusters(&nc, k_probe, &cfg);
			cudaMemset(d, 0, 4);
			cudaError_t el = cudaLaunchKernelEx(&cfg, k_probe, d);
			cudaError_t es = cudaDeviceSynchronize();
			int h = 0; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
			printf("smem %zu cluster %s: maxActiveClusters %d (%s) launch %s sync %s sum %d\n", s, shape ? "(1,2,1)" : "(2,1,1)", nc,
			       cudaGetErrorName(eo), cudaGetErrorName(el), cudaGetErrorName(es), h);
			cudaGetLastError();
		}
	return 0;
}
