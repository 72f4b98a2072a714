// Random aligned gathers of G bytes (32 / 64 / 128) from a working set of W bytes: how many requests per second does the
// memory system deliver when the set is L2-resident, and when it is not?  (Ceiling of the band-major projection kernel.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_gather_bench tools/l2_gather_bench.cu && tools/l2_gather_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// LPG lanes share one gather of LPG * 16 bytes; UNROLL independent gathers in flight per lane
template <int LPG, int UNROLL>
__global__ void __launch_bounds__(256) k_gather(const float4 *buf, uint32_t ncell_mask, int iters, float *out)
{
	const uint32_t lane = threadIdx.x & 31, g = (blockIdx.x * blockDim.x + threadIdx.x) / LPG, k = lane % LPG;
	float acc = 0.f;
	uint32_t s = g * 2654435761u + 12345u;
	for (int i = 0; i < iters; i++)
	{
		float4 v[UNROLL];
#pragma unroll
		for (int u = 0; u < UNROLL; u++)
		{
			s = hash32(s + u + 1);
			const uint32_t cell = s & ncell_mask;
			v[u] = __ldcg(buf + (size_t) cell * LPG + k);
		}
#pragma unroll
		for (int u = 0; u < UNROLL; u++) acc += v[u].x + v[u].w;
	}
	if (acc == 123.456f) out[0] = acc;
}

template <int LPG, int UNROLL>
void run(const float4 *buf, size_t wbytes, float *out, const char *name)
{
	const uint32_t ncell = (uint32_t) (wbytes / (LPG * 16));   // power of two
	int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
	const int grid = p.multiProcessorCount * 8, iters = 400;
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	k_gather<LPG, UNROLL><<<grid, 256>>>(buf, ncell - 1, 20, out);
	cudaEventRecord(a);
	k_gather<LPG, UNROLL><<<grid, 256>>>(buf, ncell - 1, iters, out);
	cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b);
	const double req = (double) grid * 256 / LPG * iters * UNROLL;
	printf("%-10s W=%7.0f MB  gather=%3d B  unroll=%d  %7.1f Greq/s  %7.2f TB/s  (%.3f ms)\n", name, wbytes / 1048576.0, LPG * 16, UNROLL,
	       req / ms * 1e-6, req * LPG * 16 / ms * 1e-9, ms);
}

int main()
{
	const size_t maxb = (size_t) 4 << 30;
	float4 *buf; float *out;
	cudaMalloc(&buf, maxb); cudaMalloc(&out, 16);
	cudaMemset(buf, 1, maxb);
	for (size_t w : {(size_t) 16 << 20, (size_t) 32 << 20, (size_t) 64 << 20, (size_t) 128 << 20, (size_t) 1 << 30, (size_t) 4 << 30})
	{
		run<2, 4>(buf, w, out, "ldg 32B");
		run<4, 4>(buf, w, out, "ldg 64B");
		run<8, 4>(buf, w, out, "ldg 128B");
		run<4, 8>(buf, w, out, "ldg 64B u8");
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
