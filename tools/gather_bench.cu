// Microbenchmark: ceiling of random 64-byte gathers / 16-byte vector reductions on this GPU (HBM-resident arrays).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu ; run on the B200 box.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t rng(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
template <int MLP, int LANES_PER_CELL>
__global__ void k_gather(const float4 *vol, size_t ncell, int iters, float *out)
{
	uint64_t s = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) / LANES_PER_CELL * 2654435761ull + 12345;
	const int k = threadIdx.x % LANES_PER_CELL;
	float acc = 0.f;
	for (int it = 0; it < iters; it++)
	{
		float4 v[MLP * (4 / LANES_PER_CELL)];
#pragma unroll
		for (int m = 0; m < MLP; m++)
		{
			size_t c = rng(s) % ncell;
#pragma unroll
			for (int j = 0; j < 4 / LANES_PER_CELL; j++) v[m * (4 / LANES_PER_CELL) + j] = __ldg(vol + 4 * c + k * (4 / LANES_PER_CELL) + j);
		}
#pragma unroll
		for (int m = 0; m < MLP * (4 / LANES_PER_CELL); m++) acc += v[m].x + v[m].w;
	}
	if (acc == 123.456f) out[0] = acc;
}
template <int MLP>
__global__ void k_red(float4 *vol, size_t ncell, int iters)
{
	uint64_t s = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) * 2654435761ull + 999;
	for (int it = 0; it < iters; it++)
	{
#pragma unroll
		for (int m = 0; m < MLP; m++)
		{
			size_t c = rng(s) % ncell;   // each "pixel": pair of adjacent 16-byte voxels (32 B), 4 such pairs
#pragma unroll
			for (int j = 0; j < 2; j++)
				asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(vol + 2 * c + j), "f"(1.f), "f"(1.f), "f"(1.f), "f"(0.f) : "memory");
		}
	}
}
template <typename F> float timeit(F f)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	f(); cudaDeviceSynchronize();
	cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main(int argc, char **argv)
{
	size_t ncell = (size_t) (argc > 1 ? atol(argv[1]) : 68000000);
	float4 *vol; cudaMalloc(&vol, ncell * 64); cudaMemset(vol, 0, ncell * 64);
	float *out; cudaMalloc(&out, 4);
	int sms = 148;
	for (int bps : {2, 4, 8})
	{
		int iters = 64;
		{
			float ms = timeit([&] { k_gather<1, 1><<<sms * bps, 256>>>(vol, ncell, iters * 4, out); });
			double bytes = (double) sms * bps * 256 * iters * 4 * 64;
			printf("gather64 lane-per-cell MLP1 blocks/SM=%d: %.3f ms  %.0f GB/s\n", bps, ms, bytes / ms / 1e6);
		}
		{
			float ms = timeit([&] { k_gather<4, 1><<<sms * bps, 256>>>(vol, ncell, iters, out); });
			double bytes = (double) sms * bps * 256 * iters * 4 * 64;
			printf("gather64 lane-per-cell MLP4 blocks/SM=%d: %.3f ms  %.0f GB/s\n", bps, ms, bytes / ms / 1e6);
		}
		{
			float ms = timeit([&] { k_gather<4, 4><<<sms * bps, 256>>>(vol, ncell, iters, out); });
			double bytes = (double) sms * bps * 64 * iters * 4 * 64;
			printf("gather64 quad-per-cell MLP4 blocks/SM=%d: %.3f ms  %.0f GB/s\n", bps, ms, bytes / ms / 1e6);
		}
		{
			float ms = timeit([&] { k_gather<8, 4><<<sms * bps, 256>>>(vol, ncell, iters, out); });
			double bytes = (double) sms * bps * 64 * iters * 8 * 64;
			printf("gather64 quad-per-cell MLP8 blocks/SM=%d: %.3f ms  %.0f GB/s\n", bps, ms, bytes / ms / 1e6);
		}
		{
			size_t npair = ncell * 2;          // 32-byte pairs over the same 4.35 GB
			float ms = timeit([&] { k_red<4><<<sms * bps, 256>>>(vol, npair, iters); });
			double bytes = (double) sms * bps * 256 * iters * 4 * 32 * 2;   // read + write of each 32 B pair
			printf("red.v4 pairs MLP4 blocks/SM=%d: %.3f ms  %.0f GB/s (RMW bytes)\n", bps, ms, bytes / ms / 1e6);
		}
	}
	return 0;
}
