// Can a thin spherical shell of the expanded reference (64-byte cells, 515 x 515 x 258) be streamed into L2 ahead of the
// band-major gathers, and do the gathers then hit?  Times, per shell radius: the random 64-byte gathers of the shell cold
// (L2 flushed) and warm (same shell again), and each way of bringing the shell in first:
//   pf64    prefetch.global.L2 on every cell of the (z, y) row runs the shell cuts out
//   bulk    one cp.async.bulk.prefetch.L2 per row run
//   touch   two 16-byte ld.global.cg per cell (both sectors), results discarded
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/shell_prefetch_bench tools/shell_prefetch_bench.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
#include <numeric>

static const int X = 258, Y = 515, Z = 515, IY = -257, IZ = -257;

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ float u01(uint32_t h) { return (h >> 8) * (1.f / 16777216.f); }

// cells of random points of the shell [R, R + dR), precomputed so that the timed kernel is nothing but the gathers
__global__ void __launch_bounds__(256) k_cells(float R, float dR, int n, uint32_t *cells, int layout, uint32_t ncompact, const uint32_t *table = nullptr, int lb = 0)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t s = hash32(i * 2654435761u + 777u);
	const float cz = 2.f * u01(s) - 1.f; s = hash32(s + 1);
	const float ph = 6.2831853f * u01(s); s = hash32(s + 1);
	const float r = R + dR * u01(s), st = sqrtf(fmaxf(0.f, 1.f - cz * cz));
	float x = r * st * cosf(ph), y = r * st * sinf(ph), z = r * cz;
	if (x < 0.f) { x = -x; y = -y; z = -z; }
	const int xi = (int) floorf(x), yi = (int) floorf(y) - IY, zi = (int) floorf(z) - IZ;
	uint32_t c = (uint32_t) (((size_t) zi * Y + yi) * X + xi);
	if (layout == 1) c = hash32(c) % ncompact;                                 // same reuse pattern, contiguous region
	if (layout == 2)                                                           // 32 x 32 x 32-cell blocks (2 MB = one page)
	{
		const uint32_t bx = xi >> 5, by = yi >> 5, bz = zi >> 5, nbx = (X + 31) >> 5, nby = (Y + 31) >> 5;
		c = (((bz * nby + by) * nbx + bx) << 15) + ((zi & 31) << 10) + ((yi & 31) << 5) + (xi & 31);
	}
	if (layout == 3)                                                           // radius-sorted blocks of (1 << lb)^3 cells, rank from a table
	{
		const int m = (1 << lb) - 1, nbx = (X + m) >> lb, nby = (Y + m) >> lb;
		const uint32_t rank = table[(((zi >> lb) * nby) + (yi >> lb)) * nbx + (xi >> lb)];
		c = (rank << (3 * lb)) + ((zi & m) << (2 * lb)) + ((yi & m) << lb) + (xi & m);
	}
	if (layout == 5)                                                           // accumulator: radius-sorted 4^3 blocks stored with their +1 halo (5^3 -> 128 voxels)
	{
		const int nbx = (X + 3) >> 2, nby = (Y + 3) >> 2;
		const uint32_t rank = table[(((zi >> 2) * nby) + (yi >> 2)) * nbx + (xi >> 2)];
		c = (rank << 7) + (zi & 3) * 25 + (yi & 3) * 5 + (xi & 3);
	}
	if (layout == 4) c = (hash32(c) % ncompact) << 15;                         // one 64-byte gather per 2 MB page, ncompact pages: TLB reach
	cells[i] = c;
}

// quad-cooperative 64-byte gathers, four in flight per lane
__global__ void __launch_bounds__(256) k_gather(const float4 *vol, const uint32_t *cells, int n, float *out)
{
	const uint32_t k = threadIdx.x & 3;
	const int nq = (gridDim.x * blockDim.x) >> 2;
	float acc = 0.f;
	for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; i + 3 * nq < n; i += 4 * nq)
	{
		float4 v[4];
#pragma unroll
		for (int u = 0; u < 4; u++) v[u] = __ldcg(vol + (size_t) __ldg(cells + i + u * nq) * 4 + k);
		acc += v[0].x + v[1].w + v[2].y + v[3].z;
	}
	if (acc == 123.456f) out[0] = acc;
}

// the store stage's scatter: two lanes per sample, each four red.global.add.v4.f32 (16-byte voxels) at +0, +sy, +sz, +sy+sz
__global__ void __launch_bounds__(256) k_scatter(float4 *acc, const uint32_t *cells, int n, uint32_t sy, uint32_t sz)
{
	const uint32_t px = threadIdx.x & 1;
	const int np = (gridDim.x * blockDim.x) >> 1;
	for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 1; i < n; i += np)
	{
		float4 *b = acc + (size_t) __ldg(cells + i) + px;
		asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(b), "f"(1.f) : "memory");
		asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(b + sy), "f"(1.f) : "memory");
		asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(b + sz), "f"(1.f) : "memory");
		asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(b + sy + sz), "f"(1.f) : "memory");
	}
}

// MODE 0: prefetch.global.L2 per cell, 1: bulk prefetch per run, 2: loads
template <int MODE>
__global__ void __launch_bounds__(256) k_shell(const char *vol, float Rlo, float Rhi, float *out)
{
	const int Rh = (int) ceilf(Rhi), side = 2 * Rh + 1, nrows = side * side;
	const float Rlo2 = Rlo * Rlo, Rhi2 = Rhi * Rhi;
	float acc = 0.f;
	for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += gridDim.x * blockDim.x)
	{
		const int zz = row / side - Rh, yy = row % side - Rh;
		const float rho2 = (float) (zz * zz + yy * yy);
		if (rho2 > Rhi2) continue;
		const int yi = yy - IY, zi = zz - IZ;
		if (yi < 0 || yi >= Y || zi < 0 || zi >= Z) continue;
		int x_hi = (int) sqrtf(Rhi2 - rho2) + 1;
		int x_lo = rho2 < Rlo2 ? (int) sqrtf(Rlo2 - rho2) - 1 : 0;
		x_lo = max(x_lo, 0); x_hi = min(x_hi, X - 1);
		const char *p = vol + (((size_t) zi * Y + yi) * X + x_lo) * 64;
		const int ncell = x_hi - x_lo + 1;
		if (MODE == 0) for (int c = 0; c < ncell; c++) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 64 * c));
		if (MODE == 1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(ncell * 64) : "memory");
		if (MODE == 2) for (int c = 0; c < ncell; c++) { const float4 a = __ldcg((const float4 *) (p + 64 * c)), b = __ldcg((const float4 *) (p + 64 * c + 32)); acc += a.x + b.x; }
	}
	if (acc == 123.456f) out[0] = acc;
}

static float *g_flush; static const size_t FLUSH = (size_t) 512 << 20;
__global__ void k_flush(float4 *b, size_t n) { for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) { float4 v = b[i]; v.x += 1.f; b[i] = v; } }
static void flush_l2() { k_flush<<<1184, 256>>>((float4 *) g_flush, FLUSH / 16); }

template <class F> static float timed(F f)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	cudaDeviceSynchronize();
	cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b);
	cudaEventDestroy(a); cudaEventDestroy(b);
	return ms;
}

int main()
{
	const size_t bytes = (size_t) 9 * 17 * 17 * 32768 * 64;   // room for the 32^3-blocked layout (5.5 GB)
	float4 *vol; float *out;
	cudaMalloc(&vol, bytes); cudaMalloc(&out, 16); cudaMalloc(&g_flush, FLUSH);
	cudaMemset(vol, 1, bytes);
	int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
	const int sms = p.multiProcessorCount;
	// TLB reach: one random 64-byte gather per page, N pages of 2 MB
	for (uint32_t npages : {16u, 32u, 64u, 96u, 128u, 192u, 256u, 512u, 1024u, 2048u})
	{
		const int n = 4 * sms * 8 * 64 * 8;
		uint32_t *c2; cudaMalloc(&c2, (size_t) n * 4);
		k_cells<<<(n + 255) / 256, 256>>>(200.f, 0.7f, n, c2, 4, npages);
		auto g2 = [&] { k_gather<<<sms * 8, 256>>>((const float4 *) vol, c2, n, out); };
		timed(g2); const float w = timed(g2);
		printf("pages %5u (one 64-byte cell in each): %.1f G/s\n", npages, n / w * 1e-6);
		cudaFree(c2);
	}
	for (float dR : {0.7f})
	for (float R : {60.f, 100.f, 160.f, 220.f, 250.f})
	{
		// cells the shell can touch: thickness dR + 2 (trilinear cell origin floor and the row-run padding)
		const double cells = 2. * M_PI * R * R * (dR + 1.);
		const int ngather = (int) (cells * 20.);
		const int grid = sms * 8, quads = grid * 64;
		const int n = (ngather + 4 * quads - 1) / (4 * quads) * (4 * quads);
		const double ng = (double) n;
		uint32_t *cells_d; cudaMalloc(&cells_d, (size_t) n * 4);
		k_cells<<<(n + 255) / 256, 256>>>(R, dR, n, cells_d, 0, 0);
		for (int lb = 2; lb <= 2; lb++)
		{
			const int m = (1 << lb) - 1, nbx = (X + m) >> lb, nby = (Y + m) >> lb, nbz = (Z + m) >> lb;
			const size_t nb = (size_t) nbx * nby * nbz;
			std::vector<float> rad(nb);
			for (int bz = 0; bz < nbz; bz++) for (int by = 0; by < nby; by++) for (int bx = 0; bx < nbx; bx++)
			{
				const float h = 0.5f * (1 << lb), cx = (bx << lb) + h, cy = (by << lb) + h + IY, cz = (bz << lb) + h + IZ;
				rad[((size_t) bz * nby + by) * nbx + bx] = sqrtf(cx * cx + cy * cy + cz * cz);
			}
			std::vector<uint32_t> order(nb), rank(nb);
			std::iota(order.begin(), order.end(), 0u);
			std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return rad[a] < rad[b]; });
			for (size_t i = 0; i < nb; i++) rank[order[i]] = (uint32_t) i;
			uint32_t *tab_d, *c2; cudaMalloc(&tab_d, nb * 4); cudaMalloc(&c2, (size_t) n * 4);
			cudaMemcpy(tab_d, rank.data(), nb * 4, cudaMemcpyHostToDevice);
			k_cells<<<(n + 255) / 256, 256>>>(R, dR, n, c2, 3, 0, tab_d, lb);
			auto g2 = [&] { k_gather<<<grid, 256>>>((const float4 *) vol, c2, n, out); };
			flush_l2(); const float c = timed(g2); const float w = timed(g2);
			printf("R=%5.0f dR=%.1f  layout radius-sorted %d^3 blocks: cold %.3f ms (%.1f G/s)  warm %.3f ms (%.1f G/s)\n", R, dR, 1 << lb, c, ng / c * 1e-6, w, ng / w * 1e-6);
			if (lb == 2)
			{
				// scatter: canonical [z][y][x] accumulator against padded radius-sorted blocks
				uint32_t *c5; cudaMalloc(&c5, (size_t) n * 4);
				k_cells<<<(n + 255) / 256, 256>>>(R, dR, n, c5, 5, 0, tab_d, 2);
				auto s0 = [&] { k_scatter<<<grid, 256>>>((float4 *) vol, cells_d, n, X, X * Y); };
				auto s5 = [&] { k_scatter<<<grid, 256>>>((float4 *) vol, c5, n, 5, 25); };
				flush_l2(); const float a0 = timed(s0); const float a1 = timed(s0);
				flush_l2(); const float b0 = timed(s5); const float b1 = timed(s5);
				printf("R=%5.0f dR=%.1f  scatter (8 x red.v4 per sample): [z][y][x] cold %.3f ms warm %.3f ms (%.1f G samples/s) | padded sorted blocks cold %.3f ms warm %.3f ms (%.1f G samples/s)\n",
				       R, dR, a0, a1, ng / a1 * 1e-6, b0, b1, ng / b1 * 1e-6);
				cudaFree(c5);
			}
			cudaFree(c2); cudaFree(tab_d);
		}
		for (int layout = 1; layout <= 2; layout++)
		{
			uint32_t *c2; cudaMalloc(&c2, (size_t) n * 4);
			k_cells<<<(n + 255) / 256, 256>>>(R, dR, n, c2, layout, (uint32_t) cells);
			auto g2 = [&] { k_gather<<<grid, 256>>>((const float4 *) vol, c2, n, out); };
			flush_l2(); const float c = timed(g2); const float w = timed(g2);
			printf("R=%5.0f dR=%.1f  layout %s: cold %.3f ms (%.1f G/s)  warm %.3f ms (%.1f G/s)\n", R, dR, layout == 1 ? "compact" : "blocked32", c, ng / c * 1e-6, w, ng / w * 1e-6);
			cudaFree(c2);
		}
		auto gather = [&] { k_gather<<<grid, 256>>>((const float4 *) vol, cells_d, n, out); };
		const float Rlo = R - 1.f, Rhi = R + dR + 1.f;
		flush_l2(); const float cold = timed(gather);
		const float warm = timed(gather);
		printf("R=%5.0f dR=%.1f  shell ~%5.1f MB  gathers %.2fM | cold %.3f ms (%.1f G/s)  warm %.3f ms (%.1f G/s)\n", R, dR, cells * 64e-6, ng * 1e-6,
		       cold, ng / cold * 1e-6, warm, ng / warm * 1e-6);
		const char *names[3] = {"pf64", "bulk", "touch"};
		for (int mode = 1; mode < 2; mode++)
		{
			auto pre = [&] {
				if (mode == 0) k_shell<0><<<sms * 4, 256>>>((const char *) vol, Rlo, Rhi, out);
				if (mode == 1) k_shell<1><<<sms * 4, 256>>>((const char *) vol, Rlo, Rhi, out);
				if (mode == 2) k_shell<2><<<sms * 4, 256>>>((const char *) vol, Rlo, Rhi, out);
			};
			flush_l2(); const float tp = timed(pre);
			const float tg = timed(gather);
			flush_l2(); const float both = timed([&] { pre(); gather(); });
			printf("    %-6s prefetch %.3f ms, gathers after it %.3f ms (%.1f G/s), back to back %.3f ms\n", names[mode], tp, tg, ng / tg * 1e-6, both);
		}
		cudaFree(cells_d);
	}
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
