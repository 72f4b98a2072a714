// Microbenchmark: does a blocked layout of the neighbourhood-expanded reference (64-byte cells) raise the HBM gather
// rate of the fine pass?  Quads walk the rows of random central slices through a 515 x 515 x 258 volume exactly like
// k_diff2_fine (one 64-byte cell per pixel, four lanes x 16 B), only the cell -> address map changes:
//   layout 0: linear  [z][y][x]
//   layout 1: blocks of 4x4x4 cells (4 KB), cells linear inside a block
//   layout 2: blocks of 8x8x8 cells (32 KB)
//   layout 3: blocks of 2x2x2 cells (512 B)
//   layout 4: blocks of 4x4x4 cells, Morton order inside
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o slice_gather_bench slice_gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

static const int VX = 258, VY = 515, VZ = 515, INIT = -257;

template <int LAYOUT>
__device__ __forceinline__ size_t cell_index(int x, int y, int z)
{
	if (LAYOUT == 0) return ((size_t) z * VY + y) * VX + x;
	constexpr int B = LAYOUT == 1 ? 4 : LAYOUT == 2 ? 8 : LAYOUT == 3 ? 2 : 4;
	constexpr int NBX = (VX + B - 1) / B, NBY = (VY + B - 1) / B;
	const size_t blk = ((size_t) (z / B) * NBY + (y / B)) * NBX + (x / B);
	int in;
	if (LAYOUT == 4)
	{
		const int a = x % B, b = y % B, c = z % B;
		in = (a & 1) | ((b & 1) << 1) | ((c & 1) << 2) | ((a & 2) << 2) | ((b & 2) << 3) | ((c & 2) << 4);
	}
	else in = ((z % B) * B + (y % B)) * B + (x % B);
	return blk * (B * B * B) + in;
}

template <int LAYOUT>
__global__ void __launch_bounds__(256, 3) k_walk(const float4 *vol, const float *eulers, int nslice, int n, float *out)
{
	const int k = threadIdx.x & 3, qd = threadIdx.x >> 2;
	const int half = n / 2;
	float acc = 0.f;
	for (int w = blockIdx.x; w < nslice; w += gridDim.x)
	{
		const float *e = eulers + 9 * (size_t) w;
		const float e0 = e[0], e1 = e[1], e3 = e[3], e4 = e[4], e6 = e[6], e7 = e[7];
		for (int r = qd; r < n - 1; r += 64)
		{
			const int y = r <= half ? r : r - n;
			const int xhi = (int) sqrtf((float) (half * half - y * y));
			float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
			for (int x = 0; x <= xhi; x++)
			{
				float xp = (e0 * x + e1 * y) * 2.f, yp = (e3 * x + e4 * y) * 2.f, zp = (e6 * x + e7 * y) * 2.f;
				if (xp < 0.f) { xp = -xp; yp = -yp; zp = -zp; }
				const int x0 = (int) floorf(xp), y0 = (int) floorf(yp) - INIT, z0 = (int) floorf(zp) - INIT;
				const float4 nxt = __ldg(vol + 4 * cell_index<LAYOUT>(x0, y0, z0) + k);
				acc += cur.x + cur.w;
				cur = nxt;
			}
			acc += cur.x + cur.w;
		}
	}
	if (acc == 123.456f) out[0] = acc;
}

int main(int argc, char **argv)
{
	const int n = 256, nslice = argc > 1 ? atoi(argv[1]) : 6000;
	const size_t ncell_alloc = (size_t) 520 * 520 * 264;   // room for the blocked variants
	float4 *vol; cudaMalloc(&vol, ncell_alloc * 64); cudaMemset(vol, 0, ncell_alloc * 64);
	float *out; cudaMalloc(&out, 4);
	std::vector<float> eul((size_t) nslice * 9);
	srand(7);
	for (int i = 0; i < nslice; i++)
	{
		// random rotation from a random unit quaternion
		double q[4], s = 0;
		for (int j = 0; j < 4; j++) { q[j] = rand() / (double) RAND_MAX - 0.5; s += q[j] * q[j]; }
		s = sqrt(s); for (int j = 0; j < 4; j++) q[j] /= s;
		const double a = q[0], b = q[1], c = q[2], d = q[3];
		float *e = &eul[(size_t) i * 9];
		e[0] = a * a + b * b - c * c - d * d; e[1] = 2 * (b * c - a * d); e[2] = 2 * (b * d + a * c);
		e[3] = 2 * (b * c + a * d); e[4] = a * a - b * b + c * c - d * d; e[5] = 2 * (c * d - a * b);
		e[6] = 2 * (b * d - a * c); e[7] = 2 * (c * d + a * b); e[8] = a * a - b * b - c * c + d * d;
	}
	float *d_eul; cudaMalloc(&d_eul, eul.size() * 4); cudaMemcpy(d_eul, eul.data(), eul.size() * 4, cudaMemcpyHostToDevice);
	// pixels per slice (same rule as the kernel)
	double npix = 0;
	for (int r = 0; r < n - 1; r++) { int y = r <= n / 2 ? r : r - n; npix += (int) sqrtf((float) (n / 2 * n / 2 - y * y)) + 1; }
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	for (int rep = 0; rep < 2; rep++)
		for (int layout = 0; layout < 5; layout++)
		{
			cudaEventRecord(a);
			switch (layout)
			{
			case 0: k_walk<0><<<148 * 3, 256>>>(vol, d_eul, nslice, n, out); break;
			case 1: k_walk<1><<<148 * 3, 256>>>(vol, d_eul, nslice, n, out); break;
			case 2: k_walk<2><<<148 * 3, 256>>>(vol, d_eul, nslice, n, out); break;
			case 3: k_walk<3><<<148 * 3, 256>>>(vol, d_eul, nslice, n, out); break;
			case 4: k_walk<4><<<148 * 3, 256>>>(vol, d_eul, nslice, n, out); break;
			}
			cudaEventRecord(b); cudaEventSynchronize(b);
			float ms; cudaEventElapsedTime(&ms, a, b);
			if (rep) printf("layout %d: %.3f ms  %.0f GB/s (64 B per pixel, %d slices x %.0f pixels)\n", layout, ms, nslice * npix * 64 / ms / 1e6, nslice, npix);
		}
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
