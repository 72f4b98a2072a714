// relion_b200 — coarse-pass cross term on the 5th-generation tensor cores (sm_100a: tcgen05 + TMEM + TMA).
//
// In GLOBAL searches every particle of a pool is compared with the same orientation grid, so the cross term of the
// coarse squared difference (cuda_kernel_diff2_coarse, /root/reference/src/acc/cuda/cuda_kernels/diff2.cuh:24-189)
//
//     diff2[p][o][t] = sum_pix c_p (|A_o|^2 + |X'_p|^2)  -  2 Re sum_pix conj(A_o) . (c_p X'_p e^{i phi_t})
//
// is a dense real contraction  D[o][(p,t)] = sum_k A[o][k] B[(p,t)][k]  with
//     A[o][2i], A[o][2i+1]   = Re, Im of the projected reference at valid pixel i            (M = orientations)
//     B[n][2i], B[n][2i+1]   = Re, Im of c_p X'_p e^{i phi_t} at valid pixel i, n = p*T + t   (N = particles x translations)
// and the norm term is a second, T times smaller contraction  base[o][p] = sum_i |A_o(i)|^2 c_p(i).
// (CTF, scale and sigma2 weights are real per pixel, so they fold into the particle operand.)
//
// FP32-equivalent accuracy on tensor cores: every operand is split into a TF32 "hi" part and a TF32 "lo" remainder
// (v = hi + lo to ~22 bits) and three MMAs accumulate hi*hi + hi*lo + lo*hi into fp32 TMEM accumulators
// ("3xTF32"); the dropped lo*lo term is ~2^-22 relative.  tests/test_gpu_parity.py::test_gemm_tf32x3 holds the kernel
// to 2e-6 of sum|a||b| against float64.
//
// Kernel: one CTA per 128 x 256 output tile.  Warp 0 = TMA producer (cp.async.bulk.tensor, 128-byte swizzle, mbarrier
// complete_tx), warp 1 = TMEM allocator + single-thread tcgen05.mma issuer (kind::tf32, M128 N256 K8, accumulators in
// 2 x 256 TMEM columns: main product and the two 3xTF32 correction products), warps 2-5 = epilogue (tcgen05.ld
// 32x32b.x32 -> registers -> diff2 -> Mweight).  Two smem stages of 96 KB (A_hi, A_lo 16 KB each; B_hi, B_lo 32 KB each)
// per K-block of 32.
#include "img_src.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>

// K-block 32 (128-byte swizzle, 2 stages of 96 KB) or 16 (64-byte swizzle, 4 stages of 48 KB): same bytes in flight, the
// finer one hides the TMA latency behind a deeper ring.  Operands are padded in K to GM_BK = 32 either way.
static const int GM_BM = 128, GM_BN = 256, GM_BK = 32;
static const int GM_MPAD = 256;      // operand rows are padded to a CTA pair's 256 (two 128-row tiles)
static const int GM_THREADS = 192;
// TWO: a pair of CTAs on the two SMs of a TPC (cluster 2x1) computes a 256 x 256 tile with tcgen05.mma.cta_group::2 -
// each CTA stages its own 128 rows of A and only HALF of the B tile, so the L2 -> shared-memory traffic per flop drops by
// a third (the contraction is bound by that traffic, not by the tensor pipe, once the correction products run in bf16).
template <int BK, bool TWO = false> struct GemmCfg {
	static const int B_ROWS = TWO ? GM_BN / 2 : GM_BN;
	static const int STAGES = (BK == 32 ? 2 : 4) * (TWO ? 3 : 2) / 2;
	static const uint32_t A_BYTES = GM_BM * BK * 4, B_BYTES = B_ROWS * BK * 4;
	static const uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
	static const size_t SMEM = (size_t) STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile(
		"{\n\t.reg .pred P1;\n\t"
		"WAIT_LOOP:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
		"@P1 bra DONE;\n\t"
		"bra WAIT_LOOP;\n\t"
		"DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
	             ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
		::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
		::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// ---- CTA-pair (cta_group::2) variants ----
__device__ __forceinline__ uint32_t cluster_ctarank()
{
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank)
{
	uint32_t r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
	return r;
}
// TMA load into this CTA's shared memory whose bytes are counted on the LEADER CTA's mbarrier (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1)
{
	asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
	             ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar)    // arrives on `bar` in both CTAs of the pair
{
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
	             ::"r"(bar), "h"((uint16_t) 3) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void tcgen05_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	if (BF16)
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"setp.ne.b32 p, %4, 0;\n\t"
			"tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
			::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
	else
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"setp.ne.b32 p, %4, 0;\n\t"
			"tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
			::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
	               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
	               "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
	               "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
	             : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle (what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B): rows of
// 128 bytes, 8-row groups 1024 bytes apart.  Descriptor fields (PTX ISA "tcgen05 shared memory descriptor"):
// start address >> 4 [0,14), leading byte offset >> 4 [16,30) (unused for swizzled K-major, 1), stride byte offset >> 4
// [32,46) = 1024 >> 4, version 1 at [46,48), layout type SWIZZLE_128B = 2 at [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr)
{
	uint64_t d = 0;
	d |= (uint64_t) ((smem_addr & 0x3FFFF) >> 4);
	d |= (uint64_t) 1 << 16;
	d |= (uint64_t) (1024 >> 4) << 32;
	d |= (uint64_t) 1 << 46;
	d |= (uint64_t) 2 << 61;
	return d;
}

// the same for rows of 64 bytes (CU_TENSOR_MAP_SWIZZLE_64B, layout type 4): 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr)
{
	uint64_t d = 0;
	d |= (uint64_t) ((smem_addr & 0x3FFFF) >> 4);
	d |= (uint64_t) 1 << 16;
	d |= (uint64_t) (512 >> 4) << 32;
	d |= (uint64_t) 1 << 46;
	d |= (uint64_t) 4 << 61;
	return d;
}

// instruction descriptor, kind::tf32: D fp32 (1 @ [4,6)), A/B TF32 (2 @ [7,10), [10,13)), both K-major, N >> 3 @ [17,23),
// M >> 4 @ [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
	return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

// kind::f16 with BF16 operands (1 @ [7,10), [10,13)), fp32 accumulator
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N)
{
	return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

__device__ __forceinline__ float tf32_round(float v)
{
	uint32_t r;
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
	return __uint_as_float(r);
}
__device__ __forceinline__ void tf32_split(float v, float &hi, float &lo)
{
	hi = tf32_round(v);
	lo = tf32_round(v - hi);
}

// mixed split: the main product stays TF32 x TF32; the two correction products (2^-11 of it) only need ~9 bits, so their
// operands are stored as bf16 (hi' = bf16(hi), lo' = bf16(v - hi)) and run at twice the TF32 rate.  `lo` is then two
// bf16 planes of `plane` elements each, [hi' | lo'], in the bytes the fp32 lo array would take.
__device__ __forceinline__ void store_split2(float *hi, float *lo, size_t plane, size_t idx, float a, float b)
{
	float2 h, l;
	tf32_split(a, h.x, l.x); tf32_split(b, h.y, l.y);
	*(float2 *) (hi + idx) = h;
	if (plane == 0) *(float2 *) (lo + idx) = l;
	else
	{
		__nv_bfloat16 *b16 = (__nv_bfloat16 *) lo;
		*(__nv_bfloat162 *) (b16 + idx) = __floats2bfloat162_rn(h.x, h.y);
		*(__nv_bfloat162 *) (b16 + plane + idx) = __floats2bfloat162_rn(a - h.x, b - h.y);
	}
}
__device__ __forceinline__ void store_split1(float *hi, float *lo, size_t plane, size_t idx, float a)
{
	float h, l;
	tf32_split(a, h, l);
	hi[idx] = h;
	if (plane == 0) lo[idx] = l;
	else
	{
		__nv_bfloat16 *b16 = (__nv_bfloat16 *) lo;
		b16[idx] = __float2bfloat16_rn(h);
		b16[plane + idx] = __float2bfloat16_rn(a - h);
	}
}

// ---------------------------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------------------------
struct GemmEpilogue {
	int mode;                 // 0: C[m*ldc + n] = D;  1: coarse diff2 into Mweight
	// mode 0
	float *C; int ldc;
	// mode 1
	const RbPartMeta *metas; RbPartState *states;
	const unsigned char *pdf_orient_zero;
	float *Mweight;
	const float *base; int ldbase;   // [o][p] norm term
	const float *x2;                 // [p] sum c |X'|^2
	int T, P, O, cls, o_first;       // o_first: first orientation of this M chunk
	int cc;                          // cross-correlation criterion: value = -cross / sqrt(norm term) (diff2.cuh:336-460), no minimum
	// common
	int M, N;                        // valid extent of this launch (rows of the chunk, columns)
	int exp;                         // timing experiments only (RB_GEMM_EXP): 1 no bf16 MMAs, 2 no tf32 MMAs, 3 grouped issue, 4 no MMAs
};

// MIXED (BK = 32 only): tmAlo / tmBlo are the bf16 hi' planes, tmAlb / tmBlb the bf16 lo' planes; the stage holds
// A_hi 16 KB (tf32, 128-byte rows) | A_hi' 8 KB | A_lo' 8 KB (bf16, 64-byte rows) | B_hi | B_hi' | B_lo' likewise.
// TWO: launched with cluster dims (2, 1, 1) on a grid (2 * n tiles, m pairs); rank 0 of the pair issues the MMAs.
template <int BK, bool MIXED, bool TWO>
__global__ void __launch_bounds__(GM_THREADS, 1)
k_gemm_tf32x3(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
              const __grid_constant__ CUtensorMap tmAlb,
              const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
              const __grid_constant__ CUtensorMap tmBlb,
              int num_kblocks, GemmEpilogue E)
{
	static_assert(!MIXED || BK == 32, "mixed split uses K-blocks of 32");
	typedef GemmCfg<BK, TWO> Cfg;
	constexpr int STAGES = Cfg::STAGES;
	constexpr uint32_t A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
	constexpr int MMA_M = TWO ? 2 * GM_BM : GM_BM;
	extern __shared__ uint8_t gm_smem_raw[];
	const uint32_t raw = smem_u32(gm_smem_raw);
	const uint32_t tiles = (raw + 1023u) & ~1023u;                       // 1024-byte aligned (128-byte swizzle atom)
	const uint32_t bars = tiles + STAGES * STAGE_BYTES;            // full[stages], empty[stages], tmem_full, tmem slot
	uint8_t *bars_generic = gm_smem_raw + (bars - raw);
	const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tmem_full = bars + 16 * STAGES, tmem_slot = tmem_full + 8;

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	// pairs: the two CTAs of a cluster must be neighbours in x (the driver refuses cta_group::2 kernels otherwise), so the
	// grid is (2 * n tiles, m pairs).  Either way x runs over the n tiles first: a wave of CTAs then shares all of B (which
	// stays in L2 from wave to wave) and streams a few rows of A, instead of streaming half of A per wave.
	const uint32_t rank = TWO ? cluster_ctarank() : 0;                   // pair: rank 1 stages its halves, rank 0 also issues
	const int n_tile = TWO ? blockIdx.x >> 1 : blockIdx.x, m_tile = TWO ? 2 * blockIdx.y + (int) rank : blockIdx.y;

	if (warp == 0 && lane == 0)
	{
		for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
		mbar_init(tmem_full, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1)
	{
		if (TWO)
		{
			asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
		}
		else
		{
			asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	if (TWO) cluster_sync_all(); else __syncthreads();                   // pair: the peer's barriers must exist before remote arrivals
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_base = *(volatile uint32_t *) (bars_generic + 16 * STAGES + 8);

	if (warp == 0)
	{
		if (lane == 0)
		{
			const int b_row = n_tile * GM_BN + (int) rank * Cfg::B_ROWS;
			for (int kb = 0; kb < num_kblocks; kb++)
			{
				const int s = kb % STAGES;
				const uint32_t ph = (kb / STAGES) & 1;
				mbar_wait(empty0 + 8 * s, ph ^ 1);                       // slot free (first round passes immediately)
				const uint32_t st = tiles + s * STAGE_BYTES;
				// pair: all bytes of both CTAs are counted on the leader's barrier, which alone is armed (for both halves)
				const uint32_t fb = TWO ? mapa_shared(full0 + 8 * s, 0) : full0 + 8 * s;
				if (!TWO) mbar_expect_tx(fb, STAGE_BYTES);
				else if (rank == 0) mbar_expect_tx(full0 + 8 * s, 2 * STAGE_BYTES);
				auto load = [&](uint32_t dst, const CUtensorMap *map, int row)
				{
					if (TWO) tma_load_2d_pair(dst, map, fb, kb * BK, row); else tma_load_2d(dst, map, fb, kb * BK, row);
				};
				load(st, &tmAhi, m_tile * GM_BM);
				if (MIXED)
				{
					load(st + A_BYTES, &tmAlo, m_tile * GM_BM);
					load(st + A_BYTES + A_BYTES / 2, &tmAlb, m_tile * GM_BM);
					load(st + 2 * A_BYTES, &tmBhi, b_row);
					load(st + 2 * A_BYTES + B_BYTES, &tmBlo, b_row);
					load(st + 2 * A_BYTES + B_BYTES + B_BYTES / 2, &tmBlb, b_row);
				}
				else
				{
					load(st + A_BYTES, &tmAlo, m_tile * GM_BM);
					load(st + 2 * A_BYTES, &tmBhi, b_row);
					load(st + 2 * A_BYTES + B_BYTES, &tmBlo, b_row);
				}
			}
		}
	}
	else if (warp == 1)
	{
		if (lane == 0 && rank == 0)
		{
			const uint32_t idesc = umma_idesc_tf32(MMA_M, GM_BN), idesc16 = umma_idesc_bf16(MMA_M, GM_BN);
			auto mma32 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc)
			{
				if (TWO) tcgen05_mma_pair<false>(d, a, b, idesc, acc); else tcgen05_mma_tf32(d, a, b, idesc, acc);
			};
			auto mma16 = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc)
			{
				if (TWO) tcgen05_mma_pair<true>(d, a, b, idesc16, acc); else tcgen05_mma_bf16(d, a, b, idesc16, acc);
			};
			for (int kb = 0; kb < num_kblocks; kb++)
			{
				const int s = kb % STAGES;
				const uint32_t ph = (kb / STAGES) & 1;
				mbar_wait(full0 + 8 * s, ph);                            // TMA bytes have landed (in both CTAs of a pair)
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t st = tiles + s * STAGE_BYTES;
				// the two correction products (2^-11 of the main one) go to their own accumulator: the tensor core truncates
				// every accumulate to fp32, and three times fewer accumulations into the large sum means three times
				// less truncation drift; the epilogue adds the two accumulators
				if (MIXED)
				{
					const uint64_t a32 = umma_desc_k_sw128(st), b32 = umma_desc_k_sw128(st + 2 * A_BYTES);
					const uint64_t ah = umma_desc_k_sw64(st + A_BYTES), al = umma_desc_k_sw64(st + A_BYTES + A_BYTES / 2);
					const uint64_t bh = umma_desc_k_sw64(st + 2 * A_BYTES + B_BYTES), bl = umma_desc_k_sw64(st + 2 * A_BYTES + B_BYTES + B_BYTES / 2);
					if (E.exp == 0)
					{
#pragma unroll
						for (int ks = 0; ks < 2; ks++)                   // 16 K values per step: one bf16 K16 = two tf32 K8
						{
							const uint64_t adv16 = (uint64_t) ((ks * 16 * 2) >> 4);
							mma16(tmem_base + GM_BN, al + adv16, bh + adv16, (kb | ks) != 0);
							mma16(tmem_base + GM_BN, ah + adv16, bl + adv16, 1);
							const uint64_t adv0 = (uint64_t) ((ks * 16 * 4) >> 4), adv1 = (uint64_t) (((ks * 16 + 8) * 4) >> 4);
							mma32(tmem_base, a32 + adv0, b32 + adv0, (kb | ks) != 0);
							mma32(tmem_base, a32 + adv1, b32 + adv1, 1);
						}
					}
					else
					{
						if (E.exp == 2 || E.exp == 3)
#pragma unroll
							for (int ks = 0; ks < 2; ks++)
							{
								const uint64_t adv16 = (uint64_t) ((ks * 16 * 2) >> 4);
								mma16(tmem_base + GM_BN, al + adv16, bh + adv16, (kb | ks) != 0);
								mma16(tmem_base + GM_BN, ah + adv16, bl + adv16, 1);
							}
						if (E.exp == 1 || E.exp == 3)
#pragma unroll
							for (int ks = 0; ks < 4; ks++)
							{
								const uint64_t adv0 = (uint64_t) ((ks * 8 * 4) >> 4);
								mma32(tmem_base, a32 + adv0, b32 + adv0, (kb | ks) != 0);
							}
					}
				}
				else
				{
					const uint64_t ahi = BK == 32 ? umma_desc_k_sw128(st) : umma_desc_k_sw64(st);
					const uint64_t alo = BK == 32 ? umma_desc_k_sw128(st + A_BYTES) : umma_desc_k_sw64(st + A_BYTES);
					const uint64_t bhi = BK == 32 ? umma_desc_k_sw128(st + 2 * A_BYTES) : umma_desc_k_sw64(st + 2 * A_BYTES);
					const uint64_t blo = BK == 32 ? umma_desc_k_sw128(st + 2 * A_BYTES + B_BYTES) : umma_desc_k_sw64(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
					for (int ks = 0; ks < BK / 8; ks++)
					{
						const uint64_t adv = (uint64_t) ((ks * 8 * 4) >> 4);  // 32 bytes per K step inside the swizzle atom
						mma32(tmem_base + GM_BN, alo + adv, bhi + adv, (kb | ks) != 0);
						mma32(tmem_base + GM_BN, ahi + adv, blo + adv, 1);
						mma32(tmem_base, ahi + adv, bhi + adv, (kb | ks) != 0);
					}
				}
				// frees the smem slot (in both CTAs) when these MMAs retire
				if (TWO) tcgen05_commit_pair(empty0 + 8 * s); else tcgen05_commit(empty0 + 8 * s);
			}
			if (TWO) tcgen05_commit_pair(tmem_full); else tcgen05_commit(tmem_full);   // accumulators complete
		}
	}
	else
	{
		// epilogue warps 2..5: TMEM lanes 32*(warp%4) .. +31 <-> rows of the tile
		const int q = warp & 3;
		const int row = m_tile * GM_BM + q * 32 + lane;
		mbar_wait(tmem_full, 0);
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		const uint32_t trow = tmem_base + ((uint32_t) (q * 32) << 16);
		const int n0 = n_tile * GM_BN;
		if (E.mode == 0)
		{
			for (int c = 0; c < GM_BN / 32; c++)
			{
				uint32_t v[32], v2[32];
				tmem_ld32(trow + c * 32, v);
				tmem_ld32(trow + GM_BN + c * 32, v2);
#pragma unroll
				for (int j = 0; j < 32; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
				if (row < E.M)
				{
#pragma unroll
					for (int j = 0; j < 32; j++)
					{
						const int n = n0 + c * 32 + j;
						if (n < E.N) E.C[(size_t) row * E.ldc + n] = __uint_as_float(v[j]);
					}
				}
			}
		}
		else
		{
			// column n = p*T + t.  All lanes of a warp walk the same columns, so particle boundaries are warp-uniform:
			// one warp-reduced atomicMin per (warp, particle).
			const int o = E.o_first + row;                               // orientation within the class
			const bool row_ok = row < E.M;
			int p = n0 / E.T, t = n0 - p * E.T;
			float bmin = FLT_MAX;
			// pvalid: the entry is computed; pwrite && !pvalid: orientation without prior (or class with pdf_class == 0), the
			// entry gets the lowest() that Mweight is initialised with in the reference (:3849) — the epilogue covers every
			// entry of the pool's Mweight, so no separate fill pass is needed on this path
			bool pvalid = false, pwrite = false; float bsum = 0.f, xi2 = 0.f; long long woff = 0;
			auto load_particle = [&](int pp)
			{
				pvalid = false; pwrite = false;
				if (pp < E.P && row_ok)
				{
					pwrite = true;
					const RbPartMeta m = E.metas[pp];
					const long long oc = (long long) E.cls * E.O + o;
					pvalid = !E.pdf_orient_zero[m.prior_off + oc];
					bsum = E.base[(size_t) row * E.ldbase + pp] + (E.cc ? 0.f : E.x2[pp]);
					if (E.cc) bsum = sqrtf(bsum);                                        // CC: the norm term alone, as its root
					xi2 = m.xi2_half;
					woff = m.coarse_off + oc * E.T;
				}
			};
			auto flush_min = [&](int pp)
			{
				const float wm = -warp_max(-bmin);
				if (lane == 0 && pp < E.P && wm < FLT_MAX && !E.cc) rb_atomic_min_pos(&E.states[pp].min_diff2_bits, wm);   // CC values are negative: k_weights_cc_coarse takes the minimum
				bmin = FLT_MAX;
			};
			load_particle(p);
			for (int c = 0; c < GM_BN / 32; c++)
			{
				uint32_t v[32], v2[32];
				tmem_ld32(trow + c * 32, v);
				tmem_ld32(trow + GM_BN + c * 32, v2);
#pragma unroll
				for (int j = 0; j < 32; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
				if (E.cc)
				{
#pragma unroll
					for (int j = 0; j < 32; j++)
					{
						if (pvalid) E.Mweight[woff + t] = -(__uint_as_float(v[j]) / bsum);         // diff2.h:729-735
						else if (pwrite) E.Mweight[woff + t] = RB_LOWEST;
						if (++t == E.T) { t = 0; p++; load_particle(p); }
					}
					continue;
				}
#pragma unroll
				for (int j = 0; j < 32; j++)
				{
					if (pvalid)
					{
						const float d = fmaxf(bsum - 2.f * __uint_as_float(v[j]), 0.f) + xi2;     // diff2.cuh:170-186, :1290-1296
						E.Mweight[woff + t] = d;
						bmin = fminf(bmin, d);
					}
					else if (pwrite) E.Mweight[woff + t] = RB_LOWEST;
					if (++t == E.T) { flush_min(p); t = 0; p++; load_particle(p); }
				}
			}
			flush_min(p);
		}
		asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	}
	if (TWO) cluster_sync_all(); else __syncthreads();                   // pair: neither CTA may leave while the other still reads it
	if (warp == 1)
	{
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		if (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
		else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
	}
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tmap(CUtensorMap *map, const void *base, size_t rows, size_t kpad, int box_rows, int bk = 32, bool bf16 = false)
{
	static PFN_encodeTiled fn = nullptr;
	if (!fn)
	{
		void *p = nullptr;
		cudaDriverEntryPointQueryResult qr;
		RB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
		if (!p || qr != cudaDriverEntryPointSuccess) { rb_set_error("cuTensorMapEncodeTiled not available from the driver"); return RB_ERR_CUDA; }
		fn = (PFN_encodeTiled) p;
	}
	const size_t esize = bf16 ? 2 : 4;
	const cuuint64_t dims[2] = {(cuuint64_t) kpad, (cuuint64_t) rows};
	const cuuint64_t strides[1] = {(cuuint64_t) kpad * esize};
	const cuuint32_t box[2] = {(cuuint32_t) bk, (cuuint32_t) box_rows};
	const cuuint32_t estr[2] = {1, 1};
	CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *) base, dims, strides, box, estr,
	                CU_TENSOR_MAP_INTERLEAVE_NONE, bk * esize == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
	                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { rb_set_error("cuTensorMapEncodeTiled failed (%d) rows=%zu kpad=%zu", (int) r, rows, kpad); return RB_ERR_CUDA; }
	return RB_OK;
}

static inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// RB_GEMM_MODE: bit 0 = bf16 operands for the two correction products (see store_split2), bit 1 = CTA pairs
// (cta_group::2, 256 x 256 tiles).  Default 3; 0 is the single-CTA 3xTF32 kernel (K-block from RB_GEMM_BK).
static int gemm_mode()
{
	static int m = -1;
	if (m < 0) { const char *e = getenv("RB_GEMM_MODE"); m = e ? atoi(e) & 3 : 3; }
	return m;
}
static bool gemm_mixed() { return gemm_mode() & 1; }

template <int BK, bool MIXED, bool TWO>
static int launch_gemm_variant(rb_ctx *ctx, dim3 grid, const CUtensorMap (&t)[6], int nkb, const GemmEpilogue &E)
{
	typedef GemmCfg<BK, TWO> Cfg;
	static bool configured_dev[RB_MAX_DEVICES] = {};
	bool &configured = configured_dev[ctx->device % RB_MAX_DEVICES];
	if (!configured)
	{
		RB_CUDA(cudaFuncSetAttribute(k_gemm_tf32x3<BK, MIXED, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Cfg::SMEM));
		configured = true;
	}
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid; cfg.blockDim = dim3(GM_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = ctx->stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = TWO ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	RB_CUDA(cudaLaunchKernelEx(&cfg, k_gemm_tf32x3<BK, MIXED, TWO>, t[0], t[1], t[2], t[3], t[4], t[5], nkb, E));
	RB_LAUNCH_CHECK(ctx);
	return RB_OK;
}

// A_hi/A_lo: [Mpad][Kpad], B_hi/B_lo: [Npad][Kpad] (Mpad % 256 == 0, Npad % 256 == 0, Kpad % 32 == 0, zero padded); with the
// mixed split the lo arrays hold the two bf16 planes instead
static int launch_gemm(rb_ctx *ctx, const float *Ahi, const float *Alo, size_t Mpad, const float *Bhi, const float *Blo, size_t Npad,
                       size_t Kpad, const GemmEpilogue &E_in)
{
	static int bk = 0;
	if (!bk) { const char *e = getenv("RB_GEMM_BK"); bk = (e && atoi(e) == 32) ? 32 : 16; }
	const int mode = gemm_mode();
	const bool mixed = mode & 1, two = mode & 2;
	const int kb = mode ? 32 : bk;
	const int b_rows = two ? GM_BN / 2 : GM_BN;
	if (Mpad % GM_MPAD || Npad % GM_BN || Kpad % GM_BK) { rb_set_error("launch_gemm: operands not padded"); return RB_ERR_ARG; }
	dim3 grid((unsigned) (Npad / GM_BN), (unsigned) (Mpad / GM_BM));
	if (two) grid = dim3((unsigned) (2 * (Npad / GM_BN)), (unsigned) (Mpad / GM_MPAD));
	if (grid.y > 65535) { rb_set_error("launch_gemm: too many tiles (%u) for one launch", grid.y); return RB_ERR_ARG; }
	CUtensorMap t[6];
	RB_CHECK(make_tmap(&t[0], Ahi, Mpad, Kpad, GM_BM, kb)); RB_CHECK(make_tmap(&t[3], Bhi, Npad, Kpad, b_rows, kb));
	if (mixed)
	{
		const __nv_bfloat16 *Ab = (const __nv_bfloat16 *) Alo, *Bb = (const __nv_bfloat16 *) Blo;
		RB_CHECK(make_tmap(&t[1], Ab, Mpad, Kpad, GM_BM, 32, true)); RB_CHECK(make_tmap(&t[2], Ab + Mpad * Kpad, Mpad, Kpad, GM_BM, 32, true));
		RB_CHECK(make_tmap(&t[4], Bb, Npad, Kpad, b_rows, 32, true)); RB_CHECK(make_tmap(&t[5], Bb + Npad * Kpad, Npad, Kpad, b_rows, 32, true));
	}
	else
	{
		RB_CHECK(make_tmap(&t[1], Alo, Mpad, Kpad, GM_BM, kb)); t[2] = t[1];
		RB_CHECK(make_tmap(&t[4], Blo, Npad, Kpad, b_rows, kb)); t[5] = t[4];
	}
	const int nkb = (int) (Kpad / kb);
	static int exp = -1;
	if (exp < 0) { const char *e = getenv("RB_GEMM_EXP"); exp = e ? atoi(e) : 0; }
	GemmEpilogue E = E_in;
	E.exp = exp;
	switch (mode)
	{
	case 3: return launch_gemm_variant<32, true, true>(ctx, grid, t, nkb, E);
	case 2: return launch_gemm_variant<32, false, true>(ctx, grid, t, nkb, E);
	case 1: return launch_gemm_variant<32, true, false>(ctx, grid, t, nkb, E);
	default: return kb == 32 ? launch_gemm_variant<32, false, false>(ctx, grid, t, nkb, E) : launch_gemm_variant<16, false, false>(ctx, grid, t, nkb, E);
	}
}

// ---------------------------------------------------------------------------------------------
// operand builders
// ---------------------------------------------------------------------------------------------
// generic split of a row-major [rows][cols] fp32 matrix into zero-padded hi/lo [rows_pad][kpad]
__global__ void k_split_pad(const float *src, int rows, int cols, float *hi, float *lo, size_t rows_pad, size_t kpad, size_t plane)
{
	const size_t n = rows_pad * kpad;
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
	{
		const size_t r = i / kpad, c = i - r * kpad;
		store_split1(hi, lo, plane, i, (r < (size_t) rows && c < (size_t) cols) ? src[r * cols + c] : 0.f);
	}
}

// A operands of one class for orientations [o_first, o_first + rows): projections at the coarse window
// (AccProjectorKernel::project3Dmodel, acc_projectorkernel_impl.h:161-231), interleaved (re, im) per valid pixel, and
// their squared moduli for the norm term.
// projs != nullptr: rows of ALL classes stacked, row r = class r / o_per_class, orientation r % o_per_class (pj is ignored).
__global__ void __launch_bounds__(256)
k_gemm_build_A(RbProjector pj, const float *coarse_eulers, const uint32_t *pix, int npix, int n, int o_first, int rows, int rows_pad,
               float *Ahi, float *Alo, size_t kpad, float *A2hi, float *A2lo, size_t k2pad, int mixed,
               const RbProjector *projs, int o_per_class)
{
	const int r = blockIdx.y;
	const size_t plane = mixed ? (size_t) rows_pad * kpad : 0, plane2 = mixed ? (size_t) rows_pad * k2pad : 0;
	const int imgX = n / 2 + 1;
	const bool live = r < rows;
	int o = o_first + r;
	if (projs && live) { pj = projs[r / o_per_class]; o = r % o_per_class; }
	const RbProjK pk = rb_make_projk2(pj, imgX);
	float e0 = 0, e1 = 0, e3 = 0, e4 = 0, e6 = 0, e7 = 0;
	if (live)
	{
		const float *eu = coarse_eulers + (size_t) o * 9;
		e0 = eu[0]; e1 = eu[1]; e3 = eu[3]; e4 = eu[4]; e6 = eu[6]; e7 = eu[7];
	}
	const int half = (int) (kpad / 2);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < half; i += gridDim.x * blockDim.x)
	{
		float2 ref = make_float2(0.f, 0.f);
		if (live && i < npix)
		{
			const uint32_t pkx = __ldg(pix + i);
			ref = rb_project3d_xp(pk, pj.mdl2, rb_pix_x(pkx), rb_pix_y(pkx), e0, e1, e3, e4, e6, e7);
		}
		store_split2(Ahi, Alo, plane, (size_t) r * kpad + 2 * i, ref.x, ref.y);
		if ((size_t) i < k2pad) store_split1(A2hi, A2lo, plane2, (size_t) r * k2pad + i, ref.x * ref.x + ref.y * ref.y);
	}
}

// B operands of the pool: column n = p*T + t holds c_p X'_p e^{i phi_t} over the valid pixels (translatePixel with the
// table factorisation of computeSincosLookupTable2D, cpu_kernels/helper.h:622-660); B2 row p holds c_p; x2[p] = sum c |X'|^2.
// tstride: rows per particle (T for the pool-wide contraction, 32 for the per-particle tiles of the fused local kernel);
// B2 / x2 outputs are optional (nullptr).
__global__ void __launch_bounds__(256)
k_gemm_build_B(const float4 *img4, const uint32_t *pix, int npix, int n, const float *tx, const float *ty, int T, int tstride, int P,
               float *Bhi, float *Blo, size_t kpad, float *B2hi, float *B2lo, size_t k2pad, float *x2, size_t plane, size_t plane2)
{
	__shared__ float red[32];
	const int col = blockIdx.x;                  // 0 .. Npad-1
	const int p = col / tstride, t = col - p * tstride;
	const bool live = p < P && t < T;
	const int imgX = n / 2 + 1;
	const float4 *img = img4 + (size_t) (live ? p : 0) * n * imgX;
	const float ttx = live ? tx[t] : 0.f, tty = live ? ty[t] : 0.f;
	const int half = (int) (kpad / 2);
	float acc = 0.f;
	for (int i = threadIdx.x; i < half; i += blockDim.x)
	{
		float2 y = make_float2(0.f, 0.f);
		float hc = 0.f;
		if (live && i < npix)
		{
			const uint32_t pkx = __ldg(pix + i);
			const int x = rb_pix_x(pkx), yy = rb_pix_y(pkx);
			const float4 im = __ldg(img + rb_src_index(x, yy, n));
			hc = im.z;
			float sx, cx, sy, cy;
			sincosf(x * ttx, &sx, &cx);
			sincosf((yy < 0 ? -yy : yy) * tty, &sy, &cy);
			if (yy < 0) sy = -sy;
			const float ss = sx * cy + cx * sy, cc = cx * cy - sx * sy;
			y.x = hc * (im.x * cc - im.y * ss);
			y.y = hc * (im.y * cc + im.x * ss);
			if (t == 0) acc += hc * (im.x * im.x + im.y * im.y);
		}
		store_split2(Bhi, Blo, plane, (size_t) col * kpad + 2 * i, y.x, y.y);
		if (B2hi && t == 0 && live && (size_t) i < k2pad) store_split1(B2hi, B2lo, plane2, (size_t) p * k2pad + i, hc);
	}
	if (t == 0 && live)
	{
		acc = block_sum(acc, red);
		if (threadIdx.x == 0) x2[p] = acc;
	}
}

// ---------------------------------------------------------------------------------------------
// pool driver: global-search coarse pass on the tensor cores
// ---------------------------------------------------------------------------------------------
bool rbk_coarse_gemm_applicable(rb_ctx *ctx, const PoolSlot &s)
{
	if (s.has_priors) return false;                       // local searches: per-particle orientation lists, no shared A
	const char *e = getenv("RB_COARSE_GEMM");
	const int mode = e ? atoi(e) : 1;                     // 0: never, 1: from 32 orientations, 2: always
	if (mode == 0) return false;
	// measured: already at 60 orientations (2D classification, one quarter-filled 128-row tile per class) the contraction
	// beats the SIMT kernel 7x (4.5 ms vs 33 ms for 2000 particles x 10 classes x 64 px)
	const long long O = (long long) ctx->d_samp.n_dir * ctx->d_samp.n_psi;
	return mode == 2 || O >= 32;
}

int rbk_diff2_coarse_gemm_pool(rb_ctx *ctx, PoolSlot &s, const float4 *cimg4)
{
	const RbModelDev &M = ctx->d_model;
	const RbSamplingDev &S = ctx->d_samp;
	const int P = s.P, T = S.n_trans, K = M.nr_classes;
	const int O = S.n_dir * S.n_psi;
	const int npix = M.d2_nvc, n = M.coarse_size;        // d2_*: with the CC criterion every pixel of the window's circle
	const uint32_t *pix = M.d2_pix_c;
	const size_t kpad = round_up((size_t) 2 * npix, GM_BK), k2pad = round_up((size_t) npix, GM_BK);
	const size_t Npad = round_up((size_t) P * T, GM_BN), N2pad = round_up((size_t) P, GM_BN);
	// orientation chunk: bounded operand memory (A and A2, hi + lo)
	const size_t budget = (size_t) 6 << 30;
	size_t mchunk = budget / ((kpad + k2pad) * 2 * sizeof(float));
	mchunk = std::max<size_t>(GM_MPAD, mchunk / GM_MPAD * GM_MPAD);
	mchunk = std::min<size_t>(mchunk, round_up((size_t) O, GM_MPAD));

	DevBuf &bAhi = ctx->gemm_buf[0], &bAlo = ctx->gemm_buf[1], &bA2hi = ctx->gemm_buf[2], &bA2lo = ctx->gemm_buf[3];
	DevBuf &bBhi = ctx->gemm_buf[4], &bBlo = ctx->gemm_buf[5], &bB2hi = ctx->gemm_buf[6], &bB2lo = ctx->gemm_buf[7];
	DevBuf &bBase = ctx->gemm_buf[8], &bX2 = ctx->gemm_buf[9];
	{
		// shared chunk buffers: only needed when the per-class cache below does not apply
		const size_t a_all = round_up((size_t) O, GM_MPAD) * (kpad + k2pad) * 2 * sizeof(float);
		const char *ce0 = getenv("RB_GEMM_CACHE_BYTES");
		const size_t budget0 = ce0 ? (size_t) strtoull(ce0, nullptr, 10) : ((size_t) 24 << 30);
		if (!(mchunk >= round_up((size_t) O, GM_MPAD) && a_all * (size_t) K <= budget0))
		{
			RB_CHECK(bAhi.ensure(mchunk * kpad * 4)); RB_CHECK(bAlo.ensure(mchunk * kpad * 4));
			RB_CHECK(bA2hi.ensure(mchunk * k2pad * 4)); RB_CHECK(bA2lo.ensure(mchunk * k2pad * 4));
		}
	}
	RB_CHECK(bBhi.ensure(Npad * kpad * 4)); RB_CHECK(bBlo.ensure(Npad * kpad * 4));
	RB_CHECK(bB2hi.ensure(N2pad * k2pad * 4)); RB_CHECK(bB2lo.ensure(N2pad * k2pad * 4));
	RB_CHECK(bBase.ensure(mchunk * N2pad * 4)); RB_CHECK(bX2.ensure(N2pad * 4));

	// particle operands (independent of the class)
	RB_CUDA(cudaMemsetAsync(bB2hi.p, 0, N2pad * k2pad * 4, ctx->stream));
	RB_CUDA(cudaMemsetAsync(bB2lo.p, 0, N2pad * k2pad * 4, ctx->stream));
	k_gemm_build_B<<<(unsigned) Npad, 256, 0, ctx->stream>>>(cimg4, pix, npix, n, S.ctx, S.cty, T, T, P,
		bBhi.as<float>(), bBlo.as<float>(), kpad, bB2hi.as<float>(), bB2lo.as<float>(), k2pad, bX2.as<float>(),
		gemm_mixed() ? Npad * kpad : 0, gemm_mixed() ? N2pad * k2pad : 0);
	RB_LAUNCH_CHECK(ctx);

	// The orientation operands only change with the reference, the sampling or the window sizes: when a class' whole grid
	// fits one chunk and the cache budget, they are built once per iteration and reused by every pool.
	const size_t a_bytes = round_up((size_t) O, GM_MPAD) * (kpad + k2pad) * 2 * sizeof(float);
	const char *ce = getenv("RB_GEMM_CACHE_BYTES");
	const size_t cache_budget = ce ? (size_t) strtoull(ce, nullptr, 10) : ((size_t) 24 << 30);
	const bool use_cache = mchunk >= round_up((size_t) O, GM_MPAD) && a_bytes * (size_t) K <= cache_budget;

	// Few orientations per class (2D classification: 60 in-plane rotations, K = 10 .. 200 classes): a class fills a fraction
	// of one 256-row tile.  Mweight's orientation index is class-major (iorientclass = class * O + o), so the classes' rows
	// simply stack along M: one contraction over K * O rows instead of K quarter-filled ones.
	const size_t stacked_pad = round_up((size_t) K * O, GM_MPAD);
	const char *se = getenv("RB_GEMM_STACK");                      // 0: one contraction per class (A/B)
	if ((!se || atoi(se) != 0) && K > 1 && use_cache && stacked_pad < (size_t) K * round_up((size_t) O, GM_MPAD))
	{
		const int rows = K * O, rows_pad = (int) stacked_pad;
		DevBuf *c = ctx->gemmA_all;
		RB_CHECK(c[0].ensure((size_t) rows_pad * kpad * 4)); RB_CHECK(c[1].ensure((size_t) rows_pad * kpad * 4));
		RB_CHECK(c[2].ensure((size_t) rows_pad * k2pad * 4)); RB_CHECK(c[3].ensure((size_t) rows_pad * k2pad * 4));
		RB_CHECK(bBase.ensure((size_t) rows_pad * N2pad * 4));
		long long stamp = (ctx->samp_version << 20) ^ ctx->model_version;
		for (int cls = 0; cls < K; cls++) stamp = stamp * 1000003LL + ctx->ref_version[cls];
		if (ctx->gemmA_all_stamp != stamp)
		{
			dim3 ga((unsigned) std::min<size_t>((kpad / 2 + 255) / 256, 64), (unsigned) rows_pad);
			k_gemm_build_A<<<ga, 256, 0, ctx->stream>>>(ctx->proj[0], S.coarse_eulers, pix, npix, n, 0, rows, rows_pad,
				c[0].as<float>(), c[1].as<float>(), kpad, c[2].as<float>(), c[3].as<float>(), k2pad, gemm_mixed() ? 1 : 0,
				ctx->d_proj.as<RbProjector>(), O);
			RB_LAUNCH_CHECK(ctx);
			ctx->gemmA_all_stamp = stamp;
		}
		GemmEpilogue E0;
		memset(&E0, 0, sizeof(E0));
		E0.mode = 0; E0.C = bBase.as<float>(); E0.ldc = (int) N2pad; E0.M = rows; E0.N = P;
		RB_CHECK(launch_gemm(ctx, c[2].as<float>(), c[3].as<float>(), rows_pad, bB2hi.as<float>(), bB2lo.as<float>(), N2pad, k2pad, E0));
		GemmEpilogue E1;
		memset(&E1, 0, sizeof(E1));
		E1.mode = 1; E1.metas = s.meta.as<RbPartMeta>(); E1.states = s.state.as<RbPartState>();
		E1.pdf_orient_zero = s.pdf_orient_zero.as<unsigned char>(); E1.Mweight = s.Mweight.as<float>();
		E1.base = bBase.as<float>(); E1.ldbase = (int) N2pad; E1.x2 = bX2.as<float>();
		E1.T = T; E1.P = P; E1.O = rows; E1.cls = 0; E1.o_first = 0; E1.M = rows; E1.N = P * T;   // row = iorientclass
		E1.cc = M.do_cc;
		RB_CHECK(launch_gemm(ctx, c[0].as<float>(), c[1].as<float>(), rows_pad, bBhi.as<float>(), bBlo.as<float>(), Npad, kpad, E1));
		return RB_OK;
	}
	for (int cls = 0; cls < K; cls++)
		for (int o0 = 0; o0 < O; o0 += (int) mchunk)
		{
			const int rows = std::min<int>((int) mchunk, O - o0);
			const int rows_pad = (int) round_up((size_t) rows, GM_MPAD);
			float *Ahi = bAhi.as<float>(), *Alo = bAlo.as<float>(), *A2hi = bA2hi.as<float>(), *A2lo = bA2lo.as<float>();
			bool build = true;
			if (use_cache)
			{
				DevBuf *c = ctx->gemmA[cls];
				RB_CHECK(c[0].ensure((size_t) rows_pad * kpad * 4)); RB_CHECK(c[1].ensure((size_t) rows_pad * kpad * 4));
				RB_CHECK(c[2].ensure((size_t) rows_pad * k2pad * 4)); RB_CHECK(c[3].ensure((size_t) rows_pad * k2pad * 4));
				Ahi = c[0].as<float>(); Alo = c[1].as<float>(); A2hi = c[2].as<float>(); A2lo = c[3].as<float>();
				const long long stamp = (ctx->ref_version[cls] << 40) ^ (ctx->samp_version << 20) ^ ctx->model_version;
				build = ctx->gemmA_stamp[cls] != stamp;
				ctx->gemmA_stamp[cls] = stamp;
			}
			if (build)
			{
				dim3 ga((unsigned) std::min<size_t>((kpad / 2 + 255) / 256, 64), (unsigned) rows_pad);
				k_gemm_build_A<<<ga, 256, 0, ctx->stream>>>(ctx->proj[cls], S.coarse_eulers, pix, npix, n, o0, rows, rows_pad,
					Ahi, Alo, kpad, A2hi, A2lo, k2pad, gemm_mixed() ? 1 : 0, nullptr, 0);
				RB_LAUNCH_CHECK(ctx);
			}
			// norm term base[o][p]
			GemmEpilogue E0;
			memset(&E0, 0, sizeof(E0));
			E0.mode = 0; E0.C = bBase.as<float>(); E0.ldc = (int) N2pad; E0.M = rows; E0.N = P;
			RB_CHECK(launch_gemm(ctx, A2hi, A2lo, rows_pad, bB2hi.as<float>(), bB2lo.as<float>(), N2pad, k2pad, E0));
			// cross term + diff2 epilogue
			GemmEpilogue E1;
			memset(&E1, 0, sizeof(E1));
			E1.mode = 1; E1.metas = s.meta.as<RbPartMeta>(); E1.states = s.state.as<RbPartState>();
			E1.pdf_orient_zero = s.pdf_orient_zero.as<unsigned char>(); E1.Mweight = s.Mweight.as<float>();
			E1.base = bBase.as<float>(); E1.ldbase = (int) N2pad; E1.x2 = bX2.as<float>();
			E1.T = T; E1.P = P; E1.O = O; E1.cls = cls; E1.o_first = o0; E1.M = rows; E1.N = P * T;
			E1.cc = M.do_cc;
			RB_CHECK(launch_gemm(ctx, Ahi, Alo, rows_pad, bBhi.as<float>(), bBlo.as<float>(), Npad, kpad, E1));
		}
	return RB_OK;
}


// ---------------------------------------------------------------------------------------------
// LOCAL searches: fused projection + contraction.  Every particle has its own (prior-selected) orientation list, so there
// is no pool-wide A operand; instead the CTA of (particle, 128-orientation tile) lets 8 producer warps project its
// orientations 16 pixels at a time straight into the swizzled shared-memory operand tiles (TF32 hi / lo), while one thread
// feeds the tensor core:  D[128 o][32 t] += A[128 o][32 k] . B_p[32 t][32 k]  (B_p = the particle's phase-shifted,
// weighted image for every translation, built once per pool and fetched by TMA).  The translation loop of the SIMT kernel
// (2 shared-memory loads + 10 FP32 instructions per pixel, translation and 3 orientations) disappears from the SM's
// issue slots; what is left is the gather.
// ---------------------------------------------------------------------------------------------
#ifndef RB_FU_STAGES
#define RB_FU_STAGES 2           // measured (256 px, hp4 local): see DESIGN.md section 7
#endif
static const int FU_BM = 128, FU_BN = 32, FU_PIX = 16, FU_STAGES = RB_FU_STAGES;
// producer warps per CTA: 8 (two CTAs per SM) or 16 (one CTA per SM, half the shared memory -> more L1 for the gathers)
static const uint32_t FU_A_BYTES = FU_BM * 128, FU_B_BYTES = FU_BN * 128;
static const uint32_t FU_STAGE_BYTES = 2 * FU_A_BYTES + 2 * FU_B_BYTES;            // 40 KB
static const size_t FU_SMEM = (size_t) FU_STAGES * FU_STAGE_BYTES + 1024 + 128;

struct FusedArgs {
	const RbPartMeta *metas; RbPartState *states;
	const int *dir_idx, *psi_idx;
	const unsigned char *pdf_orient_zero;
	float *Mweight;
	const float4 *img4;            // prepared images at the coarse window (for c = corr/2)
	const float *x2;               // [P] sum c |X'|^2
	const RbProjector *projs;
	const uint32_t *pix; int npix; int n;
	int T; int tiles_per_class; int num_kblocks;
	int Tstride, t0;               // Mweight row length and first translation of this pass (T <= 32 translations per pass)
	int cc;                        // cross-correlation criterion: value = -cross / sqrt(norm term), no xi2 term, no minimum (diff2.cuh:336-460)
	const int *order;              // [P] or nullptr: particle handled by CTA column blockIdx.x (PoolSlot::order)
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int NPW, int G256>
__global__ void __launch_bounds__(64 + 32 * NPW, NPW == 8 ? 2 : 1)
k_coarse_fused(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, FusedArgs A, RbSamplingDev S)
{
	constexpr int FU_PRODUCERS = 32 * NPW, FU_THREADS = 64 + FU_PRODUCERS, NJ = 64 / NPW;   // rows per producer thread
	extern __shared__ uint8_t fu_smem_raw[];
	__shared__ float s_e[FU_BM][6];
	__shared__ unsigned char s_valid[FU_BM];
	__shared__ float s_base[FU_BM];
	__shared__ float s_min[4];
	const uint32_t raw = smem_u32(fu_smem_raw);
	const uint32_t tiles = (raw + 1023u) & ~1023u;
	const uint32_t bars = tiles + FU_STAGES * FU_STAGE_BYTES;
	uint8_t *tiles_generic = fu_smem_raw + (tiles - raw);
	uint8_t *bars_generic = fu_smem_raw + (bars - raw);
	const uint32_t full0 = bars, empty0 = bars + 8 * FU_STAGES, tmem_full = bars + 16 * FU_STAGES, tmem_slot = tmem_full + 8;

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	// particle index fastest: the first tiles of all particles (always full) are dispatched before the partial last tiles,
	// which then fill the tail of the launch (tile-major order left ~10 % of the SM time idle at the end)
	const int p = A.order ? __ldg(A.order + blockIdx.x) : (int) blockIdx.x;
	const int cls = blockIdx.y / A.tiles_per_class;
	const int oi0 = (blockIdx.y - cls * A.tiles_per_class) * FU_BM;
	const RbPartMeta m = A.metas[p];
	const int no = m.nd * m.np;
	if (oi0 >= no) return;
	const int o0 = cls * no + oi0;

	// orientation table of the tile
	for (int r = threadIdx.x; r < FU_BM; r += FU_THREADS)
	{
		const int oi = oi0 + r;
		bool valid = oi < no;
		if (valid)
		{
			valid = !A.pdf_orient_zero[m.prior_off + o0 + r];
			const int idl = oi / m.np, ipl = oi - idl * m.np;
			const int gd = m.dir_off < 0 ? idl : A.dir_idx[m.dir_off + idl];
			const int gp = m.psi_off < 0 ? ipl : A.psi_idx[m.psi_off + ipl];
			const float *eu = S.coarse_eulers + ((size_t) gd * S.n_psi + gp) * 9;
			s_e[r][0] = eu[0]; s_e[r][1] = eu[1]; s_e[r][2] = eu[3]; s_e[r][3] = eu[4]; s_e[r][4] = eu[6]; s_e[r][5] = eu[7];
		}
		s_valid[r] = valid;
	}
	if (warp == 0 && lane == 0)
	{
		for (int s = 0; s < FU_STAGES; s++) { mbar_init(full0 + 8 * s, FU_PRODUCERS + 1); mbar_init(empty0 + 8 * s, 1); }
		mbar_init(tmem_full, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1)
	{
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(64) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_base = *(volatile uint32_t *) (bars_generic + 16 * FU_STAGES + 8);
	const int nkb = A.num_kblocks;

	if (warp == 0)
	{
		if (lane == 0)
		{
			for (int kb = 0; kb < nkb; kb++)
			{
				const int s = kb % FU_STAGES;
				const uint32_t ph = (kb / FU_STAGES) & 1;
				mbar_wait(empty0 + 8 * s, ph ^ 1);
				const uint32_t st = tiles + s * FU_STAGE_BYTES;
				mbar_expect_tx(full0 + 8 * s, 2 * FU_B_BYTES);
				tma_load_2d(st + 2 * FU_A_BYTES, &tmBhi, full0 + 8 * s, kb * 32, p * FU_BN);
				tma_load_2d(st + 2 * FU_A_BYTES + FU_B_BYTES, &tmBlo, full0 + 8 * s, kb * 32, p * FU_BN);
			}
		}
	}
	else if (warp == 1)
	{
		if (lane == 0)
		{
			const uint32_t idesc = umma_idesc_tf32(FU_BM, FU_BN);
			for (int kb = 0; kb < nkb; kb++)
			{
				const int s = kb % FU_STAGES;
				const uint32_t ph = (kb / FU_STAGES) & 1;
				mbar_wait(full0 + 8 * s, ph);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t st = tiles + s * FU_STAGE_BYTES;
				const uint64_t ahi = umma_desc_k_sw128(st), alo = umma_desc_k_sw128(st + FU_A_BYTES);
				const uint64_t bhi = umma_desc_k_sw128(st + 2 * FU_A_BYTES), blo = umma_desc_k_sw128(st + 2 * FU_A_BYTES + FU_B_BYTES);
#pragma unroll
				for (int ks = 0; ks < 4; ks++)
				{
					const uint64_t adv = (uint64_t) ((ks * 32) >> 4);
					tcgen05_mma_tf32(tmem_base + FU_BN, alo + adv, bhi + adv, idesc, (kb | ks) != 0);
					tcgen05_mma_tf32(tmem_base + FU_BN, ahi + adv, blo + adv, idesc, 1);
					tcgen05_mma_tf32(tmem_base, ahi + adv, bhi + adv, idesc, (kb | ks) != 0);
				}
				tcgen05_commit(empty0 + 8 * s);
			}
			tcgen05_commit(tmem_full);
		}
	}
	else
	{
		// ---- producers: 256 threads, thread -> pixel (lane & 15) of the K-block and 8 rows pw*16 + 2*j + (lane >> 4) ----
		const int pt = threadIdx.x - 64;
		const int pw = pt >> 5;
		const int i = lane & 15, rsub = lane >> 4;
		const int imgX = A.n / 2 + 1;
		// Rows are dealt to the warps in rounds of 2 NPW (round j: rows 2 NPW j + 2 pw + rsub), and only the rounds that hold rows
		// of this tile are projected: a particle's last tile (177 orientations = 128 + 49) costs its share of the rows, not a full
		// tile - with rows blocked per warp it took as long as a full one (the busy warps set the pace of every K-block).
		const int nj = (min(FU_BM, no - oi0) + 2 * NPW - 1) / (2 * NPW);
		const RbProjK pk = rb_make_projk2(A.projs[cls], imgX);
		const RbProjK8 pk8 = rb_make_projk8(A.projs[cls], imgX);
		const float4 *mdl2 = A.projs[cls].mdl2;
		const float4 *quad = A.projs[cls].quad;
		const float4 *img = A.img4 + (size_t) p * A.n * imgX;
		float bacc[NJ];
#pragma unroll
		for (int j = 0; j < NJ; j++) bacc[j] = 0.f;
		for (int kb = 0; kb < nkb; kb++)
		{
			const int s = kb % FU_STAGES;
			const uint32_t ph = (kb / FU_STAGES) & 1;
			const int ip = kb * FU_PIX + i;
			int x = 0, y = 0; float hc = 0.f;
			const bool pix_ok = ip < A.npix;
			if (pix_ok)
			{
				const uint32_t pkx = __ldg(A.pix + ip);
				x = rb_pix_x(pkx); y = rb_pix_y(pkx);
				hc = __ldg(img + rb_src_index(x, y, A.n)).z;
			}
			float2 ref[NJ];
			// G256: 2 / 3 = xy-quad gather with the loads of 4 / 2 rows in flight, 4 = one row at a time
			constexpr int FU_QUAD_MLP = G256 == 2 ? 4 : G256 == 3 ? 2 : 1;
			if constexpr (G256 == 2 || G256 == 3)
			{
				// the loads of FU_QUAD_MLP rows in flight at once (branch-free issue, then the lerps)
#pragma unroll
				for (int j0 = 0; j0 < NJ; j0 += FU_QUAD_MLP)
				{
					if (j0 >= nj)                                      // CTA-uniform: the rounds beyond a partly filled tile issue no loads at all
					{
#pragma unroll
						for (int jj = 0; jj < FU_QUAD_MLP; jj++) ref[j0 + jj] = make_float2(0.f, 0.f);
						continue;
					}
					RbQuadFetch qf[FU_QUAD_MLP];
#pragma unroll
					for (int jj = 0; jj < FU_QUAD_MLP; jj++)
					{
						const int j = j0 + jj, r = 2 * NPW * j + 2 * pw + rsub;
						rb_quad_issue(pk, quad, j < nj && pix_ok && s_valid[r], x, y, s_e[r][0], s_e[r][1], s_e[r][2], s_e[r][3], s_e[r][4], s_e[r][5], qf[jj]);
					}
#pragma unroll
					for (int jj = 0; jj < FU_QUAD_MLP; jj++)
					{
						const int j = j0 + jj;
						ref[j] = rb_quad_finish(qf[jj]);
						bacc[j] = fmaf(hc, ref[j].x * ref[j].x + ref[j].y * ref[j].y, bacc[j]);
					}
				}
			}
			else
#pragma unroll
			for (int j = 0; j < NJ; j++)
			{
				const int r = 2 * NPW * j + 2 * pw + rsub;
				ref[j] = make_float2(0.f, 0.f);
				if (j < nj && pix_ok && s_valid[r])
				{
					if constexpr (G256 >= 2) ref[j] = rb_project3d_q256(pk, quad, x, y, s_e[r][0], s_e[r][1], s_e[r][2], s_e[r][3], s_e[r][4], s_e[r][5]);
					else if constexpr (G256 == 1) ref[j] = rb_project3d_c256(pk8, x, y, s_e[r][0], s_e[r][1], s_e[r][2], s_e[r][3], s_e[r][4], s_e[r][5]);
					else ref[j] = rb_project3d_xp(pk, mdl2, x, y, s_e[r][0], s_e[r][1], s_e[r][2], s_e[r][3], s_e[r][4], s_e[r][5]);
				}
				bacc[j] = fmaf(hc, ref[j].x * ref[j].x + ref[j].y * ref[j].y, bacc[j]);
			}
			mbar_wait(empty0 + 8 * s, ph ^ 1);                       // the MMAs that read this slot have retired
			uint8_t *st = tiles_generic + (size_t) s * FU_STAGE_BYTES;
#pragma unroll
			for (int j = 0; j < NJ; j++)
			{
				if (j >= nj) continue;                                 // rows beyond the tile: their accumulator rows are never read
				const int r = 2 * NPW * j + 2 * pw + rsub;
				float2 h, l;
				tf32_split(ref[j].x, h.x, l.x); tf32_split(ref[j].y, h.y, l.y);
				const uint32_t off = (uint32_t) r * 128u + ((uint32_t) ((i >> 1) ^ (r & 7)) << 4) + (uint32_t) (i & 1) * 8u;
				*(float2 *) (st + off) = h;
				*(float2 *) (st + FU_A_BYTES + off) = l;
			}
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
			mbar_arrive(full0 + 8 * s);
		}
		// norm term per row: sum over the 16 lanes that share a row
#pragma unroll
		for (int j = 0; j < NJ; j++)
		{
			float v = bacc[j];
			v += __shfl_xor_sync(RB_FULL_MASK, v, 8); v += __shfl_xor_sync(RB_FULL_MASK, v, 4);
			v += __shfl_xor_sync(RB_FULL_MASK, v, 2); v += __shfl_xor_sync(RB_FULL_MASK, v, 1);
			if (i == 0) s_base[2 * NPW * j + 2 * pw + rsub] = v;
		}
		asm volatile("bar.sync 1, %0;" ::"n"(32 * NPW) : "memory");     // producers only
		if (pw < 4)
		{
			// ---- epilogue: warps 2..5 <-> TMEM lane quarters (warp % 4) ----
			const int q = warp & 3;
			const int r = q * 32 + lane;
			mbar_wait(tmem_full, 0);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			uint32_t v[32], v2[32];
			tmem_ld32(tmem_base + ((uint32_t) (q * 32) << 16), v);
			tmem_ld32(tmem_base + ((uint32_t) (q * 32) << 16) + FU_BN, v2);
			float bmin = FLT_MAX;
			if (s_valid[r])
			{
				const float bs = s_base[r] + A.x2[p];
				float *out = A.Mweight + m.coarse_off + (long long) (o0 + r) * A.Tstride + A.t0;
#pragma unroll
				for (int t = 0; t < 32; t++)
				{
					if (t < A.T)
					{
						const float cr = __uint_as_float(v[t]) + __uint_as_float(v2[t]);
						if (A.cc) { out[t] = -(cr / sqrtf(s_base[r])); continue; }          // diff2.h:729-735; k_weights_cc_coarse takes the minimum
						const float d = fmaxf(bs - 2.f * cr, 0.f) + m.xi2_half;           // diff2.cuh:170-186, :1290-1296
						out[t] = d;
						bmin = fminf(bmin, d);
					}
				}
			}
			bmin = -warp_max(-bmin);
			if (lane == 0) s_min[q] = bmin;
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
		}
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		const float b = fminf(fminf(s_min[0], s_min[1]), fminf(s_min[2], s_min[3]));
		if (b < FLT_MAX) rb_atomic_min_pos(&A.states[p].min_diff2_bits, b);
	}
	if (warp == 1)
	{
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
	}
}


bool rbk_coarse_fused_applicable(rb_ctx *ctx, const PoolSlot &s)
{
	const char *e = getenv("RB_COARSE_FUSED");
	const int mode = e ? atoi(e) : 1;
	if (mode == 0 || ctx->d_samp.n_trans > 8 * FU_BN) return false;   // more than 32 translations: several passes of 32 (below)
	return mode == 2 || s.has_priors;                      // default: the local-search path
}

int rbk_diff2_coarse_fused_pool(rb_ctx *ctx, PoolSlot &s, const float4 *cimg4)
{
	const RbModelDev &M = ctx->d_model;
	const RbSamplingDev &S = ctx->d_samp;
	const int P = s.P, T = S.n_trans, K = M.nr_classes;
	const int npix = M.d2_nvc, n = M.coarse_size;         // d2_*: with the CC criterion every pixel of the window's circle
	const uint32_t *pixlist = M.d2_pix_c;
	const int nkb = (npix + FU_PIX - 1) / FU_PIX;
	const size_t kpad = (size_t) nkb * 32;
	const size_t rows = (size_t) P * FU_BN;
	DevBuf &bBhi = ctx->gemm_buf[4], &bBlo = ctx->gemm_buf[5], &bX2 = ctx->gemm_buf[9];
	RB_CHECK(bBhi.ensure(rows * kpad * 4)); RB_CHECK(bBlo.ensure(rows * kpad * 4)); RB_CHECK(bX2.ensure((size_t) P * 4));
	CUtensorMap tb, tbl;
	RB_CHECK(make_tmap(&tb, bBhi.as<float>(), rows, kpad, FU_BN)); RB_CHECK(make_tmap(&tbl, bBlo.as<float>(), rows, kpad, FU_BN));
	FusedArgs A;
	memset(&A, 0, sizeof(A));
	A.metas = s.meta.as<RbPartMeta>(); A.states = s.state.as<RbPartState>();
	A.dir_idx = s.dir_idx.as<int>(); A.psi_idx = s.psi_idx.as<int>();
	A.pdf_orient_zero = s.pdf_orient_zero.as<unsigned char>(); A.Mweight = s.Mweight.as<float>();
	A.img4 = cimg4; A.x2 = bX2.as<float>(); A.projs = ctx->d_proj.as<RbProjector>();
	A.order = s.has_order ? s.order.as<int>() : nullptr;
	A.pix = pixlist; A.npix = npix; A.n = n; A.num_kblocks = nkb; A.cc = M.do_cc;
	A.tiles_per_class = (s.max_no + FU_BM - 1) / FU_BM;
	static int npw = 0;
	// measured (256 px, hp4 local): 8 warps x 2 CTAs/SM (166 KB of shared memory, 32 KB of L1 left) 7.98 ms; 16 warps x 1 CTA/SM
	// (83 KB, 128 KB of L1 for the gathers) 6.69 ms.  Loads that bypass L1 allocation: 18.6 ms - the corner rows do get reused.
	if (!npw) { const char *e = getenv("RB_FUSED_WARPS"); npw = (e && atoi(e) == 8) ? 8 : 16; }
	static bool configured_dev[RB_MAX_DEVICES] = {};
	bool &configured = configured_dev[ctx->device % RB_MAX_DEVICES];
	if (!configured)
	{
		RB_CUDA(cudaFuncSetAttribute(k_coarse_fused<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FU_SMEM));
		RB_CUDA(cudaFuncSetAttribute(k_coarse_fused<16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FU_SMEM));
		RB_CUDA(cudaFuncSetAttribute(k_coarse_fused<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FU_SMEM));
		RB_CUDA(cudaFuncSetAttribute(k_coarse_fused<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FU_SMEM));
		RB_CUDA(cudaFuncSetAttribute(k_coarse_fused<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FU_SMEM));
		RB_CUDA(cudaFuncSetAttribute(k_coarse_fused<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FU_SMEM));
		RB_CUDA(cudaFuncSetAttribute(k_coarse_fused<16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FU_SMEM));
		RB_CUDA(cudaFuncSetAttribute(k_coarse_fused<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FU_SMEM));
		configured = true;
	}
	// gather: 0 = x-pair copy, four 16-byte loads per sample; 1 = expanded cells, two 32-byte loads per sample; 2 = xy-quad copy of
	// the coarse core (when ensure_coarse_core built one), two 32-byte loads per sample
	static int g256_env = -1;
	if (g256_env < 0) { const char *e = getenv("RB_FUSED_G256"); g256_env = e ? atoi(e) : 0; }
	int g256 = g256_env;
	bool have_quad = true;
	for (int k = 0; k < K; k++) have_quad = have_quad && ctx->proj[k].quad != nullptr;
	if (have_quad) g256 = (g256 >= 2 && g256 <= 4) ? g256 : 3; else if (g256 >= 2) g256 = 0;   // measured: 4 -> 5.79 ms, 3 -> 5.75, 2 -> 6.54
	dim3 grid((unsigned) P, (unsigned) (A.tiles_per_class * K));
	// The tile holds 32 translations (TMEM columns, B operand rows): samplings with more (--offset_range 5 --offset_step 1: 81,
	// healpix_sampling.cpp:399-440) take ceil(T / 32) passes, each with its own B operand, writing its columns of Mweight.  The
	// projection is repeated per pass; the SIMT kernel this replaces is 10-40x slower than one pass.
	for (int t0 = 0; t0 < T; t0 += FU_BN)
	{
		const int Tc = std::min(FU_BN, T - t0);
		k_gemm_build_B<<<(unsigned) rows, 256, 0, ctx->stream>>>(cimg4, pixlist, npix, n, S.ctx + t0, S.cty + t0, Tc, FU_BN, P,
			bBhi.as<float>(), bBlo.as<float>(), kpad, nullptr, nullptr, 0, bX2.as<float>(), 0, 0);
		RB_LAUNCH_CHECK(ctx);
		A.T = Tc; A.Tstride = T; A.t0 = t0;
		if (npw == 16)
		{
			if (g256 == 2) k_coarse_fused<16, 2><<<grid, 64 + 32 * 16, FU_SMEM, ctx->stream>>>(tb, tbl, A, ctx->d_samp);
			else if (g256 == 3) k_coarse_fused<16, 3><<<grid, 64 + 32 * 16, FU_SMEM, ctx->stream>>>(tb, tbl, A, ctx->d_samp);
			else if (g256 == 4) k_coarse_fused<16, 4><<<grid, 64 + 32 * 16, FU_SMEM, ctx->stream>>>(tb, tbl, A, ctx->d_samp);
			else if (g256) k_coarse_fused<16, 1><<<grid, 64 + 32 * 16, FU_SMEM, ctx->stream>>>(tb, tbl, A, ctx->d_samp);
			else k_coarse_fused<16, 0><<<grid, 64 + 32 * 16, FU_SMEM, ctx->stream>>>(tb, tbl, A, ctx->d_samp);
		}
		else
		{
			if (g256 >= 2) k_coarse_fused<8, 2><<<grid, 64 + 32 * 8, FU_SMEM, ctx->stream>>>(tb, tbl, A, ctx->d_samp);
			else if (g256) k_coarse_fused<8, 1><<<grid, 64 + 32 * 8, FU_SMEM, ctx->stream>>>(tb, tbl, A, ctx->d_samp);
			else k_coarse_fused<8, 0><<<grid, 64 + 32 * 8, FU_SMEM, ctx->stream>>>(tb, tbl, A, ctx->d_samp);
		}
		RB_LAUNCH_CHECK(ctx);
	}
	return RB_OK;
}

// stage entry: C[M][N] = A[M][K] . B[N][K]^T in 3xTF32 (device pointers, row-major)
int rbk_gemm_tf32x3_stage(rb_ctx *ctx, const float *dA, const float *dB, int Mr, int Nr, int Kr, float *dC)
{
	const size_t Mpad = round_up((size_t) Mr, GM_MPAD), Npad = round_up((size_t) Nr, GM_BN), Kpad = round_up((size_t) Kr, GM_BK);
	DevBuf &bAhi = ctx->gemm_buf[0], &bAlo = ctx->gemm_buf[1], &bBhi = ctx->gemm_buf[4], &bBlo = ctx->gemm_buf[5];
	RB_CHECK(bAhi.ensure(Mpad * Kpad * 4)); RB_CHECK(bAlo.ensure(Mpad * Kpad * 4));
	RB_CHECK(bBhi.ensure(Npad * Kpad * 4)); RB_CHECK(bBlo.ensure(Npad * Kpad * 4));
	k_split_pad<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(dA, Mr, Kr, bAhi.as<float>(), bAlo.as<float>(), Mpad, Kpad, gemm_mixed() ? Mpad * Kpad : 0);
	RB_LAUNCH_CHECK(ctx);
	k_split_pad<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(dB, Nr, Kr, bBhi.as<float>(), bBlo.as<float>(), Npad, Kpad, gemm_mixed() ? Npad * Kpad : 0);
	RB_LAUNCH_CHECK(ctx);
	GemmEpilogue E;
	memset(&E, 0, sizeof(E));
	E.mode = 0; E.C = dC; E.ldc = Nr; E.M = Mr; E.N = Nr;
	return launch_gemm(ctx, bAhi.as<float>(), bAlo.as<float>(), Mpad, bBhi.as<float>(), bBlo.as<float>(), Npad, Kpad, E);
}
