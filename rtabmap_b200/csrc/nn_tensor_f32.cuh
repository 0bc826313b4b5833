// nn_tensor_f32.cuh — exact squared-L2 2-NN of FLOAT descriptors (SURF-64 / SIFT-128 sized rows) with the distance matrix on the
// 5th-generation tensor cores (sm_100a): BASELINE configs[3], "tensor cores ... if SURF/float descriptors make it a true dense GEMM".
//
// Replaces, for CV_32F descriptors: FlannIndex::knnSearch(k = 2) on a LinearIndex (corelib/src/FlannIndex.cpp:701-745 ->
// rtflann::L2<float>::operator(), rtflann/algorithms/dist.h:133-180, KNNSimpleResultSet, rtflann/util/result_set.h:151-172), whose
// result — ids AND fp32 distances in rtflann's summation order — must be reproduced bit for bit.  A tensor-core GEMM cannot reproduce
// an fp32 summation order, so it is used as a FILTER with a proven bound, followed by an exact re-rank:
//
//   1. operands rounded to fp16 (round to nearest, no scaling; |x| < 65504 is checked) in pre-tiled, 128-byte-swizzled K-major
//      images: the word image is built once per dictionary change and CACHED (half the bytes of the fp32 rows: the 1M x 64
//      vocabulary streams as 128 MB), the query image per call;
//   2. knn2_tensor_f32_kernel: tcgen05.mma kind::f16 (M128 N256 K16, fp32 accumulators in TMEM).  The word image holds -2 w and one
//      extra K-step holds |w|^2 (as an fp16 hi + lo pair against a constant column of the query image), so the accumulator IS
//      v = |w|^2 - 2 q.w (= approximate squared distance minus |q|^2): the epilogue is a running minimum over the columns (one
//      three-input FMNMX per two columns) and appends to the query's CANDIDATE LIST every row with v <= t2 + 2 eps, where t2 is the
//      second-smallest v this thread has seen so far (initialised from a shared per-query bound);
//   3. rerank_l2_kernel computes the exact rtflann-order distance of every candidate from the fp32 rows and keeps the best two by
//      (distance, row) — the same packed 64-bit keys as the exact kernel of l2_path.cuh, consumed by the same resolve kernel;
//   4. a query whose list overflowed (or whose values do not fit fp16) is redone by an exact scan (knn2_l2_fallback_kernel).
//
// Why the filter cannot lose a true neighbour.  Let d(w) be the distance rtflann computes, e1, e2 the true two nearest rows, and
// |v(w) + |q|^2 - d(w)| <= eps for every row w.  t2 is always the larger of the v's of two DISTINCT rows, hence
// t2 + |q|^2 >= d(e2) - eps; and v(e_i) + |q|^2 <= d(e_i) + eps <= d(e2) + eps.  So v(e_i) <= t2 + 2 eps whenever e_i is examined:
// both true neighbours (and every row tying with them) reach the list, and the exact re-rank then orders them as rtflann does.
// eps bounds (a) fp16 rounding of both operands: |q^.w^ - q.w| <= 2^-10 (1 + 2^-11) |q| |w| + subnormal terms (Cauchy-Schwarz on the
// element-wise relative errors 2^-11), doubled by the factor 2 in the distance; (b) fp32 accumulation in the tensor core, the fma
// the fp16 hi + lo split of |w|^2 (relative 2^-21, plus 2^-19 absolute for its subnormal tail), the float norms and rtflann's own
// rounding, all far below 2^-15 (|q|^2 + |w|^2) + 4e-6:
//     eps(q) = 1.02 * 2^-9 * |q| * Wmax + 2^-15 * (|q|^2 + Wmax^2) + 4e-6,      Wmax = largest |w| in the dictionary.
// For unit-length descriptors eps ~ 2.1e-3 against nearest-neighbour distances of 1e-2 .. 1: the lists stay a few dozen rows long.
#pragma once
#include "common.cuh"
#include "l2_path.cuh"
#include "nn_tensor.cuh"
#include <cuda_fp16.h>

namespace lcd {

constexpr int kTfBM = 128;          // queries per CTA tile (UMMA M)
constexpr int kTfBN = 256;          // words per tile (UMMA N)
constexpr int kTfEpiGroups = 4;     // column groups of a tile, one set of 4 epilogue warps each
constexpr int kTfEpiCols = kTfBN / kTfEpiGroups;
constexpr int kTfThreads = 128 + 128 * kTfEpiGroups;
constexpr int kTfCandCap = 96;      // candidate rows kept per query; more -> exact fallback for that query
constexpr int kTfMinRows = 4096;    // below this the exact CUDA-core kernel is used
constexpr int kTfPrepassTiles = 16; // rows [0, 4096): the bound-only pre-pass that initialises the per-query bound
constexpr float kTfAugScale = 64.0f; // the constant of the query image's augmentation columns; |w|^2 is stored divided by it
constexpr float kTfMaxAbs = 32752.0f; // -2 w must fit fp16
constexpr float kTfMaxNorm2 = 65504.0f * kTfAugScale;

template <int DIM>
struct TfCfg
{
	static constexpr int atoms = DIM / 64;                 // 64 halves = one 128-byte swizzle atom
	static constexpr uint32_t a_bytes = kTfBM * DIM * 2;
	static constexpr uint32_t b_bytes = kTfBN * DIM * 2;
	static constexpr uint32_t a_aug_bytes = kTfBM * 32;    // one K-step (16 halves) per row, un-swizzled core-matrix layout
	static constexpr uint32_t b_aug_bytes = kTfBN * 32;
	static constexpr int stages = DIM == 64 ? 5 : 2;
	static constexpr size_t smem = 1024 + a_bytes + a_aug_bytes + stages * (b_bytes + b_aug_bytes) + 512;
};

// order-preserving float <-> uint mapping (v can be negative), for atomicMin on the shared per-query bound
__device__ __forceinline__ uint32_t f2ord(float f)
{
	const uint32_t b = __float_as_uint(f);
	return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u)
{
	return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// ---- operand images ----------------------------------------------------------------------------------------------
// One thread converts 8 consecutive floats of a row into one 16-byte chunk.  Rows >= n_rows of the last tile are zero.
// Layout: [tile][atom][tile_rows][128 B], chunk c of row r at c ^ (r & 7) (canonical K-major SWIZZLE_128B).
// bad_row[row] (may be null) is set when a value is not finite or does not fit fp16.
template <int DIM>
__global__ void tf_expand_kernel(const float * __restrict__ src, int row_begin, int n_rows, int tile_rows, uint4 * __restrict__ dst,
                                 int * __restrict__ bad_flag, float scale)
{
	constexpr int chunks = DIM / 8;
	const size_t gid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	const int tile0 = row_begin / tile_rows;
	const size_t row = static_cast<size_t>(tile0) * tile_rows + gid / chunks;
	const int chunk = static_cast<int>(gid % chunks);
	const size_t row_end = (static_cast<size_t>(n_rows) + tile_rows - 1) / tile_rows * tile_rows;
	if (row >= row_end) return;
	uint4 o = make_uint4(0, 0, 0, 0);
	if (row < static_cast<size_t>(n_rows))
	{
		const float4 a = *reinterpret_cast<const float4 *>(src + row * DIM + chunk * 8);
		const float4 b = *reinterpret_cast<const float4 *>(src + row * DIM + chunk * 8 + 4);
		const float m = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
		                      fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
		if (!(m < kTfMaxAbs) && bad_flag) *bad_flag = 1; // also true for NaN
		// scale = -2 for the word image (exact in fp16), 1 for the query image
		const __half2 h0 = __floats2half2_rn(a.x * scale, a.y * scale), h1 = __floats2half2_rn(a.z * scale, a.w * scale),
		              h2 = __floats2half2_rn(b.x * scale, b.y * scale), h3 = __floats2half2_rn(b.z * scale, b.w * scale);
		o.x = *reinterpret_cast<const uint32_t *>(&h0);
		o.y = *reinterpret_cast<const uint32_t *>(&h1);
		o.z = *reinterpret_cast<const uint32_t *>(&h2);
		o.w = *reinterpret_cast<const uint32_t *>(&h3);
	}
	const int tile = static_cast<int>(row / tile_rows), r = static_cast<int>(row % tile_rows);
	const int atom = chunk >> 3, c = chunk & 7;
	const size_t off16 = static_cast<size_t>(tile) * (static_cast<size_t>(tile_rows) * chunks) + static_cast<size_t>(atom) * (tile_rows * 8) +
	                     static_cast<size_t>(r) * 8 + static_cast<size_t>(c ^ (r & 7));
	dst[off16] = o;
}

// The augmentation K-step of rows [row_begin rounded down to a tile, n_rows rounded up): 16 halves per row in the un-swizzled K-major
// core-matrix layout ([8-row group][k-chunk 0..1][8 rows][8 halves], 256 B per group), image = [tile][tile_rows / 8 groups].
//   words  : (hi, lo, 0...) with hi + lo = |w|^2 / kTfAugScale (+inf in the padding rows); the largest |w|^2 into *wmax2_bits
//   queries: (kTfAugScale, kTfAugScale, 0...)
template <int DIM>
__global__ void tf_aug_kernel(const float * __restrict__ src, int row_begin, int n_rows, int tile_rows, uint4 * __restrict__ aug, int is_query,
                              uint32_t * __restrict__ wmax2_bits, int * __restrict__ bad_flag)
{
	const size_t row = static_cast<size_t>(row_begin / tile_rows) * tile_rows + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	const size_t row_end = (static_cast<size_t>(n_rows) + tile_rows - 1) / tile_rows * tile_rows;
	if (row >= row_end) return;
	__half h0, h1;
	if (is_query)
	{
		h0 = __float2half_rn(kTfAugScale);
		h1 = h0;
	}
	else
	{
		float s = INFINITY;
		if (row < static_cast<size_t>(n_rows))
		{
			s = 0.0f;
			for (int i = 0; i < DIM; i += 4)
			{
				const float4 a = *reinterpret_cast<const float4 *>(src + row * DIM + i);
				s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
			}
			if (!(s < kTfMaxNorm2) && bad_flag) *bad_flag = 1;
			if (wmax2_bits && s == s) atomicMax(wmax2_bits, __float_as_uint(s)); // s >= 0: unsigned order of the bits = float order
		}
		const float t = s / kTfAugScale;
		h0 = __float2half_rn(t);
		h1 = (t < INFINITY) ? __float2half_rn(t - __half2float(h0)) : __float2half_rn(0.0f);
	}
	// row r of the image: group r / 8, core-matrix row r % 8; k-chunk 0 holds halves 0..7 (hi, lo, 0 x 6), k-chunk 1 is zero
	const size_t grp = row / 8;
	const int rr = static_cast<int>(row % 8);
	uint4 c0 = make_uint4(0, 0, 0, 0);
	c0.x = static_cast<uint32_t>(__half_as_ushort(h0)) | (static_cast<uint32_t>(__half_as_ushort(h1)) << 16);
	aug[grp * 16 + rr] = c0;
	aug[grp * 16 + 8 + rr] = make_uint4(0, 0, 0, 0);
}

// per-query state of one search: norm, empty candidate list (or "overflowed" for queries that do not fit fp16), loose bound
template <int DIM>
__global__ void tf_query_init_kernel(const float * __restrict__ q, int nq, float * __restrict__ qn, int * __restrict__ cand_count,
                                     uint32_t * __restrict__ tau)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nq) return;
	float s = 0.0f, m = 0.0f;
	for (int k = 0; k < DIM; k += 4)
	{
		const float4 a = *reinterpret_cast<const float4 *>(q + static_cast<size_t>(i) * DIM + k);
		s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
		m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
	}
	const bool ok = (m < 65504.0f) && (s == s) && (s < INFINITY);
	qn[i] = ok ? s : 0.0f;
	cand_count[i] = ok ? 0 : (kTfCandCap + 1);
	tau[i] = 0xFF800000u; // f2ord(+inf): no bound yet
}

// D[tmem] (+)= A[smem] * B[smem]^T, f16 x f16 -> f32, M128 x N256 x K16
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
		"}\n" ::"r"(tmem_d),
		"l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
		: "memory");
}
// instruction descriptor: f32 accumulator, f16 x f16, both K-major
__device__ __forceinline__ constexpr uint32_t tc_idesc_f16(int m, int n)
{
	return (1u << 4) | (0u << 7) | (0u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// shared-memory matrix descriptor of an UN-swizzled K-major operand slab of one K-step (16 halves): 8 x 16-byte core matrices, the two
// k-chunks of a row group 128 B apart (leading byte offset), row groups 256 B apart (stride byte offset)
__device__ __forceinline__ uint64_t tc_smem_desc_noswz(uint32_t saddr)
{
	uint64_t d = 0;
	d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);        // start address  [0,14)
	d |= static_cast<uint64_t>(128 >> 4) << 16;                  // leading byte offset [16,30): next core matrix along K
	d |= static_cast<uint64_t>(256 >> 4) << 32;                  // stride byte offset [32,46): next 8-row group
	d |= static_cast<uint64_t>(1) << 46;                         // descriptor version (sm_100)
	return d;                                                    // layout type 0: no swizzle
}

struct TfArgs
{
	const uint4 * word_img;      // fp16 image of -2 w
	const uint4 * word_aug;      // augmentation K-step of the words (|w|^2 as hi + lo)
	int n_rows;                  // rows of this search
	const uint4 * query_img;
	const uint4 * query_aug;
	const float * qn;            // [nq] |q|^2
	int nq;
	int tile_begin, tile_end;    // tile range of the whole launch; blockIdx.y * tiles_per_split selects the CTA's share
	int tiles_per_split;
	const uint32_t * wmax2_bits; // largest |w|^2 of the dictionary (float bits)
	uint32_t * tau;              // [nq] shared bound on the second-smallest v (ordered-uint encoding)
	uint32_t * cand;             // [nq][kTfCandCap] candidate rows (local to this engine)
	int * cand_count;            // [nq]
	int emit;                    // 0: bound-only pre-pass
};

// grid = (query tiles, splits).  Roles as in knn2_tensor_kernel (nn_tensor.cuh): warp 0 producer (bulk copies of the word tiles and
// their augmentation slabs), warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-19 epilogue (thread = one query = one TMEM lane,
// column group (warp - 4) / 4).
template <int DIM>
__global__ void __launch_bounds__(kTfThreads, 1)
knn2_tensor_f32_kernel(const TfArgs a)
{
	using Cfg = TfCfg<DIM>;
	constexpr int kStages = Cfg::stages;
	constexpr uint32_t kStageBytes = Cfg::b_bytes + Cfg::b_aug_bytes;
	extern __shared__ unsigned char smem_dyn[];
	unsigned char * smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
	unsigned char * sA = smem;                                   // query tile, then its augmentation slab
	unsigned char * sB = smem + Cfg::a_bytes + Cfg::a_aug_bytes; // stages of (word tile, augmentation slab)
	uint64_t * bars = reinterpret_cast<uint64_t *>(sB + kStages * kStageBytes);
	uint64_t * full = bars;                     // [kStages] word tile landed
	uint64_t * empty = full + kStages;          // [kStages] word tile consumed by the MMAs
	uint64_t * tfull = empty + kStages;         // [2] accumulator stage complete
	uint64_t * tempty = tfull + 2;              // [2] accumulator stage drained
	uint64_t * afull = tempty + 2;              // [1] query tile landed
	uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(afull + 1);

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int tile_begin = a.tile_begin + blockIdx.y * a.tiles_per_split;
	const int tile_end = min(a.tile_end, tile_begin + a.tiles_per_split);
	const int n_tiles = max(0, tile_end - tile_begin);
	const int qtile = blockIdx.x;

	if (tid == 0)
	{
		for (int s = 0; s < kStages; ++s)
		{
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], 1);
		}
		for (int s = 0; s < 2; ++s)
		{
			mbar_init(&tfull[s], 1);
			mbar_init(&tempty[s], 4 * kTfEpiGroups);
		}
		mbar_init(afull, 1);
		mbar_fence_init();
	}
	if (warp == 2) tc_alloc(tmem_slot, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0)
	{
		if (lane == 0 && n_tiles > 0)
		{
			mbar_arrive_expect_tx(afull, Cfg::a_bytes + Cfg::a_aug_bytes);
			bulk_g2s(sA, reinterpret_cast<const unsigned char *>(a.query_img) + static_cast<size_t>(qtile) * Cfg::a_bytes, Cfg::a_bytes, afull);
			bulk_g2s(sA + Cfg::a_bytes, reinterpret_cast<const unsigned char *>(a.query_aug) + static_cast<size_t>(qtile) * Cfg::a_aug_bytes,
			         Cfg::a_aug_bytes, afull);
			for (int t = 0; t < n_tiles; ++t)
			{
				const int s = t % kStages;
				if (t >= kStages) mbar_wait(&empty[s], ((t / kStages) - 1) & 1);
				mbar_arrive_expect_tx(&full[s], kStageBytes);
				const unsigned char * src = reinterpret_cast<const unsigned char *>(a.word_img) + static_cast<size_t>(tile_begin + t) * Cfg::b_bytes;
				unsigned char * dst = sB + s * kStageBytes;
				bulk_g2s(dst, src, Cfg::b_bytes / 2, &full[s]);
				bulk_g2s(dst + Cfg::b_bytes / 2, src + Cfg::b_bytes / 2, Cfg::b_bytes / 2, &full[s]);
				bulk_g2s(dst + Cfg::b_bytes, reinterpret_cast<const unsigned char *>(a.word_aug) + static_cast<size_t>(tile_begin + t) * Cfg::b_aug_bytes,
				         Cfg::b_aug_bytes, &full[s]);
			}
		}
	}
	else if (warp == 1)
	{
		if (lane == 0 && n_tiles > 0)
		{
			constexpr uint32_t idesc = tc_idesc_f16(kTfBM, kTfBN);
			const uint32_t a_base = smem_u32(sA);
			const uint64_t a_aug = tc_smem_desc_noswz(a_base + Cfg::a_bytes);
			mbar_wait(afull, 0);
			for (int t = 0; t < n_tiles; ++t)
			{
				const int s = t % kStages, acc = t & 1;
				if (t >= 2) mbar_wait(&tempty[acc], ((t >> 1) - 1) & 1);
				mbar_wait(&full[s], (t / kStages) & 1);
				tc_fence_after();
				const uint32_t b_base = smem_u32(sB + s * kStageBytes);
				const uint32_t d_addr = tmem_base + static_cast<uint32_t>(acc * kTfBN);
#pragma unroll
				for (int ks = 0; ks < DIM / 16; ++ks)
				{
					const uint32_t atom = ks >> 2, koff = (ks & 3) * 32; // 16 halves = 32 bytes per K step, 4 steps per 128-byte atom
					const uint64_t ad = tc_smem_desc(a_base + atom * (kTfBM * 128) + koff);
					const uint64_t bd = tc_smem_desc(b_base + atom * (kTfBN * 128) + koff);
					tc_mma_f16(d_addr, ad, bd, idesc, ks > 0 ? 1u : 0u);
				}
				// the augmentation K-step: + kTfAugScale * (hi + lo) = + |w|^2
				tc_mma_f16(d_addr, a_aug, tc_smem_desc_noswz(b_base + Cfg::b_bytes), idesc, 1u);
				tc_commit(&empty[s]);
				tc_commit(&tfull[acc]);
			}
		}
	}
	else if (warp >= 4)
	{
		const int quarter = warp & 3;
		const int cg = (warp - 4) >> 2;
		const int qi = qtile * kTfBM + quarter * 32 + lane;
		const bool live = qi < a.nq;
		float margin = 0.0f, t1 = INFINITY, t2 = INFINITY;
		if (live)
		{
			const float q2 = a.qn[qi], w2 = __uint_as_float(*a.wmax2_bits);
			const float eps = 1.02f * 0.001953125f * sqrtf(q2) * sqrtf(w2) + 3.0517578125e-05f * (q2 + w2) + 4.0e-6f;
			margin = 2.0f * eps;
			t2 = ord2f(a.tau[qi]);
		}
		const int emit = a.emit;
		for (int t = 0; t < n_tiles; ++t)
		{
			const int acc = t & 1;
			mbar_wait(&tfull[acc], (t >> 1) & 1);
			tc_fence_after();
			const int row0 = (tile_begin + t) * kTfBN + cg * kTfEpiCols;
			const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * kTfBN + cg * kTfEpiCols);
			int v[kTfEpiCols];
#pragma unroll
			for (int c0 = 0; c0 < kTfEpiCols; c0 += 32) tc_ld32(taddr + c0, *reinterpret_cast<int(*)[32]>(&v[c0]));
			tc_wait_ld();
			// the accumulator stage is free as soon as the values are in registers
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(&tempty[acc]);
			float m = INFINITY;
#pragma unroll
			for (int j = 0; j < kTfEpiCols; j += 2) m = fminf(m, fminf(__int_as_float(v[j]), __int_as_float(v[j + 1])));
			if (live && m <= t2 + margin)
			{
				// rare: some column of this tile may belong to the two nearest rows
#pragma unroll
				for (int j = 0; j < kTfEpiCols; ++j)
				{
					const float x = __int_as_float(v[j]);
					// rows past n_rows of the cached image are words that are not searchable in this call (not indexed yet)
					if (x <= t2 + margin && x < INFINITY && row0 + j < a.n_rows)
					{
						if (emit)
						{
							const int pos = atomicAdd(&a.cand_count[qi], 1);
							if (pos < kTfCandCap) a.cand[static_cast<size_t>(qi) * kTfCandCap + pos] = static_cast<uint32_t>(row0 + j);
						}
						// t2 only ever tightens: it starts as the shared bound while t1 is still unknown (+inf)
						if (x < t1)
						{
							t2 = fminf(t2, t1);
							t1 = x;
						}
						else if (x < t2) t2 = x;
					}
				}
			}
			__syncwarp();
		}
		if (live && t2 < INFINITY) atomicMin(&a.tau[qi], f2ord(t2));
	}

	tc_fence_before();
	__syncthreads();
	if (warp == 2) tc_dealloc(tmem_base, 512);
}

// ---- exact re-rank of the candidate lists: one warp per query -----------------------------------------------------------------
// partial[q] = (best key, second key) with keys (float bits of rtflann's distance << 32 | row_offset + row); queries whose list
// overflowed go to the fallback list instead.
template <int DIM>
__global__ void rerank_l2_kernel(const float * __restrict__ vocab, int row_offset, const float * __restrict__ queries, int nq,
                                 const uint32_t * __restrict__ cand, const int * __restrict__ cand_count, ulonglong2 * __restrict__ partial,
                                 int * __restrict__ fb_list, int * __restrict__ fb_count)
{
	const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const int lane = threadIdx.x & 31;
	if (qi >= nq) return;
	const int n = cand_count[qi];
	if (n > kTfCandCap)
	{
		if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = qi;
		return;
	}
	unsigned long long k1 = kKey64None, k2 = kKey64None;
	const float * q = queries + static_cast<size_t>(qi) * DIM;
	for (int c = lane; c < n; c += 32)
	{
		const uint32_t row = cand[static_cast<size_t>(qi) * kTfCandCap + c];
		const float d = l2_rtflann<DIM>(q, vocab + static_cast<size_t>(row) * DIM);
		top2_insert64(k1, k2, pack64(d, static_cast<uint32_t>(row_offset) + row));
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		const unsigned long long o1 = __shfl_down_sync(0xFFFFFFFFu, k1, o);
		const unsigned long long o2 = __shfl_down_sync(0xFFFFFFFFu, k2, o);
		top2_insert64(k1, k2, o1);
		top2_insert64(k1, k2, o2);
	}
	if (lane == 0) partial[qi] = make_ulonglong2(k1, k2);
}

// exact scan for the queries of the fallback list (list overflow, values outside fp16), spread over the whole machine: the CTAs form
// S row-splits x G query-group slots (S as large as the scratch allows: one overflowing query still uses every SM).  A CTA takes
// kFbQueries queries at a time and streams its share of the vocabulary once for all of them: 256-row tiles are loaded coalesced
// into (padded) shared memory, thread t then computes row t against every query in rtflann's summation order.  The per-split results
// are merged by knn2_l2_fallback_merge_kernel.  Dynamic shared memory: 256 * (DIM + 1) + kFbQueries * DIM floats.
constexpr int kFbQueries = 4;
template <int DIM>
__host__ __device__ constexpr size_t fallback_smem_bytes() { return (256 * (DIM + 1) + kFbQueries * DIM) * sizeof(float); }

__device__ __forceinline__ int fallback_splits(int n_fb, int n_ctas, int slots) { return max(1, min(n_ctas, slots / max(n_fb, 1))); }

template <int DIM>
__global__ void __launch_bounds__(256)
knn2_l2_fallback_kernel(const float * __restrict__ vocab, int n_rows, int row_offset, const float * __restrict__ queries,
                        const int * __restrict__ fb_list, const int * __restrict__ fb_count, ulonglong2 * __restrict__ scratch, int slots)
{
	extern __shared__ __align__(16) float fb_smem[];
	float * s_rows = fb_smem;                    // [256][DIM + 1]
	float * s_q = fb_smem + 256 * (DIM + 1);     // [kFbQueries][DIM]
	__shared__ unsigned long long s_k1[kFbQueries][8], s_k2[kFbQueries][8];
	const int n_fb = *fb_count;
	if (n_fb <= 0) return;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int S = fallback_splits(n_fb, gridDim.x, slots);
	const int split = blockIdx.x % S, gslot = blockIdx.x / S, n_gslots = gridDim.x / S;
	if (gslot >= n_gslots) return;
	const int n_groups = (n_fb + kFbQueries - 1) / kFbQueries;
	const int n_tiles = (n_rows + 255) / 256;
	for (int g = gslot; g < n_groups; g += n_gslots)
	{
		const int nqg = min(kFbQueries, n_fb - g * kFbQueries);
		__syncthreads();
		for (int i = tid; i < nqg * DIM; i += 256) s_q[i] = queries[static_cast<size_t>(fb_list[g * kFbQueries + i / DIM]) * DIM + (i % DIM)];
		unsigned long long k1[kFbQueries], k2[kFbQueries];
#pragma unroll
		for (int j = 0; j < kFbQueries; ++j) k1[j] = k2[j] = kKey64None;
		for (int t = split; t < n_tiles; t += S)
		{
			const int r0 = t * 256;
			const int nr = min(256, n_rows - r0);
			__syncthreads();
			for (int i = tid; i < nr * (DIM / 4); i += 256)
			{
				const int r = i / (DIM / 4), c = (i % (DIM / 4)) * 4;
				const float4 v = *reinterpret_cast<const float4 *>(vocab + static_cast<size_t>(r0 + r) * DIM + c);
				float * d = s_rows + r * (DIM + 1) + c;
				d[0] = v.x;
				d[1] = v.y;
				d[2] = v.z;
				d[3] = v.w;
			}
			__syncthreads();
			if (tid < nr)
			{
#pragma unroll
				for (int j = 0; j < kFbQueries; ++j)
				{
					if (j < nqg)
					{
						const float d = l2_rtflann<DIM>(s_q + j * DIM, s_rows + tid * (DIM + 1));
						top2_insert64(k1[j], k2[j], pack64(d, static_cast<uint32_t>(row_offset + r0 + tid)));
					}
				}
			}
		}
#pragma unroll
		for (int j = 0; j < kFbQueries; ++j)
		{
#pragma unroll
			for (int o = 16; o > 0; o >>= 1)
			{
				const unsigned long long o1 = __shfl_down_sync(0xFFFFFFFFu, k1[j], o);
				const unsigned long long o2 = __shfl_down_sync(0xFFFFFFFFu, k2[j], o);
				top2_insert64(k1[j], k2[j], o1);
				top2_insert64(k1[j], k2[j], o2);
			}
			if (lane == 0)
			{
				s_k1[j][warp] = k1[j];
				s_k2[j][warp] = k2[j];
			}
		}
		__syncthreads();
		if (tid < nqg)
		{
			unsigned long long a1 = kKey64None, a2 = kKey64None;
			for (int w = 0; w < 8; ++w)
			{
				top2_insert64(a1, a2, s_k1[tid][w]);
				top2_insert64(a1, a2, s_k2[tid][w]);
			}
			scratch[static_cast<size_t>(g * kFbQueries + tid) * S + split] = make_ulonglong2(a1, a2);
		}
	}
}

__global__ void knn2_l2_fallback_merge_kernel(const int * __restrict__ fb_list, const int * __restrict__ fb_count, const ulonglong2 * __restrict__ scratch,
                                              int slots, int n_ctas, ulonglong2 * __restrict__ partial)
{
	const int n_fb = *fb_count;
	const int S = fallback_splits(n_fb, n_ctas, slots);
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_fb; k += gridDim.x * blockDim.x)
	{
		unsigned long long a1 = kKey64None, a2 = kKey64None;
		for (int sp = 0; sp < S; ++sp)
		{
			const ulonglong2 p = scratch[static_cast<size_t>(k) * S + sp];
			top2_insert64(a1, a2, p.x);
			top2_insert64(a1, a2, p.y);
		}
		partial[fb_list[k]] = make_ulonglong2(a1, a2);
	}
}

} // namespace lcd
