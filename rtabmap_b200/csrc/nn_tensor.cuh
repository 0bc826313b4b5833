// nn_tensor.cuh — exact Hamming 2-NN of 256-bit descriptors on the 5th-generation tensor cores (sm_100a).
//
// The all-pairs distance matrix of VWDictionary::addNewWords / findNN (reference: corelib/src/VWDictionary.cpp:1004-1118,
// the brute-force branch, and rtflann's LinearIndex::findNeighbors, rtflann/algorithms/linear_index.h:113-123) is
// GEMM-shaped: with every descriptor bit b encoded as the signed byte 2b-1 the dot product of two descriptors is
//     a . w = 256 - 2 * hamming(a, w),
// an integer the tensor cores accumulate exactly in s32 (tcgen05.mma kind::i8).  The packed key order of
// common.cuh, (distance << 22) | row, is recovered in the epilogue as ((256 - acc) << 21) + row, so the result is
// bit-identical to the POPC kernel of nn_hamming.cuh (and to the oracle).
//
// One CTA owns a tile of 128 queries (the M dimension; the tile's accumulator rows are the 128 TMEM lanes, so
// each epilogue thread owns one query and keeps its running top-2 in registers) and a contiguous range of
// 256-word tiles (N).  Roles:
//   warp 0  producer  : cp.async.bulk (TMA engine) of the pre-tiled, pre-swizzled word image into smem (the image is cached with the
//                       dictionary); the query tile is expanded in place by the epilogue warps from its packed form
//   warp 1  MMA       : one thread issues 8 x tcgen05.mma (M128 N256 K32) per word tile into one of two TMEM
//                       accumulator stages, then tcgen05.commit to the smem-empty and accumulator-full barriers
//   warp 2  TMEM alloc/dealloc (512 columns)
//   warps 4-19 epilogue: warp w reads TMEM lanes 32*(w%4).. (its queries) and the 64 columns of column group
//                       (w-4)/4: two tcgen05.ld.x32 in flight, a max over the 64 values against the running
//                       second-best filters almost every group, only improving candidates reach the key insert
//
// Operand images (written by the expand kernels below): K-major, 128-byte swizzle (the canonical UMMA/TMA layout:
// 16-byte chunk c of row r of every 8-row x 128-byte block sits at chunk c ^ (r & 7)), split in K/128 "atoms":
//   query tile : [atom][128 rows][128 B]            (32 KB for 256-bit descriptors)
//   word tile  : [atom][256 rows][128 B]            (64 KB)
// so a tile is ONE contiguous bulk copy and no tensor map is needed.
#pragma once
#include "common.cuh"

namespace lcd {

constexpr int kTcBM = 128;       // queries per CTA tile (UMMA M)
constexpr int kTcBN = 256;       // words per tile (UMMA N)
constexpr int kTcK = 256;        // descriptor bits = int8 K extent
constexpr int kTcAtoms = kTcK / 128;
constexpr int kTcStages = 3;     // word-tile smem stages
constexpr int kTcEpiGroups = 4;  // column groups of a tile, one set of 4 epilogue warps each
constexpr int kTcEpiCols = kTcBN / kTcEpiGroups;
constexpr int kTcThreads = 128 + 128 * kTcEpiGroups;
constexpr uint32_t kTcABytes = kTcBM * kTcK;
constexpr uint32_t kTcBBytes = kTcBN * kTcK;
constexpr size_t kTcSmemBytes = 1024 /* alignment slack */ + kTcABytes + kTcStages * kTcBBytes + 256 /* barriers */;

// ---- operand expansion: bit b -> int8 (2b - 1), tiled + swizzled ---------------------------------
// One thread writes one 16-byte chunk (16 descriptor bits).  rows >= n_rows of the last tile are written as 0.
__device__ __forceinline__ uint32_t expand4(uint32_t nibble)
{
	const uint32_t t = (nibble * 0x00204081u) & 0x01010101u; // bit i -> byte i (0 / 1)
	const uint32_t m = t * 0xFFu;                            // 0x00 / 0xFF per byte
	return (~m) | t;                                         // 0 -> 0xFF (-1), 1 -> 0x01 (+1)
}

__global__ void tc_expand_kernel(const uint32_t * __restrict__ src /* [n_rows][8] */, int n_rows, int tile_rows, int tile0, int n_tiles,
                                 uint4 * __restrict__ dst)
{
	// tiles [tile0, n_tiles): the image of a frozen dictionary is built once and kept; only tiles that gained rows are rewritten
	const size_t gid = static_cast<size_t>(tile0) * tile_rows * 16 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	const size_t total = static_cast<size_t>(n_tiles) * tile_rows * 16;
	if (gid >= total) return;
	const int chunk = static_cast<int>(gid & 15);           // 16-byte chunk of the 256-byte expanded row
	const size_t row = gid >> 4;
	const int tile = static_cast<int>(row / tile_rows), r = static_cast<int>(row % tile_rows);
	uint4 v = make_uint4(0, 0, 0, 0);
	if (row < static_cast<size_t>(n_rows))
	{
		const uint32_t w = src[row * 8 + (chunk >> 1)];
		const uint32_t bits = (w >> ((chunk & 1) * 16)) & 0xFFFFu;
		v.x = expand4(bits & 15u);
		v.y = expand4((bits >> 4) & 15u);
		v.z = expand4((bits >> 8) & 15u);
		v.w = expand4((bits >> 12) & 15u);
	}
	const int atom = chunk >> 3, c = chunk & 7;
	const size_t off16 = static_cast<size_t>(tile) * (static_cast<size_t>(tile_rows) * 16) + static_cast<size_t>(atom) * (tile_rows * 8) +
	                     static_cast<size_t>(r) * 8 + static_cast<size_t>(c ^ (r & 7));
	dst[off16] = v;
}

// ---- tcgen05 / TMEM PTX wrappers ---------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t * bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_alloc(uint32_t * smem_dst, uint32_t cols)
{
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t cols)
{
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t * bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, s8 x s8 -> s32, M128 x N256 x K32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
		"}\n" ::"r"(tmem_d),
		"l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
		: "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr)
{
	uint64_t d = 0;
	d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);        // start address  [0,14)
	d |= static_cast<uint64_t>(0) << 16;                         // leading byte offset (unused for swizzled K-major)
	d |= static_cast<uint64_t>(1024 >> 4) << 32;                 // stride byte offset [32,46)
	d |= static_cast<uint64_t>(1) << 46;                         // descriptor version (sm_100)
	d |= static_cast<uint64_t>(2) << 61;                         // SWIZZLE_128B
	return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): s32 accumulator, s8 x s8, both K-major
__device__ __forceinline__ constexpr uint32_t tc_idesc_i8(int m, int n)
{
	return (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, int (&v)[32])
{
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
		: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
		  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
		  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
		  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
		: "r"(taddr)
		: "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- the kernel ------------------------------------------------------------------------------------
// neg32 = (uint32_t)-32, passed as an argument so that the key arithmetic stays an IMAD on the FMA pipe instead of being
// strength-reduced to a shift-add on the ALU pipe, which the min/max network needs.
// Work items = (query tile, word split) pairs, item = split * n_qtiles + qtile.  The grid is PERSISTENT: CTA b runs items b, b + gridDim.x,
// ... with one TMEM allocation and one set of barriers; the word-tile stream, the MMA issue and the two accumulator stages run on across
// item boundaries (all barrier phases are functions of a CTA-wide running tile count), and the epilogue warps write the next item's query
// tile into shared memory as soon as the last MMA of the current item has completed, before they read that last accumulator out.  What
// an item boundary costs is that expansion (a few hundred cycles) instead of a CTA launch, a TMEM allocation and a pipeline fill — which
// is what a word-range shard with few tiles per item was paying (24 tiles per item at 8 ranks).
// partial[(split * kTcEpiGroups + column group) * nq + query] = (best key, second key) over the rows of that split that fall in that
// column group of their tile.
__global__ void __launch_bounds__(kTcThreads, 1)
knn2_tensor_kernel(const uint4 * __restrict__ word_img, int n_rows, int row_offset, const uint32_t * __restrict__ queries /* [nq][8] packed */, int nq,
                   uint2 * __restrict__ partial, int tiles_per_split, uint32_t neg32, int n_qtiles, int n_splits)
{
	extern __shared__ unsigned char smem_dyn[];
	unsigned char * smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~static_cast<uintptr_t>(1023));
	unsigned char * sA = smem;
	unsigned char * sB = smem + kTcABytes;
	uint64_t * bars = reinterpret_cast<uint64_t *>(sB + kTcStages * kTcBBytes);
	uint64_t * full = bars;                    // [kTcStages] word tile landed
	uint64_t * empty = bars + kTcStages;       // [kTcStages] word tile consumed by the MMAs
	uint64_t * tfull = bars + 2 * kTcStages;   // [2] accumulator stage complete
	uint64_t * tempty = tfull + 2;             // [2] accumulator stage drained by the epilogue
	uint64_t * afull = tempty + 2;             // [1] query tile of the item written (phase = item count of this CTA)
	uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(afull + 1);

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int n_tiles_total = (n_rows + kTcBN - 1) / kTcBN;
	const int n_items = n_qtiles * n_splits;
	// tiles of an item (every role walks the same item sequence and keeps the same running tile count)
	auto item_tiles = [&](int item, int & tile_begin) {
		tile_begin = (item / n_qtiles) * tiles_per_split;
		return max(0, min(n_tiles_total, tile_begin + tiles_per_split) - tile_begin);
	};

	if (tid == 0)
	{
		for (int s = 0; s < kTcStages; ++s)
		{
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], 1);
		}
		for (int s = 0; s < 2; ++s)
		{
			mbar_init(&tfull[s], 1);
			mbar_init(&tempty[s], 4 * kTcEpiGroups); // one arrival per epilogue warp
		}
		mbar_init(afull, 4 * kTcEpiGroups); // the epilogue warps expand the query tile in place (one arrival per warp)
		mbar_fence_init();
	}
	if (warp == 2) tc_alloc(tmem_slot, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0)
	{
		if (lane == 0)
		{
			int g = 0; // running tile count of this CTA
			for (int item = blockIdx.x; item < n_items; item += gridDim.x)
			{
				int tile_begin;
				const int n_tiles = item_tiles(item, tile_begin);
				for (int t = 0; t < n_tiles; ++t, ++g)
				{
					const int s = g % kTcStages;
					if (g >= kTcStages) mbar_wait(&empty[s], ((g / kTcStages) - 1) & 1);
					mbar_arrive_expect_tx(&full[s], kTcBBytes);
					const unsigned char * src = reinterpret_cast<const unsigned char *>(word_img) + static_cast<size_t>(tile_begin + t) * kTcBBytes;
					bulk_g2s(sB + s * kTcBBytes, src, kTcBBytes / 2, &full[s]);
					bulk_g2s(sB + s * kTcBBytes + kTcBBytes / 2, src + kTcBBytes / 2, kTcBBytes / 2, &full[s]);
				}
			}
		}
	}
	else if (warp == 1)
	{
		if (lane == 0)
		{
			constexpr uint32_t idesc = tc_idesc_i8(kTcBM, kTcBN);
			const uint32_t a_base = smem_u32(sA);
			int g = 0, k = 0;
			for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k)
			{
				int tile_begin;
				const int n_tiles = item_tiles(item, tile_begin);
				mbar_wait(afull, k & 1);
				tc_fence_after();
				for (int t = 0; t < n_tiles; ++t, ++g)
				{
					const int s = g % kTcStages, acc = g & 1;
					if (g >= 2) mbar_wait(&tempty[acc], ((g >> 1) - 1) & 1);
					mbar_wait(&full[s], (g / kTcStages) & 1);
					tc_fence_after();
					const uint32_t b_base = smem_u32(sB + s * kTcBBytes);
					const uint32_t d_addr = tmem_base + static_cast<uint32_t>(acc * kTcBN);
#pragma unroll
					for (int ks = 0; ks < kTcK / 32; ++ks)
					{
						const uint32_t atom = ks >> 2, koff = (ks & 3) * 32;
						const uint64_t ad = tc_smem_desc(a_base + atom * (kTcBM * 128) + koff);
						const uint64_t bd = tc_smem_desc(b_base + atom * (kTcBN * 128) + koff);
						tc_mma_i8(d_addr, ad, bd, idesc, ks > 0 ? 1u : 0u);
					}
					tc_commit(&empty[s]);  // smem stage reusable once these MMAs have read it
					tc_commit(&tfull[acc]); // accumulator stage complete
				}
			}
		}
	}
	else if (warp >= 4)
	{
		// The query tile goes from its packed form (32 B per descriptor) straight into the swizzled +-1 byte image in shared memory:
		// no expanded copy of the queries ever touches HBM.  One 16-byte chunk (16 descriptor bits) per thread and step.
		const int et = tid - 128;
		uint4 * sA4 = reinterpret_cast<uint4 *>(sA);
		constexpr int kItems = kTcBM * 16 / (128 * kTcEpiGroups); // chunks per thread: same chunk column, rows 32 apart
		constexpr int kRowStep = 128 * kTcEpiGroups / 16;
		const int chunk = et & 15, r0 = et >> 4;
		uint32_t w[kItems];
		auto load_packed = [&](int qtile) { // all loads in flight before the first use: one memory round trip
#pragma unroll
			for (int j = 0; j < kItems; ++j)
			{
				const int q = qtile * kTcBM + r0 + j * kRowStep;
				w[j] = q < nq ? __ldg(queries + static_cast<size_t>(q) * 8 + (chunk >> 1)) : 0u;
			}
		};
		auto expand_tile = [&](int qtile) {
#pragma unroll
			for (int j = 0; j < kItems; ++j)
			{
				const int r = r0 + j * kRowStep;
				uint4 v = make_uint4(0, 0, 0, 0);
				if (qtile * kTcBM + r < nq)
				{
					const uint32_t bits = (w[j] >> ((chunk & 1) * 16)) & 0xFFFFu;
					v.x = expand4(bits & 15u);
					v.y = expand4((bits >> 4) & 15u);
					v.z = expand4((bits >> 8) & 15u);
					v.w = expand4((bits >> 12) & 15u);
				}
				const int atom = chunk >> 3, c = chunk & 7;
				sA4[atom * (kTcBM * 8) + r * 8 + (c ^ (r & 7))] = v;
			}
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the tensor core's async proxy
			__syncwarp();
			if (lane == 0) mbar_arrive(afull);
		};
		const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32) (hardware: warp id % 4)
		const int cg = (warp - 4) >> 2;               // column group
		int g = 0;
		if (static_cast<int>(blockIdx.x) < n_items)
		{
			load_packed(static_cast<int>(blockIdx.x) % n_qtiles);
			expand_tile(static_cast<int>(blockIdx.x) % n_qtiles);
		}
		for (int item = blockIdx.x; item < n_items; item += gridDim.x)
		{
			int tile_begin;
			const int n_tiles = item_tiles(item, tile_begin);
			const int qtile = item % n_qtiles, split = item / n_qtiles;
			const int next = item + static_cast<int>(gridDim.x);
			const bool has_next = next < n_items;
			if (has_next) load_packed(next % n_qtiles); // in registers long before the boundary
			const int qi = qtile * kTcBM + quarter * 32 + lane;
			uint32_t k1 = kKeyNone, k2 = kKeyNone;
			int thr = -100000;                            // accumulator value a candidate has to exceed to enter the top-2
			for (int t = 0; t < n_tiles; ++t, ++g)
			{
				const int acc = g & 1;
				mbar_wait(&tfull[acc], (g >> 1) & 1);
				tc_fence_after();
				// last tile of the item: every MMA that reads the query tile has completed -> write the next item's tile now, so that its
				// first MMAs run while this accumulator is read out
				if (t == n_tiles - 1 && has_next) expand_tile(next % n_qtiles);
				const int row0 = (tile_begin + t) * kTcBN + cg * kTcEpiCols;
				const int valid = n_rows - row0;              // columns of this group that are real words (may be <= 0)
				const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * kTcBN + cg * kTcEpiCols);
				int v[kTcEpiCols];
#pragma unroll
				for (int c0 = 0; c0 < kTcEpiCols; c0 += 32) tc_ld32(taddr + c0, *reinterpret_cast<int(*)[32]>(&v[c0]));
				tc_wait_ld();
				// the accumulator stage is free as soon as the values are in registers
				tc_fence_before();
				__syncwarp();
				if (lane == 0) mbar_arrive(&tempty[acc]);
				if (valid >= kTcEpiCols)
				{
					// Branch-free top-2 over the 64 columns with 16-bit keys, two columns per register:
					//   key16 = distance * 64 + j = acc * (-32) + (8192 + j)   (distance = (256 - acc) / 2 <= 256, j < 32)
					// The low half-word carries column j, the high half-word column j + 32; S = acc_hi * 65536 + acc_lo and
					// K = S * (-32) + (8192 + j) * 65537 is that pair of keys exactly (every field stays in [0, 65536), so no
					// carry crosses the halves).  Two IMADs (FMA pipe) and three VIMNMX.U16x2 (ALU pipe) per column pair.
					uint32_t p1 = 0xFFFFFFFFu, p2 = 0xFFFFFFFFu;
#pragma unroll
					for (int j = 0; j < 32; ++j)
					{
						const uint32_t sp = static_cast<uint32_t>(v[j + 32]) * 65536u + static_cast<uint32_t>(v[j]);
						const uint32_t kp = sp * neg32 + static_cast<uint32_t>(8192 + j) * 65537u;
						const uint32_t mx = __vmaxu2(p1, kp);
						p1 = __vminu2(p1, kp);
						p2 = __vminu2(p2, mx);
					}
					// the four survivors (two per half) go into the running packed keys only if they can improve them
					const uint32_t best16 = min(p1 & 0xFFFFu, p1 >> 16);
					if (static_cast<int>(best16 >> 6) * 2 < kTcK - thr)
					{
						const uint32_t base = static_cast<uint32_t>(row_offset + row0);
						const uint32_t c[4] = {p1 & 0xFFFFu, p2 & 0xFFFFu, p1 >> 16, p2 >> 16};
#pragma unroll
						for (int u = 0; u < 4; ++u)
						{
							const uint32_t key = ((c[u] >> 6) << kKeyShift) + base + (c[u] & 63u) + (u >= 2 ? 32u : 0u);
							top2_insert(k1, k2, key);
						}
						thr = k2 == kKeyNone ? -100000 : kTcK - 2 * static_cast<int>(k2 >> kKeyShift);
					}
				}
				else if (valid > 0)
				{
					// last, partly filled tile: plain insertion of the real columns
#pragma unroll
					for (int j = 0; j < kTcEpiCols; ++j)
					{
						if (j < valid)
						{
							const uint32_t key = (static_cast<uint32_t>(kTcK - v[j]) << (kKeyShift - 1)) + static_cast<uint32_t>(row_offset + row0 + j);
							top2_insert(k1, k2, key);
						}
					}
					thr = k2 == kKeyNone ? -100000 : kTcK - 2 * static_cast<int>(k2 >> kKeyShift);
				}
				__syncwarp();
			}
			if (n_tiles == 0 && has_next) expand_tile(next % n_qtiles); // (an item without tiles: nothing reads the query tile)
			if (qi < nq) partial[(static_cast<size_t>(split) * kTcEpiGroups + cg) * nq + qi] = make_uint2(k1, k2);
		}
	}

	tc_fence_before();
	__syncthreads();
	if (warp == 2) tc_dealloc(tmem_base, 512);
}

} // namespace lcd
