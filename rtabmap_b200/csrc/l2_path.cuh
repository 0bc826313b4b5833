// l2_path.cuh — the quantiser for FLOAT descriptors (SURF-64 / SURF-128 / SIFT, LCD_DESC_F32): exact squared-L2 2-NN and the
// NNDR / new-word loop, kept apart from the binary path (nn_tensor.cuh / nn_hamming.cuh / resolve.cuh), which it does not touch.
//
// Replaces, for CV_32F descriptors: FlannIndex::knnSearch on a LinearIndex (corelib/src/FlannIndex.cpp:701-745 ->
// rtflann/algorithms/linear_index.h:129-146, rtflann::L2<float>::operator() rtflann/algorithms/dist.h:133-180,
// KNNSimpleResultSet::addPoint rtflann/util/result_set.h:151-172) and the "Process results" loop of
// VWDictionary::addNewWords / findNN (corelib/src/VWDictionary.cpp:1088-1219, :1476-1546).
//
// Bit-exactness: rtflann's L2 sums the squared differences in groups of four (each group left to right, then added to the
// running total); l2_rtflann below performs the same additions in the same order without FMA contraction, so distances — and
// with them every tie-break and every NNDR decision — equal the CPU's.  Keys are 64-bit: (float bits of the distance << 32) |
// row; distances are >= 0, so unsigned order of the key is (distance, lowest row) order.
//
// This path is exact, not fast: a CUDA-core kernel (one query per thread, rows staged through shared memory) and a per-frame
// sequential replay of the new-word loop with the candidate scan spread over the CTA.  The float GEMM formulation on the
// tensor cores (BASELINE configs[3]) cannot reproduce fp32 summation order and is left for a filter + exact
// re-check design.
#pragma once
#include "common.cuh"
#include "resolve.cuh"

namespace lcd {

constexpr unsigned long long kKey64None = ~0ull;
constexpr int kL2Threads = 128;   // queries per CTA of the 2-NN kernel
constexpr int kL2TileRows = 32;   // vocabulary rows staged per shared-memory tile
constexpr int kL2ResolveThreads = 256;

__device__ __forceinline__ unsigned long long pack64(float d, uint32_t row)
{
	return (static_cast<unsigned long long>(__float_as_uint(d)) << 32) | row;
}
__device__ __forceinline__ float key64_dist(unsigned long long k) { return __uint_as_float(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ uint32_t key64_row(unsigned long long k) { return static_cast<uint32_t>(k); }
__device__ __forceinline__ void top2_insert64(unsigned long long & k1, unsigned long long & k2, unsigned long long key)
{
	const unsigned long long m = max(k1, key);
	k1 = min(k1, key);
	k2 = min(k2, m);
}

// rtflann::L2<float>::operator()(a, b, size): groups of four, each summed left to right, then added to the total
template <int DIM, class PA, class PB>
__device__ __forceinline__ float l2_rtflann(PA a, PB b)
{
	float result = 0.0f;
#pragma unroll
	for (int i = 0; i < DIM; i += 4)
	{
		const float d0 = __fsub_rn(a[i], b[i]), d1 = __fsub_rn(a[i + 1], b[i + 1]), d2 = __fsub_rn(a[i + 2], b[i + 2]),
		            d3 = __fsub_rn(a[i + 3], b[i + 3]);
		float g = __fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1));
		g = __fadd_rn(g, __fmul_rn(d2, d2));
		g = __fadd_rn(g, __fmul_rn(d3, d3));
		result = __fadd_rn(result, g);
	}
	return result;
}

// partial[split * nq + query] = (best key, second key) over the rows [split * rows_per_split, ...) of the vocabulary
template <int DIM>
__global__ void __launch_bounds__(kL2Threads)
knn2_l2_kernel(const float * __restrict__ vocab, int n_rows, int row_offset, const float * __restrict__ queries, int nq,
               ulonglong2 * __restrict__ partial, int rows_per_split)
{
	__shared__ __align__(16) float s_rows[kL2TileRows * DIM];
	const int tid = threadIdx.x;
	const int qi = blockIdx.x * kL2Threads + tid;
	const int r_begin = blockIdx.y * rows_per_split, r_end = min(n_rows, r_begin + rows_per_split);
	float q[DIM];
#pragma unroll
	for (int i = 0; i < DIM; ++i) q[i] = qi < nq ? queries[static_cast<size_t>(qi) * DIM + i] : 0.0f;
	unsigned long long k1 = kKey64None, k2 = kKey64None;
	for (int r0 = r_begin; r0 < r_end; r0 += kL2TileRows)
	{
		const int nr = min(kL2TileRows, r_end - r0);
		__syncthreads();
		for (int i = tid; i < nr * DIM; i += kL2Threads) s_rows[i] = vocab[static_cast<size_t>(r0) * DIM + i];
		__syncthreads();
		for (int r = 0; r < nr; ++r)
		{
			const float d = l2_rtflann<DIM>(q, s_rows + r * DIM);
			top2_insert64(k1, k2, pack64(d, static_cast<uint32_t>(row_offset + r0 + r)));
		}
	}
	if (qi < nq) partial[static_cast<size_t>(blockIdx.y) * nq + qi] = make_ulonglong2(k1, k2);
}

// lcd_dict_knn2 for float descriptors: merge the splits, decode to (word id, squared distance)
__global__ void knn2_l2_decode_kernel(const ulonglong2 * __restrict__ partial, int n_splits, int nq, const int * __restrict__ row_ids,
                                      int * __restrict__ id1, float * __restrict__ d1, int * __restrict__ id2, float * __restrict__ d2)
{
	const int qi = blockIdx.x * blockDim.x + threadIdx.x;
	if (qi >= nq) return;
	unsigned long long k1 = kKey64None, k2 = kKey64None;
	for (int c = 0; c < n_splits; ++c)
	{
		const ulonglong2 p = partial[static_cast<size_t>(c) * nq + qi];
		top2_insert64(k1, k2, p.x);
		top2_insert64(k1, k2, p.y);
	}
	id1[qi] = k1 == kKey64None ? 0 : row_ids[key64_row(k1)];
	d1[qi] = k1 == kKey64None ? -1.0f : key64_dist(k1);
	id2[qi] = k2 == kKey64None ? 0 : row_ids[key64_row(k2)];
	d2[qi] = k2 == kKey64None ? -1.0f : key64_dist(k2);
}

// The decision on the multimap<float,int> fullResults of one descriptor (VWDictionary.cpp:1162-1219): candidates in insertion
// order a1, a2 (index hits), n1, n2 (hits among this frame's new words); strict '<' keeps the earlier entry on ties.
__device__ __forceinline__ bool nndr_decide64(unsigned long long a1, unsigned long long a2, unsigned long long n1, unsigned long long n2,
                                              float nndr, int & best_tag)
{
	float bd = INFINITY, sd = INFINITY;
	int bt = -1, cnt = 0;
	const unsigned long long keys[4] = {a1, a2, n1, n2};
#pragma unroll
	for (int t = 0; t < 4; ++t)
	{
		if (keys[t] != kKey64None)
		{
			const float d = key64_dist(keys[t]);
			++cnt;
			if (d < bd)
			{
				sd = bd;
				bd = d;
				bt = t;
			}
			else if (d < sd) sd = d;
		}
	}
	best_tag = bt;
	return cnt < 2 || bd > __fmul_rn(nndr, sd);
}

// One CTA per frame: the reference's sequential loop, descriptor by descriptor; the scan of the words this frame has created
// so far is spread over the CTA.  Dynamic shared memory: nq_pad uint32 (TF-IDF preparation buffer).
template <int DIM>
__global__ void __launch_bounds__(kL2ResolveThreads)
resolve_l2_kernel(const ResolveArgs a, const ulonglong2 * __restrict__ partial64)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	uint32_t * sbuf = reinterpret_cast<uint32_t *>(smem_raw);
	__shared__ float s_q[DIM];
	__shared__ unsigned long long s_k1[kL2ResolveThreads / 32], s_k2[kL2ResolveThreads / 32];
	__shared__ int s_nL, s_wid;
	__shared__ uint16_t s_new[kMaxFrameQueries]; // descriptor index of the k-th word created by this frame

	const int cap = a.nq;
	int nq_pad = 32;
	while (nq_pad < cap) nq_pad <<= 1;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int frame = blockIdx.x;
	const int nq = a.nq_frame ? min(max(a.nq_frame[frame], 0), cap) : cap;
	const float * fq = reinterpret_cast<const float *>(a.queries) + static_cast<size_t>(frame) * cap * DIM;
	const size_t pbase = static_cast<size_t>(frame) * cap;
	if (tid == 0) s_nL = 0;
	__syncthreads();

	for (int i = 0; i < nq; ++i)
	{
		for (int v = tid; v < DIM; v += blockDim.x) s_q[v] = fq[static_cast<size_t>(i) * DIM + v];
		__syncthreads();
		const int nL = s_nL;
		// hits among the words created by descriptors 0..i-1 of this frame (cv::BFMatcher::knnMatch on newWords)
		unsigned long long k1 = kKey64None, k2 = kKey64None;
		if (a.incremental && a.cmp_new)
		{
			for (int k = tid; k < nL; k += blockDim.x)
			{
				const float d = l2_rtflann<DIM>(s_q, fq + static_cast<size_t>(s_new[k]) * DIM);
				top2_insert64(k1, k2, pack64(d, static_cast<uint32_t>(k)));
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1)
			{
				const unsigned long long o1 = __shfl_down_sync(0xFFFFFFFFu, k1, o);
				const unsigned long long o2 = __shfl_down_sync(0xFFFFFFFFu, k2, o);
				top2_insert64(k1, k2, o1);
				top2_insert64(k1, k2, o2);
			}
			if (lane == 0)
			{
				s_k1[warp] = k1;
				s_k2[warp] = k2;
			}
		}
		__syncthreads();
		if (tid == 0)
		{
			unsigned long long n1 = kKey64None, n2 = kKey64None;
			if (a.incremental && a.cmp_new)
			{
				for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w)
				{
					top2_insert64(n1, n2, s_k1[w]);
					top2_insert64(n1, n2, s_k2[w]);
				}
			}
			unsigned long long a1 = kKey64None, a2 = kKey64None;
			for (int c = 0; c < a.n_chunks; ++c)
			{
				const ulonglong2 p = partial64[static_cast<size_t>(c) * a.nq_total + pbase + i];
				top2_insert64(a1, a2, p.x);
				top2_insert64(a1, a2, p.y);
			}
			int wid = 0;
			uint32_t sv = kSortNone;
			if (a.incremental)
			{
				int bt;
				const bool bad = nndr_decide64(a1, a2, n1, n2, a.nndr, bt);
				if (bad)
				{
					if (!a.find_only)
					{
						wid = a.last_word_id + 1 + nL;
						s_new[nL] = static_cast<uint16_t>(i);
						s_nL = nL + 1;
					}
				}
				else if (bt >= 2) wid = a.last_word_id + 1 + static_cast<int>(key64_row(bt == 2 ? n1 : n2));
				else
				{
					wid = a.row_ids[key64_row(bt == 0 ? a1 : a2)];
					sv = static_cast<uint32_t>(wid);
				}
			}
			else if (a1 != kKey64None)
			{
				// fixed dictionary: nearest word, no NNDR (VWDictionary.cpp:1211-1218)
				wid = a.row_ids[key64_row(a1)];
				sv = static_cast<uint32_t>(wid);
			}
			if (a.word_ids_out) a.word_ids_out[pbase + i] = wid;
			sbuf[i] = sv;
		}
		__syncthreads();
	}
	const int n_new = s_nL;
	for (int i = nq + tid; i < nq_pad; i += blockDim.x)
	{
		sbuf[i] = kSortNone;
		if (i < cap && a.word_ids_out) a.word_ids_out[pbase + i] = 0; // padding rows of a short frame
	}
	// commit the created words into the not-indexed tail of the vocabulary (single-frame quantise)
	if (a.pending_desc && !a.find_only)
	{
		float * dst = reinterpret_cast<float *>(a.pending_desc);
		for (int k = warp; k < n_new; k += static_cast<int>(blockDim.x >> 5))
		{
			for (int v = lane; v < DIM; v += 32) dst[static_cast<size_t>(k) * DIM + v] = fq[static_cast<size_t>(s_new[k]) * DIM + v];
			if (lane == 0) a.pending_ids[k] = a.last_word_id + 1 + k;
		}
	}
	if (tid == 0 && a.n_new_out) a.n_new_out[frame] = n_new;
	__syncthreads();
	if (a.do_prep) score_prep(sbuf, nq_pad, cap, a, frame);
}

} // namespace lcd
