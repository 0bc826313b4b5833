// ref_flann.cpp — thin C entry points around the REFERENCE'S OWN vendored rtflann
// (compiled in place from /root/reference/corelib/src/rtflann by oracle/Makefile into
// oracle/_ref/libref_flann.so; no reference source is copied into this repository).
//
// TEST INFRASTRUCTURE ONLY.  It performs exactly the calls FlannIndex makes for
// Kp/NNStrategy=0 (corelib/src/FlannIndex.cpp:312-335 build, :718-743 knnSearch):
//   rtflann::Index<Hamming<unsigned char>> / Index<L2<float>> with LinearIndexParams,
//   knnSearch(query, indices, dists, 2, SearchParams(checks=32, eps=0, sorted=true)).
// Used to pin oracle.cpp's restated linear scan (tie-breaking and float summation order)
// and, optionally, as the "reference" CPU baseline of bench.py.
#include "rtflann/flann.hpp"
#include <cstdint>
#include <vector>

extern "C" {

// idx_out[2*nq] (-1 = none), dist_out[2*nq]
int ref_flann_knn2_hamming(const unsigned char * data, int rows, int dim_bytes, const unsigned char * queries, int nq,
                           long long * idx_out, float * dist_out)
{
	rtflann::Matrix<unsigned char> dataset(const_cast<unsigned char *>(data), rows, dim_bytes);
	rtflann::Index<rtflann::Hamming<unsigned char>> index(dataset, rtflann::LinearIndexParams());
	index.buildIndex();
	rtflann::Matrix<unsigned char> q(const_cast<unsigned char *>(queries), nq, dim_bytes);
	std::vector<size_t> indices(2 * (size_t)nq, (size_t)-1);
	std::vector<unsigned int> dists(2 * (size_t)nq, 0u);
	rtflann::Matrix<size_t> mi(indices.data(), nq, 2);
	rtflann::Matrix<unsigned int> md(dists.data(), nq, 2);
	index.knnSearch(q, mi, md, 2, rtflann::SearchParams(32, 0, true));
	for (size_t i = 0; i < 2 * (size_t)nq; ++i)
	{
		idx_out[i] = indices[i] == (size_t)-1 ? -1 : (long long)indices[i];
		dist_out[i] = (float)dists[i];
	}
	return 0;
}

int ref_flann_knn2_l2(const float * data, int rows, int dim, const float * queries, int nq, long long * idx_out, float * dist_out)
{
	rtflann::Matrix<float> dataset(const_cast<float *>(data), rows, dim);
	rtflann::Index<rtflann::L2<float>> index(dataset, rtflann::LinearIndexParams());
	index.buildIndex();
	rtflann::Matrix<float> q(const_cast<float *>(queries), nq, dim);
	std::vector<size_t> indices(2 * (size_t)nq, (size_t)-1);
	std::vector<float> dists(2 * (size_t)nq, 0.f);
	rtflann::Matrix<size_t> mi(indices.data(), nq, 2);
	rtflann::Matrix<float> md(dists.data(), nq, 2);
	index.knnSearch(q, mi, md, 2, rtflann::SearchParams(32, 0, true));
	for (size_t i = 0; i < 2 * (size_t)nq; ++i)
	{
		idx_out[i] = indices[i] == (size_t)-1 ? -1 : (long long)indices[i];
		dist_out[i] = dists[i];
	}
	return 0;
}

} // extern "C"
