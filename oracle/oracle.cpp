// oracle.cpp — CPU restatement of the reference's quantise + score path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (rtabmap_b200/lib/liblcd_b200.so) never links, imports or calls anything in oracle/.
//
// What is restated, and from where (paths under /root/reference/corelib):
//   OracleDict::update          src/VWDictionary.cpp:475-701   (Kp/NNStrategy=0, incremental FLANN:
//                                                                rows appended in ascending id, removed
//                                                                rows skipped, insertion order kept)
//   OracleDict::knn2            src/FlannIndex.cpp:701-745 -> src/rtflann/algorithms/linear_index.h:129-146,
//                               src/rtflann/util/result_set.h:151-172 (KNNSimpleResultSet::addPoint),
//                               src/rtflann/algorithms/dist.h:533-580 (Hamming), :133-180 (L2)
//   OracleDict::add_new_words   src/VWDictionary.cpp:913-1229
//   OracleDict::find_nn         src/VWDictionary.cpp:1273-1552
//   add_word_ref/remove_all_ref src/VWDictionary.cpp:880-911, src/VisualWord.cpp:51-70
//   likelihood                  src/Memory.cpp:2215-2291, :4955-4968
//   adjust_likelihood           src/Rtabmap.cpp:5691-5760, utilite UMath.h:419-431, :512-526
//
// Pinning: knn2 is checked against the reference's own compiled rtflann (oracle/_ref, built from
// /root/reference by oracle/Makefile) in tests/test_oracle_ref.py; likelihood is checked against the
// reference's only golden vector for this path, archive/2010-LoopClosure/Tests/TestComputeLikelihood.m
// (tests/golden/tfidf_golden.json, made by tests/golden/make_tfidf_golden.py).  The NNDR / new-word
// loop and adjust_likelihood have no golden vector in the reference.  The loop is checked
// against an independent replay on the reference's own primitives (its compiled rtflann + cv::BFMatcher,
// tests/test_oracle_ref.py::test_quantiser_loop_against_reference_primitives); adjust_likelihood stays
// PARITY UNPINNED (oracle-vs-CUDA only).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <list>
#include <map>
#include <set>
#include <vector>

namespace {

enum { kU8 = 0, kF32 = 1 };

struct Word
{
	int id = 0;
	std::vector<uint8_t> desc;   // raw bytes (dim bytes for U8, dim*4 for F32)
	std::map<int, int> refs;     // signature id -> count (VisualWord::_references)
};

struct OracleDict
{
	int type = kU8;
	int dim = 32;
	bool incremental = true;
	float nndr = 0.8f;
	bool cmp_new = true;
	int last_id = 0;
	std::map<int, Word> words;          // _visualWords
	std::set<int> not_indexed;          // _notIndexedWords
	std::set<int> removed_indexed;      // _removedIndexedWords
	std::vector<int> rows;              // search order: word id per row (removed rows erased at update)
	long long total_refs = 0;           // _totalActiveReferences
	std::map<int, int> ni;              // signature id -> number of words (Memory::getNi)
	size_t row_bytes() const { return type == kU8 ? (size_t)dim : (size_t)dim * 4; }
};

// rtflann::Hamming: popcount of XOR over the descriptor bytes
inline float dist_hamming(const uint8_t * a, const uint8_t * b, int bytes)
{
	unsigned int d = 0;
	int i = 0;
	for (; i + 8 <= bytes; i += 8)
	{
		uint64_t x, y;
		memcpy(&x, a + i, 8);
		memcpy(&y, b + i, 8);
		d += (unsigned int)__builtin_popcountll(x ^ y);
	}
	for (; i < bytes; ++i) d += (unsigned int)__builtin_popcount((unsigned)(a[i] ^ b[i]));
	return (float)d;
}

// rtflann::L2<float>: groups of four, each group summed left to right and then added to the total.
// volatile stops the compiler from contracting a*b+c into an FMA or reassociating.
inline float dist_l2(const float * a, const float * b, int n)
{
	volatile float result = 0.0f;
	int i = 0;
	for (; i + 4 <= n; i += 4)
	{
		volatile float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1], d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
		volatile float s0 = d0 * d0, s1 = d1 * d1, s2 = d2 * d2, s3 = d3 * d3;
		volatile float g = s0 + s1;
		g = g + s2;
		g = g + s3;
		result = result + g;
	}
	for (; i < n; ++i)
	{
		volatile float d0 = a[i] - b[i];
		volatile float s0 = d0 * d0;
		result = result + s0;
	}
	return result;
}

inline float dist(const OracleDict & d, const uint8_t * a, const uint8_t * b)
{
	return d.type == kU8 ? dist_hamming(a, b, d.dim) : dist_l2((const float *)a, (const float *)b, d.dim);
}

// KNNSimpleResultSet with capacity 2, fed in ascending row order
struct Top2
{
	float d[2];
	int idx[2];
	int count = 0;
	float worst;
	Top2()
	{
		d[0] = d[1] = std::numeric_limits<float>::max();
		idx[0] = idx[1] = -1;
		worst = std::numeric_limits<float>::max();
	}
	void add(float dist, int index)
	{
		if (dist >= worst) return;
		if (count < 2) ++count;
		int i;
		for (i = count - 1; i > 0; --i)
		{
			if (d[i - 1] > dist)
			{
				d[i] = d[i - 1];
				idx[i] = idx[i - 1];
			}
			else break;
		}
		d[i] = dist;
		idx[i] = index;
		worst = d[1];
	}
};

void knn2_rows(const OracleDict & dd, const std::vector<const uint8_t *> & rowptr, const uint8_t * q, Top2 & out)
{
	for (size_t r = 0; r < rowptr.size(); ++r) out.add(dist(dd, rowptr[r], q), (int)r);
}

std::vector<const uint8_t *> indexed_ptrs(const OracleDict & d)
{
	std::vector<const uint8_t *> p;
	p.reserve(d.rows.size());
	for (int id : d.rows) p.push_back(d.words.at(id).desc.data());
	return p;
}

void add_word_ref(OracleDict & d, int word, int sig)
{
	auto it = d.words.find(word);
	if (it == d.words.end()) return;
	it->second.refs[sig] += 1;
	d.total_refs += 1;
}

void remove_all_ref(OracleDict & d, int word, int sig)
{
	auto it = d.words.find(word);
	if (it == d.words.end()) return;
	auto r = it->second.refs.find(sig);
	if (r != it->second.refs.end())
	{
		d.total_refs -= r->second;
		it->second.refs.erase(r);
	}
}

void update(OracleDict & d)
{
	if (!d.removed_indexed.empty())
	{
		std::vector<int> kept;
		kept.reserve(d.rows.size());
		for (int id : d.rows)
			if (!d.removed_indexed.count(id)) kept.push_back(id);
		d.rows.swap(kept);
	}
	for (int id : d.not_indexed) d.rows.push_back(id); // std::set: ascending id
	d.not_indexed.clear();
	d.removed_indexed.clear();
}

// The decision on one descriptor's multimap<float,int> fullResults
struct Decision
{
	bool bad = true;
	int best_id = 0;
};

Decision decide(const OracleDict & d, const std::multimap<float, int> & full)
{
	Decision r;
	if (d.incremental)
	{
		if (full.size() >= 2)
		{
			auto first = full.begin();
			auto second = first;
			++second;
			r.bad = first->first > d.nndr * second->first;
		}
		else r.bad = true;
		if (!r.bad) r.best_id = full.begin()->second;
	}
	else if (!full.empty())
	{
		r.bad = false;
		r.best_id = full.begin()->second;
	}
	return r;
}

int add_new_words(OracleDict & d, const uint8_t * desc, int n, int sig, int * out)
{
	if (n <= 0) return 0;
	if (!d.incremental && d.words.empty()) return 0;
	const size_t rb = d.row_bytes();
	const std::vector<const uint8_t *> rowptr = indexed_ptrs(d);
	std::vector<const uint8_t *> newptr; // descriptors of the words created by this call
	std::vector<int> newid;
	int nout = 0;
	for (int i = 0; i < n; ++i)
	{
		const uint8_t * q = desc + (size_t)i * rb;
		std::multimap<float, int> full;
		if (!rowptr.empty())
		{
			Top2 t;
			knn2_rows(d, rowptr, q, t);
			for (int j = 0; j < 2; ++j)
			{
				if (t.idx[j] < 0) break; // fewer than two indexed words
				full.insert(std::make_pair(t.d[j], d.rows[t.idx[j]]));
			}
		}
		if (d.cmp_new && !newptr.empty())
		{
			Top2 t; // cv::BFMatcher::knnMatch(k = rows>1 ? 2 : 1): same (distance, lowest index) order
			knn2_rows(d, newptr, q, t);
			for (int j = 0; j < 2; ++j)
			{
				if (t.idx[j] < 0) break;
				full.insert(std::make_pair(t.d[j], newid[t.idx[j]]));
			}
		}
		Decision r = decide(d, full);
		if (d.incremental && r.bad)
		{
			Word w;
			w.id = ++d.last_id;
			w.desc.assign(q, q + rb);
			if (sig) w.refs[sig] = 1; // VisualWord(id, descriptor, signatureId) adds the first reference
			d.words[w.id] = w;
			d.not_indexed.insert(w.id);
			newptr.push_back(d.words[w.id].desc.data());
			newid.push_back(w.id);
			out[nout++] = w.id;
		}
		else if (!r.bad)
		{
			if (sig) add_word_ref(d, r.best_id, sig);
			out[nout++] = r.best_id;
		}
	}
	d.total_refs += (long long)d.not_indexed.size(); // VWDictionary.cpp:1227
	// std::map nodes are stable, but newptr was taken from d.words after insertion, so it stays valid
	return nout;
}

void find_nn(const OracleDict & d, const uint8_t * desc, int n, int * out)
{
	const size_t rb = d.row_bytes();
	const std::vector<const uint8_t *> rowptr = indexed_ptrs(d);
	std::vector<const uint8_t *> niptr;
	std::vector<int> niid;
	for (int id : d.not_indexed)
	{
		niptr.push_back(d.words.at(id).desc.data());
		niid.push_back(id);
	}
	for (int i = 0; i < n; ++i)
	{
		out[i] = 0;
		if (d.words.empty()) continue;
		const uint8_t * q = desc + (size_t)i * rb;
		std::multimap<float, int> full;
		if (!rowptr.empty())
		{
			Top2 t;
			knn2_rows(d, rowptr, q, t);
			for (int j = 0; j < 2; ++j)
				if (t.idx[j] >= 0) full.insert(std::make_pair(t.d[j], d.rows[t.idx[j]]));
		}
		if (!niptr.empty())
		{
			Top2 t;
			knn2_rows(d, niptr, q, t);
			for (int j = 0; j < 2; ++j)
				if (t.idx[j] >= 0) full.insert(std::make_pair(t.d[j], niid[t.idx[j]]));
		}
		Decision r = decide(d, full);
		if (!r.bad) out[i] = r.best_id;
	}
}

// Memory::computeLikelihood, TF-IDF branch
void likelihood(const OracleDict & d, const int * qwords, int nq, const int * ids, int ns, int N_, float * out)
{
	std::map<int, float> lik;
	for (int k = 0; k < ns; ++k) lik.insert(lik.end(), std::make_pair(ids[k], 0.0f));
	std::set<int> uniq(qwords, qwords + nq); // uUniqueKeys: ascending unique keys
	float nwi, ni, nw, N, logNnw;
	N = (float)N_;
	if (N)
	{
		for (int w : uniq)
		{
			if (w <= 0) continue;
			auto it = d.words.find(w);
			if (it == d.words.end()) continue; // reference asserts; unknown words cannot occur in parity runs
			const std::map<int, int> & refs = it->second.refs;
			nw = (float)refs.size();
			if (nw)
			{
				logNnw = log10f(N / nw);
				if (logNnw)
				{
					for (auto & r : refs)
					{
						auto li = lik.find(r.first);
						if (li != lik.end())
						{
							nwi = (float)r.second;
							auto nit = d.ni.find(r.first);
							ni = nit != d.ni.end() ? (float)nit->second : 0.0f;
							if (ni != 0)
							{
								volatile float num = nwi * logNnw;
								volatile float term = num / ni;
								volatile float acc = li->second + term;
								li->second = acc;
							}
						}
					}
				}
			}
		}
	}
	for (int k = 0; k < ns; ++k) out[k] = lik[ids[k]];
}

} // namespace

extern "C" {

void * orc_create(int desc_type, int dim, int incremental, float nndr, int cmp_new)
{
	OracleDict * d = new OracleDict();
	d->type = desc_type;
	d->dim = dim;
	d->incremental = incremental != 0;
	d->nndr = nndr;
	d->cmp_new = cmp_new != 0;
	return d;
}
void orc_destroy(void * h) { delete (OracleDict *)h; }
void orc_set_params(void * h, int incremental, float nndr, int cmp_new)
{
	OracleDict * d = (OracleDict *)h;
	d->incremental = incremental != 0;
	d->nndr = nndr;
	d->cmp_new = cmp_new != 0;
}

// VWDictionary::addWord
int orc_add_words(void * h, const int * ids, const void * desc, int n)
{
	OracleDict * d = (OracleDict *)h;
	const size_t rb = d->row_bytes();
	for (int i = 0; i < n; ++i)
	{
		Word w;
		w.id = ids[i];
		const uint8_t * p = (const uint8_t *)desc + (size_t)i * rb;
		w.desc.assign(p, p + rb);
		d->words[w.id] = w;
		d->not_indexed.insert(w.id);
	}
	return 0;
}
// VWDictionary::removeWords
int orc_remove_words(void * h, const int * ids, int n)
{
	OracleDict * d = (OracleDict *)h;
	for (int i = 0; i < n; ++i)
	{
		auto it = d->words.find(ids[i]);
		if (it == d->words.end()) continue;
		if (!d->not_indexed.erase(ids[i])) d->removed_indexed.insert(ids[i]);
		for (auto & r : it->second.refs) d->total_refs -= r.second;
		d->words.erase(it);
	}
	return 0;
}
void orc_update(void * h) { update(*(OracleDict *)h); }
int orc_size(void * h) { return (int)((OracleDict *)h)->words.size(); }
int orc_indexed_size(void * h) { return (int)((OracleDict *)h)->rows.size(); }
int orc_not_indexed_size(void * h) { return (int)((OracleDict *)h)->not_indexed.size(); }
int orc_last_word_id(void * h) { return ((OracleDict *)h)->last_id; }
void orc_set_last_word_id(void * h, int id) { ((OracleDict *)h)->last_id = id; }
long long orc_total_refs(void * h) { return ((OracleDict *)h)->total_refs; }
int orc_get_indexed_ids(void * h, int * ids, int cap)
{
	OracleDict * d = (OracleDict *)h;
	int n = std::min(cap, (int)d->rows.size());
	for (int i = 0; i < n; ++i) ids[i] = d->rows[i];
	return n;
}

// FlannIndex::knnSearch(k=2): ids (0 = none) and distances (-1 = none)
void orc_knn2(void * h, const void * queries, int nq, int * id1, float * d1, int * id2, float * d2)
{
	OracleDict * d = (OracleDict *)h;
	const std::vector<const uint8_t *> rowptr = indexed_ptrs(*d);
	const size_t rb = d->row_bytes();
	for (int i = 0; i < nq; ++i)
	{
		Top2 t;
		knn2_rows(*d, rowptr, (const uint8_t *)queries + (size_t)i * rb, t);
		id1[i] = t.idx[0] >= 0 ? d->rows[t.idx[0]] : 0;
		d1[i] = t.idx[0] >= 0 ? t.d[0] : -1.0f;
		id2[i] = t.idx[1] >= 0 ? d->rows[t.idx[1]] : 0;
		d2[i] = t.idx[1] >= 0 ? t.d[1] : -1.0f;
	}
}

// raw 2-NN over an explicit matrix (row indices), for checking against oracle/_ref
void orc_knn2_raw(int desc_type, int dim, const void * data, int rows, const void * queries, int nq, int * idx, float * dist_out)
{
	OracleDict d;
	d.type = desc_type;
	d.dim = dim;
	const size_t rb = d.row_bytes();
	std::vector<const uint8_t *> rowptr(rows);
	for (int r = 0; r < rows; ++r) rowptr[r] = (const uint8_t *)data + (size_t)r * rb;
	for (int i = 0; i < nq; ++i)
	{
		Top2 t;
		knn2_rows(d, rowptr, (const uint8_t *)queries + (size_t)i * rb, t);
		idx[2 * i] = t.idx[0];
		idx[2 * i + 1] = t.idx[1];
		dist_out[2 * i] = t.idx[0] >= 0 ? t.d[0] : -1.0f;
		dist_out[2 * i + 1] = t.idx[1] >= 0 ? t.d[1] : -1.0f;
	}
}

// VWDictionary::addNewWords; returns the number of ids written (== n except in the error cases)
int orc_add_new_words(void * h, const void * desc, int n, int sig_id, int * out_ids)
{
	OracleDict * d = (OracleDict *)h;
	int r = add_new_words(*d, (const uint8_t *)desc, n, sig_id, out_ids);
	if (sig_id > 0) d->ni[sig_id] += n;
	return r;
}
void orc_find_nn(void * h, const void * desc, int n, int * out_ids) { find_nn(*(OracleDict *)h, (const uint8_t *)desc, n, out_ids); }

void orc_add_refs(void * h, int sig_id, const int * word_ids, int n)
{
	OracleDict * d = (OracleDict *)h;
	for (int i = 0; i < n; ++i)
		if (word_ids[i] > 0) add_word_ref(*d, word_ids[i], sig_id);
	d->ni[sig_id] += n;
}
void orc_remove_sig(void * h, int sig_id)
{
	OracleDict * d = (OracleDict *)h;
	for (auto & kv : d->words) remove_all_ref(*d, kv.first, sig_id);
	d->ni.erase(sig_id);
}
void orc_set_ni(void * h, const int * sig_ids, const int * ni, int n)
{
	OracleDict * d = (OracleDict *)h;
	for (int i = 0; i < n; ++i) d->ni[sig_ids[i]] = ni[i];
}
// bulk load of references (word -> (sig,count)), CSR
void orc_load_csr(void * h, const int * word_ids, int nw, const int64_t * row_ptr, const int * sig, const int * cnt)
{
	OracleDict * d = (OracleDict *)h;
	for (int k = 0; k < nw; ++k)
	{
		Word & w = d->words.at(word_ids[k]);
		for (int64_t p = row_ptr[k]; p < row_ptr[k + 1]; ++p)
		{
			w.refs[sig[p]] += cnt[p];
			d->total_refs += cnt[p];
			d->ni[sig[p]] += cnt[p];
		}
	}
}
int orc_get_refs(void * h, int word_id, int * sig, int * cnt, int cap)
{
	OracleDict * d = (OracleDict *)h;
	auto it = d->words.find(word_id);
	if (it == d->words.end()) return 0;
	int k = 0;
	for (auto & r : it->second.refs)
	{
		if (k < cap)
		{
			sig[k] = r.first;
			cnt[k] = r.second;
		}
		++k;
	}
	return k;
}

void orc_likelihood(void * h, const int * qwords, int nq, const int * sig_ids, int ns, int n_total, float * out)
{
	likelihood(*(OracleDict *)h, qwords, nq, sig_ids, ns, n_total, out);
}

// One localisation query with roll-back: addNewWords(desc, sig) ->
// computeLikelihood(sig, ids) -> remove the query's references and the words it created.
int orc_localize(void * h, const void * desc, int n, int sig_id, const int * sig_ids, int ns, int n_total, int * out_words, float * out_like)
{
	OracleDict * d = (OracleDict *)h;
	const int last0 = d->last_id;
	const long long refs0 = d->total_refs;
	std::vector<int> ids(n);
	const int nout = add_new_words(*d, (const uint8_t *)desc, n, sig_id, ids.data());
	d->ni[sig_id] = n;
	if (out_like) likelihood(*d, ids.data(), nout, sig_ids, ns, n_total, out_like);
	if (out_words) memcpy(out_words, ids.data(), (size_t)nout * sizeof(int));
	std::set<int> uniq(ids.begin(), ids.begin() + nout);
	for (int w : uniq)
	{
		if (w > last0)
		{
			d->words.erase(w);
			d->not_indexed.erase(w);
		}
		else
		{
			auto it = d->words.find(w);
			if (it != d->words.end()) it->second.refs.erase(sig_id);
		}
	}
	d->last_id = last0;
	d->total_refs = refs0;
	d->ni.erase(sig_id);
	return nout;
}

// Read-only form of orc_localize for the multi-threaded CPU baseline: the same decisions and the
// same float arithmetic, but the words the frame would create live in local vectors and the query's
// own reference is accounted for as nw = refs.size() + 1 instead of being inserted and rolled back.
// tests/test_oracle_golden.py checks it against orc_localize.  Safe to call from many threads.
static int localize_ro_impl(void * h, const void * desc_, int n, const int * sig_ids, int ns, int n_total, int * out_words, float * out_like,
                            const long long * knn_idx, const float * knn_dist);

int orc_localize_ro(void * h, const void * desc_, int n, const int * sig_ids, int ns, int n_total, int * out_words, float * out_like)
{
	return localize_ro_impl(h, desc_, n, sig_ids, ns, n_total, out_words, out_like, nullptr, nullptr);
}

// Same, with the index search of every descriptor done by the caller — bench.py's CPU arm passes the output of the REFERENCE'S OWN
// compiled rtflann (oracle/_ref, FlannIndex::knnSearch on the LinearIndex): knn_idx[n][2] rows in search order (-1 = none),
// knn_dist[n][2].  Everything after the search (multimap, NNDR, new words, TF-IDF) is the restatement.
int orc_localize_ro_knn(void * h, const void * desc_, int n, const long long * knn_idx, const float * knn_dist, const int * sig_ids, int ns,
                        int n_total, int * out_words, float * out_like)
{
	return localize_ro_impl(h, desc_, n, sig_ids, ns, n_total, out_words, out_like, knn_idx, knn_dist);
}

static int localize_ro_impl(void * h, const void * desc_, int n, const int * sig_ids, int ns, int n_total, int * out_words, float * out_like,
                            const long long * knn_idx, const float * knn_dist)
{
	const OracleDict & d = *(const OracleDict *)h;
	const uint8_t * desc = (const uint8_t *)desc_;
	const size_t rb = d.row_bytes();
	const std::vector<const uint8_t *> rowptr = indexed_ptrs(d);
	std::vector<const uint8_t *> newptr;
	std::vector<int> newid;
	std::vector<int> ids;
	int next_id = d.last_id;
	for (int i = 0; i < n; ++i)
	{
		const uint8_t * q = desc + (size_t)i * rb;
		std::multimap<float, int> full;
		if (knn_idx)
		{
			for (int j = 0; j < 2; ++j)
			{
				const long long r = knn_idx[2 * (size_t)i + j];
				if (r < 0) break;
				full.insert(std::make_pair(knn_dist[2 * (size_t)i + j], d.rows[(size_t)r]));
			}
		}
		else if (!rowptr.empty())
		{
			Top2 t;
			knn2_rows(d, rowptr, q, t);
			for (int j = 0; j < 2; ++j)
			{
				if (t.idx[j] < 0) break;
				full.insert(std::make_pair(t.d[j], d.rows[t.idx[j]]));
			}
		}
		if (d.cmp_new && !newptr.empty())
		{
			Top2 t;
			knn2_rows(d, newptr, q, t);
			for (int j = 0; j < 2; ++j)
			{
				if (t.idx[j] < 0) break;
				full.insert(std::make_pair(t.d[j], newid[t.idx[j]]));
			}
		}
		Decision r = decide(d, full);
		if (d.incremental && r.bad)
		{
			newptr.push_back(q);
			newid.push_back(++next_id);
			ids.push_back(next_id);
		}
		else if (!r.bad) ids.push_back(r.best_id);
	}
	if (out_words) memcpy(out_words, ids.data(), ids.size() * sizeof(int));
	if (out_like)
	{
		std::map<int, float> lik;
		for (int k = 0; k < ns; ++k) lik.insert(lik.end(), std::make_pair(sig_ids[k], 0.0f));
		std::set<int> uniq(ids.begin(), ids.end());
		const float N = (float)n_total;
		for (int w : uniq)
		{
			if (w <= 0 || w > d.last_id) continue; // words created by this frame are referenced by it alone
			auto it = d.words.find(w);
			if (it == d.words.end()) continue;
			const std::map<int, int> & refs = it->second.refs;
			const float nw = (float)(refs.size() + 1);
			const float logNnw = log10f(N / nw);
			if (!logNnw) continue;
			for (auto & r : refs)
			{
				auto li = lik.find(r.first);
				if (li == lik.end()) continue;
				const float nwi = (float)r.second;
				auto nit = d.ni.find(r.first);
				const float ni = nit != d.ni.end() ? (float)nit->second : 0.0f;
				if (ni != 0)
				{
					volatile float num = nwi * logNnw;
					volatile float term = num / ni;
					volatile float acc = li->second + term;
					li->second = acc;
				}
			}
		}
		for (int k = 0; k < ns; ++k) out_like[k] = lik[sig_ids[k]];
	}
	return (int)ids.size();
}

// Rtabmap::adjustLikelihood with Rtabmap/VirtualPlaceLikelihoodRatio = 0 or 1; element 0 is the virtual place
void orc_adjust_likelihood(float * lik, int n, int virtual_place_ratio)
{
	if (n <= 0) return;
	std::list<float> values;
	for (int i = 1; i < n; ++i)
		if (lik[i] > 0) values.push_back(lik[i]);
	float mean = 0;
	if (!values.empty())
	{
		for (float v : values) mean += v;
		mean /= values.size();
	}
	float var = 0;
	if (values.size() > 1)
	{
		float sum = 0;
		for (float v : values) sum += (v - mean) * (v - mean);
		var = sum / (values.size() - 1);
	}
	float stdDev = std::sqrt(var);
	const float epsilon = 0.0001f;
	float max = 0.0f;
	for (int i = 1; i < n; ++i)
	{
		float value = lik[i];
		lik[i] = 1.0f;
		if (value > mean + stdDev)
		{
			if (virtual_place_ratio == 0 && mean) lik[i] = (value - (stdDev - epsilon)) / mean;
			else if (virtual_place_ratio != 0 && stdDev) lik[i] = (value - mean) / stdDev;
		}
		if (value > max) max = value;
	}
	if (virtual_place_ratio == 0 && stdDev > epsilon && max) lik[0] = mean / stdDev + 1.0f;
	else if (virtual_place_ratio != 0 && max > mean) lik[0] = stdDev / (max - mean) + 1.0f;
	else lik[0] = 2.0f;
}

} // extern "C"
