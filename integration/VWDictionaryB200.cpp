// VWDictionaryB200.cpp — see VWDictionaryB200.h.  Error convention of the reference: log with UERROR and return the empty container
// (VWDictionary.cpp:920-957); nothing throws across the C ABI.
#include "VWDictionaryB200.h"
#include <rtabmap/utilite/ULogger.h>

namespace rtabmap {

VWDictionaryB200::VWDictionaryB200(const ParametersMap & parameters, int device) : VWDictionary(parameters), e_(0), device_(device) {}

VWDictionaryB200::~VWDictionaryB200()
{
	lcd_destroy(e_);
}

bool VWDictionaryB200::ensureEngine(const cv::Mat & d) const
{
	if (e_) return true;
	lcd_config c;
	c.device = device_;
	c.desc_type = d.type() == CV_8U ? LCD_DESC_U8 : LCD_DESC_F32;
	c.desc_dim = d.cols;
	c.max_words = 1 << 16;
	c.max_signatures = 1 << 14;
	c.max_queries = 4096;
	c.max_batch = 1;
	e_ = lcd_create(&c);
	if (!e_) UERROR("B200 dictionary engine: %s", lcd_last_error(0));
	return e_ != 0;
}

void VWDictionaryB200::addWord(VisualWord * vw)
{
	if (!vw) return;
	VWDictionary::addWord(vw); // _visualWords, _notIndexedWords, _unusedWords, _totalActiveReferences, _lastWordId
	if (!ensureEngine(vw->getDescriptor())) return;
	const int id = vw->id();
	if (lcd_dict_add_words(e_, &id, vw->getDescriptor().data, 1) != LCD_OK) UERROR("%s", lcd_last_error(e_));
	for (std::map<int, int>::const_iterator r = vw->getReferences().begin(); r != vw->getReferences().end(); ++r)
	{
		std::vector<int> w(r->second, id); // a word loaded from the database arrives with its references (Memory.cpp:410-438)
		lcd_index_add_refs(e_, r->first, w.data(), (int)w.size());
	}
	lcd_dict_set_last_word_id(e_, _lastWordId);
}

void VWDictionaryB200::update()
{
	// the reference rebuilds / extends its search structure here; on the device that is a boundary bump (+ compaction of removed rows)
	if (e_ && lcd_dict_update(e_) != LCD_OK) UERROR("%s", lcd_last_error(e_));
	_notIndexedWords.clear();
	_removedIndexedWords.clear();
}

std::list<int> VWDictionaryB200::addNewWords(const cv::Mat & descriptorsIn, int signatureId)
{
	std::list<int> out;
	if (descriptorsIn.rows == 0 || descriptorsIn.cols == 0)
	{
		UERROR("Descriptors size is null!");
		return out;
	}
	const cv::Mat d = descriptorsIn.isContinuous() ? descriptorsIn : descriptorsIn.clone();
	if (!ensureEngine(d)) return out;
	const int before = _lastWordId;
	std::vector<int> ids(d.rows);
	int nNew = 0;
	// NN + NNDR + intra-frame new words + posting-list update for signatureId, all on the device
	if (lcd_dict_quantize(e_, d.data, d.rows, signatureId, isIncremental() ? 1 : 0, getNndrRatio(), _newWordsComparedTogether ? 1 : 0, ids.data(), &nNew) != LCD_OK)
	{
		UERROR("%s", lcd_last_error(e_));
		return out;
	}
	// mirror the decisions into the host containers Memory and DBDriver read
	for (int i = 0; i < d.rows; ++i)
	{
		const int id = ids[i];
		if (id > before && _visualWords.find(id) == _visualWords.end())
		{
			VisualWord * vw = new VisualWord(id, d.row(i), signatureId); // the first descriptor quantised to a new id IS the word
			_visualWords.insert(_visualWords.end(), std::make_pair(id, vw));
			_notIndexedWords.insert(_notIndexedWords.end(), id);
			_totalActiveReferences += 1;
		}
		else if (id > 0) VWDictionary::addWordRef(id, signatureId); // host bookkeeping only: the device reference is already in
		out.push_back(id);
	}
	_lastWordId = before + nNew;
	return out;
}

std::vector<int> VWDictionaryB200::findNN(const cv::Mat & descriptors) const
{
	std::vector<int> ids(descriptors.rows, 0);
	if (descriptors.rows == 0 || !ensureEngine(descriptors)) return ids;
	const cv::Mat d = descriptors.isContinuous() ? descriptors : descriptors.clone();
	if (lcd_dict_find_nn(e_, d.data, d.rows, isIncremental() ? 1 : 0, getNndrRatio(), ids.data()) != LCD_OK) UERROR("%s", lcd_last_error(e_));
	return ids;
}

bool VWDictionaryB200::addWordRef(int wordId, int signatureId)
{
	const bool found = VWDictionary::addWordRef(wordId, signatureId);
	if (found && e_) lcd_index_add_refs(e_, signatureId, &wordId, 1);
	return found;
}

void VWDictionaryB200::removeAllWordRef(int wordId, int signatureId)
{
	// Memory::disableWordsRef (Memory.cpp:6871-6897) calls this for every word of the signature: the device drops the whole
	// signature at the first call and ignores the rest
	VWDictionary::removeAllWordRef(wordId, signatureId);
	if (e_) lcd_index_remove_sig(e_, signatureId);
}

void VWDictionaryB200::removeWords(const std::vector<VisualWord *> & words)
{
	std::vector<int> ids;
	for (size_t i = 0; i < words.size(); ++i) ids.push_back(words[i]->id());
	VWDictionary::removeWords(words);
	if (e_ && !ids.empty()) lcd_dict_remove_words(e_, ids.data(), (int)ids.size());
}

void VWDictionaryB200::clear(bool printWarningsIfNotEmpty)
{
	VWDictionary::clear(printWarningsIfNotEmpty);
	if (e_) lcd_dict_clear(e_);
}

std::map<int, float> VWDictionaryB200::computeLikelihood(const std::vector<int> & words, const std::list<int> & ids, int nSignaturesInMemory) const
{
	std::map<int, float> likelihood;
	std::vector<int> idv(ids.begin(), ids.end());
	std::vector<float> l(idv.size(), 0.0f);
	if (e_ && !words.empty() && !idv.empty() &&
	    lcd_index_score(e_, words.data(), (int)words.size(), idv.data(), (int)idv.size(), nSignaturesInMemory, l.data()) != LCD_OK)
		UERROR("%s", lcd_last_error(e_));
	for (size_t i = 0; i < idv.size(); ++i) likelihood.insert(likelihood.end(), std::make_pair(idv[i], l[i]));
	return likelihood;
}

} // namespace rtabmap
