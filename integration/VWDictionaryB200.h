// VWDictionaryB200 — rtabmap::VWDictionary backed by liblcd_b200.so (include/lcd_b200.h).
//
// Drop-in for corelib/: Memory news its dictionary at corelib/src/Memory.cpp:144 (`_vwd = new VWDictionary(parameters)`); with
// Kp/NNStrategy=5 that line becomes `new VWDictionaryB200(parameters)`.  The host-side containers of the base class (VisualWord
// objects, their reference maps, _unusedWords, _notIndexedWords) stay the truth Memory / DBDriver read; every distance, NNDR
// decision, posting-list update and TF-IDF score is computed on the B200.  Reference-side edits this class relies on are listed
// in INTEGRATION.md §1 (five `virtual` keywords and one `private:` -> `protected:` in VWDictionary.h).
#pragma once
#include <rtabmap/core/VWDictionary.h>
#include "lcd_b200.h"

namespace rtabmap {

class VWDictionaryB200 : public VWDictionary
{
public:
	explicit VWDictionaryB200(const ParametersMap & parameters = ParametersMap(), int device = 0);
	~VWDictionaryB200() override;

	void update() override;                                                        // VWDictionary.cpp:475-701
	std::list<int> addNewWords(const cv::Mat & descriptors, int signatureId) override; // VWDictionary.cpp:913-1229
	void addWord(VisualWord * vw) override;                                        // VWDictionary.cpp:1554-1572
	std::vector<int> findNN(const cv::Mat & descriptors) const override;           // VWDictionary.cpp:1273-1552
	bool addWordRef(int wordId, int signatureId) override;                         // VWDictionary.cpp:880-897
	void removeAllWordRef(int wordId, int signatureId) override;                   // VWDictionary.cpp:899-911
	void removeWords(const std::vector<VisualWord *> & words) override;            // VWDictionary.cpp:1582-1593
	void clear(bool printWarningsIfNotEmpty = true) override;                      // VWDictionary.cpp:842-873

	// the TF-IDF branch of Memory::computeLikelihood (Memory.cpp:2215-2291): words = uKeys(signature->getWords()), ids = the signatures to score
	std::map<int, float> computeLikelihood(const std::vector<int> & words, const std::list<int> & ids, int nSignaturesInMemory) const;

	lcd_engine * engine() const { return e_; }
	bool ok() const { return e_ != 0; }

private:
	bool ensureEngine(const cv::Mat & descriptors) const;   // the descriptor type is only known with the first descriptors (VWDictionary.cpp:933-957)
	mutable lcd_engine * e_;
	int device_;
};

} // namespace rtabmap
