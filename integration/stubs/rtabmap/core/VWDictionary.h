// rtabmap::VWDictionary — the interface the B200 shim derives from (corelib/include/rtabmap/core/VWDictionary.h:46-160), with the
// header edits INTEGRATION.md §1 lists applied and marked.  Bodies are trivial stand-ins: only the declarations matter here.
#pragma once
#include <list>
#include <map>
#include <set>
#include <vector>
#include <opencv2/core.hpp>
#include <rtabmap/core/Parameters.h>
#include <rtabmap/core/VisualWord.h>
namespace rtabmap {
class VWDictionary
{
public:
	VWDictionary(const ParametersMap & parameters = ParametersMap()) { parseParameters(parameters); }   // :77
	virtual ~VWDictionary() {}                                                                          // :78
	virtual void parseParameters(const ParametersMap &) {}                                              // :80
	virtual void update() {}                                                                            // :82
	virtual std::list<int> addNewWords(const cv::Mat &, int) { return std::list<int>(); }               // :84-86
	virtual void addWord(VisualWord * vw);                                                              // :87
	/* EDIT 1 */ virtual std::vector<int> findNN(const cv::Mat &) const { return std::vector<int>(); } // :90  (+virtual)
	/* EDIT 2 */ virtual bool addWordRef(int wordId, int signatureId);                                  // :92  (+virtual)
	/* EDIT 3 */ virtual void removeAllWordRef(int wordId, int signatureId);                            // :93  (+virtual)
	const VisualWord * getWord(int id) const { auto i = _visualWords.find(id); return i == _visualWords.end() ? 0 : i->second; } // :94
	void setLastWordId(int id) { _lastWordId = id; }                                                    // :96
	const std::map<int, VisualWord *> & getVisualWords() const { return _visualWords; }                 // :97
	float getNndrRatio() const { return _nndrRatio; }                                                   // :98
	bool isIncremental() const { return _incrementalDictionary; }                                       // :106
	/* EDIT 4 */ virtual void clear(bool printWarningsIfNotEmpty = true);                               // :117 (+virtual)
	/* EDIT 5 */ virtual void removeWords(const std::vector<VisualWord *> & words);                     // :121 (+virtual)
protected:
	int getNextId() { return ++_lastWordId; }                                                           // :129
	std::map<int, VisualWord *> _visualWords;                                                           // :132
	int _totalActiveReferences = 0;                                                                     // :133
	/* EDIT 6: `private:` at :135 becomes `protected:` — the shim keeps these containers as the host-side truth */
	bool _incrementalDictionary = true;                                                                 // :136
	float _nndrRatio = 0.8f;                                                                            // :140
	bool _newWordsComparedTogether = true;                                                              // :143
	int _lastWordId = 0;                                                                                // :145
	std::map<int, VisualWord *> _unusedWords;                                                           // :153
	std::set<int> _notIndexedWords;                                                                     // :154
	std::set<int> _removedIndexedWords;                                                                 // :155
};
// stand-ins of VWDictionary.cpp:880-911, :1554-1572, :1595-1607, :842-873 (host bookkeeping only)
inline void VWDictionary::addWord(VisualWord * vw) { if (!vw) return; _visualWords[vw->id()] = vw; _notIndexedWords.insert(vw->id()); if (vw->getReferences().empty()) _unusedWords[vw->id()] = vw; else for (auto & r : vw->getReferences()) _totalActiveReferences += r.second; if (_lastWordId < vw->id()) _lastWordId = vw->id(); }
inline bool VWDictionary::addWordRef(int wordId, int signatureId) { auto i = _visualWords.find(wordId); if (i == _visualWords.end()) return false; i->second->addRef(signatureId); ++_totalActiveReferences; _unusedWords.erase(wordId); return true; }
inline void VWDictionary::removeAllWordRef(int wordId, int signatureId) { auto i = _visualWords.find(wordId); if (i == _visualWords.end()) return; _totalActiveReferences -= i->second->removeAllRef(signatureId); if (i->second->getReferences().empty()) _unusedWords[wordId] = i->second; }
inline void VWDictionary::removeWords(const std::vector<VisualWord *> & words) { for (VisualWord * w : words) { _visualWords.erase(w->id()); _unusedWords.erase(w->id()); if (_notIndexedWords.erase(w->id()) == 0) _removedIndexedWords.insert(w->id()); } }
inline void VWDictionary::clear(bool) { _visualWords.clear(); _unusedWords.clear(); _notIndexedWords.clear(); _removedIndexedWords.clear(); _totalActiveReferences = 0; _lastWordId = 0; }
} // namespace rtabmap
