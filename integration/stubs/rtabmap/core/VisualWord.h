// rtabmap::VisualWord — the members the shims use (corelib/include/rtabmap/core/VisualWord.h:40-65).
#pragma once
#include <map>
#include <opencv2/core.hpp>
namespace rtabmap {
class VisualWord
{
public:
	VisualWord(int id, const cv::Mat & descriptor, int signatureId = 0) : _id(id), _descriptor(descriptor.clone()) { if (signatureId) addRef(signatureId); } // :41
	void addRef(int signatureId) { ++_references[signatureId]; }                                                   // :44
	int removeAllRef(int signatureId) { auto i = _references.find(signatureId); int n = 0; if (i != _references.end()) { n = i->second; _references.erase(i); } return n; } // :45
	int id() const { return _id; }                                                                                   // :49
	const cv::Mat & getDescriptor() const { return _descriptor; }                                                    // :50
	const std::map<int, int> & getReferences() const { return _references; }                                         // :51
private:
	int _id;
	cv::Mat _descriptor;
	std::map<int, int> _references;
};
} // namespace rtabmap
