// Minimal stand-in for the parts of <opencv2/core.hpp> the shims use (cv::Mat header fields, KeyPoint, Point3f/2f, Rect, Size).
#pragma once
#include <cstddef>
#include <vector>
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
namespace cv {
struct Size { int width = 0, height = 0; };
struct Rect { int x = 0, y = 0, width = 0, height = 0; };
struct Point2f { float x = 0, y = 0; };
struct Point3f { float x = 0, y = 0, z = 0; };
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
class Mat
{
public:
	Mat() {}
	Mat(int r, int c, int t) : rows(r), cols(c), type_(t), store_(static_cast<size_t>(r) * c * elemSize1_(t)) { data = store_.data(); }
	Mat(int r, int c, int t, void * p) : rows(r), cols(c), data(static_cast<unsigned char *>(p)), type_(t) {}
	int rows = 0, cols = 0;
	unsigned char * data = nullptr;
	int type() const { return type_; }
	int channels() const { return 1; }
	bool empty() const { return rows == 0 || cols == 0; }
	bool isContinuous() const { return true; }
	Mat clone() const { Mat m(rows, cols, type_); for (size_t i = 0; i < m.store_.size(); ++i) m.store_[i] = data[i]; return m; }
	Mat row(int i) const { return Mat(1, cols, type_, data + static_cast<size_t>(i) * cols * elemSize1_(type_)); }
	Mat rowRange(int a, int b) const { return Mat(b - a, cols, type_, data + static_cast<size_t>(a) * cols * elemSize1_(type_)); }
	template <typename T> T * ptr(int r = 0) { return reinterpret_cast<T *>(data + static_cast<size_t>(r) * cols * elemSize1_(type_)); }
	template <typename T> const T * ptr(int r = 0) const { return reinterpret_cast<const T *>(data + static_cast<size_t>(r) * cols * elemSize1_(type_)); }
	template <typename T> T & at(int r, int c = 0) { return ptr<T>(r)[c]; }
	template <typename T> const T & at(int r, int c = 0) const { return ptr<T>(r)[c]; }
private:
	static size_t elemSize1_(int t) { return t == CV_8U ? 1 : (t == CV_64F ? 8 : 4); }
	int type_ = 0;
	std::vector<unsigned char> store_;
};
} // namespace cv
