// shim_driver.cpp — exercises the three reference-side shims against a live engine (tests/test_integration_shims.py, GPU box).
// Not a parity test (those are tests/test_gpu_*.py through the same C ABI): it checks that the shims drive the ABI the way
// rtabmap::Memory drives VWDictionary / Feature2D / util3d::solvePnPRansac, and that the host containers stay consistent.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "VWDictionaryB200.h"
#include <rtabmap/core/Features2d.h>

namespace rtabmap {
Feature2D * createOrbB200(const ParametersMap & parameters);
namespace util3d {
void solvePnPRansac(const std::vector<cv::Point3f> &, const std::vector<cv::Point2f> &, const cv::Mat &, const cv::Mat &, cv::Mat &, cv::Mat &, bool, int,
                    float, int, std::vector<int> &, int, int, float);
}
} // namespace rtabmap

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main()
{
	using namespace rtabmap;
	std::mt19937 rng(7);
	// ---- dictionary: three frames, the third re-observes the first -----------------------------------------
	VWDictionaryB200 vwd;
	const int n = 300;
	cv::Mat f1(n, 32, CV_8U), f2(n, 32, CV_8U), f3(n, 32, CV_8U);
	for (int i = 0; i < n * 32; ++i) { f1.data[i] = (unsigned char)rng(); f2.data[i] = (unsigned char)rng(); }
	for (int i = 0; i < n * 32; ++i) f3.data[i] = f1.data[i] ^ (unsigned char)((rng() % 16 == 0) ? (1u << (rng() % 8)) : 0u); // ~1.5 % bit flips
	vwd.update();
	std::list<int> w1 = vwd.addNewWords(f1, 1);
	CHECK(vwd.ok() && (int)w1.size() == n && (int)vwd.getVisualWords().size() == n); // empty dictionary: every descriptor is a new word
	vwd.update();
	std::list<int> w2 = vwd.addNewWords(f2, 2);
	vwd.update();
	std::list<int> w3 = vwd.addNewWords(f3, 3);
	int same = 0;
	std::list<int>::iterator a = w1.begin(), b = w3.begin();
	for (; a != w1.end(); ++a, ++b) same += *a == *b;
	CHECK(same > n * 9 / 10); // the revisit finds the words of frame 1
	CHECK((int)vwd.getVisualWords().size() == lcd_dict_size(vwd.engine()));
	CHECK(lcd_index_total_refs(vwd.engine()) == 3 * n);
	std::vector<int> words(w3.begin(), w3.end());
	std::list<int> ids;
	ids.push_back(1);
	ids.push_back(2);
	std::map<int, float> like = vwd.computeLikelihood(words, ids, 3);
	CHECK(like.size() == 2 && like[1] > 10 * like[2] && like[1] > 0);
	std::vector<int> nn = vwd.findNN(f3);
	CHECK((int)nn.size() == n && nn[0] == words[0]);
	// forget signature 2 (Memory::disableWordsRef), then remove the words only it used (Memory::cleanUnusedWords)
	for (std::list<int>::iterator i = w2.begin(); i != w2.end(); ++i) vwd.removeAllWordRef(*i, 2);
	CHECK(lcd_index_total_refs(vwd.engine()) == 2 * n);
	std::vector<VisualWord *> unused;
	for (std::map<int, VisualWord *>::const_iterator i = vwd.getVisualWords().begin(); i != vwd.getVisualWords().end(); ++i)
		if (i->second->getReferences().empty()) unused.push_back(i->second);
	CHECK(!unused.empty());
	vwd.removeWords(unused);
	for (size_t i = 0; i < unused.size(); ++i) delete unused[i];
	vwd.update();
	CHECK((int)vwd.getVisualWords().size() == lcd_dict_size(vwd.engine()));
	// ---- ORB: detect + describe one textured frame, with and without a mask -------------------------------
	ParametersMap pm;
	pm["Kp/MaxFeatures"] = "400";
	Feature2D * orb = createOrbB200(pm);
	cv::Mat img(240, 320, CV_8U), mask(240, 320, CV_8U);
	for (int y = 0; y < 240; ++y)
		for (int x = 0; x < 320; ++x)
		{
			img.data[y * 320 + x] = (unsigned char)((((x / 12) + (y / 9)) % 2) * 170 + rng() % 40);
			mask.data[y * 320 + x] = x < 160 ? 255 : 0;
		}
	std::vector<cv::KeyPoint> kp = orb->generateKeypoints(img);
	cv::Mat desc = orb->generateDescriptors(img, kp);
	CHECK(!kp.empty() && (int)kp.size() <= 400 && desc.rows == (int)kp.size() && desc.cols == 32);
	std::vector<cv::KeyPoint> kpm = orb->generateKeypoints(img, mask);
	CHECK(!kpm.empty());
	for (size_t i = 0; i < kpm.size(); ++i) CHECK(kpm[i].pt.x < 160.0f * 1.0f + 1.0f);
	delete orb;
	// ---- PnP: util3d::solvePnPRansac with its own argument list -------------------------------------------
	const int m = 200;
	std::vector<cv::Point3f> obj(m);
	std::vector<cv::Point2f> pix(m);
	std::uniform_real_distribution<float> ux(-2.f, 2.f), uz(1.f, 5.f), un(-0.3f, 0.3f);
	const float tx = 0.1f, ty = -0.05f, tz = 0.2f;
	for (int i = 0; i < m; ++i)
	{
		obj[i].x = ux(rng); obj[i].y = ux(rng) * 0.7f; obj[i].z = uz(rng);
		pix[i].x = 525.f * (obj[i].x + tx) / (obj[i].z + tz) + 320.f + un(rng);
		pix[i].y = 525.f * (obj[i].y + ty) / (obj[i].z + tz) + 240.f + un(rng);
		if (i % 4 == 0) { pix[i].x = (float)(rng() % 640); pix[i].y = (float)(rng() % 480); } // 25 % outliers
	}
	cv::Mat K(3, 3, CV_64F), D(1, 5, CV_64F), rvec(3, 1, CV_64F), tvec(3, 1, CV_64F);
	for (int i = 0; i < 9; ++i) K.ptr<double>()[i] = 0;
	K.at<double>(0, 0) = 525; K.at<double>(1, 1) = 525; K.at<double>(0, 2) = 320; K.at<double>(1, 2) = 240; K.at<double>(2, 2) = 1;
	for (int i = 0; i < 5; ++i) D.ptr<double>()[i] = 0;
	for (int i = 0; i < 3; ++i) { rvec.ptr<double>()[i] = 0; tvec.ptr<double>()[i] = 0; }
	std::vector<int> inliers;
	util3d::solvePnPRansac(obj, pix, K, D, rvec, tvec, true, 300, 2.0f, 20, inliers, 0, 1, 3.0f);
	CHECK((int)inliers.size() > m * 6 / 10);
	CHECK(std::fabs(tvec.ptr<double>()[0] - tx) < 0.02 && std::fabs(tvec.ptr<double>()[1] - ty) < 0.02 && std::fabs(tvec.ptr<double>()[2] - tz) < 0.05);
	std::printf("shim driver ok: %d words, %d ORB keypoints, %d PnP inliers\n", (int)vwd.getVisualWords().size(), (int)kp.size(), (int)inliers.size());
	return 0;
}
