// util3d::solvePnPRansac on the B200 — the body a maintainer swaps in at corelib/src/util3d_motion_estimation.cpp:843-990 (the free
// function RegistrationVis / estimateMotion3DTo2D call; SURVEY.md §8(b) "simplest splice point").  Same argument list, same outputs.
#include <opencv2/core.hpp>
#include <rtabmap/utilite/ULogger.h>
#include <vector>
#include "lcd_b200.h"

namespace rtabmap {
namespace util3d {

void solvePnPRansac(const std::vector<cv::Point3f> & objectPoints, const std::vector<cv::Point2f> & imagePoints, const cv::Mat & cameraMatrix,
                    const cv::Mat & distCoeffs, cv::Mat & rvec, cv::Mat & tvec, bool useExtrinsicGuess, int iterationsCount,
                    float reprojectionError, int minInliersCount, std::vector<int> & inliers, int flags, int refineIterations, float refineSigma)
{
	static lcd_engine * engine = 0; // util3d is stateless: one lazily created engine serves every call of the process
	inliers.clear();
	if (!engine)
	{
		lcd_config c;
		c.device = 0; c.desc_type = LCD_DESC_U8; c.desc_dim = 32; c.max_words = 1024; c.max_signatures = 1024; c.max_queries = 4096; c.max_batch = 1;
		engine = lcd_create(&c);
		if (!engine)
		{
			UERROR("B200 PnP engine: %s", lcd_last_error(0));
			return;
		}
	}
	const int n = (int)objectPoints.size();
	if (n == 0 || n != (int)imagePoints.size()) return;
	double K[9];
	for (int i = 0; i < 9; ++i) K[i] = cameraMatrix.at<double>(i / 3, i % 3);
	std::vector<double> D((size_t)distCoeffs.rows * distCoeffs.cols);
	for (size_t i = 0; i < D.size(); ++i) D[i] = distCoeffs.ptr<double>()[i];
	std::vector<int> idx(n);
	int nInliers = 0;
	static_assert(sizeof(cv::Point3f) == 12 && sizeof(cv::Point2f) == 8, "points are packed floats");
	const int rc = lcd_pnp_ransac(engine, &objectPoints[0].x, &imagePoints[0].x, n, K, D.empty() ? 0 : D.data(), (int)D.size(), rvec.ptr<double>(),
	                              tvec.ptr<double>(), useExtrinsicGuess ? 1 : 0, iterationsCount, reprojectionError, minInliersCount, flags,
	                              refineIterations, refineSigma, idx.data(), &nInliers);
	if (rc != LCD_OK)
	{
		UERROR("%s", lcd_last_error(engine)); // distortion / P3P / flags: the caller keeps the CPU implementation for those
		return;
	}
	inliers.assign(idx.begin(), idx.begin() + nInliers);
}

} // namespace util3d
} // namespace rtabmap
