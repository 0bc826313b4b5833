// ORB_B200 — rtabmap::Feature2D for Kp/DetectorStrategy=2 backed by liblcd_b200.so.
//
// Feature2D::create (corelib/src/Features2d.cpp:617) gets one more case: `case Feature2D::kFeatureOrb: return gpu ? new ORB_B200(p) : new ORB(p);`.
// Feature2D's public, non-virtual generateKeypoints() builds the depth mask and calls the private virtual generateKeypointsImpl()
// (Features2d.cpp:775-857); generateDescriptors() calls generateDescriptorsImpl() with the keypoints it got back (:879-903).  The
// engine computes keypoints and descriptors in ONE pass, so the descriptors of the last detection are kept and handed out when
// generateDescriptorsImpl is called for the same keypoints (what Memory::createSignature does, Memory.cpp:5529-5560).
#include <rtabmap/core/Features2d.h>
#include <rtabmap/utilite/ULogger.h>
#include "lcd_b200.h"

namespace rtabmap {

class ORB_B200 : public Feature2D
{
public:
	explicit ORB_B200(const ParametersMap & parameters = ParametersMap()) : Feature2D(parameters), e_(0), n_(0)
	{
		p_.n_features = 1000; p_.n_levels = 3; p_.scale_factor = 2.0f; p_.edge_threshold = 19; p_.fast_threshold = 20; p_.patch_size = 31;
		p_.min_depth = 0; p_.max_depth = 0; p_.depth_as_mask = 1; p_.fx = p_.fy = 1; p_.cx = p_.cy = 0;
		parseParameters(parameters);
		lcd_config c;
		c.device = 0; c.desc_type = LCD_DESC_U8; c.desc_dim = 32; c.max_words = 1024; c.max_signatures = 1024; c.max_queries = 4096; c.max_batch = 1;
		e_ = lcd_create(&c);
		if (!e_) UERROR("B200 ORB engine: %s", lcd_last_error(0));
	}
	~ORB_B200() override { lcd_destroy(e_); }
	Feature2D::Type getType() const override { return kFeatureOrb; }

	void parseParameters(const ParametersMap & parameters) override
	{
		Feature2D::parseParameters(parameters);
		Parameters::parse(parameters, "ORB/NLevels", p_.n_levels);            // Parameters.h:321
		Parameters::parse(parameters, "ORB/ScaleFactor", p_.scale_factor);    // :320
		Parameters::parse(parameters, "ORB/EdgeThreshold", p_.edge_threshold); // :322
		Parameters::parse(parameters, "ORB/PatchSize", p_.patch_size);        // :326
		Parameters::parse(parameters, "FAST/Threshold", p_.fast_threshold);   // :303
		p_.n_features = getMaxFeatures();
	}

private:
	std::vector<cv::KeyPoint> generateKeypointsImpl(const cv::Mat & image, const cv::Rect & roi, const cv::Mat & mask) override
	{
		std::vector<cv::KeyPoint> out;
		if (!e_ || image.empty()) return out;
		if (roi.x != 0 || roi.y != 0 || roi.width != image.cols || roi.height != image.rows)
		{
			UERROR("ORB_B200 works on whole images (Kp/RoiRatios and Kp/GridRows/Cols must stay at their defaults)");
			return out;
		}
		const int cap = p_.n_features + 256;
		std::vector<lcd_keypoint> kp(cap);
		desc_ = cv::Mat(cap, 32, CV_8U);
		// the mask Feature2D built from the depth image (0 / 255) goes in as it is; without one every pixel is eligible
		const int rc = lcd_orb_detect_describe(e_, 1, image.data, image.cols, image.rows, image.channels(), mask.empty() ? 0 : mask.data,
		                                       mask.empty() ? LCD_DEPTH_NONE : LCD_DEPTH_MASK_U8, &p_, cap, kp.data(), desc_.data, 0, &n_);
		if (rc != LCD_OK)
		{
			UERROR("%s", lcd_last_error(e_));
			n_ = 0;
			return out;
		}
		out.resize(n_);
		for (int i = 0; i < n_; ++i)
		{
			out[i].pt.x = kp[i].x; out[i].pt.y = kp[i].y; out[i].size = kp[i].size; out[i].angle = kp[i].angle;
			out[i].response = kp[i].response; out[i].octave = kp[i].octave;
		}
		return out;
	}

	cv::Mat generateDescriptorsImpl(const cv::Mat &, std::vector<cv::KeyPoint> & keypoints) const override
	{
		if ((int)keypoints.size() != n_) UWARN("ORB_B200: descriptors requested for %d keypoints, the last detection found %d", (int)keypoints.size(), n_);
		return n_ ? desc_.rowRange(0, n_) : cv::Mat();
	}

	lcd_engine * e_;
	lcd_orb_params p_;
	cv::Mat desc_;
	int n_;
};

Feature2D * createOrbB200(const ParametersMap & parameters) { return new ORB_B200(parameters); }

} // namespace rtabmap
