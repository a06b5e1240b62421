// Drop-in replacement of ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111): same class name, ctor,
// operator(), getters and public mvImagePyramid, implemented on the plslam_b200 C ABI (no OpenCV on the hot path).
#pragma once
#include <vector>
#include "plcv.h"
struct PLOrb;
namespace ORB_SLAM2 {
class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  ~ORBextractor();
  // Mask is ignored, as in the reference (ORBextractor.h:58).
  void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);
  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return (float)scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
  std::vector<cv::Mat> mvImagePyramid;   // filled on demand by FetchImagePyramid() (only stereo matching reads it)
  void FetchImagePyramid();
 protected:
  void EnsureHandle(int width, int height);
  int nfeatures; double scaleFactor; int nlevels, iniThFAST, minThFAST;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  PLOrb* handle = nullptr; int hw = 0, hh = 0;
};
}  // namespace ORB_SLAM2
