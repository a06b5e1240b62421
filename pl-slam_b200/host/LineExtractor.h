// Drop-in replacement of ORB_SLAM2::LINEextractor (reference include/LineExtractor.h:20-62) on the plslam_b200 C ABI.
#pragma once
#include <vector>
#include "plcv.h"
struct PLLine;
namespace ORB_SLAM2 {
class LINEextractor {
 public:
  LINEextractor(int _numOctaves, float _scale, unsigned int _nLSDFeature, double _min_line_length);
  ~LINEextractor();
  void operator()(cv::InputArray image, cv::InputArray mask, std::vector<KeyLine>& keylines, cv::OutputArray descriptors,
                  std::vector<Eigen::Vector3d>& lineVec2d);
  int inline GetLevels() { return numOctaves; }
  float inline GetScaleFactor() { return scale; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
 protected:
  int numOctaves; float scale; unsigned int nLSDFeature; double min_line_length;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  PLLine* handle = nullptr; int hw = 0, hh = 0;
};
}  // namespace ORB_SLAM2
