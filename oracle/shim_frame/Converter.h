// ORACLE shim (test infrastructure): the reference's Converter.h pulls in g2o and Eigen; Frame.cc uses one function of it
// (toDescriptorVector in ComputeBoW, src/Frame.cc:910), which no test reaches
#pragma once
#include <opencv2/core/core.hpp>
#include <vector>
namespace ORB_SLAM2 {
class Converter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& D) { std::vector<cv::Mat> v; for (int j = 0; j < D.rows; j++) v.push_back(D.row(j)); return v; }
};
}
