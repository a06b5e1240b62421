// ORACLE shim (test infrastructure): the reference's ORBVocabulary is DBoW2's templated vocabulary (needs cv::FileStorage and the
// 140 MB vocabulary file); Frame.cc only calls transform() in ComputeBoW (src/Frame.cc:906-913), which no test reaches
#pragma once
#include <opencv2/core/core.hpp>
#include <vector>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
namespace ORB_SLAM2 {
class ORBVocabulary {
 public:
  void transform(const std::vector<cv::Mat>&, DBoW2::BowVector&, DBoW2::FeatureVector&, int) const { abort(); }
};
}
