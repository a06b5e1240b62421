// The reference's matcher / optimiser classes with the reference's OWN signatures (include/ORBmatcher.h:37-102,
// include/LSDmatcher.h:22-76, include/Optimizer.h:56-65), implemented in reference_glue.cc on the plslam_b200 C ABI.
// Inside the reference tree these declarations are the reference's headers themselves (unchanged); this file only exists
// because those headers cannot be included here (OpenCV / Eigen / DBoW2 are absent): it repeats the members the glue
// defines, against the mock Frame / MapPoint / MapLine / KeyFrame / Map of reference_mock.h.
#pragma once
#ifdef PLSLAM_IN_REFERENCE_TREE
#include "Frame.h"
#include "KeyFrame.h"
#include "LSDmatcher.h"
#include "Map.h"
#include "MapLine.h"
#include "MapPoint.h"
#include "ORBmatcher.h"
#include "Optimizer.h"
#else
#include "reference_mock.h"
namespace ORB_SLAM2 {

class ORBmatcher {
 public:
  ORBmatcher(float nnratio = 0.6, bool checkOri = true);
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  // Search matches between Frame keypoints and projected MapPoints. Returns number of matches (Tracking::SearchLocalPoints)
  int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
  // Project MapPoints tracked in last frame into the current frame and search matches (Tracking::TrackWithMotionModel)
  int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
  // Matching for the Map Initialization (only used in the monocular case)
  int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);
  static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
 protected:
  float mfNNratio; bool mbCheckOrientation;
};

class LSDmatcher {
 public:
  LSDmatcher(float nnratio = 0.7, bool checkOri = true);
  int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th);
  int SearchByProjection(Frame& F, const std::vector<MapLine*>& vpMapLines, const float th = 3);
  int SearchDouble(Frame& InitialFrame, Frame& CurrentFrame, std::vector<int>& LineMatches);
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
 protected:
  float mfNNratio; bool mbCheckOrientation;
};

class Optimizer {
 public:
  int static PoseOptimization(Frame* pFrame);
  int static PoseOptimizationWithPoints(Frame* pFrame);
  int static PoseOptimizationWithLines(Frame* pFrame);
  void static LocalBundleAdjustmentWithLine(KeyFrame* pKF, bool* pbStopFlag, Map* pMap);
};

}  // namespace ORB_SLAM2
#endif
