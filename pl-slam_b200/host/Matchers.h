// C++ host classes with the reference's names for the matcher / optimiser / frame-glue part of the path, on the plslam_b200
// C ABI: ORB_SLAM2::ORBmatcher (reference include/ORBmatcher.h:37-102), ORB_SLAM2::LSDmatcher (include/LSDmatcher.h:22-76),
// ORB_SLAM2::Optimizer::PoseOptimization (include/Optimizer.h), and the Frame glue of src/Frame.cc:215-250.
// The reference's methods take Frame / KeyFrame objects; the members they read are gathered in FrameView (same member
// names), so that inside the reference tree a Frame can be passed by filling a view (or by making Frame derive from it).
#pragma once
#include <vector>
#include "plcv.h"
struct PLUndistort;
namespace ORB_SLAM2 {

struct FrameView {
  // Frame.h: N, mvKeys, mvKeysUn, mDescriptors, NL, mvKeylinesUn, mLdesc, mvKeyLineFunctions, mnMinX..mnMaxY
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
  cv::Mat mDescriptors;
  std::vector<KeyLine> mvKeylinesUn;
  cv::Mat mLdesc;
  std::vector<Eigen::Vector3d> mvKeyLineFunctions;
  float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
};

// Frame::Frame's camera glue (src/Frame.cc:215-250): undistorted image for the line extractor, mvKeysUn, image bounds
class FrameUndistorter {
 public:
  FrameUndistorter(const float K[4], const float distCoef[5], int width, int height);
  ~FrameUndistorter();
  void remap(const cv::Mat& imGray, cv::Mat& ImageGray) const;            // Frame.cc:220-222
  void UndistortKeyPoints(FrameView& F) const;                              // Frame.cc:915-945
  void ComputeImageBounds(FrameView& F) const;                              // Frame.cc:947-985
 private:
  float K_[4], D_[5]; int w_, h_;
  PLUndistort* handle_ = nullptr;
};

class ORBmatcher {
 public:
  static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;
  ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);        // ORBmatcher.cc:1764-1780 (first rows)
  // ORBmatcher.cc:455-572; vbPrevMatched in/out, vnMatches12 out
  int SearchForInitialization(FrameView& F1, FrameView& F2, std::vector<cv::Point2f>& vbPrevMatched,
                              std::vector<int>& vnMatches12, int windowSize = 10);
 protected:
  float mfNNratio; bool mbCheckOrientation;
};

class LSDmatcher {
 public:
  static const int TH_LOW = 50, TH_HIGH = 80;
  LSDmatcher(float nnratio = 0.7, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
  int SearchDouble(FrameView& InitialFrame, FrameView& CurrentFrame, std::vector<int>& LineMatches);   // LSDmatcher.cpp:440-460
  void FrameBFMatch(const cv::Mat& ldesc1, const cv::Mat& ldesc2, std::vector<int>& LineMatches, float TH);   // :462-486
 protected:
  float mfNNratio; bool mbCheckOrientation;
};

// One correspondence set of a frame, as Optimizer::PoseOptimization gathers it (Optimizer.cc:668-840)
struct PoseProblem {
  float Tcw[16];                                 // pFrame->mTcw, row-major
  float K[4];                                    // fx fy cx cy
  std::vector<float> pt_obs, pt_invSigma2, pt_Xw;   // [n][2], [n], [n][3]
  std::vector<double> line_func, line_Xw;            // [m][3], [m][6]
};
class Optimizer {
 public:
  // returns nInitialCorrespondences - nBad like the reference; Tcw is updated; outlier flags as mvbOutlier / mvbLineOutlier
  static int PoseOptimization(PoseProblem& P, std::vector<bool>& vbOutlier, std::vector<bool>& vbLineOutlier);
};
}  // namespace ORB_SLAM2
