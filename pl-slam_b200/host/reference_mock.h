// MOCK of the reference's map / frame classes: exactly the members that the methods implemented in reference_glue.cc touch,
// with the reference's names, types and semantics (include/Frame.h, MapPoint.h, MapLine.h, KeyFrame.h, Map.h of
// HarborC/PL-SLAM).  It exists so that reference_glue.cc - written against the REAL signatures
//   ORBmatcher::SearchByProjection(Frame&, const Frame&, float, bool)          include/ORBmatcher.h:37-102
//   LSDmatcher::SearchByProjection(Frame&, const std::vector<MapLine*>&, float) include/LSDmatcher.h:22-76
//   Optimizer::PoseOptimization(Frame*), LocalBundleAdjustmentWithLine(KeyFrame*, bool*, Map*)   include/Optimizer.h:56-65
// - compiles and is TESTED in this repository, where OpenCV / Eigen / the reference headers are not available.  Inside the
// reference tree reference_glue.cc is compiled with -DPLSLAM_IN_REFERENCE_TREE and includes the real headers instead of this
// file; nothing in it depends on anything the real classes do not offer.
#pragma once
#include <array>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <vector>
#include "plcv.h"

typedef std::array<double, 6> Vector6d;      // Eigen::Matrix<double,6,1> (MapLine.h:28): only operator[] / () style access is used

namespace ORB_SLAM2 {
class KeyFrame;
class Frame;
class Map;

class MapPoint {
 public:
  cv::Mat GetWorldPos() { return mWorldPos.clone(); }                     // 3x1 CV_32F
  void SetWorldPos(const cv::Mat& Pos) { mWorldPos = Pos.clone(); }
  cv::Mat GetDescriptor() { return mDescriptor.clone(); }                 // 1x32 CV_8U
  bool isBad() { return mbBad; }
  int Observations() { return nObs; }
  std::map<KeyFrame*, size_t> GetObservations() { return mObservations; }
  void EraseObservation(KeyFrame* pKF) { if (mObservations.erase(pKF)) nObs--; }
  void UpdateNormalAndDepth() { nNormalUpdates++; }                       // (MapPoint.cc:316-360: recomputed from the observations)
  long unsigned int mnId = 0;
  bool mbTrackInView = false; float mTrackProjX = 0, mTrackProjY = 0; int mnTrackScaleLevel = 0; float mTrackViewCos = 0;
  long unsigned int mnLastFrameSeen = 0, mnBALocalForKF = 0;
  static std::mutex mGlobalMutex;
  // mock state
  cv::Mat mWorldPos, mDescriptor; bool mbBad = false; int nObs = 0, nNormalUpdates = 0;
  std::map<KeyFrame*, size_t> mObservations;
};

class MapLine {
 public:
  Vector6d GetWorldPos() { return mWorldPos; }
  void SetWorldPos(const Vector6d& Pos) { mWorldPos = Pos; }
  cv::Mat GetDescriptor() { return mLDescriptor.clone(); }
  bool isBad() { return mbBad; }
  int Observations() { return nObs; }
  std::map<KeyFrame*, size_t> GetObservations() { return mObservations; }
  void EraseObservation(KeyFrame* pKF) { if (mObservations.erase(pKF)) nObs--; }
  void UpdateAverageDir() { nDirUpdates++; }
  long unsigned int mnId = 0;
  bool mbTrackInView = false; float mTrackProjX1 = 0, mTrackProjY1 = 0, mTrackProjX2 = 0, mTrackProjY2 = 0;
  int mnTrackScaleLevel = 0; float mTrackViewCos = 0;
  long unsigned int mnLastFrameSeen = 0, mnBALocalForKF = 0;
  static std::mutex mGlobalMutex;
  // mock state (the frustum test of Frame::isInFrustum(MapLine*) needs them)
  Vector6d mWorldPos{}; std::array<double, 3> mNormalVector{}; float mfMinDistance = 0, mfMaxDistance = 0;
  cv::Mat mLDescriptor; bool mbBad = false; int nObs = 0, nDirUpdates = 0;
  std::map<KeyFrame*, size_t> mObservations;
};

class Frame {
 public:
  // Frame.h:137-146, :156-230: what ORBmatcher / LSDmatcher / Optimizer read and write
  static float fx, fy, cx, cy;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
  float mb = 0;
  int N = 0, NL = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
  std::vector<float> mvuRight;
  cv::Mat mDescriptors;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  cv::Mat mLdesc;
  std::vector<KeyLine> mvKeylinesUn;
  std::vector<Eigen::Vector3d> mvKeyLineFunctions;
  std::vector<bool> mvbLineOutlier;
  std::vector<MapLine*> mvpMapLines;
  cv::Mat mTcw;                                  // 4x4 CV_32F
  cv::Mat mOw;                                   // 3x1 CV_32F (Frame::GetCameraCenter)
  int mnScaleLevels = 8; float mfLogScaleFactor = 0;
  std::vector<float> mvScaleFactors, mvInvLevelSigma2;
  float mfLogScaleFactorLine = 0;
  long unsigned int mnId = 0;
  void SetPose(cv::Mat Tcw) { mTcw = Tcw.clone(); }
};

class KeyFrame {
 public:
  long unsigned int mnId = 0, mnBALocalForKF = 0, mnBAFixedForKF = 0;
  float fx = 0, fy = 0, cx = 0, cy = 0;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<Eigen::Vector3d> mvKeyLineFunctions;
  std::vector<float> mvInvLevelSigma2;
  bool isBad() { return mbBad; }
  cv::Mat GetPose() { return Tcw.clone(); }
  void SetPose(const cv::Mat& T) { Tcw = T.clone(); }
  std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return mvpOrderedConnectedKeyFrames; }
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  std::vector<MapLine*> GetMapLineMatches() { return mvpMapLines; }
  void EraseMapPointMatch(MapPoint* pMP) { for (auto& p : mvpMapPoints) if (p == pMP) p = nullptr; }
  void EraseMapLineMatch(MapLine* pML) { for (auto& p : mvpMapLines) if (p == pML) p = nullptr; }
  // mock state
  cv::Mat Tcw; bool mbBad = false;
  std::vector<KeyFrame*> mvpOrderedConnectedKeyFrames;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<MapLine*> mvpMapLines;
};

class Map {
 public:
  std::mutex mMutexMapUpdate;
};

}  // namespace ORB_SLAM2
