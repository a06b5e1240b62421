// ORACLE — TEST INFRASTRUCTURE ONLY.  Mock Frame / KeyFrame / Map for compiling the reference's ORBmatcher.cc and MapPoint.cc
// unmodified and where they lie (oracle/Makefile target `ref` -> oracle/_ref/libref_match.so).  The compiler is run with
//   -Ioracle/shim_slam -I- -Ioracle/shim -I/root/reference -I/root/reference/include
// so the reference's own  #include "KeyFrame.h" / "Frame.h" / "Map.h"  (whose real versions pull in DBoW2's vocabulary, Eigen,
// g2o, OpenMP, Pangolin...) resolve to the three one-line headers next to this file, while ORBmatcher.h, MapPoint.h and the
// DBoW2 FeatureVector are the reference's real headers.  The mocks declare the members those two source files touch, with the
// reference's names and types (include/Frame.h, include/KeyFrame.h), and hold plain data filled by oracle/ref_match_wrap.cpp.
// Restated here (because Frame.cc / KeyFrame.cc cannot be compiled): the bucket-grid lookup GetFeaturesInArea
// (Frame.cc:713-766, KeyFrame.cc:606-650), PosInGrid / AssignFeaturesToGrid (Frame.cc:278-294, 893-903), IsInImage
// (KeyFrame.cc:760-763).  Everything ORBmatcher.cc and MapPoint.cc do themselves is the reference's code.
#pragma once
#include <opencv2/core/core.hpp>
#include <climits>
#include <map>
#include <mutex>
#include <set>
#include <vector>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std;    // the reference's Frame.h leaks these two (through LineExtractor.h:16-17); ORBmatcher.h and MapPoint.cc rely on it
using namespace cv;

namespace ORB_SLAM2 {
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
class MapPoint;
class KeyFrame;
class Frame;

class Map {
 public:
  std::mutex mMutexPointCreation;
  std::vector<MapPoint*> erased;
  void EraseMapPoint(MapPoint* p) { erased.push_back(p); }
};

// what Frame and KeyFrame share on this path: undistorted keys, descriptors, scale tables, intrinsics, image bounds, grid
class GridView {
 public:
  int N = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
  std::vector<float> mvuRight, mvDepth;
  cv::Mat mDescriptors;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
  int mnScaleLevels = 8;
  float mfScaleFactor = 1.2f, mfLogScaleFactor = 0.f;
  std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0, mThDepth = 0;
  float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;     // (static members in Frame, per-object in KeyFrame: same use)
  float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
  std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];

  bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY) const {      // Frame.cc:893-903
    posX = round((kp.pt.x - mnMinX) * mfGridElementWidthInv);
    posY = round((kp.pt.y - mnMinY) * mfGridElementHeightInv);
    if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) return false;
    return true;
  }
  void AssignFeaturesToGrid() {                                               // Frame.cc:278-294
    for (int i = 0; i < FRAME_GRID_COLS; i++) for (int j = 0; j < FRAME_GRID_ROWS; j++) mGrid[i][j].clear();
    for (int i = 0; i < N; i++) {
      int gx, gy;
      if (PosInGrid(mvKeysUn[i], gx, gy)) mGrid[gx][gy].push_back(i);
    }
  }
  std::vector<std::size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const {
    std::vector<std::size_t> vIndices;                                        // Frame.cc:713-766
    vIndices.reserve(N);
    const int nMinCellX = std::max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
    if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
    const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = std::max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
    if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
    const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
    if (nMaxCellY < 0) return vIndices;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
      for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
        const std::vector<std::size_t>& vCell = mGrid[ix][iy];
        for (std::size_t j = 0, jend = vCell.size(); j < jend; j++) {
          const cv::KeyPoint& kpUn = mvKeysUn[vCell[j]];
          if (bCheckLevels) {
            if (kpUn.octave < minLevel) continue;
            if (maxLevel >= 0 && kpUn.octave > maxLevel) continue;
          }
          const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
          if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
        }
      }
    return vIndices;
  }
  bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }   // KeyFrame.cc:760
};

class Frame : public GridView {
 public:
  long unsigned int mnId = 0;
  cv::Mat mTcw, mOw;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  cv::Mat GetCameraCenter() { return mOw.clone(); }
};

class KeyFrame : public GridView {
 public:
  long unsigned int mnId = 0, mnFrameId = 0;
  cv::Mat Tcw, Ow;                                   // 4x4 and 3x1, CV_32F
  std::vector<MapPoint*> mvpMapPoints;
  bool mbBad = false;
  struct Call { int kind; MapPoint* a; MapPoint* b; std::size_t idx; };   // map surgery recorded for the wrapper: 0 AddMapPoint, 1 Replace, 2 Erase
  std::vector<Call> calls;
  cv::Mat GetRotation() { return Tcw.rowRange(0, 3).colRange(0, 3).clone(); }
  cv::Mat GetTranslation() { return Tcw.rowRange(0, 3).col(3).clone(); }
  cv::Mat GetCameraCenter() { return Ow.clone(); }
  cv::Mat GetPose() { return Tcw.clone(); }
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  MapPoint* GetMapPoint(const std::size_t& idx) { return mvpMapPoints[idx]; }
  std::set<MapPoint*> GetMapPoints();
  void AddMapPoint(MapPoint* pMP, const std::size_t& idx) { mvpMapPoints[idx] = pMP; calls.push_back({0, pMP, nullptr, idx}); }
  void EraseMapPointMatch(const std::size_t& idx) { calls.push_back({2, mvpMapPoints[idx], nullptr, idx}); mvpMapPoints[idx] = nullptr; }
  void ReplaceMapPointMatch(const std::size_t& idx, MapPoint* pMP) { calls.push_back({1, mvpMapPoints[idx], pMP, idx}); mvpMapPoints[idx] = pMP; }
  bool isBad() { return mbBad; }
};

}  // namespace ORB_SLAM2
