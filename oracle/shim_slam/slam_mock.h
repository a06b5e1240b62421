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
#include <list>
#include <unordered_set>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "auxiliar.h"        // the reference's own (sort predicates, vector_mad, KeyLine via the line_descriptor header, Vector6d)
#include "lineIterator.h"    // the reference's own (src/lineIterator.cpp is compiled with the matchers)

using namespace std;    // the reference's Frame.h leaks these two (through LineExtractor.h:16-17); ORBmatcher.h and MapPoint.cc rely on it
using namespace cv;

namespace ORB_SLAM2 {
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
class MapPoint;
class MapLine;
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

  // ---- lines (LSDmatcher.cpp)
  int NL = 0;
  std::vector<KeyLine> mvKeyLines, mvKeylinesUn;
  std::vector<Eigen::Vector3d> mvKeyLineFunctions;
  cv::Mat mLdesc, mLineDescriptors;                 // Frame names it mLdesc, KeyFrame mLineDescriptors: both kept in step
  std::vector<MapLine*> mvpMapLines;
  cv::Mat mK, ImageGray;
  float mfLogScaleFactorLine = 0.f;
  // mvScaleFactorsLine[level]: the reference indexes this vector with an UNCLAMPED predicted level (LSDmatcher.cpp:944), out of
  // range for every level outside [0, octaves): read here as scale^level by cumulative fp32 products, the rule the table is built with (LineExtractor.cpp:7-14)
  struct ScaleTable {
    float scale = 1.2f;
    float operator[](int level) const { float sf = 1.0f; for (int k = 0; k < (level < 0 ? -level : level); k++) sf = sf * scale; return level < 0 ? 1.0f / sf : sf; }
  } mvScaleFactorsLine;
  std::vector<std::size_t> mGridForLine[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  void AssignFeaturesToGridForLine() {                                        // Frame.cc:296-320, with the reference's LineIterator
    for (int i = 0; i < FRAME_GRID_COLS; i++) for (int j = 0; j < FRAME_GRID_ROWS; j++) mGridForLine[i][j].clear();
    for (int i = 0; i < NL; i++) {
      const KeyLine& kl = mvKeylinesUn[i];
      LineIterator it(kl.startPointX * mfGridElementWidthInv, kl.startPointY * mfGridElementHeightInv, kl.endPointX * mfGridElementWidthInv,
                      kl.endPointY * mfGridElementHeightInv);
      std::pair<int, int> p;
      while (it.getNext(p))
        if (p.first >= 0 && p.first < FRAME_GRID_COLS && p.second >= 0 && p.second < FRAME_GRID_ROWS) mGridForLine[p.first][p.second].push_back(i);
    }
  }
  std::vector<std::size_t> GetFeaturesInAreaForLine(const float& x1, const float& y1, const float& x2, const float& y2, const float& r,
                                                    const int minLevel = -1, const int maxLevel = -1, const float TH = 0.998) const {
    std::vector<std::size_t> vIndices;                                        // Frame.cc:768-842
    vIndices.reserve(NL);
    std::unordered_set<std::size_t> vIndices_set;
    float x[3] = {x1, (float)((x1 + x2) / 2.0), x2};
    float y[3] = {y1, (float)((y1 + y2) / 2.0), y2};
    float delta1x = x1 - x2, delta1y = y1 - y2;
    float norm_delta1 = sqrt(delta1x * delta1x + delta1y * delta1y);
    delta1x /= norm_delta1; delta1y /= norm_delta1;
    for (int i = 0; i < 3; i++) {
      const int nMinCellX = std::max(0, (int)floor((x[i] - mnMinX - r) * mfGridElementWidthInv));
      if (nMinCellX >= FRAME_GRID_COLS) continue;
      const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x[i] - mnMinX + r) * mfGridElementWidthInv));
      if (nMaxCellX < 0) continue;
      const int nMinCellY = std::max(0, (int)floor((y[i] - mnMinY - r) * mfGridElementHeightInv));
      if (nMinCellY >= FRAME_GRID_ROWS) continue;
      const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y[i] - mnMinY + r) * mfGridElementHeightInv));
      if (nMaxCellY < 0) continue;
      for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
          const std::vector<std::size_t>& vCell = mGridForLine[ix][iy];
          for (std::size_t j = 0, jend = vCell.size(); j < jend; j++) {
            if (vIndices_set.find(vCell[j]) != vIndices_set.end()) continue;
            const KeyLine& klUn = mvKeylinesUn[vCell[j]];
            float delta2x = klUn.startPointX - klUn.endPointX, delta2y = klUn.startPointY - klUn.endPointY;
            float norm_delta2 = sqrt(delta2x * delta2x + delta2y * delta2y);
            delta2x /= norm_delta2; delta2y /= norm_delta2;
            float CosSita = abs(delta1x * delta2x + delta1y * delta2y);
            if (CosSita < TH) continue;
            Eigen::Vector3d Lfunc = mvKeyLineFunctions[vCell[j]];
            const float dist = Lfunc(0) * x[i] + Lfunc(1) * y[i] + Lfunc(2);
            if (fabs(dist) < r) { vIndices.push_back(vCell[j]); vIndices_set.insert(vCell[j]); }
          }
        }
    }
    return vIndices;
  }
  std::vector<std::size_t> GetLinesInArea(const float& x1, const float& y1, const float& x2, const float& y2, const float& r, const float TH = 0.998) const {
    std::vector<std::size_t> vIndices;                                        // KeyFrame.cc:647-682
    float delta1x = x1 - x2, delta1y = y1 - y2;
    float norm_delta1 = sqrt(delta1x * delta1x + delta1y * delta1y);
    delta1x /= norm_delta1; delta1y /= norm_delta1;
    for (std::size_t i = 0; i < mvKeyLines.size(); i++) {
      const KeyLine& keyline = mvKeyLines[i];
      float distance = (0.5 * (x1 + x2) - keyline.pt.x) * (0.5 * (x1 + x2) - keyline.pt.x) + (0.5 * (y1 + y2) - keyline.pt.y) * (0.5 * (y1 + y2) - keyline.pt.y);
      if (distance > r * r) continue;
      float delta2x = keyline.startPointX - keyline.endPointX, delta2y = keyline.startPointY - keyline.endPointY;
      float norm_delta2 = sqrt(delta2x * delta2x + delta2y * delta2y);
      delta2x /= norm_delta2; delta2y /= norm_delta2;
      float CosSita = abs(delta1x * delta2x + delta1y * delta2y);
      if (CosSita < TH) continue;
      vIndices.push_back(i);
    }
    return vIndices;
  }
  MapLine* GetMapLine(const std::size_t& idx) { return mvpMapLines[idx]; }
  void AddMapLine(MapLine* pML, const std::size_t& idx) { mvpMapLines[idx] = pML; }
};

// MapLine: plain data + the few rules of MapLine.cpp the matcher relies on (restated: MapLine.cpp needs Eigen proper)
class MapLine {
 public:
  Vector6d mWorldPos; Eigen::Vector3d mNormalVector;
  cv::Mat mLDescriptor;
  float mfMinDistance = 0, mfMaxDistance = 0;
  bool mbBad = false, mbTrackInView = false, mbInFrustum = true;
  int nObs = 0, mnTrackScaleLevel = 0;
  float mTrackViewCos = 1.f, mTrackProjX1 = 0, mTrackProjY1 = 0, mTrackProjX2 = 0, mTrackProjY2 = 0;
  std::map<KeyFrame*, std::size_t> mObservations;
  MapLine* mpReplaced = nullptr;
  Vector6d GetWorldPos() { return mWorldPos; }
  Eigen::Vector3d GetNormal() { return mNormalVector; }
  cv::Mat GetDescriptor() { return mLDescriptor.clone(); }
  bool isBad() { return mbBad; }
  int Observations() { return nObs; }
  bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
  int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
  float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }            // MapLine.cpp:383-393
  float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
  int PredictScale(const float& currentDist, const float& logScaleFactor) {    // MapLine.cpp:395-404 (no clamping)
    float ratio = mfMaxDistance / currentDist;
    return ceil(log(ratio) / logScaleFactor);
  }
  void AddObservation(KeyFrame* pKF, std::size_t idx) { if (mObservations.count(pKF)) return; mObservations[pKF] = idx; nObs++; }
  void Replace(MapLine* pML) { if (pML == this) return; mbBad = true; mpReplaced = pML; }
};

#ifndef PL_SHIM_REAL_FRAME
class Frame : public GridView {
 public:
  long unsigned int mnId = 0;
  cv::Mat mTcw, mOw;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier, mvbLineOutlier;
  cv::Mat GetCameraCenter() { return mOw.clone(); }
  // Frame::isInFrustum(MapLine*, viewingCosLimit) (Frame.cc:628-702) belongs to the frame glue (cv2-pinned projection, tests of
  // oracle_frame); here the wrapper has stored its outcome (flag + mTrackProj*) on the map line
  bool isInFrustum(MapLine* pML, float) const;
  // only LSDmatcher::SerachForInitialize (LSDmatcher.cpp:340-373) calls this; no caller of that function exists in the reference
  void lineDescriptorMAD(std::vector<std::vector<cv::DMatch>>, double&, double&) const { abort(); }
};
#endif

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

#ifndef PL_SHIM_REAL_FRAME
inline bool Frame::isInFrustum(MapLine* pML, float) const { return pML->mbInFrustum; }
#endif

}  // namespace ORB_SLAM2
