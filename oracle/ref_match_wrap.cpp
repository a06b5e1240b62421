// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C ABI around the reference's own ORBmatcher.cc and MapPoint.cc, compiled unmodified and where they lie
// (oracle/Makefile, target `ref` -> oracle/_ref/libref_match.so; nothing of the reference is copied into the repository):
//   /root/reference/src/ORBmatcher.cc   SearchForInitialization, SearchByProjection (last frame / local points), SearchByBoW,
//                                       SearchForTriangulation, Fuse, DescriptorDistance, ComputeThreeMaxima   
//   /root/reference/src/MapPoint.cc     PredictScale, ComputeDistinctiveDescriptors, Replace, AddObservation ...
// against mock Frame / KeyFrame / Map (oracle/shim_slam/slam_mock.h: plain data; the bucket-grid lookup is restated there
// because Frame.cc / KeyFrame.cc cannot be built) and the OpenCV stand-in (oracle/shim/: cv::Mat with cv::gemm's fp32
// accumulation order, pinned to cv2 by tests/golden/frame_cv2.npz).
// Every entry point takes the SAME flat arrays as the oracle function of the same name (oracle_match.cpp), builds the mock
// objects, calls the reference method and flattens its result the way the oracle reports it.
#include <opencv2/core/core.hpp>
#include <cstdint>
#include <memory>
#include "ORBmatcher.h"     // /root/reference/include

using namespace ORB_SLAM2;

namespace ORB_SLAM2 {
std::set<MapPoint*> KeyFrame::GetMapPoints() {
  std::set<MapPoint*> s;
  for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p);
  return s;
}
}  // namespace ORB_SLAM2

namespace {
struct MP : MapPoint {      // reaches the protected state of the reference's MapPoint
  MP(const cv::Mat& pos, KeyFrame* kf, Map* m) : MapPoint(pos, kf, m) {}
  void set_desc(const uint8_t* d) { mDescriptor = cv::Mat(1, 32, CV_8UC1); memcpy(mDescriptor.ptr(0), d, 32); }
  void set_normal(const float* n) { mNormalVector = cv::Mat(3, 1, CV_32F); for (int i = 0; i < 3; i++) mNormalVector.at<float>(i) = n[i]; }
  void set_dist(float mn, float mx) { mfMinDistance = mn; mfMaxDistance = mx; }
  void set_obs(int n) { nObs = n; }
  const cv::Mat& desc() const { return mDescriptor; }
};
cv::Mat vec3(const float* p) { cv::Mat m(3, 1, CV_32F); for (int i = 0; i < 3; i++) m.at<float>(i) = p[i]; return m; }
struct World {
  Map map;
  KeyFrame ref;                              // reference keyframe of the map points built here
  std::vector<std::unique_ptr<MP>> pts;
  MP* point(const float* pos, const uint8_t* desc, int nobs) {
    static const float zero[3] = {0, 0, 0};
    pts.emplace_back(new MP(vec3(pos ? pos : zero), &ref, &map));
    if (desc) pts.back()->set_desc(desc);
    pts.back()->set_obs(nobs);
    return pts.back().get();
  }
};
void set_view(GridView& g, const void* keys, const uint8_t* desc, int n, const float* bounds, const float* scaleFactors, int nlevels = 8) {
  g.N = n;
  g.mvKeys.assign((const cv::KeyPoint*)keys, (const cv::KeyPoint*)keys + n);
  g.mvKeysUn = g.mvKeys;
  g.mvuRight.assign((size_t)n, -1.f);
  g.mDescriptors = cv::Mat(n, 32, CV_8UC1);
  if (n) memcpy(g.mDescriptors.ptr(0), desc, (size_t)n * 32);
  if (bounds) {
    g.mnMinX = bounds[0]; g.mnMinY = bounds[1]; g.mnMaxX = bounds[2]; g.mnMaxY = bounds[3];
    g.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (g.mnMaxX - g.mnMinX);      // Frame.cc:256-257
    g.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (g.mnMaxY - g.mnMinY);
    g.AssignFeaturesToGrid();
  }
  g.mnScaleLevels = nlevels;
  if (scaleFactors) g.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
}
void set_K(GridView& g, const float* K) { g.fx = K[0]; g.fy = K[1]; g.cx = K[2]; g.cy = K[3]; g.invfx = 1.0f / g.fx; g.invfy = 1.0f / g.fy; }
cv::Mat pose44(const float* T12) {           // rows 0..2 from the 3x4 (stride 4) array, last row 0 0 0 1
  cv::Mat T(4, 4, CV_32F);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) T.at<float>(r, c) = T12[4 * r + c];
  T.at<float>(3, 0) = 0; T.at<float>(3, 1) = 0; T.at<float>(3, 2) = 0; T.at<float>(3, 3) = 1;
  return T;
}
void set_featvec(DBoW2::FeatureVector& fv, const unsigned* nodes, const int* start, const int* items, int nn) {
  for (int a = 0; a < nn; a++) {
    std::vector<unsigned int>& v = fv[nodes[a]];
    for (int i = start[a]; i < start[a + 1]; i++) v.push_back((unsigned)items[i]);
  }
}
}  // namespace

extern "C" {

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat ma(1, 32, CV_8UC1, const_cast<uint8_t*>(a)), mb(1, 32, CV_8UC1, const_cast<uint8_t*>(b));
  return ORBmatcher::DescriptorDistance(ma, mb);
}

int ref_search_for_initialization(const void* keys1, const uint8_t* desc1, int n1, const void* keys2, const uint8_t* desc2, int n2,
                                  const float* bounds, float* prev_matched, int* matches12, int windowSize, float nnratio, int checkOri) {
  static Frame F1, F2;                       // (static: the 64x48 grid of vectors is too large to put on the stack twice)
  F1 = Frame(); F2 = Frame();
  set_view(F1, keys1, desc1, n1, bounds, nullptr);
  set_view(F2, keys2, desc2, n2, bounds, nullptr);
  std::vector<cv::Point2f> prev((size_t)n1);
  for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
  std::vector<int> m12;
  ORBmatcher matcher(nnratio, checkOri != 0);
  const int r = matcher.SearchForInitialization(F1, F2, prev, m12, windowSize);
  for (int i = 0; i < n1; i++) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
  return r;
}

int ref_search_by_projection_last(const void* keys_cur, const uint8_t* desc_cur, int n_cur, const float* bounds, const float* Tcw,
                                  const float* K, const float* scaleFactors, int n_last, const uint8_t* last_valid, const float* last_pos,
                                  const uint8_t* last_desc, const int* last_octave, const float* last_angle, float th, int checkOri,
                                  const uint8_t* cur_preassigned, int* cur_match) {
  static Frame Cur, Last;
  Cur = Frame(); Last = Frame();
  World W;
  set_view(Cur, keys_cur, desc_cur, n_cur, bounds, scaleFactors);
  set_K(Cur, K);
  Cur.mTcw = pose44(Tcw);
  Cur.mvpMapPoints.assign((size_t)n_cur, nullptr);
  MP* pre = W.point(nullptr, nullptr, 1);                 // "already holds an observed map point"
  for (int i = 0; i < n_cur; i++) if (cur_preassigned && cur_preassigned[i]) Cur.mvpMapPoints[i] = pre;
  Last.N = n_last;
  Last.mTcw = pose44(Tcw);                               // only enters tlc (forward / backward, stereo), unused with bMono
  Last.mvKeys.resize((size_t)n_last); Last.mvKeysUn.resize((size_t)n_last);
  Last.mvpMapPoints.assign((size_t)n_last, nullptr);
  Last.mvbOutlier.assign((size_t)n_last, false);
  std::map<MapPoint*, int> index;
  for (int i = 0; i < n_last; i++) {
    Last.mvKeys[i].octave = Last.mvKeysUn[i].octave = last_octave[i];
    Last.mvKeys[i].angle = Last.mvKeysUn[i].angle = last_angle[i];
    if (last_valid[i]) { MP* p = W.point(last_pos + 3 * i, last_desc + 32 * i, 1); Last.mvpMapPoints[i] = p; index[p] = i; }
  }
  ORBmatcher matcher(0.9f, checkOri != 0);
  const int r = matcher.SearchByProjection(Cur, Last, th, true);
  for (int i = 0; i < n_cur; i++) {
    MapPoint* p = Cur.mvpMapPoints[i];
    cur_match[i] = !p ? -1 : (p == pre ? -2 : index[p]);
  }
  return r;
}

int ref_search_by_projection_points(const void* keys, const uint8_t* desc, int n, const float* bounds, const float* scaleFactors, int n_mp,
                                    const uint8_t* in_view, const float* proj, const int* level, const float* view_cos, const uint8_t* mp_desc,
                                    float th, float nnratio, const uint8_t* preassigned, int* match) {
  static Frame F;
  F = Frame();
  World W;
  set_view(F, keys, desc, n, bounds, scaleFactors);
  F.mvpMapPoints.assign((size_t)n, nullptr);
  MP* pre = W.point(nullptr, nullptr, 1);
  for (int i = 0; i < n; i++) if (preassigned && preassigned[i]) F.mvpMapPoints[i] = pre;
  std::vector<MapPoint*> mps((size_t)n_mp);
  std::map<MapPoint*, int> index;
  for (int i = 0; i < n_mp; i++) {
    MP* p = W.point(nullptr, mp_desc + 32 * i, 1);
    p->mbTrackInView = in_view[i] != 0; p->mnTrackScaleLevel = level[i]; p->mTrackViewCos = view_cos[i];
    p->mTrackProjX = proj[2 * i]; p->mTrackProjY = proj[2 * i + 1]; p->mTrackProjXR = -1;
    mps[i] = p; index[p] = i;
  }
  ORBmatcher matcher(nnratio, true);
  const int r = matcher.SearchByProjection(F, mps, th);
  for (int i = 0; i < n; i++) { MapPoint* p = F.mvpMapPoints[i]; match[i] = !p ? -1 : (p == pre ? -2 : index[p]); }
  return r;
}

int ref_search_for_triangulation(const void* keys1, const uint8_t* desc1, const uint8_t* has_mp1, int n1, const void* keys2, const uint8_t* desc2,
                                 const uint8_t* has_mp2, int n2, const unsigned* fv1_nodes, const int* fv1_start, const int* fv1_items, int nn1,
                                 const unsigned* fv2_nodes, const int* fv2_start, const int* fv2_items, int nn2, const float* F12, const float* Cw1,
                                 const float* R2w, const float* t2w, const float* K2, const float* scaleFactors2, const float* levelSigma2_2,
                                 int check_orientation, int* matches12) {
  static KeyFrame K1, Kf2;
  K1 = KeyFrame(); Kf2 = KeyFrame();
  World W;
  set_view(K1, keys1, desc1, n1, nullptr, nullptr);
  set_view(Kf2, keys2, desc2, n2, nullptr, scaleFactors2);
  Kf2.mvLevelSigma2.assign(levelSigma2_2, levelSigma2_2 + 8);
  set_K(Kf2, K2);
  MP* has = W.point(nullptr, nullptr, 1);
  K1.mvpMapPoints.assign((size_t)n1, nullptr); Kf2.mvpMapPoints.assign((size_t)n2, nullptr);
  for (int i = 0; i < n1; i++) if (has_mp1[i]) K1.mvpMapPoints[i] = has;
  for (int i = 0; i < n2; i++) if (has_mp2[i]) Kf2.mvpMapPoints[i] = has;
  set_featvec(K1.mFeatVec, fv1_nodes, fv1_start, fv1_items, nn1);
  set_featvec(Kf2.mFeatVec, fv2_nodes, fv2_start, fv2_items, nn2);
  K1.Ow = vec3(Cw1);
  Kf2.Tcw = cv::Mat(4, 4, CV_32F);
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) Kf2.Tcw.at<float>(r, c) = R2w[3 * r + c]; Kf2.Tcw.at<float>(r, 3) = t2w[r]; }
  for (int c = 0; c < 4; c++) Kf2.Tcw.at<float>(3, c) = c == 3 ? 1.f : 0.f;
  cv::Mat F(3, 3, CV_32F);
  for (int i = 0; i < 9; i++) F.at<float>(i / 3, i % 3) = F12[i];
  std::vector<std::pair<size_t, size_t>> pairs;
  ORBmatcher matcher(0.6f, check_orientation != 0);
  const int r = matcher.SearchForTriangulation(&K1, &Kf2, F, pairs, false);
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  for (auto& p : pairs) matches12[p.first] = (int)p.second;
  return r;
}

// ORBmatcher::Fuse(pKF, vpMapPoints, th) on a keyframe without map points: every hit is AddObservation + AddMapPoint, a second
// hit on the same keypoint is Replace (the point with fewer observations - the newcomer - is the one replaced).
// best_idx[i] = the keypoint the search chose for map point i (-1 none / skipped); returns nFused.
int ref_fuse_search(const void* keys, const uint8_t* desc, int n, const float* bounds, const float* Tcw, const float* Ow, const float* K,
                    const float* scaleFactors, const float* invLevelSigma2, float logScaleFactor, int nScaleLevels, int n_mp, const uint8_t* skip,
                    const float* pos, const float* normal, const float* minDist, const float* maxDist, const uint8_t* mp_desc, float th,
                    int* best_idx) {
  static KeyFrame KF;
  KF = KeyFrame();
  World W;
  set_view(KF, keys, desc, n, bounds, scaleFactors, nScaleLevels);
  set_K(KF, K);
  KF.mvInvLevelSigma2.assign(invLevelSigma2, invLevelSigma2 + nScaleLevels);
  KF.mfLogScaleFactor = logScaleFactor;
  KF.Tcw = pose44(Tcw);
  KF.Ow = vec3(Ow);
  KF.mvpMapPoints.assign((size_t)n, nullptr);
  std::vector<MapPoint*> mps((size_t)n_mp, nullptr);
  std::vector<MP*> own((size_t)n_mp, nullptr);
  for (int i = 0; i < n_mp; i++) {
    if (skip && skip[i]) continue;
    MP* p = W.point(pos + 3 * i, mp_desc + 32 * i, 0);
    p->set_normal(normal + 3 * i);
    p->set_dist(minDist[i], maxDist[i]);
    mps[i] = p; own[i] = p;
  }
  ORBmatcher matcher(0.6f, true);
  const int r = matcher.Fuse(&KF, mps, th);
  for (int i = 0; i < n_mp; i++) {
    best_idx[i] = -1;
    if (!own[i]) continue;
    MapPoint* p = own[i];
    if (p->IsInKeyFrame(&KF)) best_idx[i] = p->GetIndexInKeyFrame(&KF);
    else if (p->GetReplaced()) best_idx[i] = p->GetReplaced()->GetIndexInKeyFrame(&KF);
  }
  return r;
}

int ref_search_by_bow(const void* keysKF, const uint8_t* descKF, const uint8_t* has_mp_kf, int nKF, const void* keysF, const uint8_t* descF, int nF,
                      const unsigned* fvK_nodes, const int* fvK_start, const int* fvK_items, int nnK, const unsigned* fvF_nodes, const int* fvF_start,
                      const int* fvF_items, int nnF, float nnratio, int check_orientation, int* matchesF) {
  static KeyFrame KF; static Frame F;
  KF = KeyFrame(); F = Frame();
  World W;
  set_view(KF, keysKF, descKF, nKF, nullptr, nullptr);
  set_view(F, keysF, descF, nF, nullptr, nullptr);
  KF.mvpMapPoints.assign((size_t)nKF, nullptr);
  std::map<MapPoint*, int> index;
  for (int i = 0; i < nKF; i++) if (has_mp_kf[i]) { MP* p = W.point(nullptr, nullptr, 1); KF.mvpMapPoints[i] = p; index[p] = i; }
  set_featvec(KF.mFeatVec, fvK_nodes, fvK_start, fvK_items, nnK);
  set_featvec(F.mFeatVec, fvF_nodes, fvF_start, fvF_items, nnF);
  std::vector<MapPoint*> out;
  ORBmatcher matcher(nnratio, check_orientation != 0);
  const int r = matcher.SearchByBoW(&KF, F, out);
  for (int j = 0; j < nF; j++) matchesF[j] = out[j] ? index[out[j]] : -1;
  return r;
}

// ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (loop closing); mp1 / mp2: the keypoint holds a good map point
int ref_search_by_bow_keyframes(const void* keys1, const uint8_t* desc1, const uint8_t* mp1, int n1, const void* keys2, const uint8_t* desc2,
                                const uint8_t* mp2, int n2, const unsigned* fv1_nodes, const int* fv1_start, const int* fv1_items, int nn1,
                                const unsigned* fv2_nodes, const int* fv2_start, const int* fv2_items, int nn2, float nnratio, int check_orientation,
                                int* matches12) {
  static KeyFrame K1, K2;
  K1 = KeyFrame(); K2 = KeyFrame();
  World W;
  set_view(K1, keys1, desc1, n1, nullptr, nullptr);
  set_view(K2, keys2, desc2, n2, nullptr, nullptr);
  K1.mvpMapPoints.assign((size_t)n1, nullptr); K2.mvpMapPoints.assign((size_t)n2, nullptr);
  std::map<MapPoint*, int> index2;
  for (int i = 0; i < n1; i++) if (mp1[i]) K1.mvpMapPoints[i] = W.point(nullptr, nullptr, 1);
  for (int i = 0; i < n2; i++) if (mp2[i]) { MP* p = W.point(nullptr, nullptr, 1); K2.mvpMapPoints[i] = p; index2[p] = i; }
  set_featvec(K1.mFeatVec, fv1_nodes, fv1_start, fv1_items, nn1);
  set_featvec(K2.mFeatVec, fv2_nodes, fv2_start, fv2_items, nn2);
  std::vector<MapPoint*> out;
  ORBmatcher matcher(nnratio, check_orientation != 0);
  const int r = matcher.SearchByBoW(&K1, &K2, out);
  for (int i = 0; i < n1; i++) matches12[i] = out[i] ? index2[out[i]] : -1;
  return r;
}

// ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (relocalisation).  kf_valid[i]: the keyframe's i-th
// map point exists, is good and is not in sAlreadyFound.  Ow_out: the camera centre the function derives from Tcw (-Rcw^T tcw).
int ref_search_by_projection_keyframe(const void* keys_cur, const uint8_t* desc_cur, int n_cur, const float* bounds, const float* Tcw,
                                      const float* K, const float* scaleFactors, int nlevels, float logScaleFactor, int n_kf,
                                      const uint8_t* kf_valid, const float* pos, const uint8_t* mp_desc, const float* minDist, const float* maxDist,
                                      const float* kf_angle, float th, int ORBdist, int checkOri, const uint8_t* cur_preassigned, int* cur_match,
                                      float* Ow_out) {
  static Frame Cur; static KeyFrame KF;
  Cur = Frame(); KF = KeyFrame();
  World W;
  set_view(Cur, keys_cur, desc_cur, n_cur, bounds, scaleFactors, nlevels);
  set_K(Cur, K);
  Cur.mfLogScaleFactor = logScaleFactor;
  Cur.mTcw = pose44(Tcw);
  Cur.mvpMapPoints.assign((size_t)n_cur, nullptr);
  MP* pre = W.point(nullptr, nullptr, 1);
  for (int i = 0; i < n_cur; i++) if (cur_preassigned && cur_preassigned[i]) Cur.mvpMapPoints[i] = pre;
  KF.N = n_kf;
  KF.mvKeysUn.resize((size_t)n_kf); KF.mvKeys.resize((size_t)n_kf);
  KF.mvpMapPoints.assign((size_t)n_kf, nullptr);
  std::map<MapPoint*, int> index;
  for (int i = 0; i < n_kf; i++) {
    KF.mvKeysUn[i].angle = KF.mvKeys[i].angle = kf_angle[i];
    if (!kf_valid[i]) continue;
    MP* p = W.point(pos + 3 * i, mp_desc + 32 * i, 1);
    p->set_dist(minDist[i], maxDist[i]);
    KF.mvpMapPoints[i] = p; index[p] = i;
  }
  const std::set<MapPoint*> none;
  ORBmatcher matcher(0.9f, checkOri != 0);
  const int r = matcher.SearchByProjection(Cur, &KF, none, th, ORBdist);
  for (int i = 0; i < n_cur; i++) { MapPoint* p = Cur.mvpMapPoints[i]; cur_match[i] = !p ? -1 : (p == pre ? -2 : index[p]); }
  const cv::Mat Rcw = Cur.mTcw.rowRange(0, 3).colRange(0, 3), tcw = Cur.mTcw.rowRange(0, 3).col(3);
  const cv::Mat Ow = -Rcw.t() * tcw;
  for (int i = 0; i < 3; i++) Ow_out[i] = Ow.at<float>(i);
  return r;
}

// MapPoint::ComputeDistinctiveDescriptors for n_mp points; observation k of point m is row offsets[m] + k of desc.  The reference
// walks std::map<KeyFrame*, size_t>, i.e. keyframes in ADDRESS order: the keyframes of one point are consecutive elements of
// one array here, so that order is the row order.  chosen: n_mp x 32 bytes (the descriptor the point ends up with).
void ref_distinctive_descriptors(const uint8_t* desc, const int* offsets, int n_mp, uint8_t* chosen) {
  Map map;
  for (int m = 0; m < n_mp; m++) {
    const int N = offsets[m + 1] - offsets[m];
    std::vector<KeyFrame> kfs((size_t)std::max(N, 1));
    for (int k = 0; k < N; k++) {
      kfs[k].mnId = (unsigned long)k;
      kfs[k].mDescriptors = cv::Mat(1, 32, CV_8UC1);
      memcpy(kfs[k].mDescriptors.ptr(0), desc + 32 * (size_t)(offsets[m] + k), 32);
      kfs[k].mvuRight.assign(1, -1.f);
    }
    static const float zero[3] = {0, 0, 0};
    static const uint8_t none[32] = {0};
    MP p(vec3(zero), &kfs[0], &map);
    p.set_desc(none);
    for (int k = 0; k < N; k++) p.AddObservation(&kfs[k], 0);
    p.ComputeDistinctiveDescriptors();
    memcpy(chosen + 32 * (size_t)m, p.desc().ptr(0), 32);
  }
}

// MapPoint::PredictScale(currentDist, pKF) for n distances
void ref_predict_scale(const float* dist, const float* maxDist, int n, float logScaleFactor, int nScaleLevels, int* level) {
  Map map; KeyFrame kf;
  kf.mfLogScaleFactor = logScaleFactor; kf.mnScaleLevels = nScaleLevels;
  static const float zero[3] = {0, 0, 0};
  MP p(vec3(zero), &kf, &map);
  for (int i = 0; i < n; i++) { p.set_dist(maxDist[i] / 10.f, maxDist[i]); level[i] = p.PredictScale(dist[i], &kf); }
}
}
