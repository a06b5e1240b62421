// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C ABI around the reference's own LSDmatcher.cpp (+ lineIterator.cpp), compiled unmodified and where they lie into
// oracle/_ref/libref_match.so (oracle/Makefile, target `ref`), against the mocks of oracle/shim_slam/ (Frame / KeyFrame / MapLine:
// plain data; the line grid is filled by the reference's LineIterator, GetFeaturesInAreaForLine / GetLinesInArea are restated
// there from Frame.cc:768-842 / KeyFrame.cc:647-682) and the OpenCV stand-in (cv::BFMatcher::knnMatch below = the oracle's
// cv2-pinned bf_knn2, tests/golden/match_cv2_knn.npz).  What runs here is the reference's FrameBFMatch, lineDescriptorMAD
// (with auxiliar.h's own sort predicates), SearchDouble, both SearchByProjection overloads, SearchForTriangulation and Fuse
// (SURVEY §8 a16, a17, f.2).  Entry points take the same flat arrays as the oracle function of the same name (oracle_match.cpp).
#include <opencv2/core/core.hpp>
#include <cstdint>
#include <memory>
#include "LSDmatcher.h"     // /root/reference/include

using namespace ORB_SLAM2;

namespace {
struct FlatKL { float startX, startY, endX, endY, lineLength, angle; int octave; };     // oracle_match.cpp's flat KeyLine
KeyLine keyline(const FlatKL& f) {
  KeyLine k;
  memset(&k, 0, sizeof(k));
  k.startPointX = k.sPointInOctaveX = f.startX; k.startPointY = k.sPointInOctaveY = f.startY;
  k.endPointX = k.ePointInOctaveX = f.endX; k.endPointY = k.ePointInOctaveY = f.endY;
  k.lineLength = f.lineLength; k.angle = f.angle; k.octave = f.octave;
  k.pt = cv::Point2f((f.endX + f.startX) / 2, (f.endY + f.startY) / 2);
  return k;
}
cv::Mat desc_mat(const uint8_t* d, int n) {
  cv::Mat m(n, 32, CV_8UC1);
  if (n) memcpy(m.ptr(0), d, (size_t)n * 32);
  return m;
}
void set_lines(GridView& g, const FlatKL* k, const double* lfunc, const uint8_t* desc, int n, const float* bounds) {
  g.NL = n;
  g.mvKeyLines.resize((size_t)n);
  for (int i = 0; i < n; i++) g.mvKeyLines[i] = keyline(k[i]);
  g.mvKeylinesUn = g.mvKeyLines;
  g.mvKeyLineFunctions.resize((size_t)n);
  for (int i = 0; i < n; i++) for (int j = 0; j < 3; j++) g.mvKeyLineFunctions[i](j) = lfunc ? lfunc[3 * i + j] : 0.0;
  g.mLdesc = desc_mat(desc, n); g.mLineDescriptors = g.mLdesc;
  g.mvpMapLines.assign((size_t)n, nullptr);
  if (bounds) {
    g.mnMinX = bounds[0]; g.mnMinY = bounds[1]; g.mnMaxX = bounds[2]; g.mnMaxY = bounds[3];
    g.mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (g.mnMaxX - g.mnMinX);
    g.mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (g.mnMaxY - g.mnMinY);
    g.AssignFeaturesToGridForLine();
  }
}
struct Matcher : LSDmatcher {                 // reaches the protected FrameBFMatch
  Matcher(float r) : LSDmatcher(r, true) {}
  using LSDmatcher::FrameBFMatch;
};
cv::Mat eye4() { cv::Mat T(4, 4, CV_32F); for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T.at<float>(r, c) = r == c ? 1.f : 0.f; return T; }
}  // namespace

extern "C" {

void ref_frame_bf_match(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float TH, float nnratio, int* m) {
  std::vector<int> out;
  Matcher(nnratio).FrameBFMatch(desc_mat(d1, n1), desc_mat(d2, n2), out, TH);
  for (int i = 0; i < n1; i++) m[i] = out[i];
}

int ref_search_double(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnratio, int* matches) {
  static Frame F1, F2;
  F1 = Frame(); F2 = Frame();
  F1.NL = n1; F1.mLdesc = desc_mat(d1, n1);
  F2.NL = n2; F2.mLdesc = desc_mat(d2, n2);
  std::vector<int> out;
  LSDmatcher matcher(nnratio, true);
  const int r = matcher.SearchDouble(F1, F2, out);
  for (int i = 0; i < n1; i++) matches[i] = i < (int)out.size() ? out[i] : -1;
  return r;
}

int ref_line_search_by_projection_last(const void* kls_cur, const double* lfunc_cur, const uint8_t* desc_cur, int n_cur, const float* bounds,
                                       int n_last, const uint8_t* last_valid, const float* proj, const uint8_t* last_desc,
                                       const float* last_length, float th, const uint8_t* preassigned, int* cur_match) {
  static Frame Cur, Last;
  Cur = Frame(); Last = Frame();
  set_lines(Cur, (const FlatKL*)kls_cur, lfunc_cur, desc_cur, n_cur, bounds);
  Cur.mTcw = eye4(); Last.mTcw = eye4();
  std::vector<std::unique_ptr<MapLine>> own;
  MapLine pre; pre.nObs = 1;
  for (int i = 0; i < n_cur; i++) if (preassigned && preassigned[i]) Cur.mvpMapLines[i] = &pre;
  Last.NL = n_last;
  Last.mvKeylinesUn.resize((size_t)n_last); Last.mvpMapLines.assign((size_t)n_last, nullptr); Last.mvbLineOutlier.assign((size_t)n_last, false);
  std::map<MapLine*, int> index;
  for (int i = 0; i < n_last; i++) {
    memset(&Last.mvKeylinesUn[i], 0, sizeof(KeyLine));
    Last.mvKeylinesUn[i].lineLength = last_length[i];
    if (!last_valid[i]) continue;
    own.emplace_back(new MapLine());
    MapLine* p = own.back().get();
    p->nObs = 1; p->mbInFrustum = true; p->mLDescriptor = desc_mat(last_desc + 32 * (size_t)i, 1);
    p->mTrackProjX1 = proj[4 * i]; p->mTrackProjY1 = proj[4 * i + 1]; p->mTrackProjX2 = proj[4 * i + 2]; p->mTrackProjY2 = proj[4 * i + 3];
    Last.mvpMapLines[i] = p; index[p] = i;
  }
  LSDmatcher matcher(0.7f, true);
  const int r = matcher.SearchByProjection(Cur, Last, th);
  for (int i = 0; i < n_cur; i++) { MapLine* p = Cur.mvpMapLines[i]; cur_match[i] = !p ? -1 : (p == &pre ? -2 : index[p]); }
  return r;
}

int ref_line_search_by_projection_lines(const void* kls, const double* lfunc, const uint8_t* desc, int n, const float* bounds, int n_ml,
                                        const uint8_t* in_view, const float* proj, const float* view_cos, const uint8_t* ml_desc, float th,
                                        float nnratio, const uint8_t* preassigned, int* match) {
  static Frame F;
  F = Frame();
  set_lines(F, (const FlatKL*)kls, lfunc, desc, n, bounds);
  MapLine pre; pre.nObs = 1;
  for (int i = 0; i < n; i++) if (preassigned && preassigned[i]) F.mvpMapLines[i] = &pre;
  std::vector<std::unique_ptr<MapLine>> own;
  std::vector<MapLine*> mls((size_t)n_ml);
  std::map<MapLine*, int> index;
  for (int i = 0; i < n_ml; i++) {
    own.emplace_back(new MapLine());
    MapLine* p = own.back().get();
    p->nObs = 1; p->mbTrackInView = in_view[i] != 0; p->mTrackViewCos = view_cos[i]; p->mnTrackScaleLevel = 0;
    p->mLDescriptor = desc_mat(ml_desc + 32 * (size_t)i, 1);
    p->mTrackProjX1 = proj[4 * i]; p->mTrackProjY1 = proj[4 * i + 1]; p->mTrackProjX2 = proj[4 * i + 2]; p->mTrackProjY2 = proj[4 * i + 3];
    mls[i] = p; index[p] = i;
  }
  LSDmatcher matcher(nnratio, true);
  const int r = matcher.SearchByProjection(F, mls, th);
  for (int i = 0; i < n; i++) { MapLine* p = F.mvpMapLines[i]; match[i] = !p ? -1 : (p == &pre ? -2 : index[p]); }
  return r;
}

// LSDmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, isDouble); th must be the reference's TH_HIGH (80)
int ref_lsd_search_for_triangulation(const uint8_t* d1, const uint8_t* ml1, int n1, const uint8_t* d2, const uint8_t* ml2, int n2, float th,
                                     float nnratio, int isDouble, int* pairs) {
  if (th != (float)LSDmatcher::TH_HIGH) return -1;
  static KeyFrame K1, K2;
  K1 = KeyFrame(); K2 = KeyFrame();
  MapLine has; has.nObs = 1;
  K1.NL = n1; K1.mLineDescriptors = desc_mat(d1, n1); K1.mvKeyLines.resize((size_t)n1); K1.mvpMapLines.assign((size_t)n1, nullptr);
  K2.NL = n2; K2.mLineDescriptors = desc_mat(d2, n2); K2.mvKeyLines.resize((size_t)n2); K2.mvpMapLines.assign((size_t)n2, nullptr);
  for (int i = 0; i < n1; i++) if (ml1[i]) K1.mvpMapLines[i] = &has;
  for (int i = 0; i < n2; i++) if (ml2[i]) K2.mvpMapLines[i] = &has;
  std::vector<int> out;
  LSDmatcher matcher(nnratio, true);
  const int r = matcher.SearchForTriangulation(&K1, &K2, out, isDouble != 0);
  for (int i = 0; i < n1; i++) pairs[i] = i < (int)out.size() ? out[i] : -1;
  return r;
}

// LSDmatcher::Fuse(pKF, vpMapLines, th) on a keyframe without map lines: every hit is AddObservation + AddMapLine, a later hit on
// the same keyline is Replace.  best_idx[i] = keyline chosen for map line i (-1 none); *ret = Fuse's return value (0 = the
// `return false` on the first end point behind the camera, LSDmatcher.cpp:907).  keylines: 68-byte records.
void ref_lsd_fuse_search(const void* keylines, int nl, const uint8_t* kf_point_desc, int n_pdesc, const float* bounds, const float* Tcw,
                         const float* Ow, const float* K, float scale_line, int n_line_levels, float logScaleFactorLine, int n_ml,
                         const uint8_t* skip, const double* pos, const double* normal, const float* minDist, const float* maxDist,
                         const uint8_t* ml_desc, float th, int* best_idx, int* ret) {
  static KeyFrame KF;
  KF = KeyFrame();
  KF.NL = nl;
  KF.mvKeyLines.assign((const KeyLine*)keylines, (const KeyLine*)keylines + nl);
  KF.mvpMapLines.assign((size_t)nl, nullptr);
  KF.mDescriptors = desc_mat(kf_point_desc, n_pdesc);       // the reference compares against the POINT descriptors (LSDmatcher.cpp:966)
  KF.mLineDescriptors = desc_mat(kf_point_desc, 0);
  KF.mnMinX = bounds[0]; KF.mnMinY = bounds[1]; KF.mnMaxX = bounds[2]; KF.mnMaxY = bounds[3];
  KF.fx = K[0]; KF.fy = K[1]; KF.cx = K[2]; KF.cy = K[3];
  KF.mfLogScaleFactorLine = logScaleFactorLine; KF.mvScaleFactorsLine.scale = scale_line;
  KF.Tcw = cv::Mat(4, 4, CV_32F);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) KF.Tcw.at<float>(r, c) = Tcw[4 * r + c];
  for (int c = 0; c < 4; c++) KF.Tcw.at<float>(3, c) = c == 3 ? 1.f : 0.f;
  KF.Ow = cv::Mat(3, 1, CV_32F);
  for (int i = 0; i < 3; i++) KF.Ow.at<float>(i) = Ow[i];
  (void)n_line_levels;
  std::vector<std::unique_ptr<MapLine>> own;
  std::vector<MapLine*> mls((size_t)n_ml, nullptr);
  for (int i = 0; i < n_ml; i++) {
    if (skip && skip[i]) continue;
    own.emplace_back(new MapLine());
    MapLine* p = own.back().get();
    for (int j = 0; j < 6; j++) p->mWorldPos(j) = pos[6 * i + j];
    for (int j = 0; j < 3; j++) p->mNormalVector(j) = normal[3 * i + j];
    p->mfMinDistance = minDist[i]; p->mfMaxDistance = maxDist[i];
    p->mLDescriptor = desc_mat(ml_desc + 32 * (size_t)i, 1);
    mls[i] = p;
  }
  LSDmatcher matcher(0.7f, true);
  *ret = matcher.Fuse(&KF, mls, th);
  for (int i = 0; i < n_ml; i++) {
    best_idx[i] = -1;
    MapLine* p = mls[i];
    if (!p) continue;
    if (p->IsInKeyFrame(&KF)) best_idx[i] = p->GetIndexInKeyFrame(&KF);
    else if (p->mpReplaced) best_idx[i] = p->mpReplaced->GetIndexInKeyFrame(&KF);
  }
}
}
