// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C ABI around the reference's own Frame.cc, compiled unmodified and where it lies against the reference's REAL Frame.h
// (oracle/Makefile, target `ref` -> oracle/_ref/libref_frame.so), together with the reference's ORBextractor.cc, LineExtractor.cpp,
// the vendored line-descriptor sources, MapPoint.cc, ORBmatcher.cc and lineIterator.cpp.  Shadowed with -I- (oracle/shim_frame/):
// KeyFrame.h / Map.h / MapLine.h (the mocks of oracle/shim_slam/), ORBVocabulary.h (DBoW2's vocabulary), Converter.h (g2o),
// LocalMapping.h (included, unused).  OpenCV is the stand-in of oracle/shim/ with the oracle's cv2-pinned primitives behind it
// (oracle/ref_cv_impl.cpp: initUndistortRectifyMap + remap, undistortPoints, resize, blur, FAST, LSD, Sobel).
//
// What runs here is the reference's monocular Frame constructor itself (Frame.cc:193-276): undistortion map + remap, ORB on the raw
// image and LSD / LBD on the undistorted one in two threads, UndistortKeyPoints, ComputeImageBounds, both grid assignments - and
// Frame::isInFrustum for points and lines, Frame::GetFeaturesInArea / GetFeaturesInAreaForLine.
#include <opencv2/core/core.hpp>
#include <cstdint>
#include <memory>
#include "Frame.h"          // /root/reference/include: the real one
#include "ORBmatcher.h"

#include "ref_alloc.inc"     // allocation order = address order for the quadtree's pointer ties (see there)

using namespace ORB_SLAM2;

namespace ORB_SLAM2 {
std::set<MapPoint*> KeyFrame::GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p); return s; }
}

namespace {
struct MP : MapPoint {
  MP(const cv::Mat& pos, KeyFrame* kf, Map* m) : MapPoint(pos, kf, m) {}
  void set_normal(const float* n) { mNormalVector = cv::Mat(3, 1, CV_32F); for (int i = 0; i < 3; i++) mNormalVector.at<float>(i) = n[i]; }
  void set_dist(float mn, float mx) { mfMinDistance = mn; mfMaxDistance = mx; }
};
cv::Mat vec3(const float* p) { cv::Mat m(3, 1, CV_32F); for (int i = 0; i < 3; i++) m.at<float>(i) = p[i]; return m; }
cv::Mat Kmat(const float* K) {
  cv::Mat m(3, 3, CV_32F);
  for (int i = 0; i < 9; i++) m.at<float>(i / 3, i % 3) = 0.f;
  m.at<float>(0, 0) = K[0]; m.at<float>(1, 1) = K[1]; m.at<float>(0, 2) = K[2]; m.at<float>(1, 2) = K[3]; m.at<float>(2, 2) = 1.f;
  return m;
}
cv::Mat Dmat(const float* D) {          // Tracking.cc:91-103: 4 x 1, resized to 5 x 1 when k3 != 0
  const int n = D[4] != 0.f ? 5 : 4;
  cv::Mat m(n, 1, CV_32F);
  for (int i = 0; i < n; i++) m.at<float>(i) = D[i];
  return m;
}
std::unique_ptr<Frame> g_frame;
std::unique_ptr<ORBextractor> g_orb;
std::unique_ptr<LINEextractor> g_line;
ORBVocabulary g_voc;
template <typename G> int csr(G& grid, int* start, int* items, int cap) {
  int k = 0;
  for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
    for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
      start[ix * FRAME_GRID_ROWS + iy] = k;
      for (size_t id : grid[ix][iy]) { if (k < cap) items[k] = (int)id; k++; }
    }
  start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = k;
  return k;
}
}  // namespace

extern "C" {

// Frame(imGray, timeStamp, orbextractor, lsdextractor, voc, K, distCoef, bf, thDepth, mask).  counts[0] = N, counts[1] = NL.
// Returns 0, or -1 when an output capacity is too small.  The frame stays alive for the ref_frame_* queries below.
int ref_frame_construct(const uint8_t* img, int w, int h, const uint8_t* mask, const float* K, const float* D, int nfeatures, float scaleFactor,
                        int nlevels, int iniTh, int minTh, int nlines, double min_line_length, int cap, int capl, int* counts, void* keys,
                        void* keysUn, uint8_t* desc, void* keylines, uint8_t* ldesc, double* lfunc, float* bounds, int* grid_start, int* grid_items,
                        int* lgrid_start, int* lgrid_items, int lgrid_cap, int* lgrid_n) {
  g_orb.reset(new ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh));
  g_line.reset(new LINEextractor(1, 1.2f, (unsigned)nlines, min_line_length));
  cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(img)), m, Kc = Kmat(K), Dc = Dmat(D);
  if (mask) m = cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(mask));
  Frame::mbInitialComputations = true;             // bounds, grid cell sizes and the static intrinsics are (re)computed for this camera
  ref_arena_begin(false);                          // never reset: small members of the previous Frame may still live in the arena
  Frame* nf = new Frame(image, 0.0, g_orb.get(), g_line.get(), &g_voc, Kc, Dc, 0.0f, 0.0f, m);
  if (ref_arena_end()) { delete nf; return -2; }
  g_frame.reset(nf);
  Frame& F = *g_frame;
  counts[0] = F.N; counts[1] = F.NL;
  if (F.N > cap || F.NL > capl) return -1;
  memcpy(keys, F.mvKeys.data(), sizeof(cv::KeyPoint) * (size_t)F.N);
  memcpy(keysUn, F.mvKeysUn.data(), sizeof(cv::KeyPoint) * (size_t)F.N);
  for (int i = 0; i < F.N; i++) memcpy(desc + 32 * (size_t)i, F.mDescriptors.ptr(i), 32);
  memcpy(keylines, F.mvKeylinesUn.data(), sizeof(KeyLine) * (size_t)F.NL);
  for (int i = 0; i < F.NL; i++) { memcpy(ldesc + 32 * (size_t)i, F.mLdesc.ptr(i), 32); for (int j = 0; j < 3; j++) lfunc[3 * i + j] = F.mvKeyLineFunctions[i](j); }
  bounds[0] = Frame::mnMinX; bounds[1] = Frame::mnMinY; bounds[2] = Frame::mnMaxX; bounds[3] = Frame::mnMaxY;
  csr(F.mGrid, grid_start, grid_items, cap);
  *lgrid_n = csr(F.mGridForLine, lgrid_start, lgrid_items, lgrid_cap);
  return *lgrid_n > lgrid_cap ? -1 : 0;
}
// Frame::GetFeaturesInArea / GetFeaturesInAreaForLine on the frame built last
int ref_frame_features_in_area(float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
  const std::vector<size_t> v = g_frame->GetFeaturesInArea(x, y, r, minLevel, maxLevel);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
  return (int)v.size();
}
int ref_frame_features_in_area_line(float x1, float y1, float x2, float y2, float r, float TH, int* out, int cap) {
  const std::vector<size_t> v = g_frame->GetFeaturesInAreaForLine(x1, y1, x2, y2, r, -1, -1, TH);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
  return (int)v.size();
}

static void set_camera(Frame& F, const float* Tcw16, const float* K, const float* bounds, float logScaleFactor, int nScaleLevels, float* Ow_out) {
  Frame::fx = K[0]; Frame::fy = K[1]; Frame::cx = K[2]; Frame::cy = K[3]; Frame::invfx = 1.0f / K[0]; Frame::invfy = 1.0f / K[1];
  Frame::mnMinX = bounds[0]; Frame::mnMinY = bounds[1]; Frame::mnMaxX = bounds[2]; Frame::mnMaxY = bounds[3];
  F.mnScaleLevels = nScaleLevels; F.mfLogScaleFactor = logScaleFactor; F.mfLogScaleFactorLine = logScaleFactor; F.mbf = 0.f;
  cv::Mat T(4, 4, CV_32F);
  for (int i = 0; i < 16; i++) T.at<float>(i / 4, i % 4) = Tcw16[i];
  F.SetPose(T);
  const cv::Mat Ow = F.GetCameraCenter();                 // -Rcw^T tcw as Frame::UpdatePoseMatrices computes it
  for (int i = 0; i < 3; i++) Ow_out[i] = Ow.at<float>(i);
}
// Frame::isInFrustum(MapPoint*, viewingCosLimit) for n points; Ow_out: the camera centre the frame derived from the pose
void ref_frame_is_in_frustum_points(const float* Tcw16, const float* K, const float* bounds, float logScaleFactor, int nScaleLevels, float cosLimit, int n,
                                    const float* pos, const float* normal, const float* minDist, const float* maxDist, uint8_t* inview, float* proj,
                                    int* level, float* viewcos, float* Ow_out) {
  Frame F; Map map; KeyFrame kf;
  set_camera(F, Tcw16, K, bounds, logScaleFactor, nScaleLevels, Ow_out);
  for (int i = 0; i < n; i++) {
    MP p(vec3(pos + 3 * i), &kf, &map);
    p.set_normal(normal + 3 * i); p.set_dist(minDist[i], maxDist[i]);
    inview[i] = F.isInFrustum(&p, cosLimit) ? 1 : 0;
    if (inview[i]) { proj[2 * i] = p.mTrackProjX; proj[2 * i + 1] = p.mTrackProjY; level[i] = p.mnTrackScaleLevel; viewcos[i] = p.mTrackViewCos; }
  }
}
// Frame::isInFrustum(MapLine*, viewingCosLimit): pos = 6 doubles, normal = 3 doubles per line
void ref_frame_is_in_frustum_lines(const float* Tcw16, const float* K, const float* bounds, float logScaleFactor, float cosLimit, int n, const double* pos,
                                   const double* normal, const float* minDist, const float* maxDist, uint8_t* inview, float* proj, int* level,
                                   float* viewcos, float* Ow_out) {
  Frame F;
  set_camera(F, Tcw16, K, bounds, logScaleFactor, 8, Ow_out);
  for (int i = 0; i < n; i++) {
    MapLine l;
    for (int j = 0; j < 6; j++) l.mWorldPos(j) = pos[6 * i + j];
    for (int j = 0; j < 3; j++) l.mNormalVector(j) = normal[3 * i + j];
    l.mfMinDistance = minDist[i]; l.mfMaxDistance = maxDist[i];
    inview[i] = F.isInFrustum(&l, cosLimit) ? 1 : 0;
    if (inview[i]) {
      proj[4 * i] = l.mTrackProjX1; proj[4 * i + 1] = l.mTrackProjY1; proj[4 * i + 2] = l.mTrackProjX2; proj[4 * i + 3] = l.mTrackProjY2;
      level[i] = l.mnTrackScaleLevel; viewcos[i] = l.mTrackViewCos;
    }
  }
}
}
