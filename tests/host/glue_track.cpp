// Drives reference_glue.cc - the reference's ORBmatcher / LSDmatcher / Optimizer signatures on mock Frame / MapPoint /
// MapLine objects - through Tracking::TrackWithMotionModel's call sequence (reference src/Tracking.cc:1331-1372):
//   matcher.SearchByProjection(mCurrentFrame, mLastFrame, th = 15, mono)      :1345
//   lmatcher.SearchByProjection(mCurrentFrame, mLastFrame, th)                :1347
//   Optimizer::PoseOptimization(&mCurrentFrame)                               :1372
// then Tracking::SearchLocalPoints' matchers on a "local map" (:1799, :1855) and LocalBundleAdjustmentWithLine on a small window.
// Everything the calls read is dumped as flat arrays next to what they wrote, so tests/test_host_cpp.py can recompute each
// call with the CPU oracle on identical inputs.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include "../../include/plslam_b200.h"
#include "../../pl-slam_b200/host/LineExtractor.h"
#include "../../pl-slam_b200/host/ORBextractor.h"
#include "../../pl-slam_b200/host/reference_glue.h"
using namespace ORB_SLAM2;

static std::vector<uint8_t> slurp(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> b((size_t)n);
  if (fread(b.data(), 1, b.size(), f) != b.size()) exit(2);
  fclose(f);
  return b;
}
static FILE* g_out;
template <typename T> static void put(const T* p, size_t n) { fwrite(p, sizeof(T), n, g_out); }
static void puti(int v) { put(&v, 1); }
static cv::Mat mat31(float a, float b, float c) { cv::Mat m(3, 1, CV_32F); m.at<float>(0) = a; m.at<float>(1) = b; m.at<float>(2) = c; return m; }

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: glue_track f1.raw f2.raw W H out.bin\n"); return 2; }
  const int W = atoi(argv[3]), H = atoi(argv[4]);
  std::vector<uint8_t> img[2] = {slurp(argv[1]), slurp(argv[2])};
  const float K[4] = {517.306408f, 516.469215f, 318.643040f, 255.313989f};              // Examples/Monocular/TUM1.yaml
  const float D[5] = {0.262383f, -0.953104f, -0.005358f, 0.002628f, 1.163314f};
  Frame::fx = K[0]; Frame::fy = K[1]; Frame::cx = K[2]; Frame::cy = K[3];
  float bounds[4];
  if (pl_frame_image_bounds(K, D, W, H, bounds) != PL_OK) return 3;
  Frame::mnMinX = bounds[0]; Frame::mnMinY = bounds[1]; Frame::mnMaxX = bounds[2]; Frame::mnMaxY = bounds[3];
  PLUndistort* und = nullptr;
  if (pl_undistort_create(K, D, W, H, &und) != PL_OK) return 3;
  ORBextractor orb(1000, 1.2f, 8, 20, 7);
  LINEextractor lsd(1, 1.2f, 200, 0.0);
  Frame F[2];                                                       // Frame::Frame (mono), src/Frame.cc:215-250
  cv::Mat none;
  for (int k = 0; k < 2; k++) {
    cv::Mat im(H, W, CV_8UC1, img[k].data()), u(H, W, CV_8UC1);
    orb(im, none, F[k].mvKeys, F[k].mDescriptors);
    if (pl_undistort_remap(und, im.ptr(0), W, u.ptr(0), W) != PL_OK) return 3;
    lsd(u, none, F[k].mvKeylinesUn, F[k].mLdesc, F[k].mvKeyLineFunctions);
    F[k].N = (int)F[k].mvKeys.size(); F[k].NL = (int)F[k].mvKeylinesUn.size();
    F[k].mvKeysUn.resize(F[k].N);
    if (F[k].N && pl_undistort_keypoints(und, (const PLKeyPoint*)F[k].mvKeys.data(), F[k].N, (PLKeyPoint*)F[k].mvKeysUn.data()) != PL_OK) return 3;
    F[k].mvuRight.assign(F[k].N, -1.f);
    F[k].mvpMapPoints.assign(F[k].N, nullptr); F[k].mvbOutlier.assign(F[k].N, false);
    F[k].mvpMapLines.assign(F[k].NL, nullptr); F[k].mvbLineOutlier.assign(F[k].NL, false);
    F[k].mvScaleFactors = orb.GetScaleFactors(); F[k].mvInvLevelSigma2 = orb.GetInverseScaleSigmaSquares();
    F[k].mnScaleLevels = 8; F[k].mfLogScaleFactor = logf(1.2f); F[k].mfLogScaleFactorLine = logf(1.2f);
    F[k].mnId = k;
  }
  Frame& Last = F[0];
  Frame& Cur = F[1];
  // the map seen by the last frame: every keypoint / keyline back-projected to a deterministic depth (last pose = identity)
  cv::Mat I(4, 4, CV_32F);
  for (int i = 0; i < 4; i++) I.at<float>(i, i) = 1.f;
  Last.mTcw = I; Last.mOw = mat31(0, 0, 0);
  std::vector<std::unique_ptr<MapPoint>> mps;
  std::vector<std::unique_ptr<MapLine>> mls;
  for (int i = 0; i < Last.N; i++) {
    if (i % 10 == 3) continue;                                      // keypoints without a map point
    const float z = 2.0f + (float)((i * 37) % 100) / 25.0f;
    auto mp = std::make_unique<MapPoint>();
    mp->mnId = (unsigned long)i;
    mp->mWorldPos = mat31((Last.mvKeysUn[i].pt.x - K[2]) / K[0] * z, (Last.mvKeysUn[i].pt.y - K[3]) / K[1] * z, z);
    mp->mDescriptor = Last.mDescriptors.row(i).clone();
    mp->nObs = (i % 7 == 0) ? 0 : 2;
    Last.mvpMapPoints[i] = mp.get();
    Last.mvbOutlier[i] = (i % 11 == 5);
    mps.push_back(std::move(mp));
  }
  for (int i = 0; i < Last.NL; i++) {
    const KeyLine& kl = Last.mvKeylinesUn[i];
    if (i % 9 == 4 || kl.lineLength == 0) continue;
    const float z1 = 2.5f + (float)((i * 13) % 50) / 20.0f, z2 = z1 + 0.1f * (float)((i % 5) - 2);
    auto ml = std::make_unique<MapLine>();
    ml->mnId = (unsigned long)i;
    const double s[3] = {(kl.startPointX - K[2]) / K[0] * z1, (kl.startPointY - K[3]) / K[1] * z1, z1};
    const double e[3] = {(kl.endPointX - K[2]) / K[0] * z2, (kl.endPointY - K[3]) / K[1] * z2, z2};
    ml->mWorldPos = Vector6d{s[0], s[1], s[2], e[0], e[1], e[2]};
    const double m[3] = {0.5 * (s[0] + e[0]), 0.5 * (s[1] + e[1]), 0.5 * (s[2] + e[2])};
    const double nm = std::sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    ml->mNormalVector = {m[0] / nm, m[1] / nm, m[2] / nm};
    ml->mfMaxDistance = (float)(nm * 1.3); ml->mfMinDistance = (float)(nm / 3.0);
    ml->mLDescriptor = Last.mLdesc.row(i).clone();
    ml->nObs = 2;
    Last.mvpMapLines[i] = ml.get();
    Last.mvbLineOutlier[i] = (i % 13 == 6);
    mls.push_back(std::move(ml));
  }
  // constant-velocity prediction of the current pose (Tracking.cc:1332): a small motion
  cv::Mat Tcw = I.clone();
  const float a = 0.004f;
  Tcw.at<float>(0, 0) = cosf(a); Tcw.at<float>(0, 2) = sinf(a); Tcw.at<float>(2, 0) = -sinf(a); Tcw.at<float>(2, 2) = cosf(a);
  Tcw.at<float>(0, 3) = 0.012f; Tcw.at<float>(1, 3) = -0.006f; Tcw.at<float>(2, 3) = 0.02f;
  Cur.SetPose(Tcw);
  // Frame::UpdatePoseMatrices: mOw = -Rcw^T tcw
  float Ow[3];
  for (int r = 0; r < 3; r++) { Ow[r] = 0; for (int k = 0; k < 3; k++) Ow[r] -= Tcw.at<float>(k, r) * Tcw.at<float>(k, 3); }
  Cur.mOw = mat31(Ow[0], Ow[1], Ow[2]);

  g_out = fopen(argv[5], "wb");
  if (!g_out) return 2;
  // ---- inputs, flat (what pytest feeds to the oracle)
  puti(Last.N); puti(Last.NL); puti(Cur.N); puti(Cur.NL);
  put(bounds, 4);
  for (int r = 0; r < 4; r++) put(Tcw.ptr<float>(r), 4);
  put(Ow, 3);
  put(Last.mvKeys.data(), Last.N); put(Last.mvKeysUn.data(), Last.N);
  for (int i = 0; i < Last.N; i++) put(Last.mDescriptors.ptr(i), 32);
  for (int i = 0; i < Last.N; i++) {
    MapPoint* p = Last.mvpMapPoints[i];
    const uint8_t valid = p && !Last.mvbOutlier[i], has = p != nullptr;
    float X[3] = {0, 0, 0};
    if (p) for (int k = 0; k < 3; k++) X[k] = p->mWorldPos.at<float>(k);
    put(&valid, 1); put(&has, 1); put(X, 3);
  }
  put(Last.mvKeylinesUn.data(), Last.NL);
  for (int i = 0; i < Last.NL; i++) put(Last.mLdesc.ptr(i), 32);
  for (int i = 0; i < Last.NL; i++) {
    MapLine* p = Last.mvpMapLines[i];
    const uint8_t cand = p && !Last.mvbLineOutlier[i], has = p != nullptr;
    double P[6] = {0}, n[3] = {0}; float md[2] = {0, 0};
    if (p) { for (int k = 0; k < 6; k++) P[k] = p->mWorldPos[k]; for (int k = 0; k < 3; k++) n[k] = p->mNormalVector[k]; md[0] = p->mfMinDistance; md[1] = p->mfMaxDistance; }
    put(&cand, 1); put(&has, 1); put(P, 6); put(n, 3); put(md, 2);
  }
  put(Cur.mvKeysUn.data(), Cur.N);
  for (int i = 0; i < Cur.N; i++) put(Cur.mDescriptors.ptr(i), 32);
  put(Cur.mvKeylinesUn.data(), Cur.NL);
  for (int i = 0; i < Cur.NL; i++) put(Cur.mLdesc.ptr(i), 32);
  for (int i = 0; i < Cur.NL; i++) put(Cur.mvKeyLineFunctions[i].data(), 3);

  // ---- Tracking::TrackWithMotionModel, src/Tracking.cc:1334-1372
  ORBmatcher matcher(0.9, true);
  LSDmatcher lmatcher;
  std::fill(Cur.mvpMapPoints.begin(), Cur.mvpMapPoints.end(), static_cast<MapPoint*>(nullptr));
  std::fill(Cur.mvpMapLines.begin(), Cur.mvpMapLines.end(), static_cast<MapLine*>(nullptr));
  const int th = 15;
  int nmatches = matcher.SearchByProjection(Cur, Last, (float)th, true);
  const int lmatches = lmatcher.SearchByProjection(Cur, Last, (float)th);
  if (nmatches < 20) {
    std::fill(Cur.mvpMapPoints.begin(), Cur.mvpMapPoints.end(), static_cast<MapPoint*>(nullptr));
    nmatches = matcher.SearchByProjection(Cur, Last, (float)(2 * th), true);
  }
  puti(nmatches); puti(lmatches);
  for (int i = 0; i < Cur.N; i++) puti(Cur.mvpMapPoints[i] ? (int)Cur.mvpMapPoints[i]->mnId : -1);
  for (int i = 0; i < Cur.NL; i++) puti(Cur.mvpMapLines[i] ? (int)Cur.mvpMapLines[i]->mnId : -1);
  const int inl = Optimizer::PoseOptimization(&Cur);
  puti(inl);
  for (int r = 0; r < 4; r++) put(Cur.mTcw.ptr<float>(r), 4);
  for (int i = 0; i < Cur.N; i++) { const uint8_t o = Cur.mvbOutlier[i]; put(&o, 1); }
  for (int i = 0; i < Cur.NL; i++) { const uint8_t o = Cur.mvbLineOutlier[i]; put(&o, 1); }
  fclose(g_out);
  pl_undistort_destroy(und);
  return 0;
}
