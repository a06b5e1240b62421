// Drives the C++ host classes the way Tracking drives the reference for a frame pair (Frame::Frame, Tracking::
// MonocularInitialization's matcher calls, Tracking::TrackWithMotionModel's PoseOptimization): two raw frames and one pose
// problem in, matches / pose out, for comparison with the oracle (tests/test_host_cpp.py).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pl-slam_b200/host/ORBextractor.h"
#include "../../pl-slam_b200/host/LineExtractor.h"
#include "../../pl-slam_b200/host/Matchers.h"
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
template <typename T> static void take(const uint8_t*& p, T* dst, size_t n) { memcpy(dst, p, n * sizeof(T)); p += n * sizeof(T); }
int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: host_pipeline f1.raw f2.raw W H problem.bin out.bin\n"); return 2; }
  const int W = atoi(argv[3]), H = atoi(argv[4]);
  std::vector<uint8_t> b1 = slurp(argv[1]), b2 = slurp(argv[2]), pb = slurp(argv[5]);
  const float K[4] = {517.306408f, 516.469215f, 318.643040f, 255.313989f};              // Examples/Monocular/TUM1.yaml
  const float D[5] = {0.262383f, -0.953104f, -0.005358f, 0.002628f, 1.163314f};
  ORBextractor orb(1000, 1.2f, 8, 20, 7);
  LINEextractor lsd(1, 1.2f, 200, 0.0);
  FrameUndistorter cam(K, D, W, H);
  FrameView F[2];
  cv::Mat none;
  for (int k = 0; k < 2; k++) {                                   // Frame::Frame (mono), src/Frame.cc:215-250
    cv::Mat im(H, W, CV_8UC1, k ? b2.data() : b1.data()), und;
    orb(im, none, F[k].mvKeys, F[k].mDescriptors);
    cam.remap(im, und);
    lsd(und, none, F[k].mvKeylinesUn, F[k].mLdesc, F[k].mvKeyLineFunctions);
    cam.UndistortKeyPoints(F[k]);
    cam.ComputeImageBounds(F[k]);
  }
  std::vector<cv::Point2f> prev(F[0].mvKeysUn.size());
  for (size_t i = 0; i < prev.size(); i++) prev[i] = F[0].mvKeysUn[i].pt;
  std::vector<int> m12, lm;
  ORBmatcher matcher(0.9f, true);
  const int nm = matcher.SearchForInitialization(F[0], F[1], prev, m12, 100);
  LSDmatcher lmatcher(0.7f);
  const int nlm = lmatcher.SearchDouble(F[0], F[1], lm);
  PoseProblem P;
  const uint8_t* p = pb.data();
  int np = 0, nl = 0;
  take(p, P.Tcw, 16); take(p, P.K, 4); take(p, &np, 1);
  P.pt_obs.resize(2 * np); P.pt_invSigma2.resize(np); P.pt_Xw.resize(3 * np);
  take(p, P.pt_obs.data(), 2 * np); take(p, P.pt_invSigma2.data(), np); take(p, P.pt_Xw.data(), 3 * np);
  take(p, &nl, 1);
  P.line_func.resize(3 * nl); P.line_Xw.resize(6 * nl);
  take(p, P.line_func.data(), 3 * nl); take(p, P.line_Xw.data(), 6 * nl);
  std::vector<bool> po, lo;
  const int inl = Optimizer::PoseOptimization(P, po, lo);
  FILE* o = fopen(argv[6], "wb");
  const int n1 = (int)m12.size(), nl1 = (int)lm.size(), n2 = (int)F[1].mvKeysUn.size();
  fwrite(&n1, 4, 1, o); fwrite(&nm, 4, 1, o); fwrite(m12.data(), 4, n1, o);
  fwrite(&nl1, 4, 1, o); fwrite(&nlm, 4, 1, o); fwrite(lm.data(), 4, nl1, o);
  fwrite(&n2, 4, 1, o); fwrite(F[1].mvKeysUn.data(), sizeof(cv::KeyPoint), n2, o);
  float bounds[4] = {F[1].mnMinX, F[1].mnMinY, F[1].mnMaxX, F[1].mnMaxY};
  fwrite(bounds, 4, 4, o);
  fwrite(&inl, 4, 1, o); fwrite(P.Tcw, 4, 16, o);
  for (int i = 0; i < np; i++) { uint8_t v = po[i]; fwrite(&v, 1, 1, o); }
  for (int i = 0; i < nl; i++) { uint8_t v = lo[i]; fwrite(&v, 1, 1, o); }
  fclose(o);
  printf("host_pipeline: %d point matches, %d line matches, %d inliers, d(ORB dist of first rows) = %d\n", nm, nlm, inl,
         F[0].mDescriptors.rows && F[1].mDescriptors.rows ? ORBmatcher::DescriptorDistance(F[0].mDescriptors, F[1].mDescriptors) : -1);
  return 0;
}
