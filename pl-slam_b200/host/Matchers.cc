#include "Matchers.h"
#include <stdexcept>
#include <string>
#include "../../include/plslam_b200.h"
namespace ORB_SLAM2 {
static void fail() { throw std::runtime_error(std::string("plslam_b200: ") + pl_last_error()); }
static std::vector<uint8_t> rows32(const cv::Mat& m) {
  std::vector<uint8_t> d((size_t)m.rows * 32);
  for (int i = 0; i < m.rows; i++) memcpy(&d[(size_t)i * 32], m.ptr(i), 32);
  return d;
}

FrameUndistorter::FrameUndistorter(const float K[4], const float distCoef[5], int width, int height) : w_(width), h_(height) {
  memcpy(K_, K, sizeof(K_)); memcpy(D_, distCoef, sizeof(D_));
  if (pl_undistort_create(K_, D_, width, height, &handle_) != PL_OK) fail();
}
FrameUndistorter::~FrameUndistorter() { pl_undistort_destroy(handle_); }
void FrameUndistorter::remap(const cv::Mat& imGray, cv::Mat& ImageGray) const {
  if (imGray.cols != w_ || imGray.rows != h_) throw std::runtime_error("plslam_b200: image size does not match the camera");
  ImageGray.create(h_, w_, CV_8UC1);
  if (pl_undistort_remap(handle_, imGray.ptr(0), (int)imGray.step, ImageGray.ptr(0), (int)ImageGray.step) != PL_OK) fail();
}
void FrameUndistorter::UndistortKeyPoints(FrameView& F) const {
  static_assert(sizeof(cv::KeyPoint) == sizeof(PLKeyPoint), "layout");
  F.mvKeysUn.resize(F.mvKeys.size());
  if (F.mvKeys.empty()) return;
  if (pl_undistort_keypoints(handle_, (const PLKeyPoint*)F.mvKeys.data(), (int)F.mvKeys.size(), (PLKeyPoint*)F.mvKeysUn.data()) != PL_OK) fail();
}
void FrameUndistorter::ComputeImageBounds(FrameView& F) const {
  float b[4];
  if (pl_frame_image_bounds(K_, D_, w_, h_, b) != PL_OK) fail();
  F.mnMinX = b[0]; F.mnMinY = b[1]; F.mnMaxX = b[2]; F.mnMaxY = b[3];
}

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
  int d = 0;
  if (pl_descriptor_distance_batch(a.ptr(0), b.ptr(0), 1, &d) != PL_OK) fail();
  return d;
}
int ORBmatcher::SearchForInitialization(FrameView& F1, FrameView& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                        std::vector<int>& vnMatches12, int windowSize) {
  const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
  vnMatches12.assign(n1, -1);
  if (n1 == 0 || n2 == 0) return 0;
  if ((int)vbPrevMatched.size() != n1) throw std::runtime_error("plslam_b200: vbPrevMatched must have F1.N entries");
  const float bounds[4] = {F2.mnMinX, F2.mnMinY, F2.mnMaxX, F2.mnMaxY};
  const std::vector<uint8_t> d1 = rows32(F1.mDescriptors), d2 = rows32(F2.mDescriptors);
  static_assert(sizeof(cv::Point2f) == 8, "layout");
  const int nm = pl_orb_search_for_initialization((const PLKeyPoint*)F1.mvKeysUn.data(), d1.data(), n1, (const PLKeyPoint*)F2.mvKeysUn.data(),
                                                  d2.data(), n2, bounds, (float*)vbPrevMatched.data(), vnMatches12.data(), windowSize,
                                                  mfNNratio, mbCheckOrientation ? 1 : 0);
  if (nm < 0) fail();
  return nm;
}

void LSDmatcher::FrameBFMatch(const cv::Mat& ldesc1, const cv::Mat& ldesc2, std::vector<int>& LineMatches, float TH) {
  LineMatches.assign(ldesc1.rows, -1);
  if (ldesc1.rows == 0) return;
  const std::vector<uint8_t> d1 = rows32(ldesc1), d2 = rows32(ldesc2);
  if (pl_lsd_frame_bf_match(d1.data(), ldesc1.rows, d2.data(), ldesc2.rows, TH, mfNNratio, LineMatches.data()) < 0) fail();
}
int LSDmatcher::SearchDouble(FrameView& InitialFrame, FrameView& CurrentFrame, std::vector<int>& LineMatches) {
  const int n1 = InitialFrame.mLdesc.rows, n2 = CurrentFrame.mLdesc.rows;
  LineMatches.assign(n1, -1);
  if (n1 == 0 || n2 == 0) return 0;
  const std::vector<uint8_t> d1 = rows32(InitialFrame.mLdesc), d2 = rows32(CurrentFrame.mLdesc);
  const int nm = pl_lsd_search_double(d1.data(), n1, d2.data(), n2, mfNNratio, LineMatches.data());
  if (nm < 0) fail();
  return nm;
}

int Optimizer::PoseOptimization(PoseProblem& P, std::vector<bool>& vbOutlier, std::vector<bool>& vbLineOutlier) {
  const int np = (int)P.pt_invSigma2.size(), nl = (int)P.line_func.size() / 3;
  std::vector<uint8_t> po(np + 1), lo(nl + 1);
  float Tout[16];
  static const float zf[3] = {0, 0, 0}; static const double zd[6] = {0, 0, 0, 0, 0, 0};
  const int n = pl_pose_optimization(0, P.Tcw, P.K, np, np ? P.pt_obs.data() : zf, np ? P.pt_invSigma2.data() : zf, np ? P.pt_Xw.data() : zf, nl,
                                     nl ? P.line_func.data() : zd, nl ? P.line_Xw.data() : zd, Tout, po.data(), lo.data(), nullptr);
  if (n < 0) fail();
  vbOutlier.assign(po.begin(), po.begin() + np); vbLineOutlier.assign(lo.begin(), lo.begin() + nl);
  if (np >= 3) memcpy(P.Tcw, Tout, sizeof(Tout));     // the reference leaves the pose untouched below 3 correspondences (Optimizer.cc:846)
  return n;
}
}  // namespace ORB_SLAM2
