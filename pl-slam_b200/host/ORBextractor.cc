#include "ORBextractor.h"
#include <stdexcept>
#include <string>
#include "../../include/plslam_b200.h"
namespace ORB_SLAM2 {
ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
  // scale tables as ORBextractor.cc:410-431 (the device side computes the same tables; these serve the getters)
  mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
  mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) { mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor); mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i]; }
  for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
  mvImagePyramid.resize(nlevels);
}
ORBextractor::~ORBextractor() { pl_orb_destroy(handle); }
void ORBextractor::EnsureHandle(int width, int height) {
  if (handle && width == hw && height == hh) return;
  pl_orb_destroy(handle); handle = nullptr;
  PLOrbConfig cfg = {width, height, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, 1, 0};
  if (pl_orb_create(&cfg, &handle) != PL_OK) throw std::runtime_error(std::string("plslam_b200: ") + pl_last_error());
  hw = width; hh = height;
}
void ORBextractor::operator()(cv::InputArray _image, cv::InputArray, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors) {
  if (_image.empty()) return;                                    // ORBextractor.cc:1046-1047
  cv::Mat image = _image.getMat();
  EnsureHandle(image.cols, image.rows);
  const int cap = pl_orb_capacity(handle);
  std::vector<PLKeyPoint> kps(cap);
  std::vector<uint8_t> desc((size_t)cap * 32);
  int n = 0;
  if (pl_orb_extract(handle, image.ptr(0), (int)image.step, kps.data(), desc.data(), &n) != PL_OK)
    throw std::runtime_error(std::string("plslam_b200: ") + pl_last_error());
  _keypoints.resize(n);
  static_assert(sizeof(cv::KeyPoint) == sizeof(PLKeyPoint), "layout");
  if (n) memcpy((void*)_keypoints.data(), kps.data(), (size_t)n * sizeof(PLKeyPoint));
  if (n == 0) _descriptors.release();
  else {
    _descriptors.create(n, 32, CV_8U);
    cv::Mat d = _descriptors.getMat();          // cv::OutputArray has no ptr(): rows are written through the Mat it wraps
    for (int i = 0; i < n; i++) memcpy(d.ptr(i), &desc[(size_t)i * 32], 32);
  }
}
void ORBextractor::FetchImagePyramid() {
  if (!handle) return;
  std::vector<int> lw(nlevels), lh(nlevels);
  pl_orb_tables(handle, nullptr, nullptr, nullptr, nullptr, nullptr, lw.data(), lh.data());
  for (int l = 0; l < nlevels; l++) {
    mvImagePyramid[l].create(lh[l] + 38, lw[l] + 38, CV_8UC1);   // with the 19-px border, like ComputePyramid's `temp`
    pl_orb_get_level(handle, 0, l, mvImagePyramid[l].ptr(0), 1);
  }
}
}  // namespace ORB_SLAM2
