#include "LineExtractor.h"
#include <stdexcept>
#include <string>
#include "../../include/plslam_b200.h"
namespace ORB_SLAM2 {
LINEextractor::LINEextractor(int _numOctaves, float _scale, unsigned int _nLSDFeature, double _min_line_length)
    : numOctaves(_numOctaves), scale(_scale), nLSDFeature(_nLSDFeature), min_line_length(_min_line_length) {
  mvScaleFactor.resize(numOctaves); mvLevelSigma2.resize(numOctaves); mvInvScaleFactor.resize(numOctaves); mvInvLevelSigma2.resize(numOctaves);
  mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
  for (int i = 1; i < numOctaves; i++) { mvScaleFactor[i] = mvScaleFactor[i - 1] * scale; mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i]; }
  for (int i = 0; i < numOctaves; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
  // the detector receives (int)scale and numOctaves (LSDDetector_custom.cpp:105): every shipped config gives 1 and 1
  if (numOctaves != 1 || (int)scale != 1) throw std::runtime_error("plslam_b200: LINEextractor supports numOctaves == 1, (int)scale == 1");
}
LINEextractor::~LINEextractor() { pl_line_destroy(handle); }
void LINEextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<KeyLine>& _keylines, cv::OutputArray _descriptors,
                               std::vector<Eigen::Vector3d>& _lineVec2d) {
  if (_image.empty()) return;
  cv::Mat image = _image.getMat();
  cv::Mat mask = _mask.getMat();
  if (mask.data != nullptr && (mask.rows != image.rows || mask.cols != image.cols))
    throw std::runtime_error("Mask error while detecting lines: please check its dimensions and that data type is CV_8UC1");
  if (!handle || hw != image.cols || hh != image.rows) {
    pl_line_destroy(handle); handle = nullptr;
    PLLineConfig cfg = {image.cols, image.rows, (int)nLSDFeature, min_line_length, 1, 0, 0};
    if (pl_line_create(&cfg, &handle) != PL_OK) throw std::runtime_error(std::string("plslam_b200: ") + pl_last_error());
    hw = image.cols; hh = image.rows;
  }
  const int cap = pl_line_capacity(handle);
  std::vector<KeyLine> kl(cap);
  std::vector<uint8_t> desc((size_t)cap * 32);
  std::vector<double> lf((size_t)cap * 3);
  int n = 0;
  std::vector<uint8_t> packed;   // the C ABI wants a dense mask
  const uint8_t* mptr = nullptr;
  if (mask.data) { packed.resize((size_t)mask.rows * mask.cols); for (int r = 0; r < mask.rows; r++) memcpy(&packed[(size_t)r * mask.cols], mask.ptr(r), mask.cols); mptr = packed.data(); }
  if (pl_line_extract(handle, image.ptr(0), (int)image.step, mptr, kl.data(), desc.data(), lf.data(), &n) != PL_OK)
    throw std::runtime_error(std::string("plslam_b200: ") + pl_last_error());
  _keylines.assign(kl.begin(), kl.begin() + n);
  if (n == 0) { _descriptors.release(); return; }
  _descriptors.create(n, 32, CV_8U);
  cv::Mat d = _descriptors.getMat();
  for (int i = 0; i < n; i++) memcpy(d.ptr(i), &desc[(size_t)i * 32], 32);
  _lineVec2d.clear();
  for (int i = 0; i < n; i++) { Eigen::Vector3d v; v[0] = lf[3 * i]; v[1] = lf[3 * i + 1]; v[2] = lf[3 * i + 2]; _lineVec2d.push_back(v); }
}
}  // namespace ORB_SLAM2
