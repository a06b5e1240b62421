// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C ABI around the line-extraction sources of the reference tree, compiled unmodified and where they lie
// (oracle/Makefile, target `ref` -> oracle/_ref/libref_line.so; nothing of the reference is copied into the repository):
//   /root/reference/src/LineExtractor.cpp                                          LINEextractor::operator()
//   /root/reference/Thirdparty/line_descriptor/src/LSDDetector_custom.cpp         LSDDetectorC::detect      (KeyLines, §8 a10)
//   /root/reference/Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp   BinaryDescriptor::compute (LBD, §8 a11)
// (LineExtractor.cpp includes <opencv2/line_descriptor/descriptor.hpp> of opencv_contrib, which is absent from /root/reference;
// Thirdparty/line_descriptor is the vendored copy of the same module and the source the oracle cites.)
// OpenCV itself is not in this image: oracle/shim/ declares what these files use, and oracle/ref_cv_impl.cpp implements the image
// primitives behind it on the oracle's restatements (liboracle.so), pinned bit for bit to cv2 4.13: GaussianBlur 5x5 sigma 1,
// Sobel 3x3 8U -> 16S, the full LineSegmentDetector (ordered segment lists, tests/golden/lsd_cv2_*.npz).
//
// What this pins: the LBD band arithmetic, weights, normalisation and binary conversion, the KeyLine record construction, and
// LINEextractor's sort / nfeatures(+1) truncation / class ids / line equations are the reference tree's own code (the contrib
// header name and the three Eigen operations LineExtractor.cpp uses come from oracle/shim/).  Not pinned here: the OpenCV
// primitives (pinned to cv2 separately).
#include <opencv2/core/core.hpp>
#include <cstdint>
#include "line_descriptor_custom.hpp"   // /root/reference/Thirdparty/line_descriptor/include (-I on the command line)
#include "LineExtractor.h"              // /root/reference/include: the reference's LINEextractor (src/LineExtractor.cpp is compiled too)

using cv::line_descriptor::KeyLine;
static_assert(sizeof(KeyLine) == 68, "KeyLine layout");

extern "C" {
// LSDDetectorC::detect(image, keylines, scale, numOctaves, mask): returns the number of KeyLines (68-byte records)
int ref_lsd_keylines(const uint8_t* img, int w, int h, const uint8_t* mask, int scale, int num_octaves, void* keylines_out, int cap) {
  cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(img)), m;
  if (mask) m = cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(mask));
  std::vector<KeyLine> kls;
  cv::Ptr<cv::line_descriptor::LSDDetectorC> lsd = cv::line_descriptor::LSDDetectorC::createLSDDetectorC();
  lsd->detect(image, kls, scale, num_octaves, m);
  const int n = (int)kls.size();
  memcpy(keylines_out, kls.data(), sizeof(KeyLine) * (size_t)std::min(n, cap));
  return n;
}
// LINEextractor(1, 1.2f, nfeatures, min_line_length)(image, mask, keylines, descriptors, lineVec2d); returns the KeyLine count.
// Only defined for frames with MORE lines than nfeatures: otherwise the reference appends a default-constructed (uninitialised)
// KeyLine and describes it (LineExtractor.cpp:64), which is not reproducible.
int ref_line_extract(const uint8_t* img, int w, int h, const uint8_t* mask, int nfeatures, double min_line_length, void* keylines_out,
                     uint8_t* desc_out, double* linefunc_out, int cap) {
  cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(img)), m, d;
  if (mask) m = cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(mask));
  std::vector<KeyLine> kls;
  std::vector<Eigen::Vector3d> lf;
  ORB_SLAM2::LINEextractor ex(1, 1.2f, (unsigned)nfeatures, min_line_length);
  ex(image, m, kls, d, lf);
  const int n = (int)kls.size();
  if (n > cap || d.rows != n || (int)lf.size() != n) return -1;
  memcpy(keylines_out, kls.data(), sizeof(KeyLine) * (size_t)n);
  for (int i = 0; i < n; i++) { memcpy(desc_out + 32 * (size_t)i, d.ptr(i), 32); for (int j = 0; j < 3; j++) linefunc_out[3 * i + j] = lf[i](j); }
  return n;
}
// BinaryDescriptor::createBinaryDescriptor()->compute(image, keylines, descriptors[, returnFloatDescr]).
// desc: n x 32 bytes; desvec (optional): n x 72 floats (a second compute() with returnFloatDescr = true)
int ref_lbd_compute(const uint8_t* img, int w, int h, const void* keylines, int n, uint8_t* desc, float* desvec) {
  cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(img));
  std::vector<KeyLine> kls((const KeyLine*)keylines, (const KeyLine*)keylines + n);
  cv::Ptr<cv::line_descriptor::BinaryDescriptor> lbd = cv::line_descriptor::BinaryDescriptor::createBinaryDescriptor();
  cv::Mat d;
  lbd->compute(image, kls, d);
  if (d.rows != n || d.cols != 32) return -1;
  for (int i = 0; i < n; i++) memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
  if (desvec) {
    cv::Mat f;
    lbd->compute(image, kls, f, true);
    if (f.rows != n || f.cols != 72) return -2;
    for (int i = 0; i < n; i++) memcpy(desvec + 72 * (size_t)i, f.ptr(i), 72 * sizeof(float));
  }
  return n;
}
}
