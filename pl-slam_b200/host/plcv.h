// Minimal stand-ins for the OpenCV / Eigen types that cross the hot-path boundary, used only when
// the host classes are built WITHOUT OpenCV (this repo's tests).  Inside the reference tree define
// PLSLAM_WITH_OPENCV and the real <opencv2/...> / <Eigen/Core> types are used instead; layouts are identical
// (cv::KeyPoint 28 B, KeyLine 68 B).
#pragma once
#ifdef PLSLAM_WITH_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/line_descriptor/descriptor.hpp>
#include <Eigen/Core>
namespace plcv = cv;
using cv::line_descriptor::KeyLine;
#else
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>
namespace cv {
struct Point2f { float x = 0, y = 0; };
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");
enum { CV_8UC1 = 0, CV_8U = 0 };
// 8-bit single-channel matrix view / owner (all the hot path exchanges)
struct Mat {
  int rows = 0, cols = 0; size_t step = 0; uint8_t* data = nullptr;
  std::vector<uint8_t> store;
  Mat() {}
  Mat(int r, int c, int /*type*/) { create(r, c, 0); }
  Mat(int r, int c, int, void* p, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c), data((uint8_t*)p) {}
  void create(int r, int c, int) { rows = r; cols = c; step = (size_t)c; store.assign((size_t)r * c, 0); data = store.data(); }
  void release() { rows = cols = 0; step = 0; store.clear(); data = nullptr; }
  bool empty() const { return rows == 0 || cols == 0 || !data; }
  int type() const { return CV_8UC1; }
  uint8_t* ptr(int r = 0) { return data + (size_t)r * step; }
  const uint8_t* ptr(int r = 0) const { return data + (size_t)r * step; }
  Mat getMat() const { Mat m(rows, cols, 0, data, step); return m; }
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
namespace line_descriptor {
struct KeyLine {
  float angle = 0; int class_id = 0; int octave = 0; Point2f pt; float response = 0; float size = 0;
  float startPointX = 0, startPointY = 0, endPointX = 0, endPointY = 0;
  float sPointInOctaveX = 0, sPointInOctaveY = 0, ePointInOctaveX = 0, ePointInOctaveY = 0;
  float lineLength = 0; int numOfPixels = 0;
};
static_assert(sizeof(KeyLine) == 68, "KeyLine layout");
}  // namespace line_descriptor
}  // namespace cv
using cv::line_descriptor::KeyLine;
namespace Eigen { typedef std::array<double, 3> Vector3d; }
#endif
