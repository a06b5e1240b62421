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
}  // namespace cv
// type codes are MACROS in OpenCV (CV_8U is 0, CV_32F is 5): the stand-in defines the same macros, so host code spells them the same way
#ifndef CV_8U
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#endif
namespace cv {
// 8-bit or float single-channel matrix view / owner (all the hot path exchanges: images, descriptor rows, 4x4 poses, 3x1 points)
struct Mat {
  int rows = 0, cols = 0; size_t step = 0; uint8_t* data = nullptr; int type_ = CV_8UC1;
  std::vector<uint8_t> store;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* p, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c * (type == CV_32F ? 4 : 1)), data((uint8_t*)p), type_(type) {}
  Mat(const Mat& o) { *this = o; }
  Mat& operator=(const Mat& o) {      // owners copy their storage, views stay views (cv::Mat is reference counted: same observable behaviour here)
    rows = o.rows; cols = o.cols; step = o.step; type_ = o.type_; store = o.store;
    data = o.store.empty() ? o.data : store.data();
    return *this;
  }
  size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
  void create(int r, int c, int type) { rows = r; cols = c; type_ = type; step = (size_t)c * elemSize(); store.assign((size_t)r * step, 0); data = store.data(); }
  void release() { rows = cols = 0; step = 0; store.clear(); data = nullptr; }
  bool empty() const { return rows == 0 || cols == 0 || !data; }
  int type() const { return type_; }
  uint8_t* ptr(int r = 0) { return data + (size_t)r * step; }
  const uint8_t* ptr(int r = 0) const { return data + (size_t)r * step; }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
  template <typename T> T& at(int r, int c = 0) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c = 0) const { return ptr<T>(r)[c]; }
  Mat row(int r) const { return Mat(1, cols, type_, data + (size_t)r * step, step); }
  Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; r++) memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize()); return m; }
};
// Same shape as OpenCV's proxies (core/mat.hpp): only getMat / empty / create / release - NO ptr() - so code that compiles
// against the stand-in compiles against the real cv::InputArray / cv::OutputArray.
class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  Mat getMat() const { return Mat(m_->rows, m_->cols, m_->type(), m_->data, m_->step); }
  bool empty() const { return m_->empty(); }
 private:
  const Mat* m_;
};
class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void release() const { m_->release(); }
  Mat getMat() const { return Mat(m_->rows, m_->cols, m_->type(), m_->data, m_->step); }
 private:
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
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
