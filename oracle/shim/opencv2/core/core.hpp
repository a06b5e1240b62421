// ORACLE — TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the OpenCV headers the reference's ORBextractor.cc includes
// (/root/reference/src/ORBextractor.cc:57-60, include/ORBextractor.h:26), so that the reference's OWN source file compiles
// here, unmodified and where it lies, into oracle/_ref/libref_orb.so (recipe: oracle/Makefile target `ref`).
// This is not OpenCV: it declares exactly the types and functions that file uses.  The image primitives behind the
// declarations (resize, GaussianBlur, FAST, copyMakeBorder, fastAtan2) are implemented in oracle/ref_orb_wrap.cpp on top of
// the oracle's restatements, which are pinned bit for bit to cv2 4.13 (tests/test_oracle_orb.py).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

// OpenCV's cvRound is round-half-to-even (cvtsd2si / lrint), cvFloor / cvCeil are the exact integer floor / ceil
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename S> Point_& operator*=(S s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };

struct KeyPoint {   // 28 bytes, the layout of cv::KeyPoint
  Point2f pt; float size, angle, response; int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatZeros { int rows, cols, type; };   // what Mat::zeros() returns: assigning it to a Mat is create() + fill, as in OpenCV

// 8-bit single-channel matrix header over a shared buffer; headers made by rowRange / colRange / operator()(Rect) alias it
class Mat {
 public:
  int rows, cols; size_t step; uchar* data;
  Mat() : rows(0), cols(0), step(0), data(nullptr) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
  Mat(int r, int c, int, void* ext, size_t st = 0) : rows(r), cols(c), step(st ? st : (size_t)c), data((uchar*)ext) {}   // caller's memory
  Mat(const MatZeros& z) : Mat() { *this = z; }
  // Mat::create keeps the buffer when the shape already matches (the reference relies on it: resize() into a pyramid ROI,
  // `descriptors = Mat::zeros(...)` into a rowRange of the output, ORBextractor.cc:1037,1120)
  void create(int r, int c, int) {
    if (data && r == rows && c == cols) return;
    rows = r; cols = c; step = (size_t)c;
    buf_ = std::shared_ptr<uchar>(new uchar[(size_t)std::max(r, 0) * std::max(c, 0) + 1], std::default_delete<uchar[]>());
    data = buf_.get();
  }
  Mat& operator=(const MatZeros& z) {
    create(z.rows, z.cols, z.type);
    for (int y = 0; y < rows; y++) memset(data + (size_t)y * step, 0, (size_t)cols);
    return *this;
  }
  static MatZeros zeros(int r, int c, int type) { return MatZeros{r, c, type}; }
  Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
  Mat colRange(int a, int b) const { Mat m(*this); m.data = data + a; m.cols = b - a; return m; }
  Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
  Mat clone() const {
    Mat m(rows, cols, 0);
    for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols);
    return m;
  }
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  size_t step1() const { return step; }
  size_t elemSize() const { return 1; }
  int type() const { return CV_8UC1; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
 private:
  std::shared_ptr<uchar> buf_;
};

// InputArray / OutputArray: a view of the caller's Mat
class _InputArray {
 public:
  _InputArray() : m_(nullptr) {}
  _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
  bool empty() const { return !m_ || m_->empty(); }
  Mat getMat() const { return m_ ? *m_ : Mat(); }
 protected:
  Mat* m_;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() {}
  _OutputArray(Mat& m) : _InputArray(m) {}
  void create(int r, int c, int type) const { if (m_) m_->create(r, c, type); }
  void release() const { if (m_) m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline _InputArray noArray() { return _InputArray(); }

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

// implemented in oracle/ref_orb_wrap.cpp on the oracle's cv2-pinned primitives
float fastAtan2(float y, float x);
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType);
struct KeyPointsFilter { static void retainBest(std::vector<KeyPoint>& keypoints, int npoints); };

}  // namespace cv
