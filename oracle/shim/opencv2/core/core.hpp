// ORACLE — TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the OpenCV headers that a few reference source files include, so that
// those files compile here unmodified and where they lie (recipe: oracle/Makefile target `ref` -> oracle/_ref/*.so):
//   src/ORBextractor.cc, src/LineExtractor.cpp, src/ORBmatcher.cc, src/LSDmatcher.cpp, src/MapPoint.cc, src/Frame.cc, src/lineIterator.cpp,
//   Thirdparty/line_descriptor/src/{binary_descriptor_custom,LSDDetector_custom}.cpp, Thirdparty/DBoW2/DBoW2/{BowVector,FeatureVector}.cpp
// This is not OpenCV: it declares the types and functions those files use, nothing else.  cv::Mat arithmetic on CV_32F follows
// cv::gemm's small-matrix fp32 order (pinned to cv2 by tests/golden/frame_cv2.npz through the oracle's gemm3); the image primitives
// behind the declarations (resize, GaussianBlur, FAST, copyMakeBorder, fastAtan2, Sobel, LineSegmentDetector, initUndistortRectifyMap,
// remap, undistortPoints, BFMatcher::knnMatch) are implemented in oracle/ref_cv_impl.cpp on top of the oracle's restatements, which
// are pinned bit for bit to cv2 4.13 (tests/test_oracle_{orb,line,frame,match}.py); everything never executed on the tested paths
// (EDLine's helpers, colour conversion, pyrDown for more than one octave, SVD, the stereo matcher's helpers) aborts or is a plain loop.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef signed char schar;
typedef unsigned short ushort;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> 3) & 511) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8SC1 CV_MAKETYPE(CV_8S, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_WRAP
#define CV_OUT
#define CV_IN_OUT
#define CV_Assert(x) do { if (!(x)) { fprintf(stderr, "CV_Assert failed: %s\n", #x); abort(); } } while (0)

// OpenCV's cvRound is round-half-to-even (cvtsd2si / lrint), cvFloor / cvCeil are the exact integer floor / ceil
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

typedef std::string String;
using std::max; using std::min; using std::swap; using std::sqrt; using std::exp; using std::pow; using std::log;   // as cvstd.hpp does
template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename S> Point_& operator*=(S s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } static Scalar all(double v) { return Scalar(v, v, v, v); } };
template <typename T, int N> struct Vec { T val[N]; Vec() { for (int i = 0; i < N; i++) val[i] = T(); } T& operator[](int i) { return val[i]; } const T& operator[](int i) const { return val[i]; } };
typedef Vec<float, 4> Vec4f;
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {} };

struct KeyPoint {   // 28 bytes, the layout of cv::KeyPoint
  Point2f pt; float size, angle, response; int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatZeros { int rows, cols, type; };   // what Mat::zeros() returns: assigning it to a Mat is create() + fill, as in OpenCV
class _InputArray;
class _OutputArray;

// dense 2-D matrix header over a shared buffer; headers made by rowRange / colRange / operator()(Rect) alias it
class Mat {
 public:
  int rows, cols; size_t step; uchar* data;
  Mat() : rows(0), cols(0), step(0), data(nullptr), type_(0) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* ext, size_t st = 0) : rows(r), cols(c), step(0), data((uchar*)ext), type_(type) { step = st ? st : (size_t)c * elemSize(); }   // caller's memory
  Mat(const MatZeros& z) : Mat() { *this = z; }
  // Mat::create keeps the buffer when shape and type already match (the reference relies on it: resize() into a pyramid ROI,
  // `descriptors = Mat::zeros(...)` into a rowRange of the output, ORBextractor.cc:1037,1120)
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;
    rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
    buf_ = std::shared_ptr<uchar>(new uchar[(size_t)std::max(r, 0) * step + 16], std::default_delete<uchar[]>());
    data = buf_.get();
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  Mat& operator=(const MatZeros& z) {
    create(z.rows, z.cols, z.type);
    for (int y = 0; y < rows; y++) memset(data + (size_t)y * step, 0, (size_t)cols * elemSize());
    return *this;
  }
  static MatZeros zeros(int r, int c, int type) { return MatZeros{r, c, type}; }
  Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
  Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
  Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
  Mat row(int y) const { return rowRange(y, y + 1); }
  Mat col(int x) const { return colRange(x, x + 1); }
  // element i of a vector (3x1 or 1xN), as cv::Mat::at(int) addresses it
  template <typename T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
  template <typename T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
  Mat inv() const {                     // small dense inverse (Gauss-Jordan in fp64); LSDmatcher::ComputeF12 only, not on the tested paths
    const int n = rows;
    std::vector<double> a((size_t)n * 2 * n, 0.0);
    for (int y = 0; y < n; y++) { for (int x = 0; x < n; x++) a[(size_t)y * 2 * n + x] = get(y, x); a[(size_t)y * 2 * n + n + y] = 1.0; }
    for (int c = 0; c < n; c++) {
      int piv = c;
      for (int y = c + 1; y < n; y++) if (std::fabs(a[(size_t)y * 2 * n + c]) > std::fabs(a[(size_t)piv * 2 * n + c])) piv = y;
      for (int x = 0; x < 2 * n; x++) std::swap(a[(size_t)c * 2 * n + x], a[(size_t)piv * 2 * n + x]);
      const double d = a[(size_t)c * 2 * n + c];
      for (int x = 0; x < 2 * n; x++) a[(size_t)c * 2 * n + x] /= d;
      for (int y = 0; y < n; y++) if (y != c) { const double f = a[(size_t)y * 2 * n + c]; for (int x = 0; x < 2 * n; x++) a[(size_t)y * 2 * n + x] -= f * a[(size_t)c * 2 * n + x]; }
    }
    Mat m(n, n, type_);
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) m.put(y, x, a[(size_t)y * 2 * n + n + x]);
    return m;
  }
  Mat cross(const Mat& b) const {
    Mat m(rows, cols, type_);
    const double a0 = get(0, 0), a1 = cols == 1 ? get(1, 0) : get(0, 1), a2 = cols == 1 ? get(2, 0) : get(0, 2);
    const double b0 = b.get(0, 0), b1 = b.cols == 1 ? b.get(1, 0) : b.get(0, 1), b2 = b.cols == 1 ? b.get(2, 0) : b.get(0, 2);
    const double c[3] = {a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0};
    for (int i = 0; i < 3; i++) { if (cols == 1) m.put(i, 0, c[i]); else m.put(0, i, c[i]); }
    return m;
  }
  double dot(const Mat& m) const {     // cv::Mat::dot on CV_32F: products and sum in fp64, row-major order
    double s = 0;
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) s += get(y, x) * m.get(y, x);
    return s;
  }
  Mat clone() const { Mat m; copyTo(m); return m; }
  void copyTo(Mat& m) const {
    if (empty()) { m.release(); return; }
    m.create(rows, cols, type_);
    for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
  }
  void copyTo(Mat&& view) const {           // into an existing view of the same shape (Tcw.colRange(0, 3) as a destination)
    if (view.rows != rows || view.cols != cols || view.type() != type_) abort();
    for (int y = 0; y < rows; y++) memcpy(view.data + (size_t)y * view.step, data + (size_t)y * step, (size_t)cols * elemSize());
  }
  void copyTo(const _OutputArray& o) const;
  Mat reshape(int cn) const {               // same data, another channel count (N x 2 one-channel <-> N x 1 two-channel)
    Mat m(*this);
    const int per_row = cols * channels();
    if (per_row % cn) abort();
    m.type_ = CV_MAKETYPE(depth(), cn); m.cols = per_row / cn;
    return m;
  }
  void convertTo(Mat& m, int type) const { Mat r = convertedTo(type); m = r; }
  static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.setTo(1.0); return m; }
  Mat& setTo(double v) {
    for (int y = 0; y < rows; y++)
      for (int x = 0; x < cols * channels(); x++) put(y, x, v);
    return *this;
  }
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
  size_t elemSize1() const { static const int s[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)s[CV_MAT_DEPTH(type_)]; }
  size_t elemSize() const { return elemSize1() * (size_t)channels(); }
  size_t step1() const { return step / elemSize1(); }
  int type() const { return type_; }
  int depth() const { return CV_MAT_DEPTH(type_); }
  int channels() const { return CV_MAT_CN(type_); }
  Size size() const { return Size(cols, rows); }
  size_t total() const { return (size_t)rows * cols; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
  // element access as double, for the generic helpers below (single channel)
  double get(int y, int x) const {
    const uchar* p = data + (size_t)y * step + (size_t)x * elemSize1();
    switch (depth()) { case CV_8U: return *p; case CV_8S: return *(const schar*)p; case CV_16U: return *(const ushort*)p; case CV_16S: return *(const short*)p;
                       case CV_32S: return *(const int*)p; case CV_32F: return *(const float*)p; default: return *(const double*)p; }
  }
  void put(int y, int x, double v) {
    uchar* p = data + (size_t)y * step + (size_t)x * elemSize1();
    switch (depth()) { case CV_8U: *p = (uchar)std::min(255, std::max(0, cvRound(v))); break; case CV_8S: *(schar*)p = (schar)std::min(127, std::max(-128, cvRound(v))); break;
                       case CV_16U: *(ushort*)p = (ushort)std::min(65535, std::max(0, cvRound(v))); break; case CV_16S: *(short*)p = (short)std::min(32767, std::max(-32768, cvRound(v))); break;
                       case CV_32S: *(int*)p = cvRound(v); break; case CV_32F: *(float*)p = (float)v; break; default: *(double*)p = v; }
  }
  Mat convertedTo(int type) const {
    Mat m(rows, cols, type);
    for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) m.put(y, x, get(y, x));
    return m;
  }
  Mat t() const { Mat m(cols, rows, type_); for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) m.put(x, y, get(y, x)); return m; }
 protected:
  int type_;
  std::shared_ptr<uchar> buf_;
};
// Matrix arithmetic.  CV_32F operands compute in fp32, the product with cv::gemm's small-matrix order
// ((a0*b0 + a1*b1) + a2*b2, separately rounded; then "+ c" as its own rounding) - the order the oracle's gemm3 restates and
// tests/golden/frame_cv2.npz pins against cv2; other depths (EDLine's line fit, never on the tested paths) go through double.
inline Mat operator*(const Mat& a, const Mat& b) {
  Mat m(a.rows, b.cols, a.type());
  const bool f32 = a.depth() == CV_32F && b.depth() == CV_32F;
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < b.cols; x++) {
    if (f32) { float s = a.at<float>(y, 0) * b.at<float>(0, x); for (int k = 1; k < a.cols; k++) s = s + a.at<float>(y, k) * b.at<float>(k, x); m.at<float>(y, x) = s; }
    else { double s = 0; for (int k = 0; k < a.cols; k++) s += a.get(y, k) * b.get(k, x); m.put(y, x, s); }
  }
  return m;
}
#define PL_SHIM_ELEMWISE(NAME, EXPRF, EXPRD)                                                            \
  inline Mat NAME(const Mat& a, const Mat& b) {                                                         \
    Mat m(a.rows, a.cols, a.type());                                                                    \
    const bool f32 = a.depth() == CV_32F && b.depth() == CV_32F;                                        \
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) {                                 \
      if (f32) { const float p = a.at<float>(y, x), q = b.at<float>(y, x); m.at<float>(y, x) = EXPRF; } \
      else { const double p = a.get(y, x), q = b.get(y, x); m.put(y, x, EXPRD); }                       \
    }                                                                                                   \
    return m;                                                                                           \
  }
PL_SHIM_ELEMWISE(operator+, p + q, p + q)
PL_SHIM_ELEMWISE(operator-, p - q, p - q)
#undef PL_SHIM_ELEMWISE
// scaling: cv evaluates  alpha * M  through convertTo; for CV_32F the work type is float (cvtScale_<float, float, float>): v * (float)alpha
inline Mat operator*(double s, const Mat& a) {
  Mat m(a.rows, a.cols, a.type());
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) {
    if (a.depth() == CV_32F) m.at<float>(y, x) = a.at<float>(y, x) * (float)s; else m.put(y, x, a.get(y, x) * s);
  }
  return m;
}
inline Mat operator*(const Mat& a, double s) { return s * a; }
inline Mat operator-(const Mat& a) { return -1.0 * a; }
inline double norm(const Mat& a, const Mat& b, int type) {
  if (type != 2 /* NORM_L1 */) abort();
  double s = 0;
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) s += std::fabs(a.get(y, x) - b.get(y, x));
  return s;
}
inline double norm(const Mat& a) {      // NORM_L2: squares and sum in fp64
  double s = 0;
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) { const double v = a.get(y, x); s += v * v; }
  return std::sqrt(s);
}

inline Mat operator/(const Mat& a, double d) { return (1.0 / d) * a; }     // cv: M / s is M * (1 / s)
inline Mat& operator/=(Mat& a, double d) { Mat m = a / d; m.copyTo(a); return a; }

template <typename T> struct DataDepth;
template <> struct DataDepth<uchar> { enum { value = CV_8U }; };
template <> struct DataDepth<short> { enum { value = CV_16S }; };
template <> struct DataDepth<int> { enum { value = CV_32S }; };
template <> struct DataDepth<float> { enum { value = CV_32F }; };
template <> struct DataDepth<double> { enum { value = CV_64F }; };
template <typename T> class Mat_ : public Mat {
 public:
  Mat_() : Mat() { type_ = DataDepth<T>::value; }
  Mat_(int r, int c) : Mat(r, c, DataDepth<T>::value) {}
  Mat_(const Mat& m) : Mat() { *this = m; }
  static Mat_ eye(int r, int c) { Mat_ m(r, c); for (int y = 0; y < r; y++) for (int x = 0; x < c; x++) m[y][x] = (T)(x == y); return m; }
  Mat_& operator=(const Mat& m) {   // shares the data when the type matches, converts otherwise (as OpenCV does)
    if (m.type() == DataDepth<T>::value || m.empty()) { Mat::operator=(m); type_ = DataDepth<T>::value; }
    else Mat::operator=(m.convertedTo(DataDepth<T>::value));
    return *this;
  }
  T* operator[](int y) { return (T*)(data + (size_t)y * step); }
  const T* operator[](int y) const { return (const T*)(data + (size_t)y * step); }
};
// (Mat_<float>(3, 1) << a, b, c): row-major fill
template <typename T> struct MatCommaInitializer_ {
  Mat_<T> m; int i;
  MatCommaInitializer_(const Mat_<T>& m_, T v) : m(m_), i(0) { put(v); }
  void put(T v) { m[i / m.cols][i % m.cols] = v; i++; }
  template <typename S> MatCommaInitializer_& operator,(S v) { put((T)v); return *this; }
  operator Mat() const { return m; }
  operator Mat_<T>() const { return m; }
};
template <typename T, typename S> MatCommaInitializer_<T> operator<<(const Mat_<T>& m, S v) { return MatCommaInitializer_<T>(m, (T)v); }

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
  Mat* target() const { return m_; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline _InputArray noArray() { return _InputArray(); }
inline void Mat::copyTo(const _OutputArray& o) const { if (o.target()) copyTo(*o.target()); }

// persistence: declared because Algorithm / Params::read / write mention them; never used on the tested paths
class FileNode { public: FileNode operator[](const char*) const { return FileNode(); } operator int() const { return 0; } };
class FileStorage {};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
class Algorithm { public: virtual ~Algorithm() {} virtual void read(const FileNode&) {} virtual void write(FileStorage&) const {} };

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { COLOR_BGR2GRAY = 6 };
enum { THRESH_BINARY = 0, THRESH_TOZERO = 3 };
enum { CMP_EQ = 0, CMP_GT = 1, CMP_GE = 2, CMP_LT = 3, CMP_LE = 4, CMP_NE = 5 };
enum { NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };

// ---- implemented in oracle/ref_orb_wrap.cpp / oracle/ref_line_wrap.cpp on the oracle's cv2-pinned primitives
float fastAtan2(float y, float x);
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType);
void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize = 3);
void pyrDown(const Mat& src, Mat& dst, Size dstsize = Size());
void cvtColor(const Mat& src, Mat& dst, int code);
struct KeyPointsFilter { static void retainBest(std::vector<KeyPoint>& keypoints, int npoints); };
class LineSegmentDetector { public: virtual ~LineSegmentDetector() {} virtual void detect(const Mat& image, std::vector<Vec4f>& lines) = 0; };
Ptr<LineSegmentDetector> createLineSegmentDetector(int refine = 1, double scale = 0.8, double sigma_scale = 0.6, double quant = 2.0, double ang_th = 22.5,
                                                   double log_eps = 0, double density_th = 0.7, int n_bins = 1024);
void initUndistortRectifyMap(const Mat& K, const Mat& D, const Mat& R, const Mat& newK, Size size, int m1type, Mat& map1, Mat& map2);
void remap(const Mat& src, Mat& dst, const Mat& map1, const Mat& map2, int interpolation);
void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& D, const Mat& R, const Mat& P);
// cv::SVD::compute and norm(a, b, type): Frame::ComputeLine3D / ComputeStereoMatches only (stereo; never reached on the tested paths)
struct SVD { enum { MODIFY_A = 1, FULL_UV = 4 }; static void compute(const Mat&, Mat&, Mat&, Mat&, int = 0) { abort(); } };
// cv::LineIterator(img, p1, p2): only .count is used (LSDDetector_custom.cpp:184): 8-connected, end points rounded half-to-even
class LineIterator { public: LineIterator(const Mat& img, Point2f p1, Point2f p2); int count; };

// k-nearest-neighbour matching of binary descriptors (implemented in oracle/ref_lsd_wrap.cpp on the oracle's cv2-pinned bf_knn2)
class BFMatcher {
 public:
  BFMatcher(int normType = NORM_HAMMING, bool crossCheck = false) { if (normType != NORM_HAMMING || crossCheck) abort(); }
  void knnMatch(const Mat& query, const Mat& train, std::vector<std::vector<DMatch>>& matches, int k) const;
};
// debug drawing inside the reference's matchers: no-ops
#define CV_AA 16
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
inline bool imwrite(const String&, const Mat&) { return true; }

// element-wise helpers of EDLine's edge drawing (binary_descriptor_custom.cpp:1484-1492; not on the tested paths)
inline Mat abs(const Mat& a) { Mat m(a.rows, a.cols, a.type()); for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) m.put(y, x, std::fabs(a.get(y, x))); return m; }
inline void add(const Mat& a, const Mat& b, Mat& d) { d = a + b; }
inline double threshold(const Mat& s, Mat& d, double th, double, int type) {
  if (type != THRESH_TOZERO) abort();
  Mat m(s.rows, s.cols, s.type());
  for (int y = 0; y < s.rows; y++) for (int x = 0; x < s.cols; x++) { const double v = s.get(y, x); m.put(y, x, v > th ? v : 0); }
  d = m; return th;
}
inline void compare(const Mat& a, const Mat& b, Mat& d, int op) {
  if (op != CMP_LT) abort();
  Mat m(a.rows, a.cols, CV_8UC1);
  for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) m.put(y, x, a.get(y, x) < b.get(y, x) ? 255 : 0);
  d = m;
}

}  // namespace cv
