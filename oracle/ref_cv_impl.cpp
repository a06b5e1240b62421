// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// The image primitives behind the OpenCV stand-in of oracle/shim/ (linked into every oracle/_ref/libref_*.so), implemented on the
// oracle's restatements (liboracle.so), which are pinned bit for bit to the cv2 4.13 wheel (tests/test_oracle_{orb,line,frame}.py,
// tests/golden/*cv2*.npz): resize 8U INTER_LINEAR, GaussianBlur 7x7 sigma 2 / 5x5 sigma 1, FAST 9/16 + NMS, copyMakeBorder
// REFLECT_101, fastAtan2, Sobel 3x3 8U -> 16S, the LineSegmentDetector, initUndistortRectifyMap + remap INTER_LINEAR,
// undistortPoints, BFMatcher::knnMatch (k = 2, Hamming).  Everything the compiled reference files never reach on the tested paths aborts.
#include <opencv2/core/core.hpp>
#include <cstdint>

extern "C" {   // liboracle.so
void oracle_resize_linear_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
void oracle_blur_u8(const uint8_t* src, int w, int h, uint8_t* dst, int ksize);
float oracle_fast_atan2(float y, float x);
int oracle_fast_detect(const uint8_t* img, int w, int h, int threshold, void* out, int cap);
void oracle_sobel3_u8(const uint8_t* img, int w, int h, int16_t* dx, int16_t* dy);
int oracle_lsd_detect(const uint8_t* img, int w, int h, int order_mode, float* lines, int cap);
void oracle_undistort_map(const float* K, const float* D, int w, int h, float* mx, float* my);
void oracle_remap(const uint8_t* src, int w, int h, const float* mx, const float* my, uint8_t* dst);
void oracle_undistort_keypoints(const void* kps, int n, const float* K, const float* D, void* out);
void oracle_bf_knn2(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* idx, int* dist);
}

namespace cv {

static std::vector<uchar> packed(const Mat& m) {
  std::vector<uchar> v((size_t)m.rows * m.cols + 1);
  for (int y = 0; y < m.rows; y++) memcpy(v.data() + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
  return v;
}
static void unpack(const std::vector<uchar>& v, Mat& m) {
  for (int y = 0; y < m.rows; y++) memcpy(m.ptr(y), v.data() + (size_t)y * m.cols, (size_t)m.cols);
}

float fastAtan2(float y, float x) { return oracle_fast_atan2(y, x); }

// cv::FAST(roi, kps, th, true): the ROI is an image of its own (no pixels outside it are read)
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  if (!nonmaxSuppression) abort();                       // ORBextractor.cc always asks for NMS
  keypoints.clear();
  if (image.rows < 7 || image.cols < 7) return;
  const std::vector<uchar> img = packed(image);
  std::vector<KeyPoint> out((size_t)image.rows * image.cols);
  const int n = oracle_fast_detect(img.data(), image.cols, image.rows, threshold, out.data(), (int)out.size());
  keypoints.assign(out.begin(), out.begin() + n);
}

// ORBextractor.cc:1086: GaussianBlur(m, m, Size(7, 7), 2, 2, BORDER_REFLECT_101); binary_descriptor_custom.cpp:358: GaussianBlur(m, m, Size(5, 5), 1)
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
  const bool k7 = ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2;
  const bool k5 = ksize.width == 5 && ksize.height == 5 && sigmaX == 1 && (sigmaY == 0 || sigmaY == 1);
  if (src.type() != CV_8UC1 || !(k7 || k5) || borderType != BORDER_REFLECT_101) abort();
  const std::vector<uchar> in = packed(src);
  std::vector<uchar> out(in.size());
  oracle_blur_u8(in.data(), src.cols, src.rows, out.data(), k7 ? 7 : 5);
  dst.create(src.rows, src.cols, CV_8UC1);
  unpack(out, dst);
}

void resize(const Mat& src, Mat& dst, Size dsize, double, double, int interpolation) {
  if (interpolation != INTER_LINEAR || dsize.width <= 0 || dsize.height <= 0) abort();   // (the fx / fy form is EDLine's, never reached)
  const std::vector<uchar> in = packed(src);
  std::vector<uchar> out((size_t)dsize.width * dsize.height + 1);
  oracle_resize_linear_u8(in.data(), src.cols, src.rows, out.data(), dsize.width, dsize.height);
  dst.create(dsize.height, dsize.width, CV_8UC1);      // keeps the pyramid ROI (same shape)
  unpack(out, dst);
}

static int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}
// copyMakeBorder(..., BORDER_REFLECT_101 [+ BORDER_ISOLATED]).  Both call sites of ORBextractor.cc:1122-1128 pass a source whose
// pixels outside the ROI must not be used (level > 0: ISOLATED; level 0: a whole image), so the border is always the reflection
// of the source itself.  The source may be the interior of dst (level > 0): read it through a packed copy.
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType) {
  if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) abort();
  const std::vector<uchar> in = packed(src);
  const int w = src.cols, h = src.rows;
  dst.create(h + top + bottom, w + left + right, CV_8UC1);
  for (int y = 0; y < dst.rows; y++) {
    const uchar* row = in.data() + (size_t)reflect101(y - top, h) * w;
    uchar* o = dst.ptr(y);
    for (int x = 0; x < dst.cols; x++) o[x] = row[reflect101(x - left, w)];
  }
}

// only ComputeKeyPointsOld() uses it, which operator() never calls (ORBextractor.cc:1057); defined so the file links
void KeyPointsFilter::retainBest(std::vector<KeyPoint>& keypoints, int npoints) {
  if (npoints < 0 || (int)keypoints.size() <= npoints) return;
  std::stable_sort(keypoints.begin(), keypoints.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
  keypoints.resize((size_t)npoints);
}

void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize) {
  if (src.type() != CV_8UC1 || ddepth != CV_16SC1 || ksize != 3 || dx + dy != 1) abort();   // binary_descriptor_custom.cpp:395-396
  const std::vector<uchar> in = packed(src);
  std::vector<int16_t> gx((size_t)src.rows * src.cols + 1), gy(gx.size());
  oracle_sobel3_u8(in.data(), src.cols, src.rows, gx.data(), gy.data());
  dst.create(src.rows, src.cols, CV_16SC1);
  const std::vector<int16_t>& g = dx ? gx : gy;
  for (int y = 0; y < dst.rows; y++) memcpy(dst.ptr(y), g.data() + (size_t)y * dst.cols, (size_t)dst.cols * 2);
}

// the reference runs one octave (LINEextractor passes numOctaves = 1) on grey images: never reached on the tested paths
void pyrDown(const Mat&, Mat&, Size) { fprintf(stderr, "ref: pyrDown (more than one octave) is not provided\n"); abort(); }
void cvtColor(const Mat&, Mat&, int) { fprintf(stderr, "ref: cvtColor is not provided (grey input only)\n"); abort(); }

namespace {
class Lsd : public LineSegmentDetector {
 public:
  void detect(const Mat& image, std::vector<Vec4f>& lines) override {
    const std::vector<uchar> in = packed(image);
    std::vector<float> out((size_t)4 * 65536);
    const int n = oracle_lsd_detect(in.data(), image.cols, image.rows, /*order_mode: cv2's*/ 1, out.data(), 65536);
    if (n > 65536) abort();
    lines.resize((size_t)n);
    for (int i = 0; i < n; i++) for (int k = 0; k < 4; k++) lines[i][k] = out[(size_t)4 * i + k];
  }
};
}  // namespace
Ptr<LineSegmentDetector> createLineSegmentDetector(int refine, double scale, double sigma_scale, double quant, double ang_th, double log_eps,
                                                   double density_th, int n_bins) {
  // LSD_REFINE_STD and the defaults: the only configuration the reference uses (LSDDetector_custom.cpp:150)
  if (refine != 1 || scale != 0.8 || sigma_scale != 0.6 || quant != 2.0 || ang_th != 22.5 || log_eps != 0 || density_th != 0.7 || n_bins != 1024) abort();
  return Ptr<LineSegmentDetector>(new Lsd());
}

// cv::LineIterator(img, Point2f, Point2f): the end points convert to Point by saturate_cast<int> (= cvRound), 8-connected count;
// both end points lie inside the image here (checkLineExtremes, LSDDetector_custom.cpp:77-103), so no clipping happens
LineIterator::LineIterator(const Mat&, Point2f p1, Point2f p2) {
  const int x0 = cvRound(p1.x), y0 = cvRound(p1.y), x1 = cvRound(p2.x), y1 = cvRound(p2.y);
  count = std::max(std::abs(x1 - x0), std::abs(y1 - y0)) + 1;
}

// ---- camera model (Frame.cc:220-222, 915-975): K 3x3 CV_32F, distCoef 4x1 or 5x1 CV_32F (k1 k2 p1 p2 [k3])
static void cam_arrays(const Mat& K, const Mat& D, float* k4, float* d5) {
  k4[0] = K.at<float>(0, 0); k4[1] = K.at<float>(1, 1); k4[2] = K.at<float>(0, 2); k4[3] = K.at<float>(1, 2);
  for (int i = 0; i < 5; i++) d5[i] = i < (int)D.total() ? D.at<float>(i) : 0.f;
}
void initUndistortRectifyMap(const Mat& K, const Mat& D, const Mat& R, const Mat& newK, Size size, int m1type, Mat& map1, Mat& map2) {
  if (m1type != CV_32F) abort();
  for (int y = 0; y < 3; y++) for (int x = 0; x < 3; x++) if (R.get(y, x) != (x == y ? 1.0 : 0.0) || newK.get(y, x) != K.get(y, x)) abort();   // R = I, newK = K
  float k4[4], d5[5];
  cam_arrays(K, D, k4, d5);
  map1.create(size.height, size.width, CV_32FC1); map2.create(size.height, size.width, CV_32FC1);
  oracle_undistort_map(k4, d5, size.width, size.height, map1.ptr<float>(0), map2.ptr<float>(0));
}
void remap(const Mat& src, Mat& dst, const Mat& map1, const Mat& map2, int interpolation) {
  if (interpolation != INTER_LINEAR || src.type() != CV_8UC1) abort();       // BORDER_CONSTANT 0, as cv::remap defaults
  const std::vector<uchar> in = packed(src);
  std::vector<uchar> out(in.size());
  oracle_remap(in.data(), src.cols, src.rows, map1.ptr<float>(0), map2.ptr<float>(0), out.data());
  dst.create(src.rows, src.cols, CV_8UC1);
  unpack(out, dst);
}
// cv::undistortPoints(src, dst, K, D, noArray(), P = K) on an N x 1 two-channel float matrix (in place in Frame.cc)
void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& D, const Mat& R, const Mat& P) {
  if (src.depth() != CV_32F || src.channels() != 2 || !R.empty()) abort();
  for (int y = 0; y < 3; y++) for (int x = 0; x < 3; x++) if (P.get(y, x) != K.get(y, x)) abort();
  float k4[4], d5[5];
  cam_arrays(K, D, k4, d5);
  const int n = src.rows * src.cols;
  std::vector<KeyPoint> in((size_t)n), out((size_t)n);
  for (int i = 0; i < n; i++) { const float* p = src.ptr<float>(i / src.cols) + 2 * (i % src.cols); in[i].pt.x = p[0]; in[i].pt.y = p[1]; }
  oracle_undistort_keypoints(in.data(), n, k4, d5, out.data());
  Mat r(src.rows, src.cols, src.type());
  for (int i = 0; i < n; i++) { float* p = r.ptr<float>(i / r.cols) + 2 * (i % r.cols); p[0] = out[i].pt.x; p[1] = out[i].pt.y; }
  dst = r;
}


// BFMatcher(NORM_HAMMING).knnMatch(query, train, matches, 2): per query the two nearest train rows, ties -> lower index (cv2-pinned);
// with fewer than two train rows OpenCV returns shorter lists, which the reference then indexes out of range ([i][1]): see the tests
void BFMatcher::knnMatch(const Mat& query, const Mat& train, std::vector<std::vector<DMatch>>& matches, int k) const {
  if (k != 2) abort();
  const int n1 = query.rows, n2 = train.rows;
  matches.assign((size_t)n1, std::vector<DMatch>());
  if (n1 == 0 || n2 < 2) abort();
  std::vector<uint8_t> q((size_t)n1 * 32), t((size_t)n2 * 32);
  for (int i = 0; i < n1; i++) memcpy(q.data() + 32 * (size_t)i, query.ptr(i), 32);
  for (int i = 0; i < n2; i++) memcpy(t.data() + 32 * (size_t)i, train.ptr(i), 32);
  std::vector<int> idx((size_t)n1 * 2), dist((size_t)n1 * 2);
  oracle_bf_knn2(q.data(), n1, t.data(), n2, idx.data(), dist.data());
  for (int i = 0; i < n1; i++)
    for (int j = 0; j < 2; j++) { DMatch m; m.queryIdx = i; m.trainIdx = idx[2 * i + j]; m.imgIdx = 0; m.distance = (float)dist[2 * i + j]; matches[i].push_back(m); }
}

}  // namespace cv
