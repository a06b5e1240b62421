// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C ABI around the REFERENCE's own ORBextractor: /root/reference/src/ORBextractor.cc is compiled unmodified, where it lies,
// next to this file (oracle/Makefile, target `ref` -> oracle/_ref/libref_orb.so; nothing of the reference is copied into the
// repository).  The reference needs OpenCV, which this image does not have; oracle/shim/ declares the handful of OpenCV types
// and functions that one source file uses, and oracle/ref_cv_impl.cpp implements the image primitives behind them on the oracle's
// restatements (liboracle.so: oracle_resize_linear_u8, oracle_blur_u8, oracle_fast_detect, oracle_fast_atan2), which are pinned
// bit for bit to the cv2 4.13 wheel (tests/test_oracle_orb.py, tests/golden/orb_cv2_primitives.npz).
//
// What this pins: everything ORBextractor.cc itself does - constructor tables, ComputePyramid's geometry, the per-cell FAST
// loop with its threshold fallback, DistributeOctTree / DivideNode (including the pair<int, ExtractorNode*> sort whose ties
// break on heap addresses), IC_Angle, the steered-BRIEF descriptor, operator()'s level order and scaling - is the reference's
// own code.  What it does not pin: the OpenCV primitives (those are pinned to cv2 separately).
#include <opencv2/core/core.hpp>
#include <cstdint>
#include <cstdlib>
#include "ORBextractor.h"   // /root/reference/include (-I on the command line)

// ---- allocation order = address order ---------------------------------------------------------------------------------------
// DistributeOctTree sorts pair<int, ExtractorNode*> (ORBextractor.cc:684): nodes holding the same number of keypoints are ordered
// by their HEAP ADDRESS, which the C++ program does not define (with glibc's malloc the freed list nodes are handed out again
// last-in-first-out, so the order depends on the allocator's history).  ref_alloc.inc gives this library an operator new whose
// list-node-sized blocks have increasing addresses: one legal execution of the reference program, and the tie rule the oracle
// documents ("the later-created node counts as the larger pointer").  ref_orb_set_bump(0) restores malloc, for measuring how far
// glibc's order moves the result (tests/test_oracle_orb_ref.py).
#include "ref_alloc.inc"
extern "C" void ref_orb_set_bump(int on) { g_bump = on; }

extern "C" {
void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
  return new ORB_SLAM2::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void ref_orb_destroy(void* h) { delete (ORB_SLAM2::ORBextractor*)h; }
// ORBextractor::operator()(image, Mat(), keypoints, descriptors); returns the keypoint count (which may exceed cap)
int ref_orb_extract(void* h, const uint8_t* img, int w, int hh, int stride, void* kps, uint8_t* desc, int cap) {
  int n;
  ref_arena_begin(true);           // every arena block of the previous call is dead (they are all temporaries of operator())
  {
    cv::Mat image(hh, w, CV_8UC1, const_cast<uint8_t*>(img), (size_t)stride), d;
    std::vector<cv::KeyPoint> k;
    (*(ORB_SLAM2::ORBextractor*)h)(image, cv::Mat(), k, d);
    n = (int)k.size();
    const int m = n < cap ? n : cap;
    memcpy(kps, k.data(), sizeof(cv::KeyPoint) * (size_t)m);
    for (int i = 0; i < m; i++) memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
  }
  return ref_arena_end() ? -1 : n;
}
void ref_orb_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2) {
  ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
  const int n = e->GetLevels();
  std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(), dd = e->GetInverseScaleSigmaSquares();
  for (int i = 0; i < n; i++) { scale[i] = a[i]; invScale[i] = b[i]; sigma2[i] = c[i]; invSigma2[i] = dd[i]; }
}
// level image of the last extraction, without its border
int ref_orb_level(void* h, int l, uint8_t* out, int* w, int* hh) {
  const cv::Mat& m = ((ORB_SLAM2::ORBextractor*)h)->mvImagePyramid[l];
  *w = m.cols; *hh = m.rows;
  if (out) for (int y = 0; y < m.rows; y++) memcpy(out + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
  return 0;
}
}
