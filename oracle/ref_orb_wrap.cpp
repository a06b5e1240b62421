// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// C ABI around the REFERENCE's own ORBextractor: /root/reference/src/ORBextractor.cc is compiled unmodified, where it lies,
// next to this file (oracle/Makefile, target `ref` -> oracle/_ref/libref_orb.so; nothing of the reference is copied into the
// repository).  The reference needs OpenCV, which this image does not have; oracle/shim/ declares the handful of OpenCV types
// and functions that one source file uses, and THIS file implements the image primitives behind them on the oracle's
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
// last-in-first-out, so the order depends on the allocator's history).  Inside this library (linked -Bsymbolic, so only the
// reference code compiled here is affected) operator new serves list-node-sized blocks from a bump arena that never reuses
// memory: a later allocation has a higher address.  That is one legal execution of the reference program, and it is the tie
// rule the oracle documents ("the later-created node counts as the larger pointer").  ref_orb_set_bump(0) restores malloc,
// for measuring how far glibc's order moves the result (tests/test_oracle_orb_ref.py).
#include <new>
#include <sys/mman.h>
static int g_bump = 1;                 // ref_orb_set_bump()
static bool g_in_call = false;         // only allocations made inside operator() go to the arena (all of them are temporaries of the call)
static char* g_base = nullptr;
static const size_t kArena = (size_t)64 << 20;
static size_t g_off = 0;
static int g_overflow = 0;
static inline bool in_arena(void* p) { return g_base && (char*)p >= g_base && (char*)p < g_base + kArena; }
static void* pl_alloc(size_t n) {
  if (g_bump && g_in_call && n >= 64 && n <= 128) {          // std::list<ExtractorNode> nodes are 88 bytes
    const size_t need = (n + 15) & ~(size_t)15;
    if (!g_base) {
      void* m = mmap(nullptr, kArena, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (m != MAP_FAILED) g_base = (char*)m;
    }
    if (g_base && g_off + need <= kArena) { void* r = g_base + g_off; g_off += need; return r; }
    g_overflow = 1;                                            // reported by ref_orb_extract (negative return)
  }
  void* r = malloc(n ? n : 1);
  if (!r) throw std::bad_alloc();
  return r;
}
void* operator new(size_t n) { return pl_alloc(n); }
void* operator new[](size_t n) { return pl_alloc(n); }
void operator delete(void* p) noexcept { if (p && !in_arena(p)) free(p); }
void operator delete[](void* p) noexcept { if (p && !in_arena(p)) free(p); }
void operator delete(void* p, size_t) noexcept { if (p && !in_arena(p)) free(p); }
void operator delete[](void* p, size_t) noexcept { if (p && !in_arena(p)) free(p); }
extern "C" void ref_orb_set_bump(int on) { g_bump = on; }

extern "C" {
void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
  return new ORB_SLAM2::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void ref_orb_destroy(void* h) { delete (ORB_SLAM2::ORBextractor*)h; }
// ORBextractor::operator()(image, Mat(), keypoints, descriptors); returns the keypoint count (which may exceed cap)
int ref_orb_extract(void* h, const uint8_t* img, int w, int hh, int stride, void* kps, uint8_t* desc, int cap) {
  int n;
  g_off = 0; g_overflow = 0;       // every arena block of the previous call is dead (they are all temporaries of operator())
  g_in_call = true;
  {
    cv::Mat image(hh, w, CV_8UC1, const_cast<uint8_t*>(img), (size_t)stride), d;
    std::vector<cv::KeyPoint> k;
    (*(ORB_SLAM2::ORBextractor*)h)(image, cv::Mat(), k, d);
    n = (int)k.size();
    const int m = n < cap ? n : cap;
    memcpy(kps, k.data(), sizeof(cv::KeyPoint) * (size_t)m);
    for (int i = 0; i < m; i++) memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
  }
  g_in_call = false;
  return g_overflow ? -1 : n;
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
