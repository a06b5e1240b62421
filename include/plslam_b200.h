/* plslam_b200 — C ABI of the B200-native PL-SLAM front-end / LM hot path.
 *
 * Every entry point replaces one C++ interface of the reference (HarborC/PL-SLAM); the
 * reference has no FFI of its own, so these are what a thin C++ class with the
 * reference's signature binds (see pl-slam_b200/host/ and INTEGRATION.md).
 *
 * Conventions: plain pointers and sizes; `_dev` variants take DEVICE pointers (inputs resident
 * in HBM) plus a cudaStream_t passed as void* (NULL = the handle's own stream) and are
 * asynchronous; the plain variants take HOST pointers, copy in/out and synchronise.
 * Return value: 0 = ok, <0 = error (pl_last_error() gives the text).  There is NO CPU fallback:
 * without a usable sm_100 device every compute entry point fails with PL_ERR_CUDA.
 */
#ifndef PLSLAM_B200_H
#define PLSLAM_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PL_OK 0
#define PL_ERR_ARG (-1)
#define PL_ERR_CUDA (-2)
#define PL_ERR_CAPACITY (-3)

const char* pl_last_error(void);
int pl_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
unsigned long long pl_launch_count(void);

/* ------------------------------------------------------------------ ORB extraction
 * replaces ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111,
 * src/ORBextractor.cc:410-470 ctor, :1043-1105 operator()).                         */
typedef struct PLKeyPoint { /* byte-compatible with cv::KeyPoint (28 B) */
  float x, y, size, angle, response;
  int32_t octave, class_id;
} PLKeyPoint;

typedef struct PLOrbConfig {
  int width, height;   /* frame size (fixed per handle)                               */
  int nfeatures;       /* ORBextractor.nFeatures                                      */
  float scale_factor;  /* ORBextractor.scaleFactor                                    */
  int nlevels;         /* ORBextractor.nLevels (<= 12)                                */
  int ini_th_fast;     /* ORBextractor.iniThFAST                                      */
  int min_th_fast;     /* ORBextractor.minThFAST                                      */
  int max_batch;       /* frames per call upper bound (device buffers are sized once) */
  int cell_slot_cap;   /* max NMS maxima kept per FAST cell; 0 = default 128          */
} PLOrbConfig;

typedef struct PLOrb PLOrb;

int pl_orb_create(const PLOrbConfig* cfg, PLOrb** out);
void pl_orb_destroy(PLOrb* h);
/* max keypoints one frame can return (nfeatures + 3 per level overshoot) */
int pl_orb_capacity(const PLOrb* h);
/* ORBextractor::Get{ScaleFactors,InverseScaleFactors,ScaleSigmaSquares,InverseScaleSigmaSquares},
 * mnFeaturesPerLevel and the level sizes; each array has nlevels entries (NULL = skip). */
int pl_orb_tables(const PLOrb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                  int* features_per_level, int* level_w, int* level_h);
/* ORBextractor::operator()(image, mask, keypoints, descriptors) for ONE host frame.
 * kps: capacity pl_orb_capacity(); desc: capacity*32 bytes; *n receives the count.       */
int pl_orb_extract(PLOrb* h, const uint8_t* img, int stride, PLKeyPoint* kps, uint8_t* desc, int* n);
/* B host frames (frame b at imgs + b*frame_stride); outputs are [B][capacity] arrays. */
int pl_orb_extract_batch(PLOrb* h, const uint8_t* imgs, int stride, size_t frame_stride, int B,
                         PLKeyPoint* kps, uint8_t* desc, int* n);
/* Same with every pointer a device pointer; asynchronous on `stream`. */
int pl_orb_extract_batch_dev(PLOrb* h, const uint8_t* imgs, int stride, size_t frame_stride, int B,
                             PLKeyPoint* kps, uint8_t* desc, int* n, void* stream);
/* ORBextractor::mvImagePyramid[level] of frame `frame` of the LAST call, copied to host;
 * with_border != 0 adds the 19-px BORDER_REFLECT_101 frame (reference ORBextractor.cc:1107-1132). */
int pl_orb_get_level(PLOrb* h, int frame, int level, uint8_t* out, int with_border);
/* Debug / parity taps of the LAST call: pre-quadtree FAST candidates of (frame, level) in the
 * reference's order, coordinates relative to the level's (16,16) detection origin.  Returns count. */
int pl_orb_debug_candidates(PLOrb* h, int frame, int level, PLKeyPoint* out, int cap);

#ifdef __cplusplus
}
#endif
#endif
