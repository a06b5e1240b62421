// Line-feature extraction (LSD segments + LBD descriptors) for batches of frames on sm_100a.
//
// Replaces LINEextractor::operator() (reference src/LineExtractor.cpp:26-93) and what it calls:
//   LSDDetector::detect           opencv_contrib line_descriptor; spec copy Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:56-215
//   cv::LineSegmentDetector       OpenCV imgproc lsd.cpp (defaults: REFINE_STD, scale .8, sigma_scale .6, quant 2, 22.5 deg, density .7)
//   BinaryDescriptor::compute     spec copy Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:350-398,539-687,1026-1372
//
// Kernel map
//   k_lsd_scale     7x7 sigma .75 Gaussian (8.8 fixed point) fused with the 0.8x INTER_LINEAR_EXACT resize, smem tiles
//   k_lsd_grad      2x2 gradient -> one 16-byte record per pixel (angle, cos, sin, squared magnitude), per-frame max
//   k_lsd_hist/scan/scatter   stable counting sort of the defined pixels into 1024 magnitude bins (descending),
//                   equal bins keep row-major order == OpenCV 4.13's seed order (pinned in the oracle tests)
//   k_lsd_grow      region growing + rectangle fit + density refinement; inherently ordered (a pixel consumed by an
//                   earlier seed is unavailable to later ones) -> ONE warp per frame walks the seeds in order; the 32
//                   lanes test the 3x3 neighbourhood, evaluate angles and reduce the rectangle moments in parallel.
//   k_keylines      KeyLine records, mask filter, response sort (bitonic, ties keep detection order), truncation
//                   quirk of LineExtractor.cpp:44-67, normalised 2-D line equations
//   k_lbd_sobel     5x5 sigma 1 Gaussian (8.8 fixed point) fused with the 3x3 Sobel pair -> int16 dx, dy
//   k_lbd_describe  one CTA per line: 63 support rows in parallel (each row accumulates along the line in the
//                   reference's order, fp32 without FMA), band statistics, 72-float LBD, 32-byte binarisation

#include "common.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

namespace pl {

constexpr double kPI = 3.14159265358979323846;
constexpr double kDegToRads = kPI / 180;
constexpr int kBins = 1024;
constexpr int kChunkRows = 8;
constexpr int kRing = 512;

struct PLKeyLineRec {  // cv::line_descriptor::KeyLine, 68 bytes
  float angle; int class_id; int octave; float ptx, pty; float response; float size;
  float startPointX, startPointY, endPointX, endPointY;
  float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
  float lineLength; int numOfPixels;
};
static_assert(sizeof(PLKeyLineRec) == 68, "KeyLine layout");

struct LineParams {
  int w, h, sw, sh, npx;        // image, scaled image, sw*sh
  int nchunk;                   // ceil((sh-1)/kChunkRows)
  int s_th;                     // gradient defined  <=>  gx^2+gy^2 > s_th
  int min_reg_size;
  int seg_cap, capL, nfeatures;
  double min_line_length;
  double prec, prec_hi, p, density_th;   // prec_hi: see region_grow_t
};

// ---------------------------------------------------------------------------------------------- shared helpers
__device__ __forceinline__ float fast_atan2_deg_l(float y, float x) {  // cv::fastAtan2, no FMA
  const float k = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
  const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
  const float eps = 2.220446049250313e-16f;
  // ax >= ay ? ay/(ax+eps) : ax/(ay+eps)  ==  min/(max+eps): one division, no branch (this sits on the serial commit loop)
  const float ax = fabsf(x), ay = fabsf(y);
  const float c = __fdiv_rn(fminf(ax, ay), __fadd_rn(fmaxf(ax, ay), eps)), c2 = __fmul_rn(c, c);
  float a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  if (ax < ay) a = __fsub_rn(90.f, a);
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------- K_A scale
// Output tile 32x32 of the 0.8x image <- 40x40 blurred pixels <- 44x44 raw pixels (taps at +-3 are zero).
__global__ void __launch_bounds__(256) k_lsd_scale(LineParams P, const uint8_t* __restrict__ imgs, int stride,
                                                   long long frame_stride, uint8_t* __restrict__ scaled) {
  __shared__ uint8_t raw[44][48];
  __shared__ uint16_t hp[44][40];
  __shared__ uint8_t bl[40][40];
  const int tid = threadIdx.x;
  const int X0 = blockIdx.x * 32, Y0 = blockIdx.y * 32;
  const int bx0 = (5 * X0) >> 2, by0 = (5 * Y0) >> 2;
  const uint8_t* img = imgs + (long long)blockIdx.z * frame_stride;
  for (int i = tid; i < 44 * 44; i += 256) {
    int r = i / 44, c = i - r * 44;
    int gy = reflect101(min(by0 - 2 + r, 2 * P.h - 2), P.h), gx = reflect101(min(bx0 - 2 + c, 2 * P.w - 2), P.w);
    raw[r][c] = img[(long long)gy * stride + gx];
  }
  __syncthreads();
  for (int i = tid; i < 44 * 40; i += 256) {
    int r = i / 40, c = i - r * 40;
    const uint8_t* p = &raw[r][c];
    hp[r][c] = (uint16_t)(4 * (p[0] + p[4]) + 56 * (p[1] + p[3]) + 136 * p[2]);
  }
  __syncthreads();
  for (int i = tid; i < 40 * 40; i += 256) {
    int r = i / 40, c = i - r * 40;
    uint32_t s = 4u * (hp[r][c] + hp[r + 4][c]) + 56u * (hp[r + 1][c] + hp[r + 3][c]) + 136u * hp[r + 2][c];
    bl[r][c] = (uint8_t)((s + 32768u) >> 16);
  }
  __syncthreads();
  uint8_t* out = scaled + (long long)blockIdx.z * P.npx;
  for (int i = tid; i < 32 * 32; i += 256) {
    int ty = i >> 5, tx = i & 31;
    int x = X0 + tx, y = Y0 + ty;
    if (x >= P.sw || y >= P.sh) continue;
    int sx = (10 * x + 1) >> 3, xf = ((10 * x + 1) & 7) * 32;
    int sy = (10 * y + 1) >> 3, yf = ((10 * y + 1) & 7) * 32;
    if (sx >= P.w - 1) { sx = P.w - 1; xf = 0; }
    if (sy >= P.h - 1) { sy = P.h - 1; yf = 0; }
    int sx1 = min(sx + 1, P.w - 1), sy1 = min(sy + 1, P.h - 1);
    int lx = sx - bx0, lx1 = sx1 - bx0, ly = sy - by0, ly1 = sy1 - by0;
    int h0 = bl[ly][lx] * (256 - xf) + bl[ly][lx1] * xf;
    int h1 = bl[ly1][lx] * (256 - xf) + bl[ly1][lx1] * xf;
    out[(long long)y * P.sw + x] = (uint8_t)((h0 * (256 - yf) + h1 * yf + 32768) >> 16);
  }
}

// ---------------------------------------------------------------------------------------------- K_B gradient
// Per scaled pixel, what region growing needs, as separate arrays (the 4-byte angle word is the hot one):
//   ANG = level-line angle in degrees (cv::fastAtan2(gx, -gy)); -1024 = NOTDEF (border or magnitude <= rho)
//   CS  = cos/sin of float(angle_rad) rounded to fp32 (what region_grow adds to sumdx/sumdy)
//   S2  = s = gx^2+gy^2 (modgrad = sqrt(s/4), recomputed in fp64 where the weights are used); seedcs: see grad_record
constexpr float kNotDefDeg = -1024.f;
// The level-line record of a pixel depends only on its integer gradient (gx, gy) in [-510, 510]^2: the angle in degrees
// (cv::fastAtan2(gx, -gy)), cos/sin of that angle as region_grow adds them, and the cos/sin a region SEEDED there starts
// from.  The fp64 sincos behind them is the whole cost of the gradient pass, so it is evaluated once per (gx, gy) into a
// 25 MB table at handle creation (L2-resident, the hot entries are the small gradients) and the per-frame kernel is loads.
constexpr int kGradR = 510, kGradN = 2 * kGradR + 1;
struct GradRec { float deg, c, s, pad; };
__device__ __forceinline__ void grad_record(int gx, int gy, GradRec& rec, float2& scs) {
  const float deg = fast_atan2_deg_l((float)gx, (float)(-gy));
  const double af = (double)(float)((double)deg * kDegToRads);
  double sn, cs;
  sincos(af, &sn, &cs);
  rec.deg = deg; rec.c = (float)cs; rec.s = (float)sn; rec.pad = 0.f;
  // a region SEEDED here starts from cos/sin of the fp64 angle ad = af + dl, |dl| <= half an fp32 ulp (< 4e-7):
  // angle-addition with cos(dl) = 1 - dl^2/2, sin(dl) = dl is exact to ~1e-27, far below fp64 rounding
  const double ad = (double)deg * kDegToRads, dl = ad - af, h2 = 1.0 - 0.5 * dl * dl;
  scs = make_float2((float)(cs * h2 - sn * dl), (float)(sn * h2 + cs * dl));
}
__global__ void __launch_bounds__(256) k_lsd_grad_table(GradRec* __restrict__ T, float2* __restrict__ TS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kGradN * kGradN) return;
  GradRec rec; float2 scs;
  grad_record(i / kGradN - kGradR, i % kGradN - kGradR, rec, scs);
  T[i] = rec; TS[i] = scs;
}
// One thread = 4 consecutive pixels of a row: two 8-byte row reads, four table lookups, 16-byte stores (the pass is bound by
// the number of memory instructions, not by bytes).  kVec needs sw % 4 == 0 (rows of every output array 16-byte aligned).
template <bool kVec>
__global__ void __launch_bounds__(256) k_lsd_grad(LineParams P, const uint8_t* __restrict__ scaled, const float4* __restrict__ T,
                                                  const float2* __restrict__ TS,
                                                  float* __restrict__ ANG, float2* __restrict__ CS, int* __restrict__ S2, float2* __restrict__ seedcs, int* __restrict__ maxs) {
  __shared__ int smax;
  if (threadIdx.x == 0) smax = 0;
  __syncthreads();
  const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int f = blockIdx.z;
  int smx = 0;
  if (x0 < P.sw && y < P.sh) {
    const uint8_t* S = scaled + (long long)f * P.npx + (long long)y * P.sw + x0;
    const bool lastrow = (y >= P.sh - 1);
    int r0[5], r1[5];
    if (kVec) {       // x0 % 4 == 0 and sw % 4 == 0: one aligned word + one byte per row
      const unsigned w0 = *reinterpret_cast<const unsigned*>(S), w1 = lastrow ? 0u : *reinterpret_cast<const unsigned*>(S + P.sw);
#pragma unroll
      for (int k = 0; k < 4; k++) { r0[k] = (w0 >> (8 * k)) & 0xff; r1[k] = (w1 >> (8 * k)) & 0xff; }
      const bool in4 = (x0 + 4 < P.sw);
      r0[4] = in4 ? S[4] : 0; r1[4] = (in4 && !lastrow) ? S[P.sw + 4] : 0;
    } else {
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const bool in = (x0 + k < P.sw);
        r0[k] = in ? S[k] : 0;
        r1[k] = (in && !lastrow) ? S[P.sw + k] : 0;
      }
    }
    float ang[4], sc[4][2], se[4][2];
    int sq[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      ang[k] = kNotDefDeg; sc[k][0] = sc[k][1] = se[k][0] = se[k][1] = 0.f; sq[k] = 0;
      if (x0 + k < P.sw - 1 && !lastrow) {
        const int DA = r1[k + 1] - r0[k], BC = r0[k + 1] - r1[k];
        const int gx = DA + BC, gy = DA - BC;
        const int s = gx * gx + gy * gy;
        sq[k] = s;
        if (s > P.s_th) {
          const int ti = (gx + kGradR) * kGradN + (gy + kGradR);
          const float4 rec = __ldg(&T[ti]);
          const float2 scs = __ldg(&TS[ti]);
          ang[k] = rec.x; sc[k][0] = rec.y; sc[k][1] = rec.z; se[k][0] = scs.x; se[k][1] = scs.y;
          smx = max(smx, s);
        }
      }
    }
    const long long o = (long long)f * P.npx + (long long)y * P.sw + x0;
    if (kVec) {
      *reinterpret_cast<float4*>(ANG + o) = make_float4(ang[0], ang[1], ang[2], ang[3]);
      *reinterpret_cast<int4*>(S2 + o) = make_int4(sq[0], sq[1], sq[2], sq[3]);
      float4* c4 = reinterpret_cast<float4*>(CS + o);
      c4[0] = make_float4(sc[0][0], sc[0][1], sc[1][0], sc[1][1]); c4[1] = make_float4(sc[2][0], sc[2][1], sc[3][0], sc[3][1]);
      float4* e4 = reinterpret_cast<float4*>(seedcs + o);
      e4[0] = make_float4(se[0][0], se[0][1], se[1][0], se[1][1]); e4[1] = make_float4(se[2][0], se[2][1], se[3][0], se[3][1]);
    } else {
      for (int k = 0; k < 4 && x0 + k < P.sw; k++) {
        ANG[o + k] = ang[k]; S2[o + k] = sq[k]; CS[o + k] = make_float2(sc[k][0], sc[k][1]); seedcs[o + k] = make_float2(se[k][0], se[k][1]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) smx = max(smx, __shfl_xor_sync(0xffffffffu, smx, o));
  if ((threadIdx.x & 31) == 0 && smx > 0) atomicMax(&smax, smx);
  __syncthreads();
  if (threadIdx.x == 0 && smax > 0) atomicMax(&maxs[f], smax);
}
__device__ __forceinline__ double s_norm(int s) { return sqrt((double)s / 4.0); }
__device__ __forceinline__ int s_bin(int s, double bin_coef) { return (int)(s_norm(s) * bin_coef); }

// K_C per-chunk histograms of the defined pixels (chunk = kChunkRows image rows)
__global__ void __launch_bounds__(256) k_lsd_hist(LineParams P, const float* __restrict__ ANG, const int* __restrict__ S2, const int* __restrict__ maxs,
                                                  unsigned short* __restrict__ counts /*[B][kBins][nchunk]*/) {
  __shared__ int hist[kBins];
  const int chunk = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < kBins; i += 256) hist[i] = 0;
  __syncthreads();
  const int ms = maxs[f];
  const double max_grad = ms > 0 ? sqrt((double)ms / 4.0) : -1.0;
  const double bin_coef = (max_grad > 0) ? (double)(kBins - 1) / max_grad : 0.0;
  const int y0 = chunk * kChunkRows, y1 = min(y0 + kChunkRows, P.sh - 1);
  const float* G = ANG + (long long)f * P.npx;
  const int* SS = S2 + (long long)f * P.npx;
  for (int i = tid; i < (y1 - y0) * P.sw; i += 256) {
    int y = y0 + i / P.sw, x = i % P.sw;
    if (G[y * P.sw + x] != kNotDefDeg) atomicAdd(&hist[s_bin(SS[y * P.sw + x], bin_coef)], 1);
  }
  __syncthreads();
  for (int i = tid; i < kBins; i += 256) counts[((long long)f * kBins + i) * P.nchunk + chunk] = (unsigned short)hist[i];
}

// K_D offsets[bin][chunk] = number of defined pixels that precede (bin desc, chunk asc); ndef = total
__global__ void __launch_bounds__(kBins) k_lsd_scan(LineParams P, const unsigned short* __restrict__ counts,
                                                    int* __restrict__ offsets, int* __restrict__ ndef) {
  __shared__ int wsum[32];
  const int f = blockIdx.x, t = threadIdx.x, bin = kBins - 1 - t, lane = t & 31, wid = t >> 5;
  const unsigned short* c = counts + ((long long)f * kBins + bin) * P.nchunk;
  int tot = 0;
  for (int k = 0; k < P.nchunk; k++) tot += c[k];
  int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int v = wsum[lane], in2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, in2, o); if (lane >= o) in2 += u; }
    wsum[lane] = in2 - v;
    if (lane == 31) ndef[f] = in2;
  }
  __syncthreads();
  int base = wsum[wid] + incl - tot;
  int* o = offsets + ((long long)f * kBins + bin) * P.nchunk;
  for (int k = 0; k < P.nchunk; k++) { o[k] = base; base += c[k]; }
}

// K_E stable scatter: one warp per chunk walks its pixels in row-major order
__global__ void __launch_bounds__(128) k_lsd_scatter(LineParams P, const float* __restrict__ ANG, const int* __restrict__ S2, const int* __restrict__ maxs,
                                                     const int* __restrict__ offsets, unsigned* __restrict__ order) {
  __shared__ unsigned short cnt[4][kBins];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * 4 + wid, f = blockIdx.y;
  for (int i = lane; i < kBins; i += 32) cnt[wid][i] = 0;
  __syncwarp();
  if (chunk >= P.nchunk) return;
  const int ms = maxs[f];
  const double max_grad = ms > 0 ? sqrt((double)ms / 4.0) : -1.0;
  const double bin_coef = (max_grad > 0) ? (double)(kBins - 1) / max_grad : 0.0;
  const int y0 = chunk * kChunkRows, y1 = min(y0 + kChunkRows, P.sh - 1);
  const float* G = ANG + (long long)f * P.npx;
  const int* SS = S2 + (long long)f * P.npx;
  const int* off = offsets + (long long)f * kBins * P.nchunk;
  unsigned* O = order + (long long)f * P.npx;
  const int n = (y1 - y0) * P.sw;
  const unsigned lt = (1u << lane) - 1u;
  for (int i0 = 0; i0 < n; i0 += 32) {
    int i = i0 + lane, bin = -1, pix = 0;
    if (i < n) {
      int y = y0 + i / P.sw, x = i % P.sw;
      pix = x | (y << 16);                       // packed (x, y): the grow kernel never divides
      if (G[y * P.sw + x] != kNotDefDeg) bin = s_bin(SS[y * P.sw + x], bin_coef);
    }
    unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0) O[off[bin * P.nchunk + chunk] + cnt[wid][bin] + __popc(peers & lt)] = (unsigned)pix;
    __syncwarp();
    if (bin >= 0 && (peers & lt) == 0) cnt[wid][bin] += (unsigned short)__popc(peers);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- K_F region growing
struct GrowCtx {
  int* ANG; const float2* CS; const int* SQ; const float2* S2; unsigned* R; unsigned* ring;
  int sw, sh, s_th;
};
// The USED flag of LSD lives in the SIGN BIT of the pixel's angle word (degrees, >= 0 when defined): one 4-byte load
// tells a candidate's used state, definedness (NOTDEF = -1024 is negative too) and angle; setting / clearing it is a
// plain store by whichever lane owns the pixel (no bitmap word shared between lanes, no read-modify-write).
constexpr int kUsedBit = (int)0x80000000;
__device__ __forceinline__ bool used_get(const GrowCtx& C, int idx) { return C.ANG[idx] < 0; }     // seeds are always defined
__device__ __forceinline__ float pixel_angle(const GrowCtx& C, int idx) { return __int_as_float(C.ANG[idx] & ~kUsedBit); }   // of a defined pixel
__device__ __forceinline__ void used_clear(const GrowCtx& C, int idx) { C.ANG[idx] &= ~kUsedBit; }  // one lane per pixel
struct RectD { double x1, y1, x2, y2, width; };

__device__ __forceinline__ double angle_diff_signed(double a, double b) {
  double diff = a - b;
  while (diff <= -kPI) diff += 2 * kPI;
  while (diff > kPI) diff -= 2 * kPI;
  return diff;
}
__device__ __forceinline__ bool is_aligned(double a, double theta, double prec) {
  // branch-free form of LineSegmentDetectorImpl::isAligned (same values: |theta-a|, folded once at 3pi/2)
  const double n1 = fabs(theta - a);
  const double n2 = fabs(n1 - 2 * kPI);
  return ((n1 > (3 * kPI) / 2) ? n2 : n1) <= prec;
}

// LineSegmentDetectorImpl::region_grow — exact visiting order; returns the region size, region in C.R[0..n).
// Four queue entries are expanded per step: lanes 8g..8g+7 fetch the 8 neighbours of entry i+g (used flag, then the
// 4-byte angle and the cos/sin pair of the unused ones), then the candidates are committed in the reference's order
// (queue order, then row-major inside the 3x3); a pixel added earlier in the same step invalidates its duplicates in
// the later neighbourhoods, so the result equals the one-entry-at-a-time loop.
// The commit loop is the serial spine of the whole front end (one trip per added pixel), so it carries only what the
// next decision needs: one fp64 subtract + two compares per lane, three shuffles, the atan2.  The USED bits, the region
// list and the ring are written after the loop by the accepted lanes themselves.
// kFast: prec < pi/2, isAligned folded to  n <= prec || n >= prec_hi  (prec_hi = smallest double with 2pi-n <= prec;
// 2pi-n is exact for n in [pi,4pi], so the two forms agree bit for bit).
#ifndef GROW_ISOLATED
#define GROW_ISOLATED 0     // batch-parallel retirement of one-pixel regions: bit-exact, measured SLOWER (220 vs 197 ms): the 8 scattered loads per free seed cost more than the 36 % of regions they retire
#endif
#ifndef GROW_SPEC
#define GROW_SPEC 0         // speculative batched commit: bit-exact, but measured SLOWER on B200 (195 vs 168 ms at B=4736)
#endif
#ifndef GROW_INLINE
#define GROW_INLINE 1
#endif
#ifndef GROW_CS_EAGER
#define GROW_CS_EAGER 1
#endif
#ifndef GROW_PREFETCH
#define GROW_PREFETCH 1     // bit 0: at seed-batch load, bit 1: at publication
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// the rows a pixel's 3x3 neighbourhood will touch when it is expanded: angle words and cos/sin pairs (x-1 and x+1 ends)
template <bool kL1>
__device__ __forceinline__ void prefetch_neighbourhood(const GrowCtx& C, int idx, int npx, bool centre_row) {
  const int up = max(idx - C.sw, 1), dn = min(idx + C.sw, npx - 2), ce = min(max(idx, 1), npx - 2);
  if (kL1) {
    prefetch_l1(&C.ANG[up - 1]); prefetch_l1(&C.ANG[up + 1]); prefetch_l1(&C.ANG[dn - 1]); prefetch_l1(&C.ANG[dn + 1]);
    prefetch_l1(&C.CS[up - 1]); prefetch_l1(&C.CS[up + 1]); prefetch_l1(&C.CS[dn - 1]); prefetch_l1(&C.CS[dn + 1]);
    prefetch_l1(&C.CS[ce - 1]); prefetch_l1(&C.CS[ce + 1]);
    if (centre_row) { prefetch_l1(&C.ANG[ce - 1]); prefetch_l1(&C.ANG[ce + 1]); }
  } else {
    prefetch_l2(&C.ANG[up - 1]); prefetch_l2(&C.ANG[up + 1]); prefetch_l2(&C.ANG[dn - 1]); prefetch_l2(&C.ANG[dn + 1]);
    prefetch_l2(&C.CS[up - 1]); prefetch_l2(&C.CS[up + 1]); prefetch_l2(&C.CS[dn - 1]); prefetch_l2(&C.CS[dn + 1]);
    prefetch_l2(&C.CS[ce - 1]); prefetch_l2(&C.CS[ce + 1]);
    if (centre_row) { prefetch_l2(&C.ANG[ce - 1]); prefetch_l2(&C.ANG[ce + 1]); }
  }
}
template <bool kFast>
__device__ __forceinline__ int region_grow_t(const GrowCtx& C, unsigned seed, double prec, double prec_hi, double& reg_angle_out, int lane) {
  const int sidx = (int)(seed >> 16) * C.sw + (int)(seed & 0xffffu);
  const float2 s0 = __ldg(&C.S2[sidx]);
  const int sbits = C.ANG[sidx];                  // the seed is unused here, so this is its angle
  double reg_angle = (double)__int_as_float(sbits) * kDegToRads;
  float sumdx = s0.x, sumdy = s0.y;
  bool dirty = false;          // reg_angle lags the sums (it is the seed's own angle until the first pixel is added)
  if (lane == 0) { C.R[0] = seed; C.ring[0] = seed; C.ANG[sidx] = sbits | kUsedBit; }
  int cnt = 1;
  __syncwarp();
  // lane -> (queue entry lane/8, neighbour lane%8); the centre of a 3x3 is always USED, so 8 neighbours suffice
  const int grp = lane >> 3, kk8 = lane & 7, kk = kk8 + (kk8 >= 4);
  const int ox = kk % 3 - 1, oy = kk / 3 - 1;
  const unsigned lt = (1u << lane) - 1u;
  for (int i = 0; i < cnt;) {
    const int m = min(4, cnt - i);
    bool valid = false;
    int idx = -1;
    unsigned pk = 0xffff0000u | (unsigned)lane;      // unique per lane unless it names a real pixel
    int ab = -1;
    float2 csv = make_float2(0.f, 0.f);
    if (grp < m) {
      const int qi = i + grp;
      const unsigned p = (cnt - qi <= kRing) ? C.ring[qi & (kRing - 1)] : C.R[qi];
      const int xx = (int)(p & 0xffffu) + ox, yy = (int)(p >> 16) + oy;
      if (xx >= 0 && yy >= 0 && xx < C.sw && yy < C.sh) {
        idx = yy * C.sw + xx;
        ab = C.ANG[idx];
#if GROW_CS_EAGER
        csv = __ldg(&C.CS[idx]);           // issued together with the angle word: one memory round trip per step
#endif
        if (ab >= 0) {                     // defined and not USED
#if !GROW_CS_EAGER
          csv = __ldg(&C.CS[idx]);
#endif
          valid = true;
          pk = (unsigned)xx | ((unsigned)yy << 16);
        }
      }
    }
    i += m;
    unsigned live = __ballot_sync(0xffffffffu, valid);
    if (live == 0u) continue;
    const double a = (double)__int_as_float(ab) * kDegToRads;
    // Commit in order.  Every remaining candidate is tested against the CURRENT region angle at once; the first
    // aligned one (lowest lane = reference order) is added, which changes the angle, and the candidates after it
    // are tested again.  Candidates skipped before the committed lane were tested with the angle they would
    // have seen in the sequential loop, so they are never revisited.
    int mypos = -1;
    auto aligned_now = [&](double th) {
      if (kFast) { const double n1 = fabs(th - a); return (n1 <= prec) || (n1 >= prec_hi); }
      return is_aligned(a, th, prec);
    };
    while (live) {
      if (dirty) { reg_angle = (double)fast_atan2_deg_l(sumdy, sumdx) * kDegToRads; dirty = false; }
      const unsigned A = __ballot_sync(0xffffffffu, aligned_now(reg_angle)) & live;
      if (!A) break;
#if GROW_SPEC
      if (A & (A - 1u)) {
        // Several candidates are aligned with the current angle.  Speculate that they are all accepted in order: walk them
        // once accumulating the cos/sin sums sequentially (the only part that has to be serial for bit-exact fp32 sums),
        // every lane keeping the sums as they stand just BEFORE its own turn; then each lane evaluates the region angle it
        // would have seen (one atan2 per lane, in parallel, instead of one per accepted pixel in sequence) and re-checks its
        // own decision.  Everything before the first lane whose decision differs from the speculation is final.
        float tx = sumdx, ty = sumdy, sx = sumdx, sy = sumdy;
        unsigned Aw = A, acc = 0u;
        int killer = 64;                         // lowest speculated-accepted lane naming my pixel (64 = none)
        while (Aw) {
          const int k = __ffs(Aw) - 1;
          tx = __fadd_rn(tx, __shfl_sync(0xffffffffu, csv.x, k));
          ty = __fadd_rn(ty, __shfl_sync(0xffffffffu, csv.y, k));
          if (lane > k) { sx = tx; sy = ty; }
          acc |= 1u << k;
          const unsigned d = __ballot_sync(0xffffffffu, pk == __shfl_sync(0xffffffffu, pk, k)) & ~((2u << k) - 1u);
          if (((d >> lane) & 1u) && killer == 64) killer = k;
          Aw &= ~(d | (1u << k));
        }
        const unsigned dead = __ballot_sync(0xffffffffu, killer != 64);
        double th = reg_angle;
        if (acc & lt) th = (double)fast_atan2_deg_l(sy, sx) * kDegToRads;
        const bool alj = aligned_now(th);
        const unsigned almask = __ballot_sync(0xffffffffu, alj);
        const unsigned mism = (almask ^ acc) & live & ~dead;    // acc == speculated decisions of the lanes still in play
        if (!mism) {
          if ((acc >> lane) & 1u) mypos = cnt + __popc(acc & lt);
          cnt += __popc(acc);
          sumdx = tx; sumdy = ty; dirty = true;
          live = 0u;
          break;
        }
        const int f = __ffs(mism) - 1;
        unsigned accf = acc & ((1u << f) - 1u);
        float bx = __shfl_sync(0xffffffffu, sx, f), by = __shfl_sync(0xffffffffu, sy, f);
        unsigned deadf = __ballot_sync(0xffffffffu, killer < f);
        if ((almask >> f) & 1u) {                // f was skipped by the speculation but is aligned at its turn: accept it
          bx = __fadd_rn(bx, __shfl_sync(0xffffffffu, csv.x, f));
          by = __fadd_rn(by, __shfl_sync(0xffffffffu, csv.y, f));
          accf |= 1u << f;
          deadf |= __ballot_sync(0xffffffffu, pk == __shfl_sync(0xffffffffu, pk, f));
        }
        if ((accf >> lane) & 1u) mypos = cnt + __popc(accf & lt);
        if (accf) { cnt += __popc(accf); sumdx = bx; sumdy = by; dirty = true; }
        live &= ~(((2u << f) - 1u) | deadf);
        continue;
      }
#endif
      const int k = __ffs(A) - 1;
      if (lane == k) mypos = cnt;
      cnt++;
      sumdx = __fadd_rn(sumdx, __shfl_sync(0xffffffffu, csv.x, k));
      sumdy = __fadd_rn(sumdy, __shfl_sync(0xffffffffu, csv.y, k));
      dirty = true;
      // everything up to k has been decided; the same pixel in a later 3x3 is now USED
      live &= ~(((2u << k) - 1u) | __ballot_sync(0xffffffffu, pk == __shfl_sync(0xffffffffu, pk, k)));
    }
    if (mypos >= 0) {      // publish: every accepted lane owns its pixel
      C.ANG[idx] = ab | kUsedBit;
      C.R[mypos] = pk;
      C.ring[mypos & (kRing - 1)] = pk;
      if (GROW_PREFETCH & 2) prefetch_neighbourhood<true>(C, idx, C.sw * C.sh, false);   // it will be expanded a few steps from now
    }
    __syncwarp();
  }
  if (dirty) reg_angle = (double)fast_atan2_deg_l(sumdy, sumdx) * kDegToRads;
  reg_angle_out = reg_angle;
  return cnt;
}
__device__ __noinline__ int region_grow_hot(const GrowCtx& C, unsigned seed, double prec, double prec_hi, double& reg_angle, int lane) {
  return region_grow_t<true>(C, seed, prec, prec_hi, reg_angle, lane);
}
// refine's regrow with the adaptive tolerance tau (any value): generic isAligned, kept out of line
__device__ __noinline__ int region_grow_cold(const GrowCtx& C, unsigned seed, double prec, double& reg_angle, int lane) {
  return region_grow_t<false>(C, seed, prec, 0.0, reg_angle, lane);
}

// LineSegmentDetectorImpl::region2rect (+get_theta); sums are reduced lane-strided then by butterfly
__device__ __noinline__ void region2rect(const GrowCtx& C, int n, double reg_angle, double prec, RectD& rec, int lane) {
  double sx = 0, sy = 0, sw_ = 0;
  for (int i = lane; i < n; i += 32) {
    const unsigned p = C.R[i];
    const int px = (int)(p & 0xffffu), py = (int)(p >> 16);
    const double w = s_norm(__ldg(&C.SQ[py * C.sw + px]));
    sx += (double)px * w;
    sy += (double)py * w;
    sw_ += w;
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sw_ = warp_sum(sw_);
  const double x = sx / sw_, y = sy / sw_;
  double Ixx = 0, Iyy = 0, Ixy = 0;
  for (int i = lane; i < n; i += 32) {
    const unsigned p = C.R[i];
    const int px = (int)(p & 0xffffu), py = (int)(p >> 16);
    const double w = s_norm(__ldg(&C.SQ[py * C.sw + px]));
    const double dx = (double)px - x, dy = (double)py - y;
    Ixx += dy * dy * w; Iyy += dx * dx * w; Ixy -= dx * dy * w;
  }
  Ixx = warp_sum(Ixx); Iyy = warp_sum(Iyy); Ixy = warp_sum(Ixy);
  const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg_l((float)(lambda - Ixx), (float)Ixy)
                                         : (double)fast_atan2_deg_l((float)Ixy, (float)(lambda - Iyy));
  theta *= kDegToRads;
  if (fabs(angle_diff_signed(theta, reg_angle)) > prec) theta += kPI;
  const double dx = cos(theta), dy = sin(theta);
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
  for (int i = lane; i < n; i += 32) {
    const unsigned p = C.R[i];
    const double rdx = (double)(int)(p & 0xffffu) - x, rdy = (double)(int)(p >> 16) - y;
    const double l = rdx * dx + rdy * dy, w = -rdx * dy + rdy * dx;
    l_max = fmax(l_max, l); l_min = fmin(l_min, l);
    w_max = fmax(w_max, w); w_min = fmin(w_min, w);
  }
  l_max = warp_max_d(l_max); l_min = warp_min_d(l_min); w_max = warp_max_d(w_max); w_min = warp_min_d(w_min);
  rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
  rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
  rec.width = w_max - w_min;
  if (rec.width < 1.0) rec.width = 1.0;
}
__device__ __forceinline__ double dist_d(double x1, double y1, double x2, double y2) {
  return sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
}

// LineSegmentDetectorImpl::refine + reduce_region_radius; n is updated; returns false if the region is rejected
__device__ __noinline__ bool refine(const GrowCtx& C, int& n, double reg_angle, double prec, RectD& rec, double density_th, int lane, bool& released) {
  double density = (double)n / (dist_d(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  if (density >= density_th) return true;
  released = true;              // from here on USED flags are cleared
  const unsigned p0 = C.R[0];
  const double xc = (double)(int)(p0 & 0xffffu), yc = (double)(int)(p0 >> 16);
  const double ang_c = (double)pixel_angle(C, (int)(p0 >> 16) * C.sw + (int)(p0 & 0xffffu)) * kDegToRads;
  double sum = 0, s_sum = 0;
  int cnt = 0;
  for (int i = lane; i < n; i += 32) {
    const unsigned p = C.R[i];
    const int pidx = (int)(p >> 16) * C.sw + (int)(p & 0xffffu);
    used_clear(C, pidx);
    const double px = (double)(int)(p & 0xffffu), py = (double)(int)(p >> 16);
    if (dist_d(xc, yc, px, py) < rec.width) {
      const double ang_d = angle_diff_signed((double)pixel_angle(C, pidx) * kDegToRads, ang_c);
      sum += ang_d; s_sum += ang_d * ang_d; ++cnt;
    }
  }
  sum = warp_sum(sum); s_sum = warp_sum(s_sum); cnt = warp_sum(cnt);
  __syncwarp();
  const double mean_angle = sum / (double)cnt;
  const double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)cnt + mean_angle * mean_angle);
  n = region_grow_cold(C, p0, tau, reg_angle, lane);
  if (n < 2) return false;
  region2rect(C, n, reg_angle, prec, rec, lane);
  density = (double)n / (dist_d(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  if (density >= density_th) return true;
  // reduce_region_radius
  const double r1 = (rec.x1 - xc) * (rec.x1 - xc) + (rec.y1 - yc) * (rec.y1 - yc);
  const double r2 = (rec.x2 - xc) * (rec.x2 - xc) + (rec.y2 - yc) * (rec.y2 - yc);
  double radSq = r1 > r2 ? r1 : r2;
  const unsigned lt = (1u << lane) - 1u;
  while (density < density_th) {
    radSq *= 0.75 * 0.75;
    int kept = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {   // order-preserving compaction (the reference swap-removes; only sums follow)
      const int i = i0 + lane;
      unsigned p = 0;
      bool keep = false;
      if (i < n) {
        p = C.R[i];
        const double px = (double)(int)(p & 0xffffu), py = (double)(int)(p >> 16);
        keep = !((px - xc) * (px - xc) + (py - yc) * (py - yc) > radSq);
        if (!keep) used_clear(C, (int)(p >> 16) * C.sw + (int)(p & 0xffffu));
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      __syncwarp();
      if (keep) C.R[kept + __popc(m & lt)] = p;
      kept += __popc(m);
      __syncwarp();
    }
    n = kept;
    if (n < 2) return false;
    region2rect(C, n, reg_angle, prec, rec, lane);
    density = (double)n / (dist_d(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
  }
  return true;
}

#ifndef GROW_WARPS
#define GROW_WARPS 1        // frames (warps) per CTA; with GROW_MIN_CTAS it sets the register budget / resident warps per SM
#endif
#ifndef GROW_MIN_CTAS
#define GROW_MIN_CTAS 32
#endif
__global__ void __launch_bounds__(32 * GROW_WARPS, GROW_MIN_CTAS) k_lsd_grow(LineParams P, int* ANG, const float2* __restrict__ CS, const int* __restrict__ SQ, const float2* __restrict__ seedcs,
                                                 const unsigned* __restrict__ order, const int* __restrict__ ndef,
                                                 unsigned* __restrict__ reg, float4* __restrict__ segs,
                                                 int* __restrict__ nseg, int* __restrict__ overflow, int nframes) {
  __shared__ unsigned rings[GROW_WARPS][kRing];
  const int lane = threadIdx.x & 31;
  unsigned* ring = rings[threadIdx.x >> 5];
  for (int f = blockIdx.x * GROW_WARPS + (threadIdx.x >> 5); f < nframes; f += gridDim.x * GROW_WARPS) {   // persistent: the grid size caps the resident warps per SM
  // const object: the cold out-of-line callees take it by reference, the inlined hot loop keeps its fields in registers
  const GrowCtx C = {ANG + (long long)f * P.npx, CS + (long long)f * P.npx, SQ + (long long)f * P.npx, seedcs + (long long)f * P.npx,
                     reg + (long long)f * P.npx, ring, P.sw, P.sh, P.s_th};
  const unsigned* O = order + (long long)f * P.npx;
  float4* S = segs + (long long)f * P.seg_cap;
  const int n = ndef[f];
  int ns = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const unsigned pix = (i < n) ? O[i] : 0u;
    const int pidx = (int)(pix >> 16) * P.sw + (int)(pix & 0xffffu);
    const int sab = (i < n) ? C.ANG[pidx] : -1;          // angle word of my seed (negative: USED)
    unsigned todo = __ballot_sync(0xffffffffu, sab >= 0);
    // A seed none of whose 8 neighbours is free, defined and aligned with the seed's own angle grows a region of exactly
    // one pixel (the first step of region_grow finds no candidate): below min_reg_size, so its only effect is the seed's
    // USED bit.  36 % of all regions are like that.  The test is evaluated for the whole batch at once, one seed per
    // lane; it stays valid until the seed's turn because USED flags only increase — except when refine releases pixels,
    // after which the remaining lanes are evaluated again.
    auto isolated = [&](bool active) {
      bool iso = active;
      if (active) {
        const int sx = (int)(pix & 0xffffu), sy = (int)(pix >> 16);
        const double th = (double)__int_as_float(sab) * kDegToRads;
#pragma unroll
        for (int kk = 0; kk < 9; kk++) {
          if (kk == 4) continue;
          const int xx = sx + kk % 3 - 1, yy = sy + kk / 3 - 1;
          if (xx < 0 || yy < 0 || xx >= P.sw || yy >= P.sh) continue;
          const int ab = C.ANG[yy * P.sw + xx];
          if (ab < 0) continue;
          const double n1 = fabs(th - (double)__int_as_float(ab) * kDegToRads);
          if ((n1 <= P.prec) || (n1 >= P.prec_hi)) iso = false;
        }
      }
      return __ballot_sync(0xffffffffu, iso);
    };
    unsigned isomask = GROW_ISOLATED ? isolated((todo >> lane) & 1u) : 0u;
    if (GROW_PREFETCH & 1) {
      // this batch's seeds that will really grow: their seed record and 3x3 rows; and the next batch's flag words
      if (((todo & ~isomask) >> lane) & 1u) { prefetch_l2(&C.S2[pidx]); prefetch_neighbourhood<false>(C, pidx, P.npx, false); }
      if (i + 32 < n) { const unsigned pn = O[i + 32]; prefetch_l2(&C.ANG[(int)(pn >> 16) * P.sw + (int)(pn & 0xffffu)]); }
    }
    while (todo) {
      const int k = __ffs(todo) - 1;
      if ((isomask >> k) & 1u) {            // one-pixel region: mark the seed and move on
        if (lane == k) C.ANG[pidx] = sab | kUsedBit;
        todo &= todo - 1u;
        continue;
      }
      const unsigned seed = __shfl_sync(0xffffffffu, pix, k);
      double reg_angle;
      bool released = false;
#if GROW_INLINE
      int cnt = region_grow_t<true>(C, seed, P.prec, P.prec_hi, reg_angle, lane);
#else
      int cnt = region_grow_hot(C, seed, P.prec, P.prec_hi, reg_angle, lane);
#endif
      if (cnt >= P.min_reg_size) {
        RectD rec;
        region2rect(C, cnt, reg_angle, P.prec, rec, lane);
        if (refine(C, cnt, reg_angle, P.prec, rec, P.density_th, lane, released)) {
          if (lane == 0 && ns < P.seg_cap)
            S[ns] = make_float4((float)((rec.x1 + 0.5) / 0.8), (float)((rec.y1 + 0.5) / 0.8), (float)((rec.x2 + 0.5) / 0.8),
                                (float)((rec.y2 + 0.5) / 0.8));
          ns++;
        }
      }
      __syncwarp();
      // seeds later in this batch may have been consumed (or released by refine): re-read their flags
      const bool free_now = i < n && lane > k && !used_get(C, pidx);
      todo = __ballot_sync(0xffffffffu, free_now);
      if (released) isomask = GROW_ISOLATED ? isolated(free_now) : 0u;     // pixels came back: earlier verdicts may be stale
      else isomask &= todo;
    }
  }
  if (lane == 0) { nseg[f] = min(ns, P.seg_cap); if (ns > P.seg_cap) atomicExch(overflow, 1); }
  __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- K_G keylines
__global__ void __launch_bounds__(256) k_keylines(LineParams P, const float4* __restrict__ segs, const int* __restrict__ nseg,
                                                  const uint8_t* __restrict__ mask, PLKeyLineRec* __restrict__ kls,
                                                  double* __restrict__ linefunc, int* __restrict__ nl) {
  extern __shared__ unsigned long long keys[];   // seg_cap rounded to a power of two
  __shared__ int s_cnt;
  const int f = blockIdx.x, tid = threadIdx.x;
  const float4* S = segs + (long long)f * P.seg_cap;
  const int n = nseg[f];
  int cap2 = 1;
  while (cap2 < max(n, 1)) cap2 <<= 1;
  auto clampseg = [&](float4 e) {
    if (e.x < 0) e.x = 0; if (e.x >= P.w) e.x = (float)P.w - 1.0f;
    if (e.z < 0) e.z = 0; if (e.z >= P.w) e.z = (float)P.w - 1.0f;
    if (e.y < 0) e.y = 0; if (e.y >= P.h) e.y = (float)P.h - 1.0f;
    if (e.w < 0) e.w = 0; if (e.w >= P.h) e.w = (float)P.h - 1.0f;
    return e;
  };
  auto seglen = [&](float4 e) {
    const double a = (double)__fsub_rn(e.x, e.z), b = (double)__fsub_rn(e.y, e.w);
    return (float)sqrt(a * a + b * b);
  };
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  for (int i = tid; i < cap2; i += 256) {
    unsigned long long key = ~0ull;
    if (i < n) {
      const float4 e = clampseg(S[i]);
      bool drop = false;
      if (mask) drop = mask[(long long)(int)e.y * P.w + (int)e.x] == 0 && mask[(long long)(int)e.w * P.w + (int)e.z] == 0;
      if (!drop) {
        const float resp = __fdiv_rn(seglen(e), (float)max(P.w, P.h));
        key = ((unsigned long long)(~__float_as_uint(resp)) << 32) | (unsigned)i;   // response desc, detection order asc
        atomicAdd(&s_cnt, 1);
      }
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= cap2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < cap2; i += 256) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], b = keys[ixj];
          bool up = ((i & k) == 0);
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  const int size = s_cnt;
  // LineExtractor.cpp:44-67 truncation (total/index quirk)
  int total = size > P.nfeatures ? P.nfeatures : size, index = total;
  __shared__ int s_index;
  if (tid == 0) {
    if (total > 0) {
      const float lastLen = seglen(clampseg(S[(unsigned)(keys[total - 1] & 0xffffffffu)]));
      if ((double)lastLen < P.min_line_length) {
        for (int i = 0; i < total - 1; i++) {
          const float l0 = seglen(clampseg(S[(unsigned)(keys[i] & 0xffffffffu)]));
          const float l1 = seglen(clampseg(S[(unsigned)(keys[i + 1] & 0xffffffffu)]));
          if ((double)l0 >= P.min_line_length && (double)l1 < P.min_line_length) { index = i; break; }
        }
      }
    }
    s_index = index;
  }
  __syncthreads();
  index = s_index;
  const int nout = index + 1;
  PLKeyLineRec* K = kls + (long long)f * P.capL;
  double* LF = linefunc + (long long)f * P.capL * 3;
  for (int i = tid; i < nout && i < P.capL; i += 256) {
    PLKeyLineRec kl;
    if (i < size) {
      const float4 e = clampseg(S[(unsigned)(keys[i] & 0xffffffffu)]);
      kl.startPointX = e.x; kl.startPointY = e.y; kl.endPointX = e.z; kl.endPointY = e.w;
      kl.sPointInOctaveX = e.x; kl.sPointInOctaveY = e.y; kl.ePointInOctaveX = e.z; kl.ePointInOctaveY = e.w;
      kl.lineLength = seglen(e);
      const int x0 = __float2int_rn(e.x), y0 = __float2int_rn(e.y), x1 = __float2int_rn(e.z), y1 = __float2int_rn(e.w);
      kl.numOfPixels = max(abs(x1 - x0), abs(y1 - y0)) + 1;
      kl.angle = (float)atan2((double)__fsub_rn(e.w, e.y), (double)__fsub_rn(e.z, e.x));
      kl.octave = 0;
      kl.size = __fmul_rn(__fsub_rn(e.z, e.x), __fsub_rn(e.w, e.y));
      kl.response = __fdiv_rn(kl.lineLength, (float)max(P.w, P.h));
      kl.ptx = __fdiv_rn(__fadd_rn(e.z, e.x), 2.f); kl.pty = __fdiv_rn(__fadd_rn(e.w, e.y), 2.f);
    } else {
      memset(&kl, 0, sizeof(kl));   // the KeyLine appended by resize(index+1)
    }
    kl.class_id = i;
    K[i] = kl;
    const double sx = kl.startPointX, sy = kl.startPointY, ex = kl.endPointX, ey = kl.endPointY;
    const double lx = sy * 1.0 - 1.0 * ey, ly = 1.0 * ex - sx * 1.0, lz = sx * ey - sy * ex;
    const double nn = sqrt(lx * lx + ly * ly);
    LF[3 * i] = lx / nn; LF[3 * i + 1] = ly / nn; LF[3 * i + 2] = lz / nn;
  }
  if (tid == 0) nl[f] = min(nout, P.capL);
}

// ---------------------------------------------------------------------------------------------- K_H LBD blur + Sobel
// GaussianBlur 5x5 sigma 1 (8.8 fixed point rows [14 62 104 62 14]) fused with the Sobel pair (dx, dy as int16).
// Tile: 64x64 blurred pixels <- 68x68 raw pixels in shared memory -> 62x62 outputs.  The raw tile is loaded at
// reflect-101 coordinates; because the kernel is symmetric, the blur of the reflected image at column -1 equals the
// blurred value at column +1, i.e. exactly what Sobel's own BORDER_REFLECT_101 of the BLURRED image needs, so no tap
// ever reflects again.  One thread walks down one column: the last five horizontal sums live in registers (blur), then
// a sliding 3x3 window of blurred bytes (Sobel).
constexpr int kSobT = 64, kSobOut = kSobT - 2, kSobRawP = kSobT + 8, kSobSeg = kSobT / 4;
__global__ void __launch_bounds__(256) k_lbd_sobel(LineParams P, const uint8_t* __restrict__ imgs, int stride,
                                                   long long frame_stride, short2* __restrict__ dxy) {
  __shared__ uint8_t raw[(kSobT + 4) * kSobRawP];
  __shared__ uint8_t bl[kSobT * (kSobT + 4)];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int X0 = blockIdx.x * kSobOut, Y0 = blockIdx.y * kSobOut;   // first output pixel of the tile
  const uint8_t* img = imgs + (long long)blockIdx.z * frame_stride;
  auto r101 = [](int p, int n) { if (p < 0) p = -p; if (p >= n) p = 2 * (n - 1) - p; return min(max(p, 0), n - 1); };
  // raw rows Y0-3 .. Y0+64, columns X0-3 .. X0+64
  for (int r = ty; r < kSobT + 4; r += 4) {
    const uint8_t* row = img + (long long)r101(Y0 - 3 + r, P.h) * stride;
    raw[r * kSobRawP + tx] = row[r101(X0 - 3 + tx, P.w)];
    if (tx < 4) raw[r * kSobRawP + kSobT + tx] = row[r101(X0 - 3 + kSobT + tx, P.w)];
  }
  __syncthreads();
  {  // blurred pixel (bx, by) of the tile = image pixel (X0-1+bx, Y0-1+by); thread: column tx, rows ty*16 .. +15
    const uint8_t* p = raw + (ty * kSobSeg) * kSobRawP + tx;
    auto hrow = [&](const uint8_t* q) { return 14 * (q[0] + q[4]) + 62 * (q[1] + q[3]) + 104 * q[2]; };
    int w0 = hrow(p), w1 = hrow(p + kSobRawP), w2 = hrow(p + 2 * kSobRawP), w3 = hrow(p + 3 * kSobRawP);
    p += 4 * kSobRawP;
    uint8_t* o = bl + (ty * kSobSeg) * (kSobT + 4) + tx;
#pragma unroll 4
    for (int r = 0; r < kSobSeg; r++, p += kSobRawP, o += kSobT + 4) {
      const int w4 = hrow(p);
      const unsigned acc = 14u * (unsigned)(w0 + w4) + 62u * (unsigned)(w1 + w3) + 104u * (unsigned)w2;
      *o = (uint8_t)((acc + 32768u) >> 16);
      w0 = w1; w1 = w2; w2 = w3; w3 = w4;
    }
  }
  __syncthreads();
  const int x = X0 + tx - 1;                       // output column of blurred column tx (1..62 produce outputs)
  if (tx < 1 || tx > kSobOut || x >= P.w) return;
  const int rbeg = max(ty * kSobSeg, 1), rend = min(ty * kSobSeg + kSobSeg - 1, kSobOut);   // blurred rows of my outputs
  const uint8_t* b = bl + (rbeg - 1) * (kSobT + 4) + tx;
  int a00 = b[-1], a01 = b[0], a02 = b[1];
  b += kSobT + 4;
  int a10 = b[-1], a11 = b[0], a12 = b[1];
  short2* D = dxy + (long long)blockIdx.z * P.w * P.h;
  for (int r = rbeg; r <= rend; r++) {
    b += kSobT + 4;
    const int a20 = b[-1], a21 = b[0], a22 = b[1];
    const int y = Y0 + r - 1;
    if (y < P.h)
      D[(long long)y * P.w + x] = make_short2((short)((a02 - a00) + 2 * (a12 - a10) + (a22 - a20)),
                                              (short)((a20 - a00) + 2 * (a21 - a01) + (a22 - a02)));
    a00 = a10; a01 = a11; a02 = a12; a10 = a20; a11 = a21; a12 = a22;
  }
}

// ---------------------------------------------------------------------------------------------- K_I LBD describe
__constant__ float c_gaussG[63];
__constant__ float c_gaussL[21];
__constant__ unsigned char c_comb[64];

__global__ void __launch_bounds__(64) k_lbd_describe(LineParams P, const PLKeyLineRec* __restrict__ kls, const int* __restrict__ nl,
                                                     const short2* __restrict__ dxyi, uint8_t* __restrict__ desc) {
  __shared__ float rs[63][8];
  __shared__ float band[8][9];
  __shared__ float des[72];
  const int li = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  if (li >= nl[f]) return;
  const PLKeyLineRec kl = kls[(long long)f * P.capL + li];
  const short2* DXY = dxyi + (long long)f * P.w * P.h;
  const short realWidth = (short)P.w, imageWidth = (short)(P.w - 1), imageHeight = (short)(P.h - 1);
  const short lengthOfLSP = (short)kl.numOfPixels;
  const short halfHeight = (63 - 1) / 2, halfWidth = (short)((lengthOfLSP - 1) / 2);
  const float midX = (float)(0.5 * (double)__fadd_rn(kl.sPointInOctaveX, kl.ePointInOctaveX));
  const float midY = (float)(0.5 * (double)__fadd_rn(kl.sPointInOctaveY, kl.ePointInOctaveY));
  __shared__ float s_dL[2], s_gL[21], s_norm2[2];
  if (tid < 21) s_gL[tid] = c_gaussL[tid];
  if (tid == 0) { s_dL[0] = (float)cos((double)kl.angle); s_dL[1] = (float)sin((double)kl.angle); }   // fp64 libm once per line
  __syncthreads();
  const float dL0 = s_dL[0], dL1 = s_dL[1];
  const float dO0 = -dL1, dO1 = dL0;
  if (tid < 63) {
    const short hID = (short)tid;
    // sCorX0/Y0 after hID updates "sCorX0 -= dL[1]; sCorY0 += dL[0]" applied sequentially (fp32, same order)
    float sCorX0 = __fadd_rn(__fadd_rn(__fmul_rn(-dL0, (float)halfWidth), __fmul_rn(dL1, (float)halfHeight)), midX);
    float sCorY0 = __fadd_rn(__fsub_rn(__fmul_rn(-dL1, (float)halfWidth), __fmul_rn(dL0, (float)halfHeight)), midY);
    for (short k = 0; k < hID; k++) { sCorX0 = __fsub_rn(sCorX0, dL1); sCorY0 = __fadd_rn(sCorY0, dL0); }
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdL = 0, ngdL = 0, pgdO = 0, ngdO = 0;
    for (short wID = 0; wID < lengthOfLSP; wID++) {
      short t = (short)roundf(sCorX);
      const short xCor = (t < 0) ? 0 : (t > imageWidth) ? imageWidth : t;
      t = (short)roundf(sCorY);
      const short yCor = (t < 0) ? 0 : (t > imageHeight) ? imageHeight : t;
      const short2 g2 = __ldg(&DXY[(int)yCor * realWidth + xCor]);
      const short ddx = g2.x, ddy = g2.y;
      const float gDL = __fadd_rn(__fmul_rn((float)ddx, dL0), __fmul_rn((float)ddy, dL1));
      const float gDO = __fadd_rn(__fmul_rn((float)ddx, dO0), __fmul_rn((float)ddy, dO1));
      if (gDL > 0) pgdL = __fadd_rn(pgdL, gDL); else ngdL = __fsub_rn(ngdL, gDL);
      if (gDO > 0) pgdO = __fadd_rn(pgdO, gDO); else ngdO = __fsub_rn(ngdO, gDO);
      sCorX = __fadd_rn(sCorX, dL0); sCorY = __fadd_rn(sCorY, dL1);
    }
    const float c = c_gaussG[hID];
    pgdL = __fmul_rn(c, pgdL); ngdL = __fmul_rn(c, ngdL); pgdO = __fmul_rn(c, pgdO); ngdO = __fmul_rn(c, ngdO);
    rs[hID][0] = pgdL; rs[hID][1] = ngdL; rs[hID][2] = __fmul_rn(pgdL, pgdL); rs[hID][3] = __fmul_rn(ngdL, ngdL);
    rs[hID][4] = pgdO; rs[hID][5] = ngdO; rs[hID][6] = __fmul_rn(pgdO, pgdO); rs[hID][7] = __fmul_rn(ngdO, ngdO);
  }
  __syncthreads();
  // Band sums: band[q][k] is its own accumulator, fed in row order by the 7 rows of band k-1 (Gaussian taps 0..6), of band
  // k (taps 7..13) and of band k+1 (taps 14..20) — the order in which the reference's row loop touches it.  72 accumulators
  // in parallel, <= 21 ordered fp32 adds each.
  for (int a = tid; a < 72; a += 64) {
    const int q = a / 9, k = a - q * 9;
    const bool sq = (q == 2 || q == 3 || q == 6 || q == 7);
    float b = 0.f;
    for (int B = max(k - 1, 0); B <= min(k + 1, 8); B++) {
      const int off = (B - k + 1) * 7;
      for (int j = 0; j < 7; j++) {
        const float cg = s_gL[j + off], v = rs[B * 7 + j][q];
        b = __fadd_rn(b, sq ? __fmul_rn(__fmul_rn(cg, cg), v) : __fmul_rn(cg, v));
      }
    }
    band[q][k] = b;
  }
  __syncthreads();
  if (tid < 36) {       // mean / stddev of the four quantities of band bb
    const int bb = tid >> 2, c = tid & 3;
    const int qm = (c & 1) + ((c & 2) << 1), qs = qm + 2;          // {0,1,4,5} and {2,3,6,7}
    const float invN = (bb == 0 || bb == 8) ? (float)(1.0 / (7 * 2.0)) : (float)(1.0 / (7 * 3.0));
    const float temp = __fmul_rn(band[qm][bb], invN);
    des[8 * bb + c] = temp;
    des[8 * bb + 4 + c] = sqrtf(__fsub_rn(__fmul_rn(band[qs][bb], invN), __fmul_rn(temp, temp)));
  }
  __syncthreads();
  if (tid == 0 || tid == 32) {   // the two ordered norms (means: k = 0..3, stddevs: k = 4..7), one warp each
    const int k0 = tid ? 4 : 0;
    float acc = 0;
    for (int i = 0; i < 72; i += 8)
      for (int k = k0; k < k0 + 4; k++) acc = __fadd_rn(acc, __fmul_rn(des[i + k], des[i + k]));
    s_norm2[tid ? 1 : 0] = __fdiv_rn(1.f, sqrtf(acc));
  }
  __syncthreads();
  for (int i = tid; i < 72; i += 64) {
    float v = __fmul_rn(des[i], s_norm2[(i & 7) >> 2]);
    if ((double)v > 0.4) v = (float)0.4;
    des[i] = v;
  }
  __syncthreads();
  if (tid == 0) {
    float temp = 0;
    for (int i = 0; i < 72; i++) temp = __fadd_rn(temp, __fmul_rn(des[i], des[i]));
    s_norm2[0] = __fdiv_rn(1.f, sqrtf(temp));
  }
  __syncthreads();
  for (int i = tid; i < 72; i += 64) des[i] = __fmul_rn(des[i], s_norm2[0]);
  __syncthreads();
  if (tid < 32) {
    const float* f1 = &des[8 * c_comb[2 * tid]];
    const float* f2 = &des[8 * c_comb[2 * tid + 1]];
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) if (f1[i] > f2[i]) r += (1u << i);
    desc[((long long)f * P.capL + li) * 32 + tid] = (uint8_t)r;
  }
}

}  // namespace pl

// ================================================================================================ host side
using namespace pl;

struct PLLine {
  PLLineConfig cfg;
  LineParams P;
  cudaStream_t stream = nullptr;
  uint8_t* d_scaled = nullptr;
  float2* d_seedcs = nullptr;
  int grow_grid_cap = 1 << 30;   // max CTAs of the persistent grow kernel (env PLSLAM_LSD_GROW_CTAS_PER_SM x #SMs)
  float* d_ang = nullptr; float2* d_cs = nullptr; int* d_sq = nullptr;
  GradRec* d_gtab = nullptr; float2* d_gtab_seed = nullptr;   // (gx, gy) -> level-line record, built once (k_lsd_grad_table)
  unsigned short* d_counts = nullptr;
  int *d_offsets = nullptr, *d_ndef = nullptr, *d_maxs = nullptr, *d_nseg = nullptr, *d_overflow = nullptr;
  unsigned *d_order = nullptr, *d_reg = nullptr;
  float4* d_segs = nullptr;
  short2* d_dxy = nullptr;
  // host-pointer API staging
  uint8_t* d_img = nullptr; PLKeyLineRec* d_kls = nullptr; uint8_t* d_desc = nullptr; double* d_lf = nullptr; int* d_nl = nullptr;
  uint8_t* d_mask = nullptr;
  size_t key_smem = 0;
  int last_B = 0;
  // optional device timing of the dominant kernel (bench.py roofline): events on the launching stream
  int timing = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

static const unsigned char h_comb[64] = {0, 1, 0, 2, 0, 3, 0, 4, 0, 5, 0, 6, 1, 2, 1, 3, 1, 4, 1, 5, 1, 6, 2, 3, 2, 4, 2, 5, 2, 6, 2, 7,
                                         2, 8, 3, 4, 3, 5, 3, 6, 3, 7, 3, 8, 4, 5, 4, 6, 4, 7, 4, 8, 5, 6, 5, 7, 5, 8, 6, 7, 6, 8, 7, 8};

extern "C" void pl_line_destroy(PLLine* h) {
  if (!h) return;
  cudaFree(h->d_gtab); cudaFree(h->d_gtab_seed); cudaFree(h->d_scaled); cudaFree(h->d_seedcs); cudaFree(h->d_ang); cudaFree(h->d_cs); cudaFree(h->d_sq); cudaFree(h->d_counts); cudaFree(h->d_offsets);
  cudaFree(h->d_ndef); cudaFree(h->d_maxs); cudaFree(h->d_nseg); cudaFree(h->d_overflow); cudaFree(h->d_order);
  cudaFree(h->d_reg); cudaFree(h->d_segs); cudaFree(h->d_dxy); cudaFree(h->d_img); cudaFree(h->d_kls);
  cudaFree(h->d_desc); cudaFree(h->d_lf); cudaFree(h->d_nl); cudaFree(h->d_mask);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int pl_line_create(const PLLineConfig* cfg, PLLine** out) {
  PL_ARG(cfg && out);
  PL_ARG(cfg->width >= 64 && cfg->height >= 64 && cfg->width < 8000 && cfg->height < 8000 && cfg->nfeatures > 0 && cfg->max_batch >= 1);
  int rc = require_device();
  if (rc) return rc;
  PLLine* h = new PLLine;
  h->cfg = *cfg;
  LineParams& P = h->P;
  P.w = cfg->width; P.h = cfg->height;
  P.sw = (int)lrint(P.w * 0.8); P.sh = (int)lrint(P.h * 0.8);
  P.npx = P.sw * P.sh;
  P.nchunk = (P.sh - 1 + kChunkRows - 1) / kChunkRows;
  const double ANG_TH = 22.5, QUANT = 2.0;
  P.prec = kPI * ANG_TH / 180; P.p = ANG_TH / 180; P.density_th = 0.7;
  {  // smallest double n with (2pi - n) <= prec, the subtraction being exact in that range
    const double twopi = 2 * kPI;
    double c = twopi - P.prec;
    while ((twopi - nextafter(c, 0.0)) <= P.prec) c = nextafter(c, 0.0);
    while (!((twopi - c) <= P.prec)) c = nextafter(c, 10.0);
    P.prec_hi = c;
  }
  const double rho = QUANT / sin(P.prec);
  int s = 0;
  while (sqrt((double)(s + 1) / 4.0) <= rho) s++;   // largest s with sqrt(s/4) <= rho
  P.s_th = s;
  const double LOG_NT = 5 * (log10((double)P.sw) + log10((double)P.sh)) / 2 + log10(11.0);
  P.min_reg_size = (int)(size_t)(-LOG_NT / log10(P.p));
  P.seg_cap = cfg->segment_cap > 0 ? cfg->segment_cap : 8192;
  P.nfeatures = cfg->nfeatures; P.capL = cfg->nfeatures + 1; P.min_line_length = cfg->min_line_length;
  { size_t c2 = 1; while (c2 < (size_t)P.seg_cap) c2 <<= 1; h->key_smem = c2 * 8; }
  const size_t B = cfg->max_batch, npx = P.npx;
#define LN_TRY(e) do { int _r = (e); if (_r) { pl_line_destroy(h); return _r; } } while (0)
#define LN_CUDA(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { set_error("%s -> %s", #e, cudaGetErrorString(_e)); pl_line_destroy(h); return PL_ERR_CUDA; } } while (0)
  LN_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  LN_TRY(dev_alloc(&h->d_scaled, npx * B)); LN_TRY(dev_alloc(&h->d_seedcs, npx * B)); LN_TRY(dev_alloc(&h->d_ang, npx * B));
  LN_TRY(dev_alloc(&h->d_cs, npx * B)); LN_TRY(dev_alloc(&h->d_sq, npx * B));
  LN_TRY(dev_alloc(&h->d_counts, (size_t)kBins * P.nchunk * B)); LN_TRY(dev_alloc(&h->d_offsets, (size_t)kBins * P.nchunk * B));
  LN_TRY(dev_alloc(&h->d_ndef, B)); LN_TRY(dev_alloc(&h->d_maxs, B)); LN_TRY(dev_alloc(&h->d_nseg, B)); LN_TRY(dev_alloc(&h->d_overflow, 1));
  LN_TRY(dev_alloc(&h->d_order, npx * B)); LN_TRY(dev_alloc(&h->d_reg, npx * B)); LN_TRY(dev_alloc(&h->d_segs, (size_t)P.seg_cap * B));
  LN_TRY(dev_alloc(&h->d_dxy, (size_t)P.w * P.h * B));
  LN_CUDA(cudaMemset(h->d_overflow, 0, sizeof(int)));
  LN_TRY(dev_alloc(&h->d_gtab, (size_t)kGradN * kGradN)); LN_TRY(dev_alloc(&h->d_gtab_seed, (size_t)kGradN * kGradN));
  k_lsd_grad_table<<<(kGradN * kGradN + 255) / 256, 256, 0, h->stream>>>(h->d_gtab, h->d_gtab_seed);
  LN_CUDA(cudaGetLastError());
  LN_CUDA(cudaStreamSynchronize(h->stream));
  count_launch();
  {  // LBD weights (binary_descriptor_custom.cpp:217-259), integer divisions as in the reference
    float gG[63], gL[21];
    double u = (7 * 3 - 1) / 2, sigma = (7 * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < 21; i++) { double d = i - u; gL[i] = (float)exp(d * d * inv); }
    u = (9 * 7 - 1) / 2; sigma = u; inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < 63; i++) { double d = i - u; gG[i] = (float)exp(d * d * inv); }
    LN_CUDA(cudaMemcpyToSymbol(c_gaussG, gG, sizeof(gG)));
    LN_CUDA(cudaMemcpyToSymbol(c_gaussL, gL, sizeof(gL)));
    LN_CUDA(cudaMemcpyToSymbol(c_comb, h_comb, sizeof(h_comb)));
  }
  LN_CUDA(cudaFuncSetAttribute(k_keylines, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->key_smem));
  // cfg->lsd_used_in_global is accepted for ABI compatibility and ignored: the USED flag is the sign bit of the angle word
  { const char* e = getenv("PLSLAM_LSD_GROW_CTAS_PER_SM"); int per = e ? atoi(e) : 0;
    if (per > 0) { int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); h->grow_grid_cap = per * sms; } }
  *out = h;
  return PL_OK;
}

extern "C" int pl_line_capacity(const PLLine* h) { return h ? h->P.capL : PL_ERR_ARG; }

// Device timing of k_lsd_grow (the dominant kernel): enable, run, then read the duration of the LAST launch.
extern "C" int pl_line_set_timing(PLLine* h, int on) {
  PL_ARG(h);
  if (on && !h->ev0) { PL_CUDA(cudaEventCreate(&h->ev0)); PL_CUDA(cudaEventCreate(&h->ev1)); }
  h->timing = on;
  return PL_OK;
}
extern "C" int pl_line_grow_ms(PLLine* h, float* ms) {
  PL_ARG(h && ms && h->ev0);
  PL_CUDA(cudaEventSynchronize(h->ev1));
  PL_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return PL_OK;
}
/* algorithmic bytes k_lsd_grow must move for one frame: per scaled pixel: record (16) + seed cos/sin (8)
 * + seed order entry (4); the USED map lives in shared memory */
extern "C" long long pl_line_grow_bytes_per_frame(const PLLine* h) { return h ? (long long)h->P.npx * 28 : 0; }

extern "C" int pl_line_extract_batch_dev(PLLine* h, const uint8_t* imgs, int stride, size_t frame_stride, int B,
                                         const uint8_t* mask, void* keylines, uint8_t* desc, double* linefunc, int* n,
                                         void* stream_) {
  PL_ARG(h && imgs && keylines && desc && linefunc && n && B >= 1 && B <= h->cfg.max_batch && stride >= h->cfg.width);
  cudaStream_t st = stream_ ? (cudaStream_t)stream_ : h->stream;
  const LineParams& P = h->P;
  h->last_B = B;
  PL_CUDA(cudaMemsetAsync(h->d_maxs, 0, sizeof(int) * B, st));
  k_lsd_scale<<<dim3((P.sw + 31) / 32, (P.sh + 31) / 32, B), 256, 0, st>>>(P, imgs, stride, (long long)frame_stride, h->d_scaled);
  PL_LAUNCH_CHECK();
  {
    const dim3 grd((P.sw + 255) / 256, (P.sh + 3) / 4, B);
    if (P.sw % 4 == 0)
      k_lsd_grad<true><<<grd, 256, 0, st>>>(P, h->d_scaled, reinterpret_cast<const float4*>(h->d_gtab), h->d_gtab_seed, h->d_ang, h->d_cs, h->d_sq, h->d_seedcs, h->d_maxs);
    else
      k_lsd_grad<false><<<grd, 256, 0, st>>>(P, h->d_scaled, reinterpret_cast<const float4*>(h->d_gtab), h->d_gtab_seed, h->d_ang, h->d_cs, h->d_sq, h->d_seedcs, h->d_maxs);
  }
  PL_LAUNCH_CHECK();
  k_lsd_hist<<<dim3(P.nchunk, B), 256, 0, st>>>(P, h->d_ang, h->d_sq, h->d_maxs, h->d_counts);
  PL_LAUNCH_CHECK();
  k_lsd_scan<<<B, kBins, 0, st>>>(P, h->d_counts, h->d_offsets, h->d_ndef);
  PL_LAUNCH_CHECK();
  k_lsd_scatter<<<dim3((P.nchunk + 3) / 4, B), 128, 0, st>>>(P, h->d_ang, h->d_sq, h->d_maxs, h->d_offsets, h->d_order);
  PL_LAUNCH_CHECK();
  if (h->timing) PL_CUDA(cudaEventRecord(h->ev0, st));
  k_lsd_grow<<<std::min((B + GROW_WARPS - 1) / GROW_WARPS, h->grow_grid_cap), 32 * GROW_WARPS, 0, st>>>(P, reinterpret_cast<int*>(h->d_ang), h->d_cs, h->d_sq, h->d_seedcs, h->d_order, h->d_ndef, h->d_reg, h->d_segs, h->d_nseg, h->d_overflow, B);
  PL_LAUNCH_CHECK();
  if (h->timing) PL_CUDA(cudaEventRecord(h->ev1, st));
  k_keylines<<<B, 256, h->key_smem, st>>>(P, h->d_segs, h->d_nseg, mask, (PLKeyLineRec*)keylines, linefunc, n);
  PL_LAUNCH_CHECK();
  k_lbd_sobel<<<dim3((P.w + kSobOut - 1) / kSobOut, (P.h + kSobOut - 1) / kSobOut, B), 256, 0, st>>>(P, imgs, stride, (long long)frame_stride, h->d_dxy);
  PL_LAUNCH_CHECK();
  k_lbd_describe<<<dim3(P.capL, B), 64, 0, st>>>(P, (const PLKeyLineRec*)keylines, n, h->d_dxy, desc);
  PL_LAUNCH_CHECK();
  return PL_OK;
}

static int line_staging(PLLine* h) {
  if (h->d_img) return PL_OK;
  const size_t B = h->cfg.max_batch;
  int rc;
  if ((rc = dev_alloc(&h->d_img, (size_t)h->P.w * h->P.h * B))) return rc;
  if ((rc = dev_alloc(&h->d_kls, (size_t)h->P.capL * B))) return rc;
  if ((rc = dev_alloc(&h->d_desc, (size_t)h->P.capL * 32 * B))) return rc;
  if ((rc = dev_alloc(&h->d_lf, (size_t)h->P.capL * 3 * B))) return rc;
  if ((rc = dev_alloc(&h->d_nl, B))) return rc;
  if ((rc = dev_alloc(&h->d_mask, (size_t)h->P.w * h->P.h))) return rc;
  return PL_OK;
}

extern "C" int pl_line_extract_batch(PLLine* h, const uint8_t* imgs, int stride, size_t frame_stride, int B,
                                     const uint8_t* mask, void* keylines, uint8_t* desc, double* linefunc, int* n) {
  PL_ARG(h && imgs && keylines && desc && linefunc && n && B >= 1 && B <= h->cfg.max_batch && stride >= h->cfg.width);
  int rc = line_staging(h);
  if (rc) return rc;
  const int W = h->P.w, H = h->P.h;
  for (int b = 0; b < B; b++)
    PL_CUDA(cudaMemcpy2DAsync(h->d_img + (size_t)b * W * H, W, imgs + (size_t)b * frame_stride, stride, W, H, cudaMemcpyHostToDevice, h->stream));
  if (mask) PL_CUDA(cudaMemcpyAsync(h->d_mask, mask, (size_t)W * H, cudaMemcpyHostToDevice, h->stream));
  rc = pl_line_extract_batch_dev(h, h->d_img, W, (size_t)W * H, B, mask ? h->d_mask : nullptr, h->d_kls, h->d_desc, h->d_lf, h->d_nl, h->stream);
  if (rc) return rc;
  const size_t cap = h->P.capL;
  PL_CUDA(cudaMemcpyAsync(keylines, h->d_kls, cap * B * sizeof(PLKeyLineRec), cudaMemcpyDeviceToHost, h->stream));
  PL_CUDA(cudaMemcpyAsync(desc, h->d_desc, cap * B * 32, cudaMemcpyDeviceToHost, h->stream));
  PL_CUDA(cudaMemcpyAsync(linefunc, h->d_lf, cap * B * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  PL_CUDA(cudaMemcpyAsync(n, h->d_nl, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  PL_CUDA(cudaStreamSynchronize(h->stream));
  int ov = 0;
  PL_CUDA(cudaMemcpy(&ov, h->d_overflow, sizeof(int), cudaMemcpyDeviceToHost));
  if (ov) { cudaMemset(h->d_overflow, 0, sizeof(int)); set_error("LSD produced more than segment_cap=%d segments", h->P.seg_cap); return PL_ERR_CAPACITY; }
  return PL_OK;
}

extern "C" int pl_line_extract(PLLine* h, const uint8_t* img, int stride, const uint8_t* mask, void* keylines,
                               uint8_t* desc, double* linefunc, int* n) {
  return pl_line_extract_batch(h, img, stride, 0, 1, mask, keylines, desc, linefunc, n);
}

// parity taps of the LAST call
extern "C" int pl_line_debug_segments(PLLine* h, int frame, float* out, int cap) {
  PL_ARG(h && frame >= 0 && frame < h->last_B);
  int n = 0;
  PL_CUDA(cudaStreamSynchronize(h->stream));
  PL_CUDA(cudaMemcpy(&n, h->d_nseg + frame, sizeof(int), cudaMemcpyDeviceToHost));
  if (out && n) PL_CUDA(cudaMemcpy(out, h->d_segs + (size_t)frame * h->P.seg_cap, sizeof(float4) * std::min(n, cap), cudaMemcpyDeviceToHost));
  return n;
}
extern "C" int pl_line_debug_scaled(PLLine* h, int frame, uint8_t* out, int* sw, int* sh) {
  PL_ARG(h && frame >= 0 && frame < h->last_B && sw && sh);
  *sw = h->P.sw; *sh = h->P.sh;
  PL_CUDA(cudaStreamSynchronize(h->stream));
  if (out) PL_CUDA(cudaMemcpy(out, h->d_scaled + (size_t)frame * h->P.npx, h->P.npx, cudaMemcpyDeviceToHost));
  return PL_OK;
}
extern "C" int pl_line_debug_sobel(PLLine* h, int frame, short* dx, short* dy) {
  PL_ARG(h && frame >= 0 && frame < h->last_B && dx && dy);
  const size_t n = (size_t)h->P.w * h->P.h;
  PL_CUDA(cudaStreamSynchronize(h->stream));
  std::vector<short2> tmp(n);
  PL_CUDA(cudaMemcpy(tmp.data(), h->d_dxy + frame * n, n * sizeof(short2), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; i++) { dx[i] = tmp[i].x; dy[i] = tmp[i].y; }
  return PL_OK;
}
extern "C" int pl_line_debug_order(PLLine* h, int frame, unsigned* out, int cap) {
  PL_ARG(h && frame >= 0 && frame < h->last_B);
  int n = 0;
  PL_CUDA(cudaStreamSynchronize(h->stream));
  PL_CUDA(cudaMemcpy(&n, h->d_ndef + frame, sizeof(int), cudaMemcpyDeviceToHost));
  if (out && n) {
    PL_CUDA(cudaMemcpy(out, h->d_order + (size_t)frame * h->P.npx, sizeof(unsigned) * std::min(n, cap), cudaMemcpyDeviceToHost));
    for (int i = 0; i < std::min(n, cap); i++) out[i] = (out[i] >> 16) * (unsigned)h->P.sw + (out[i] & 0xffffu);   // packed (x,y) -> y*sw+x
  }
  return n;
}
