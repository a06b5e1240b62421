// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_orb.cpp header for the rules).
//
// CPU restatement of the Frame glue around the hot path:
//   Frame::Frame (mono)  initUndistortRectifyMap + remap, every frame      src/Frame.cc:220-222
//   Frame::UndistortKeyPoints (cv::undistortPoints, 5 iterations)          src/Frame.cc:915-945
//   Frame::ComputeImageBounds                                             src/Frame.cc:947-985
//   Frame::isInFrustum(MapPoint*) / (MapLine*)                             src/Frame.cc:560-702
//   MapPoint::PredictScale / MapLine::PredictScale                         src/MapPoint.cc:413-428, src/MapLine.cpp:395-404
// OpenCV arithmetic (not vendored), PINNED against cv2 4.13 in tests/test_oracle_frame.py + tests/golden/frame_cv2.npz:
//   initUndistortRectifyMap (fp64 per-pixel model, fp32 maps), remap INTER_LINEAR 8U (1/32-pixel fixed point, 2x2 table of
//   15-bit weights, BORDER_CONSTANT 0), undistortPoints (5 fixed-point iterations in fp64), 3x3*3x1 fp32 gemm
//   (((a0*b0 + a1*b1) + a2*b2) + c, each operation rounded to fp32), cv::norm / Mat::dot (fp64 accumulation).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
struct Cam { double fx, fy, cx, cy, k1, k2, p1, p2, k3; };
Cam make_cam(const float* K, const float* D) { return Cam{(double)K[0], (double)K[1], (double)K[2], (double)K[3], (double)D[0], (double)D[1], (double)D[2], (double)D[3], (double)D[4]}; }

void undistort_map(const Cam& c, int w, int h, float* mx, float* my) {
  // ir = inv(K) for R = I, newK = K
  const double ir0 = 1.0 / c.fx, ir2 = -c.cx / c.fx, ir4 = 1.0 / c.fy, ir5 = -c.cy / c.fy;
  for (int i = 0; i < h; i++) {
    double _x = i * 0.0 + ir2, _y = i * ir4 + ir5, _w = i * 0.0 + 1.0;
    for (int j = 0; j < w; j++, _x += ir0, _y += 0.0, _w += 0.0) {
      double ww = 1. / _w, x = _x * ww, y = _y * ww;
      double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
      double kr = (1 + ((c.k3 * r2 + c.k2) * r2 + c.k1) * r2) / (1 + ((0 * r2 + 0) * r2 + 0) * r2);
      double xd = (x * kr + c.p1 * _2xy + c.p2 * (r2 + 2 * x2)), yd = (y * kr + c.p1 * (r2 + 2 * y2) + c.p2 * _2xy);
      mx[(size_t)i * w + j] = (float)(xd * c.fx + c.cx);
      my[(size_t)i * w + j] = (float)(yd * c.fy + c.cy);
    }
  }
}
void remap_table(int* tab /*[1024][4]*/) {   // int entries: 32768 (pure copy) does not fit the int16 OpenCV uses; the
  // saturated 32767 + 1-on-the-smallest-tap form OpenCV ends up with gives the same pixel for every input
  float t1[32][2];
  for (int i = 0; i < 32; i++) { float x = (float)i * (1.f / 32); t1[i][0] = 1.f - x; t1[i][1] = x; }
  for (int i = 0; i < 32; i++)
    for (int j = 0; j < 32; j++) {
      float wf[4] = {t1[i][0] * t1[j][0], t1[i][0] * t1[j][1], t1[i][1] * t1[j][0], t1[i][1] * t1[j][1]};
      int iw[4], isum = 0;
      for (int k = 0; k < 4; k++) { iw[k] = (int)lrintf(wf[k] * 32768.f); isum += iw[k]; }
      if (isum != 32768) {
        int diff = isum - 32768, mn = 0, mxk = 0;
        for (int k = 1; k < 4; k++) { if (iw[k] < iw[mn]) mn = k; if (iw[k] > iw[mxk]) mxk = k; }
        if (diff < 0) iw[mxk] -= diff; else iw[mn] -= diff;
      }
      for (int k = 0; k < 4; k++) tab[(i * 32 + j) * 4 + k] = iw[k];
    }
}
void remap_u8(const uint8_t* src, int w, int h, const float* mx, const float* my, uint8_t* dst) {
  static int tab[1024 * 4];
  static bool init = false;
  if (!init) { remap_table(tab); init = true; }
  for (int i = 0; i < w * h; i++) {
    int sx = (int)lrintf(mx[i] * 32.f), sy = (int)lrintf(my[i] * 32.f);
    int ix = sx >> 5, iy = sy >> 5;
    const int* t = tab + (((sy & 31) * 32) + (sx & 31)) * 4;
    auto px = [&](int y, int x) { return (x >= 0 && x < w && y >= 0 && y < h) ? (int)src[(size_t)y * w + x] : 0; };
    int acc = px(iy, ix) * t[0] + px(iy, ix + 1) * t[1] + px(iy + 1, ix) * t[2] + px(iy + 1, ix + 1) * t[3];
    dst[i] = (uint8_t)((acc + (1 << 14)) >> 15);
  }
}
void undistort_point(const Cam& c, float u, float v, float* ou, float* ov) {
  const double ifx = 1. / c.fx, ify = 1. / c.fy;
  double x = ((double)u - c.cx) * ifx, y = ((double)v - c.cy) * ify, x0 = x, y0 = y;
  for (int j = 0; j < 5; j++) {
    double r2 = x * x + y * y;
    double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((c.k3 * r2 + c.k2) * r2 + c.k1) * r2);
    if (icdist < 0) { x = ((double)u - c.cx) * ifx; y = ((double)v - c.cy) * ify; break; }
    double dX = 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x * x), dY = c.p1 * (r2 + 2 * y * y) + 2 * c.p2 * x * y;
    x = (x0 - dX) * icdist; y = (y0 - dY) * icdist;
  }
  *ou = (float)(x * c.fx + c.cx); *ov = (float)(y * c.fy + c.cy);
}
inline void gemm3(const float* R /*3x3 rows at stride 4 of Tcw*/, const float* X, const float* t, float* out) {
  for (int i = 0; i < 3; i++) out[i] = ((R[4 * i] * X[0] + R[4 * i + 1] * X[1]) + R[4 * i + 2] * X[2]) + t[i];
}
}  // namespace

extern "C" {
void oracle_undistort_map(const float* K, const float* D, int w, int h, float* mx, float* my) { undistort_map(make_cam(K, D), w, h, mx, my); }
void oracle_remap(const uint8_t* src, int w, int h, const float* mx, const float* my, uint8_t* dst) { remap_u8(src, w, h, mx, my, dst); }
void oracle_undistort_remap(const uint8_t* src, int w, int h, const float* K, const float* D, uint8_t* dst) {
  std::vector<float> mx((size_t)w * h), my((size_t)w * h);
  undistort_map(make_cam(K, D), w, h, mx.data(), my.data());
  remap_u8(src, w, h, mx.data(), my.data(), dst);
}
// kps: 28-byte records; only pt.x/pt.y change (UndistortKeyPoints); D[0]==0 -> copy
void oracle_undistort_keypoints(const void* kps, int n, const float* K, const float* D, void* out) {
  memcpy(out, kps, (size_t)n * 28);
  if (D[0] == 0.0f) return;
  Cam c = make_cam(K, D);
  for (int i = 0; i < n; i++) {
    const float* p = (const float*)((const char*)kps + 28 * i);
    float* o = (float*)((char*)out + 28 * i);
    undistort_point(c, p[0], p[1], &o[0], &o[1]);
  }
}
void oracle_image_bounds(const float* K, const float* D, int w, int h, float* b /*minX,minY,maxX,maxY*/) {
  if (D[0] != 0.0f) {
    Cam c = make_cam(K, D);
    float m[4][2];
    const float pts[4][2] = {{0, 0}, {(float)w, 0}, {0, (float)h}, {(float)w, (float)h}};
    for (int i = 0; i < 4; i++) undistort_point(c, pts[i][0], pts[i][1], &m[i][0], &m[i][1]);
    b[0] = std::fmin(m[0][0], m[2][0]); b[2] = std::fmax(m[1][0], m[3][0]);
    b[1] = std::fmin(m[0][1], m[1][1]); b[3] = std::fmax(m[2][1], m[3][1]);
  } else { b[0] = 0; b[1] = 0; b[2] = (float)w; b[3] = (float)h; }
}
// Frame::isInFrustum(MapPoint*, viewingCosLimit) for n points
void oracle_is_in_frustum_points(const float* Tcw, const float* Ow, const float* K, const float* bounds, float logScaleFactor,
                                 int nScaleLevels, float viewingCosLimit, int n, const float* pos, const float* normal,
                                 const float* minDist, const float* maxDist, uint8_t* inview, float* proj, int* level, float* viewcos) {
  const float t[3] = {Tcw[3], Tcw[7], Tcw[11]};
  for (int i = 0; i < n; i++) {
    inview[i] = 0; proj[2 * i] = proj[2 * i + 1] = 0; level[i] = 0; viewcos[i] = 0;
    const float* P = pos + 3 * i;
    float Pc[3];
    gemm3(Tcw, P, t, Pc);
    if (Pc[2] < 0.0f) continue;
    const float invz = 1.0f / Pc[2];
    const float u = K[0] * Pc[0] * invz + K[2], v = K[1] * Pc[1] * invz + K[3];
    if (u < bounds[0] || u > bounds[2]) continue;
    if (v < bounds[1] || v > bounds[3]) continue;
    const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
    const float dist = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
    if (dist < 0.8f * minDist[i] || dist > 1.2f * maxDist[i]) continue;   // MapPoint::Get{Min,Max}DistanceInvariance (MapPoint.cc:384-394)
    const float* Pn = normal + 3 * i;
    const float viewCos = (float)(((double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2]) / dist);
    if (viewCos < viewingCosLimit) continue;
    const float ratio = maxDist[i] / dist;
    int nScale = (int)ceilf(logf(ratio) / logScaleFactor);
    if (nScale < 0) nScale = 0; else if (nScale >= nScaleLevels) nScale = nScaleLevels - 1;
    inview[i] = 1; proj[2 * i] = u; proj[2 * i + 1] = v; level[i] = nScale; viewcos[i] = viewCos;
  }
}
// Frame::isInFrustum(MapLine*, viewingCosLimit): pos = 6 doubles (mWorldPos), normal = 3 doubles
void oracle_is_in_frustum_lines(const float* Tcw, const float* Ow, const float* K, const float* bounds, float logScaleFactor,
                                float viewingCosLimit, int n, const double* pos, const double* normal, const float* minDist,
                                const float* maxDist, uint8_t* inview, float* proj /*[n][4]*/, int* level, float* viewcos) {
  const float t[3] = {Tcw[3], Tcw[7], Tcw[11]};
  for (int i = 0; i < n; i++) {
    inview[i] = 0; for (int k = 0; k < 4; k++) proj[4 * i + k] = 0; level[i] = 0; viewcos[i] = 0;
    const float SP[3] = {(float)pos[6 * i], (float)pos[6 * i + 1], (float)pos[6 * i + 2]};
    const float EP[3] = {(float)pos[6 * i + 3], (float)pos[6 * i + 4], (float)pos[6 * i + 5]};
    float S[3], E[3];
    gemm3(Tcw, SP, t, S); gemm3(Tcw, EP, t, E);
    if (S[2] < 0.0f || E[2] < 0.0f) continue;
    const float invz1 = 1.0f / S[2], u1 = K[0] * S[0] * invz1 + K[2], v1 = K[1] * S[1] * invz1 + K[3];
    if (u1 < bounds[0] || u1 > bounds[2]) continue;
    if (v1 < bounds[1] || v1 > bounds[3]) continue;
    const float invz2 = 1.0f / E[2], u2 = K[0] * E[0] * invz2 + K[2], v2 = K[1] * E[1] * invz2 + K[3];
    if (u2 < bounds[0] || u2 > bounds[2]) continue;
    if (v2 < bounds[1] || v2 > bounds[3]) continue;
    float OM[3];   // 0.5*(SP+EP) - mOw as a cv::MatExpr: addWeighted-style fp32 evaluation
    for (int k = 0; k < 3; k++) OM[k] = (float)(0.5 * (double)(SP[k] + EP[k])) - Ow[k];
    const float dist = (float)std::sqrt((double)OM[0] * OM[0] + (double)OM[1] * OM[1] + (double)OM[2] * OM[2]);
    if (dist < 0.8f * minDist[i] || dist > 1.2f * maxDist[i]) continue;   // MapLine.cpp:383-393
    const float pn[3] = {(float)normal[3 * i], (float)normal[3 * i + 1], (float)normal[3 * i + 2]};
    const float viewCos = (float)(((double)OM[0] * pn[0] + (double)OM[1] * pn[1] + (double)OM[2] * pn[2]) / dist);
    if (viewCos < viewingCosLimit) continue;
    const float ratio = maxDist[i] / dist;
    inview[i] = 1; proj[4 * i] = u1; proj[4 * i + 1] = v1; proj[4 * i + 2] = u2; proj[4 * i + 3] = v2;
    level[i] = (int)ceilf(logf(ratio) / logScaleFactor); viewcos[i] = viewCos;
  }
}
}
