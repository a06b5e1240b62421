// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_orb.cpp header for the rules).
//
// CPU restatement of the reference's descriptor matching path on flat arrays:
//   ORBmatcher::DescriptorDistance            src/ORBmatcher.cc:1764-1780   (== LSDmatcher.cpp:654-670)
//   Frame::AssignFeaturesToGrid / PosInGrid   src/Frame.cc:278-294, :893-903
//   Frame::AssignFeaturesToGridForLine        src/Frame.cc:296-320 + src/lineIterator.cpp
//   Frame::GetFeaturesInArea                  src/Frame.cc:713-766
//   Frame::GetFeaturesInAreaForLine           src/Frame.cc:768-842
//   ORBmatcher::SearchForInitialization       src/ORBmatcher.cc:455-572
//   ORBmatcher::SearchByProjection(F,Last)    src/ORBmatcher.cc:1441-1585   (mono branch)
//   ORBmatcher::SearchByProjection(F,points)  src/ORBmatcher.cc:56-152
//   ORBmatcher::ComputeThreeMaxima            src/ORBmatcher.cc:1718-1759
//   LSDmatcher::FrameBFMatch / lineDescriptorMAD / SearchDouble   src/LSDmatcher.cpp:440-486, :627-652
//   LSDmatcher::SearchByProjection (both)     src/LSDmatcher.cpp:72-176, :221-338
// cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) is third-party (OpenCV, not vendored): restated as "two smallest
// distances, ties -> lower train index first" and pinned against cv2 4.13 in tests/test_oracle_match.py.
// Parity status: knnMatch PINNED to cv2 4.13; everything else PARITY UNPINNED — the reference holds no golden vectors or
// tests for its matchers; each function follows the cited source lines, with known-answer tests on synthetic
// two-view geometry (tests/test_oracle_match.py, tests/test_oracle_localmap.py).
// Waived: the debug JPEG writes (LSDmatcher.cpp:67,173,422) and stereo branches (mvuRight) are not restated.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <climits>
#include <vector>
#include <algorithm>
#include <unordered_set>

namespace {
const int GRID_COLS = 64, GRID_ROWS = 48, HISTO_LENGTH = 30;

struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };
struct KeyLine {  // the fields of cv::line_descriptor::KeyLine the matchers read (flat copy)
  float startX, startY, endX, endY, lineLength, angle;
  int octave;
};

int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  const int32_t* pa = (const int32_t*)a;
  const int32_t* pb = (const int32_t*)b;
  int dist = 0;
  for (int i = 0; i < 8; i++, pa++, pb++) {
    unsigned int v = *pa ^ *pb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

struct Grid {
  float minX, minY, maxX, maxY, invW, invH;
  std::vector<int> cell[GRID_COLS][GRID_ROWS];
  void init(const float* b) {
    minX = b[0]; minY = b[1]; maxX = b[2]; maxY = b[3];
    invW = (float)GRID_COLS / (maxX - minX);
    invH = (float)GRID_ROWS / (maxY - minY);
  }
};

void assign_points(Grid& g, const KeyPoint* k, int n) {
  for (int i = 0; i < n; i++) {
    int px = (int)roundf((k[i].x - g.minX) * g.invW);
    int py = (int)roundf((k[i].y - g.minY) * g.invH);
    if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
    g.cell[px][py].push_back(i);
  }
}

void assign_lines(Grid& g, const KeyLine* kl, int n) {
  for (int i = 0; i < n; i++) {
    double x1 = kl[i].startX * g.invW, y1 = kl[i].startY * g.invH;  // float products widened to double
    double x2 = kl[i].endX * g.invW, y2 = kl[i].endY * g.invH;
    bool steep = std::abs(y2 - y1) > std::abs(x2 - x1);
    if (steep) { std::swap(x1, y1); std::swap(x2, y2); }
    if (x1 > x2) { std::swap(x1, x2); std::swap(y1, y2); }
    double dx = x2 - x1, dy = std::abs(y2 - y1), error = dx / 2.0;
    int ystep = (y1 < y2) ? 1 : -1;
    int x = (int)x1, y = (int)y1, maxX = (int)x2;
    while (x <= maxX) {
      int px = steep ? y : x, py = steep ? x : y;
      if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) g.cell[px][py].push_back(i);
      error -= dy;
      if (error < 0) { y += ystep; error += dx; }
      x++;
    }
  }
}

void features_in_area(const Grid& g, const KeyPoint* keys, float x, float y, float r, int minLevel, int maxLevel,
                      std::vector<int>& out) {
  out.clear();
  const int nMinCellX = std::max(0, (int)floorf((x - g.minX - r) * g.invW));
  if (nMinCellX >= GRID_COLS) return;
  const int nMaxCellX = std::min(GRID_COLS - 1, (int)ceilf((x - g.minX + r) * g.invW));
  if (nMaxCellX < 0) return;
  const int nMinCellY = std::max(0, (int)floorf((y - g.minY - r) * g.invH));
  if (nMinCellY >= GRID_ROWS) return;
  const int nMaxCellY = std::min(GRID_ROWS - 1, (int)ceilf((y - g.minY + r) * g.invH));
  if (nMaxCellY < 0) return;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
      for (int id : g.cell[ix][iy]) {
        const KeyPoint& kp = keys[id];
        if (bCheckLevels) {
          if (kp.octave < minLevel) continue;
          if (maxLevel >= 0 && kp.octave > maxLevel) continue;
        }
        const float distx = kp.x - x, disty = kp.y - y;
        if (fabsf(distx) < r && fabsf(disty) < r) out.push_back(id);
      }
}

void features_in_area_line(const Grid& g, const KeyLine* kls, const double* lfunc, float x1, float y1, float x2,
                           float y2, float r, float TH, std::vector<int>& out) {
  out.clear();
  std::unordered_set<int> seen;
  float x[3] = {x1, (float)((x1 + x2) / 2.0), x2};
  float y[3] = {y1, (float)((y1 + y2) / 2.0), y2};
  float d1x = x1 - x2, d1y = y1 - y2;
  float n1 = sqrtf(d1x * d1x + d1y * d1y);
  d1x /= n1; d1y /= n1;
  for (int i = 0; i < 3; i++) {
    const int nMinCellX = std::max(0, (int)floorf((x[i] - g.minX - r) * g.invW));
    if (nMinCellX >= GRID_COLS) continue;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)ceilf((x[i] - g.minX + r) * g.invW));
    if (nMaxCellX < 0) continue;
    const int nMinCellY = std::max(0, (int)floorf((y[i] - g.minY - r) * g.invH));
    if (nMinCellY >= GRID_ROWS) continue;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)ceilf((y[i] - g.minY + r) * g.invH));
    if (nMaxCellY < 0) continue;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
      for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
        for (int id : g.cell[ix][iy]) {
          if (seen.count(id)) continue;
          const KeyLine& kl = kls[id];
          float d2x = kl.startX - kl.endX, d2y = kl.startY - kl.endY;
          float n2 = sqrtf(d2x * d2x + d2y * d2y);
          d2x /= n2; d2y /= n2;
          float cosSita = fabsf(d1x * d2x + d1y * d2y);
          if (cosSita < TH) continue;
          const double* L = lfunc + 3 * id;
          const float dist = (float)(L[0] * x[i] + L[1] * y[i] + L[2]);
          if (fabsf(dist) < r) { out.push_back(id); seen.insert(id); }
        }
  }
}

void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

inline int rot_bin(float a1, float a2) {
  const float factor = 1.0f / HISTO_LENGTH;
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)roundf(rot * factor);
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

// BFMatcher(NORM_HAMMING).knnMatch(k=2): idx/dist [n1][2]; entries beyond n2 are -1
void bf_knn2(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* idx, int* dist) {
  for (int q = 0; q < n1; q++) {
    int b0 = INT_MAX, b1 = INT_MAX, i0 = -1, i1 = -1;
    for (int t = 0; t < n2; t++) {
      int d = descriptor_distance(d1 + 32 * q, d2 + 32 * t);
      if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = t; }
      else if (d < b1) { b1 = d; i1 = t; }
    }
    idx[2 * q] = i0; idx[2 * q + 1] = i1;
    dist[2 * q] = i0 >= 0 ? b0 : -1; dist[2 * q + 1] = i1 >= 0 ? b1 : -1;
  }
}

// LSDmatcher::FrameBFMatch with lineDescriptorMAD.  NL<2 on either side => no matches.
void frame_bf_match(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float TH, float nnratio, int* matches) {
  for (int i = 0; i < n1; i++) matches[i] = -1;
  if (n1 < 1 || n2 < 2) return;
  std::vector<int> idx(2 * n1), dist(2 * n1);
  bf_knn2(d1, n1, d2, n2, idx.data(), dist.data());
  // nn12 MAD: sort DESCENDING by (d1-d0), take element size/2; then |.-median| ascending, element size/2
  std::vector<float> d12(n1);
  for (int i = 0; i < n1; i++) d12[i] = (float)dist[2 * i + 1] - (float)dist[2 * i];
  std::vector<float> s = d12;
  std::sort(s.begin(), s.end(), [](float a, float b) { return a > b; });
  double med = s[n1 / 2];
  std::vector<float> dev(n1);
  for (int i = 0; i < n1; i++) dev[i] = fabsf((float)((double)d12[i] - med));
  std::sort(dev.begin(), dev.end());
  double nn12 = 1.4826 * dev[n1 / 2];
  nn12 = nn12 * 0.5;
  for (int i = 0; i < n1; i++) {
    double dist_12 = (double)((float)dist[2 * i + 1] - (float)dist[2 * i]);
    if (dist_12 > nn12 && (float)dist[2 * i] < TH && (float)dist[2 * i] < nnratio * (float)dist[2 * i + 1])
      matches[i] = idx[2 * i];
  }
}
}  // namespace

extern "C" {
int oracle_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// CSR export of Frame::mGrid / mGridForLine: cell index = ix*48+iy; start[3073], items[cap]; returns #items
int oracle_assign_grid(const void* keys, int n, const float* bounds, int* start, int* items) {
  Grid g; g.init(bounds);
  assign_points(g, (const KeyPoint*)keys, n);
  int k = 0;
  for (int ix = 0; ix < GRID_COLS; ix++)
    for (int iy = 0; iy < GRID_ROWS; iy++) {
      start[ix * GRID_ROWS + iy] = k;
      for (int id : g.cell[ix][iy]) items[k++] = id;
    }
  start[GRID_COLS * GRID_ROWS] = k;
  return k;
}
int oracle_assign_grid_lines(const void* kls, int n, const float* bounds, int* start, int* items, int cap) {
  Grid g; g.init(bounds);
  assign_lines(g, (const KeyLine*)kls, n);
  int k = 0;
  for (int ix = 0; ix < GRID_COLS; ix++)
    for (int iy = 0; iy < GRID_ROWS; iy++) {
      start[ix * GRID_ROWS + iy] = k;
      for (int id : g.cell[ix][iy]) { if (k < cap) items[k] = id; k++; }
    }
  start[GRID_COLS * GRID_ROWS] = k;
  return k;
}
int oracle_features_in_area(const void* keys, int n, const float* bounds, float x, float y, float r, int minLevel,
                            int maxLevel, int* out) {
  Grid g; g.init(bounds);
  assign_points(g, (const KeyPoint*)keys, n);
  std::vector<int> v;
  features_in_area(g, (const KeyPoint*)keys, x, y, r, minLevel, maxLevel, v);
  memcpy(out, v.data(), v.size() * sizeof(int));
  return (int)v.size();
}

// ORBmatcher::SearchForInitialization.  prev_matched: [n1][2] in/out; matches12: [n1] out.
int oracle_search_for_initialization(const void* keys1_, const uint8_t* desc1, int n1, const void* keys2_,
                                     const uint8_t* desc2, int n2, const float* bounds, float* prev_matched,
                                     int* matches12, int windowSize, float nnratio, int checkOri) {
  const KeyPoint* k1 = (const KeyPoint*)keys1_;
  const KeyPoint* k2 = (const KeyPoint*)keys2_;
  Grid g; g.init(bounds);
  assign_points(g, k2, n2);
  const int TH_LOW = 50;
  int nmatches = 0;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  std::vector<int> matchedDist(n2, INT_MAX), matches21(n2, -1), cand;
  for (int i1 = 0; i1 < n1; i1++) {
    if (k1[i1].octave > 0) continue;
    features_in_area(g, k2, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)windowSize, 0, 0, cand);
    if (cand.empty()) continue;
    int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
    for (int i2 : cand) {
      int dist = descriptor_distance(desc1 + 32 * i1, desc2 + 32 * i2);
      if (matchedDist[i2] <= dist) continue;
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
      else if (dist < bestDist2) bestDist2 = dist;
    }
    if (bestDist <= TH_LOW) {
      if (bestDist < (float)bestDist2 * nnratio) {
        if (matches21[bestIdx2] >= 0) { matches12[matches21[bestIdx2]] = -1; nmatches--; }
        matches12[i1] = bestIdx2;
        matches21[bestIdx2] = i1;
        matchedDist[bestIdx2] = bestDist;
        nmatches++;
        if (checkOri) rotHist[rot_bin(k1[i1].angle, k2[bestIdx2].angle)].push_back(i1);
      }
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i])
        if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
    }
  }
  for (int i1 = 0; i1 < n1; i1++)
    if (matches12[i1] >= 0) { prev_matched[2 * i1] = k2[matches12[i1]].x; prev_matched[2 * i1 + 1] = k2[matches12[i1]].y; }
  return nmatches;
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono=true).
// cur_match [n_cur]: in = 1 where mvpMapPoints[i] already holds an observed point (else -1/0 -> free);
// out = index of the last-frame point assigned, -1 none, -2 = pre-assigned (kept).
int oracle_search_by_projection_last(const void* keys_cur_, const uint8_t* desc_cur, int n_cur, const float* bounds,
                                     const float* Tcw, const float* K /*fx fy cx cy*/, const float* scaleFactors,
                                     int n_last, const uint8_t* last_valid, const float* last_pos,
                                     const uint8_t* last_desc, const int* last_octave, const float* last_angle,
                                     float th, int checkOri, const uint8_t* cur_preassigned, int* cur_match) {
  const KeyPoint* kc = (const KeyPoint*)keys_cur_;
  Grid g; g.init(bounds);
  assign_points(g, kc, n_cur);
  const int TH_HIGH = 100;
  int nmatches = 0;
  for (int i = 0; i < n_cur; i++) cur_match[i] = (cur_preassigned && cur_preassigned[i]) ? -2 : -1;
  std::vector<int> rotHist[HISTO_LENGTH], cand;
  for (int i = 0; i < n_last; i++) {
    if (!last_valid[i]) continue;
    const float* X = last_pos + 3 * i;
    // x3Dc = Rcw*x3Dw + tcw in fp32 (cv::Mat 3x3 * 3x1), separately rounded products and sums
    float xc = Tcw[0] * X[0] + Tcw[1] * X[1] + Tcw[2] * X[2] + Tcw[3];
    float yc = Tcw[4] * X[0] + Tcw[5] * X[1] + Tcw[6] * X[2] + Tcw[7];
    float zc = Tcw[8] * X[0] + Tcw[9] * X[1] + Tcw[10] * X[2] + Tcw[11];
    const float invzc = (float)(1.0 / zc);
    if (invzc < 0) continue;
    float u = K[0] * xc * invzc + K[2];
    float v = K[1] * yc * invzc + K[3];
    if (u < g.minX || u > g.maxX) continue;
    if (v < g.minY || v > g.maxY) continue;
    int oct = last_octave[i];
    float radius = th * scaleFactors[oct];
    features_in_area(g, kc, u, v, radius, oct - 1, oct + 1, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (cur_match[i2] != -1) continue;  // already holds a map point
      int dist = descriptor_distance(last_desc + 32 * i, desc_cur + 32 * i2);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= TH_HIGH) {
      cur_match[bestIdx2] = i;
      nmatches++;
      if (checkOri) rotHist[rot_bin(last_angle[i], kc[bestIdx2].angle)].push_back(bestIdx2);
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int j : rotHist[i]) { cur_match[j] = -1; nmatches--; }
  }
  return nmatches;
}

// ORBmatcher::SearchByProjection(F, vpMapPoints, th).  Per map point: in_view flag (mbTrackInView && !isBad),
// proj (mTrackProjX,Y), level (mnTrackScaleLevel), view_cos (mTrackViewCos), descriptor.
int oracle_search_by_projection_points(const void* keys_, const uint8_t* desc, int n, const float* bounds,
                                       const float* scaleFactors, int n_mp, const uint8_t* in_view,
                                       const float* proj, const int* level, const float* view_cos,
                                       const uint8_t* mp_desc, float th, float nnratio,
                                       const uint8_t* preassigned, int* match) {
  const KeyPoint* k = (const KeyPoint*)keys_;
  Grid g; g.init(bounds);
  assign_points(g, k, n);
  const int TH_HIGH = 100;
  int nmatches = 0;
  const bool bFactor = th != 1.0;
  for (int i = 0; i < n; i++) match[i] = (preassigned && preassigned[i]) ? -2 : -1;
  std::vector<int> cand;
  for (int iMP = 0; iMP < n_mp; iMP++) {
    if (!in_view[iMP]) continue;
    const int lvl = level[iMP];
    float r = (view_cos[iMP] > 0.998) ? 2.5f : 4.0f;
    if (bFactor) r *= th;
    features_in_area(g, k, proj[2 * iMP], proj[2 * iMP + 1], r * scaleFactors[lvl], lvl - 1, lvl, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : cand) {
      if (match[idx] != -1) continue;
      const int dist = descriptor_distance(mp_desc + 32 * iMP, desc + 32 * idx);
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = k[idx].octave; bestIdx = idx; }
      else if (dist < bestDist2) { bestLevel2 = k[idx].octave; bestDist2 = dist; }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      match[bestIdx] = iMP;
      nmatches++;
    }
  }
  return nmatches;
}

void oracle_bf_knn2(const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* idx, int* dist) {
  bf_knn2(d1, n1, d2, n2, idx, dist);
}
void oracle_frame_bf_match(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float TH, float nnratio, int* m) {
  frame_bf_match(d1, n1, d2, n2, TH, nnratio, m);
}
// LSDmatcher::SearchDouble(InitialFrame, CurrentFrame, LineMatches)
int oracle_search_double(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnratio, int* matches) {
  for (int i = 0; i < n1; i++) matches[i] = -1;
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<int> m1(n1), m2(n2);
  frame_bf_match(d1, n1, d2, n2, 50.f, nnratio, m1.data());
  frame_bf_match(d2, n2, d1, n1, 50.f, nnratio, m2.data());
  int nm = 0;
  for (int i = 0; i < n1; i++) {
    int j = m1[i];
    if (j >= 0) { if (m2[j] != i) m1[i] = -1; else nm++; }
  }
  memcpy(matches, m1.data(), n1 * sizeof(int));
  return nm;
}

// LSDmatcher::SearchByProjection(CurrentFrame, LastFrame, th): projected endpoints/in-frustum flags are inputs
// (Frame::isInFrustum(MapLine*) is Frame glue).  proj: [n_last][4] = X1,Y1,X2,Y2.
int oracle_line_search_by_projection_last(const void* kls_cur_, const double* lfunc_cur, const uint8_t* desc_cur,
                                          int n_cur, const float* bounds, int n_last, const uint8_t* last_valid,
                                          const float* proj, const uint8_t* last_desc, const float* last_length,
                                          float th, const uint8_t* preassigned, int* cur_match) {
  const KeyLine* kc = (const KeyLine*)kls_cur_;
  Grid g; g.init(bounds);
  assign_lines(g, kc, n_cur);
  const int TH_HIGH = 80;
  int nmatches = 0;
  for (int i = 0; i < n_cur; i++) cur_match[i] = (preassigned && preassigned[i]) ? -2 : -1;
  std::vector<int> cand;
  for (int i = 0; i < n_last; i++) {
    if (!last_valid[i]) continue;
    const float* p = proj + 4 * i;
    features_in_area_line(g, kc, lfunc_cur, p[0], p[1], p[2], p[3], th, 0.96f, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (cur_match[i2] != -1) continue;
      const int dist = descriptor_distance(last_desc + 32 * i, desc_cur + 32 * i2);
      float mx = std::max(last_length[i], kc[i2].lineLength), mn = std::min(last_length[i], kc[i2].lineLength);
      if (mn / mx < 0.75) continue;
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= TH_HIGH) { cur_match[bestIdx2] = i; nmatches++; }
  }
  return nmatches;
}

// LSDmatcher::SearchByProjection(F, vpMapLines, th)
int oracle_line_search_by_projection_lines(const void* kls_, const double* lfunc, const uint8_t* desc, int n,
                                           const float* bounds, int n_ml, const uint8_t* in_view, const float* proj,
                                           const float* view_cos, const uint8_t* ml_desc, float th, float nnratio,
                                           const uint8_t* preassigned, int* match) {
  const KeyLine* k = (const KeyLine*)kls_;
  Grid g; g.init(bounds);
  assign_lines(g, k, n);
  const int TH_HIGH = 80;
  int nmatches = 0;
  const bool bFactor = th != 1.0;
  for (int i = 0; i < n; i++) match[i] = (preassigned && preassigned[i]) ? -2 : -1;
  std::vector<int> cand;
  for (int iML = 0; iML < n_ml; iML++) {
    if (!in_view[iML]) continue;
    float r = (view_cos[iML] > 0.998) ? 5.0f : 8.0f;  // LSDmatcher::RadiusByViewingCos (LSDmatcher.cpp:1004-1010)
    if (bFactor) r *= th;
    const float* p = proj + 4 * iML;
    features_in_area_line(g, k, lfunc, p[0], p[1], p[2], p[3], r, 0.998f, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : cand) {
      if (match[idx] != -1) continue;
      const int dist = descriptor_distance(ml_desc + 32 * iML, desc + 32 * idx);
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = k[idx].octave; bestIdx = idx; }
      else if (dist < bestDist2) { bestLevel2 = k[idx].octave; bestDist2 = dist; }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      match[bestIdx] = iML;
      nmatches++;
    }
  }
  return nmatches;
}

// ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:720-911), monocular (bOnlyStereo = false, mvuRight < 0).
// DBoW2 feature vectors are inputs (node ids ascending as in std::map, items in insertion order): DBoW2 is a third-party
// dependency (Thirdparty/DBoW2) outside the path.  Literal quirk kept: this reference never sets vbMatched2, so every
// idx1 picks its best idx2 independently.  matches12[i] = idx2 or -1; returns nmatches.
int oracle_search_for_triangulation(const void* keys1_, const uint8_t* desc1, const uint8_t* has_mp1, int n1,
                                    const void* keys2_, const uint8_t* desc2, const uint8_t* has_mp2, int n2,
                                    const unsigned* fv1_nodes, const int* fv1_start, const int* fv1_items, int nn1,
                                    const unsigned* fv2_nodes, const int* fv2_start, const int* fv2_items, int nn2,
                                    const float* F12, const float* Cw1, const float* R2w, const float* t2w, const float* K2,
                                    const float* scaleFactors2, const float* levelSigma2_2, int check_orientation,
                                    int* matches12) {
  const KeyPoint* k1 = (const KeyPoint*)keys1_;
  const KeyPoint* k2 = (const KeyPoint*)keys2_;
  const int TH_LOW = 50;
  // epipole in image 2: C2 = R2w*Cw + t2w (cv::gemm fp32 order), ex = fx*C2x*invz + cx
  float C2[3];
  for (int i = 0; i < 3; i++) C2[i] = ((R2w[3 * i] * Cw1[0] + R2w[3 * i + 1] * Cw1[1]) + R2w[3 * i + 2] * Cw1[2]) + t2w[i];
  const float invz = 1.0f / C2[2];
  const float ex = K2[0] * C2[0] * invz + K2[2], ey = K2[1] * C2[1] * invz + K2[3];
  int nmatches = 0;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  int a = 0, b = 0;
  while (a < nn1 && b < nn2) {
    if (fv1_nodes[a] == fv2_nodes[b]) {
      for (int i1 = fv1_start[a]; i1 < fv1_start[a + 1]; i1++) {
        const int idx1 = fv1_items[i1];
        if (has_mp1[idx1]) continue;
        const KeyPoint& kp1 = k1[idx1];
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int i2 = fv2_start[b]; i2 < fv2_start[b + 1]; i2++) {
          const int idx2 = fv2_items[i2];
          if (has_mp2[idx2]) continue;
          const int dist = descriptor_distance(desc1 + 32 * idx1, desc2 + 32 * idx2);
          if (dist > TH_LOW || dist > bestDist) continue;
          const KeyPoint& kp2 = k2[idx2];
          const float distex = ex - kp2.x, distey = ey - kp2.y;
          if (distex * distex + distey * distey < 100 * scaleFactors2[kp2.octave]) continue;
          // CheckDistEpipolarLine (ORBmatcher.cc:155-172)
          const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
          const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
          const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
          const float num = la * kp2.x + lb * kp2.y + lc;
          const float den = la * la + lb * lb;
          if (den == 0) continue;
          const float dsqr = num * num / den;
          if (dsqr < 3.84 * levelSigma2_2[kp2.octave]) { bestIdx2 = idx2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) {
          matches12[idx1] = bestIdx2;
          nmatches++;
          if (check_orientation) rotHist[rot_bin(kp1.angle, k2[bestIdx2].angle)].push_back(idx1);
        }
      }
      a++; b++;
    } else if (fv1_nodes[a] < fv2_nodes[b]) {
      while (a < nn1 && fv1_nodes[a] < fv2_nodes[b]) a++;      // lower_bound
    } else {
      while (b < nn2 && fv2_nodes[b] < fv1_nodes[a]) b++;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i]) { matches12[idx1] = -1; nmatches--; }
    }
  }
  return nmatches;
}

// MapPoint::PredictScale(currentDist, pKF) (MapPoint.cc:396-411) on its own: log() of a float is libm's logf, the quotient is fp32
void oracle_predict_scale(const float* dist, const float* maxDist, int n, float logScaleFactor, int nScaleLevels, int* level) {
  for (int i = 0; i < n; i++) {
    const float ratio = maxDist[i] / dist[i];
    int lvl = (int)ceilf(logf(ratio) / logScaleFactor);
    if (lvl < 0) lvl = 0; else if (lvl >= nScaleLevels) lvl = nScaleLevels - 1;
    level[i] = lvl;
  }
}

// The search half of ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>, th) (src/ORBmatcher.cc:914-1065): for every map
// point the best keypoint of the keyframe (bestIdx, bestDist; -1/256 when it is skipped or nothing qualifies).  The map
// surgery that follows (Replace / AddObservation, :1036-1061) is sequential map logic outside the path; `skip` carries
// its loop-entry conditions (NULL, isBad, IsInKeyFrame).  Tcw row-major 4x4, Ow = camera centre.
void oracle_fuse_search(const void* keys_, const uint8_t* desc, int n, const float* bounds, const float* Tcw, const float* Ow,
                        const float* K, const float* scaleFactors, const float* invLevelSigma2, float logScaleFactor,
                        int nScaleLevels, int n_mp, const uint8_t* skip, const float* pos, const float* normal,
                        const float* minDist, const float* maxDist, const uint8_t* mp_desc, float th, int* best_idx,
                        int* best_dist) {
  const KeyPoint* k = (const KeyPoint*)keys_;
  Grid g; g.init(bounds);
  assign_points(g, k, n);
  std::vector<int> cand;
  for (int i = 0; i < n_mp; i++) {
    best_idx[i] = -1; best_dist[i] = 256;
    if (skip && skip[i]) continue;
    const float* P = pos + 3 * i;
    float Pc[3];
    for (int r = 0; r < 3; r++) Pc[r] = ((Tcw[4 * r] * P[0] + Tcw[4 * r + 1] * P[1]) + Tcw[4 * r + 2] * P[2]) + Tcw[4 * r + 3];
    if (Pc[2] < 0.0f) continue;
    const float invz = 1 / Pc[2];
    const float x = Pc[0] * invz, y = Pc[1] * invz;
    const float u = K[0] * x + K[2], v = K[1] * y + K[3];
    if (!(u >= bounds[0] && u < bounds[2] && v >= bounds[1] && v < bounds[3])) continue;       // KeyFrame::IsInImage
    const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
    const float dist3D = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
    if (dist3D < 0.8f * minDist[i] || dist3D > 1.2f * maxDist[i]) continue;   // Get{Min,Max}DistanceInvariance
    const float* Pn = normal + 3 * i;
    const double dot = (double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2];
    if (dot < 0.5 * dist3D) continue;
    const float ratio = maxDist[i] / dist3D;                                                   // MapPoint::PredictScale(dist, pKF)
    int lvl = (int)ceilf(logf(ratio) / logScaleFactor);
    if (lvl < 0) lvl = 0; else if (lvl >= nScaleLevels) lvl = nScaleLevels - 1;
    const float radius = th * scaleFactors[lvl];
    features_in_area(g, k, u, v, radius, -1, -1, cand);                                        // KeyFrame::GetFeaturesInArea
    int bestDist = 256, bestIdx = -1;
    for (int idx : cand) {
      const KeyPoint& kp = k[idx];
      if (kp.octave < lvl - 1 || kp.octave > lvl) continue;
      const float ex = u - kp.x, ey = v - kp.y;
      const float e2 = ex * ex + ey * ey;
      if (e2 * invLevelSigma2[kp.octave] > 5.99) continue;
      const int dist = descriptor_distance(mp_desc + 32 * i, desc + 32 * idx);
      if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
    best_idx[i] = bestIdx; best_dist[i] = bestDist;
  }
}

// LSDmatcher::SearchForTriangulation(pKF1, pKF2, vector<int>& vMatchedPairs, bool isDouble) (src/LSDmatcher.cpp:727-776)
int oracle_lsd_search_for_triangulation(const uint8_t* d1, const uint8_t* ml1, int n1, const uint8_t* d2, const uint8_t* ml2,
                                        int n2, float th, float nnratio, int isDouble, int* pairs) {
  for (int i = 0; i < n1; i++) pairs[i] = -1;
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<int> m1(n1), m2(n2);
  frame_bf_match(d1, n1, d2, n2, th, nnratio, m1.data());
  frame_bf_match(d2, n2, d1, n1, th, nnratio, m2.data());
  int nm = 0;
  for (int i = 0; i < n1; i++) {
    const int j = m1[i];
    if (j < 0) continue;
    if (isDouble && m2[j] != i) continue;
    if (ml1[i] || ml2[j]) continue;
    pairs[i] = j; nm++;
  }
  return nm;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (src/ORBmatcher.cc:187-327).
// has_mp_kf[i] = (vpMapPointsKF[i] && !isBad()).  matchesF[j] = index of the KF feature whose MapPoint the frame feature j
// receives (vpMapPointMatches[j] = vpMapPointsKF[matchesF[j]]), or -1.  Returns nmatches.
int oracle_search_by_bow(const void* keysKF_, const uint8_t* descKF, const uint8_t* has_mp_kf, int nKF, const void* keysF_,
                         const uint8_t* descF, int nF, const unsigned* fvK_nodes, const int* fvK_start, const int* fvK_items,
                         int nnK, const unsigned* fvF_nodes, const int* fvF_start, const int* fvF_items, int nnF, float nnratio,
                         int check_orientation, int* matchesF) {
  const KeyPoint* kK = (const KeyPoint*)keysKF_;
  const KeyPoint* kF = (const KeyPoint*)keysF_;
  const int TH_LOW = 50;
  (void)nKF;
  for (int j = 0; j < nF; j++) matchesF[j] = -1;
  int nmatches = 0;
  std::vector<int> rotHist[HISTO_LENGTH];
  int a = 0, b = 0;
  while (a < nnK && b < nnF) {
    if (fvK_nodes[a] == fvF_nodes[b]) {
      for (int iK = fvK_start[a]; iK < fvK_start[a + 1]; iK++) {
        const int idxK = fvK_items[iK];
        if (!has_mp_kf[idxK]) continue;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int iF = fvF_start[b]; iF < fvF_start[b + 1]; iF++) {
          const int idxF = fvF_items[iF];
          if (matchesF[idxF] >= 0) continue;
          const int dist = descriptor_distance(descKF + 32 * idxK, descF + 32 * idxF);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = idxF; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= TH_LOW) {
          if ((float)bestDist1 < nnratio * (float)bestDist2) {
            matchesF[bestIdxF] = idxK;
            if (check_orientation) rotHist[rot_bin(kK[idxK].angle, kF[bestIdxF].angle)].push_back(bestIdxF);
            nmatches++;
          }
        }
      }
      a++; b++;
    } else if (fvK_nodes[a] < fvF_nodes[b]) {
      while (a < nnK && fvK_nodes[a] < fvF_nodes[b]) a++;
    } else {
      while (b < nnF && fvF_nodes[b] < fvK_nodes[a]) b++;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int j : rotHist[i]) { matchesF[j] = -1; nmatches--; }
    }
  }
  return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
// (src/ORBmatcher.cc:1587-1716, Tracking::Relocalization).  kf_valid[i] = pMP && !isBad() && !sAlreadyFound.count(pMP);
// cur_preassigned[i2] = CurrentFrame.mvpMapPoints[i2] != NULL.  Unlike the LastFrame overload there is no invzc < 0 test,
// the level comes from MapPoint::PredictScale(dist3D, &CurrentFrame) and the threshold is ORBdist.
int oracle_search_by_projection_keyframe(const void* keys_cur_, const uint8_t* desc_cur, int n_cur, const float* bounds,
                                         const float* Tcw, const float* Ow, const float* K, const float* scaleFactors, int nlevels,
                                         float logScaleFactor, int n_kf, const uint8_t* kf_valid, const float* pos,
                                         const uint8_t* mp_desc, const float* minDist, const float* maxDist,
                                         const float* kf_angle, float th, int ORBdist, int checkOri,
                                         const uint8_t* cur_preassigned, int* cur_match) {
  const KeyPoint* kc = (const KeyPoint*)keys_cur_;
  Grid g; g.init(bounds);
  assign_points(g, kc, n_cur);
  int nmatches = 0;
  for (int i = 0; i < n_cur; i++) cur_match[i] = (cur_preassigned && cur_preassigned[i]) ? -2 : -1;
  std::vector<int> rotHist[HISTO_LENGTH], cand;
  for (int i = 0; i < n_kf; i++) {
    if (!kf_valid[i]) continue;
    const float* X = pos + 3 * i;
    float xc = Tcw[0] * X[0] + Tcw[1] * X[1] + Tcw[2] * X[2] + Tcw[3];
    float yc = Tcw[4] * X[0] + Tcw[5] * X[1] + Tcw[6] * X[2] + Tcw[7];
    float zc = Tcw[8] * X[0] + Tcw[9] * X[1] + Tcw[10] * X[2] + Tcw[11];
    const float invzc = (float)(1.0 / zc);
    float u = K[0] * xc * invzc + K[2];
    float v = K[1] * yc * invzc + K[3];
    if (u < g.minX || u > g.maxX) continue;
    if (v < g.minY || v > g.maxY) continue;
    const float PO[3] = {X[0] - Ow[0], X[1] - Ow[1], X[2] - Ow[2]};
    const float dist3D = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
    if (dist3D < 0.8f * minDist[i] || dist3D > 1.2f * maxDist[i]) continue;   // Get{Min,Max}DistanceInvariance
    const float ratio = maxDist[i] / dist3D;                                   // PredictScale: raw mfMaxDistance (MapPoint.cc:396-411)
    int lvl = (int)ceilf(logf(ratio) / logScaleFactor);
    if (lvl < 0) lvl = 0; else if (lvl >= nlevels) lvl = nlevels - 1;
    const float radius = th * scaleFactors[lvl];
    features_in_area(g, kc, u, v, radius, lvl - 1, lvl + 1, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : cand) {
      if (cur_match[i2] != -1) continue;
      const int dist = descriptor_distance(mp_desc + 32 * i, desc_cur + 32 * i2);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= ORBdist) {
      cur_match[bestIdx2] = i;
      nmatches++;
      if (checkOri) rotHist[rot_bin(kf_angle[i], kc[bestIdx2].angle)].push_back(bestIdx2);
    }
  }
  if (checkOri) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int j : rotHist[i]) { cur_match[j] = -1; nmatches--; }
  }
  return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (src/ORBmatcher.cc:574-709).
int oracle_search_by_bow_keyframes(const void* keys1_, const uint8_t* desc1, const uint8_t* mp1, int n1, const void* keys2_,
                                   const uint8_t* desc2, const uint8_t* mp2, int n2, const unsigned* fv1_nodes, const int* fv1_start,
                                   const int* fv1_items, int nn1, const unsigned* fv2_nodes, const int* fv2_start,
                                   const int* fv2_items, int nn2, float nnratio, int check_orientation, int* matches12) {
  const KeyPoint* k1 = (const KeyPoint*)keys1_;
  const KeyPoint* k2 = (const KeyPoint*)keys2_;
  const int TH_LOW = 50;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  std::vector<char> matched2(n2 > 0 ? n2 : 1, 0);
  std::vector<int> rotHist[HISTO_LENGTH];
  int nmatches = 0, a = 0, b = 0;
  while (a < nn1 && b < nn2) {
    if (fv1_nodes[a] == fv2_nodes[b]) {
      for (int i1 = fv1_start[a]; i1 < fv1_start[a + 1]; i1++) {
        const int idx1 = fv1_items[i1];
        if (!mp1[idx1]) continue;
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int i2 = fv2_start[b]; i2 < fv2_start[b + 1]; i2++) {
          const int idx2 = fv2_items[i2];
          if (matched2[idx2] || !mp2[idx2]) continue;
          const int dist = descriptor_distance(desc1 + 32 * idx1, desc2 + 32 * idx2);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 < TH_LOW) {
          if ((float)bestDist1 < nnratio * (float)bestDist2) {
            matches12[idx1] = bestIdx2; matched2[bestIdx2] = 1;
            if (check_orientation) rotHist[rot_bin(k1[idx1].angle, k2[bestIdx2].angle)].push_back(idx1);
            nmatches++;
          }
        }
      }
      a++; b++;
    } else if (fv1_nodes[a] < fv2_nodes[b]) {
      while (a < nn1 && fv1_nodes[a] < fv2_nodes[b]) a++;
    } else {
      while (b < nn2 && fv2_nodes[b] < fv1_nodes[a]) b++;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i]) { matches12[idx1] = -1; nmatches--; }
    }
  }
  return nmatches;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:249-314) for n_mp map points: descriptors of point m are rows
// offsets[m] .. offsets[m+1]) of desc (the observations in std::map order, bad keyframes already dropped by the caller).
// best[m] = index inside the point's list of the descriptor with the least median distance to the others (first wins),
// -1 when the list is empty (the reference returns early and keeps mDescriptor).  median = sorted[int(0.5 * (N - 1))].
void oracle_distinctive_descriptors(const uint8_t* desc, const int* offsets, int n_mp, int* best) {
  for (int m = 0; m < n_mp; m++) {
    const int N = offsets[m + 1] - offsets[m];
    best[m] = -1;
    if (N <= 0) continue;
    const uint8_t* d = desc + (size_t)offsets[m] * 32;
    std::vector<int> D((size_t)N * N, 0);
    for (int i = 0; i < N; i++)
      for (int j = i + 1; j < N; j++) { const int v = descriptor_distance(d + 32 * i, d + 32 * j); D[(size_t)i * N + j] = v; D[(size_t)j * N + i] = v; }
    int BestMedian = INT_MAX, BestIdx = 0;
    for (int i = 0; i < N; i++) {
      std::vector<int> v(D.begin() + (size_t)i * N, D.begin() + (size_t)(i + 1) * N);
      std::sort(v.begin(), v.end());
      const int median = v[(size_t)(0.5 * (N - 1))];
      if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    best[m] = BestIdx;
  }
}

// The search half of LSDmatcher::Fuse(pKF, vpMapLines, th) (src/LSDmatcher.cpp:860-1011), quirks kept:
//  * the FIRST map line with an end point behind the camera ends the whole call with `return false` (:907): stop_at = its
//    index (n_ml if none); lines from there on are not looked at, and the caller returns 0 instead of nFused;
//  * candidates come from KeyFrame::GetLinesInArea (KeyFrame.cc:647-682: midpoint within r, |cos| of the directions >= 0.998),
//    filtered by kl.octave in [level-1, level] with level = MapLine::PredictScale (unclamped ceil, MapLine.cpp:395-404);
//  * the map line's LBD descriptor is compared with row idx of pKF->mDescriptors - the keyframe's POINT descriptors, indexed
//    with the LINE index (:966).  Rows beyond the point descriptors do not exist (cv::Mat::row would be out of range): skipped;
//  * mvScaleFactorsLine[level] (:944) is read out of range for level outside [0, n_levels) (one octave: every level != 0):
//    restated as scale^level, the rule the table is built with (LineExtractor.cpp:7-14).
// skip[i] = !pML || isBad() || IsInKeyFrame(pKF); min/max_dist = raw mfMinDistance / mfMaxDistance.  best_idx = -1 /
// best_dist = 256 when skipped or nothing qualifies; the caller applies :986-1006 to lines with best_dist <= TH_LOW (50).
void oracle_lsd_fuse_search(const void* keylines_, int nl, const uint8_t* kf_point_desc, int n_pdesc, const float* bounds, const float* Tcw,
                            const float* Ow, const float* K, float scale_line, int n_line_levels, float logScaleFactorLine, int n_ml,
                            const uint8_t* skip, const double* pos, const double* normal, const float* minDist, const float* maxDist,
                            const uint8_t* ml_desc, float th, int* best_idx, int* best_dist, int* stop_at) {
  struct KL { float angle; int class_id; int octave; float ptx, pty; float response; float size; float sx, sy, ex, ey; float o[4]; float len; int npx; };
  static_assert(sizeof(KL) == 68, "KeyLine");
  const KL* kl = static_cast<const KL*>(keylines_);
  for (int i = 0; i < n_ml; i++) { best_idx[i] = -1; best_dist[i] = 256; }
  *stop_at = n_ml;
  const float t[3] = {Tcw[3], Tcw[7], Tcw[11]};
  for (int i = 0; i < n_ml; i++) {
    if (skip[i]) continue;
    const float SP[3] = {(float)pos[6 * i], (float)pos[6 * i + 1], (float)pos[6 * i + 2]};
    const float EP[3] = {(float)pos[6 * i + 3], (float)pos[6 * i + 4], (float)pos[6 * i + 5]};
    float S[3], E[3];
    for (int r = 0; r < 3; r++) {
      S[r] = Tcw[4 * r] * SP[0] + Tcw[4 * r + 1] * SP[1] + Tcw[4 * r + 2] * SP[2] + t[r];
      E[r] = Tcw[4 * r] * EP[0] + Tcw[4 * r + 1] * EP[1] + Tcw[4 * r + 2] * EP[2] + t[r];
    }
    if (S[2] < 0.0f || E[2] < 0.0f) { *stop_at = i; return; }
    const float invz1 = 1.0f / S[2], u1 = K[0] * S[0] * invz1 + K[2], v1 = K[1] * S[1] * invz1 + K[3];
    if (!(u1 >= bounds[0] && u1 < bounds[2] && v1 >= bounds[1] && v1 < bounds[3])) continue;      // KeyFrame::IsInImage
    const float invz2 = 1.0f / E[2], u2 = K[0] * E[0] * invz2 + K[2], v2 = K[1] * E[1] * invz2 + K[3];
    if (!(u2 >= bounds[0] && u2 < bounds[2] && v2 >= bounds[1] && v2 < bounds[3])) continue;
    float OM[3];
    for (int k = 0; k < 3; k++) OM[k] = (float)(0.5 * (double)(SP[k] + EP[k])) - Ow[k];
    const float dist = (float)std::sqrt((double)OM[0] * OM[0] + (double)OM[1] * OM[1] + (double)OM[2] * OM[2]);
    if (dist < 0.8f * minDist[i] || dist > 1.2f * maxDist[i]) continue;
    const float pn[3] = {(float)normal[3 * i], (float)normal[3 * i + 1], (float)normal[3 * i + 2]};
    const double dot = (double)OM[0] * pn[0] + (double)OM[1] * pn[1] + (double)OM[2] * pn[2];
    if (dot < 0.5 * dist) continue;
    const float ratio = maxDist[i] / dist;
    const int lvl = (int)ceilf(logf(ratio) / logScaleFactorLine);
    float sf = 1.0f;                                        // mvScaleFactorsLine[lvl]: cumulative fp32 products, 1/x below zero
    if (lvl >= 0) for (int k = 0; k < lvl; k++) sf = sf * scale_line;
    else { for (int k = 0; k < -lvl; k++) sf = sf * scale_line; sf = 1.0f / sf; }
    (void)n_line_levels;
    const float radius = th * sf;
    // KeyFrame::GetLinesInArea(u1, v1, u2, v2, radius, 0.998)
    float d1x = u1 - u2, d1y = v1 - v2;
    const float n1 = std::sqrt(d1x * d1x + d1y * d1y);
    d1x /= n1; d1y /= n1;
    int bestDist = 256, bestIdx = -1;
    for (int j = 0; j < nl; j++) {
      const float distance = (float)((double)((0.5 * (double)(u1 + u2) - (double)kl[j].ptx) * (0.5 * (double)(u1 + u2) - (double)kl[j].ptx)) +
                                     (double)((0.5 * (double)(v1 + v2) - (double)kl[j].pty) * (0.5 * (double)(v1 + v2) - (double)kl[j].pty)));
      if (distance > radius * radius) continue;
      float d2x = kl[j].sx - kl[j].ex, d2y = kl[j].sy - kl[j].ey;
      const float n2 = std::sqrt(d2x * d2x + d2y * d2y);
      d2x /= n2; d2y /= n2;
      const float CosSita = std::fabs(d1x * d2x + d1y * d2y);
      if (CosSita < 0.998f) continue;
      const int kpLevel = kl[j].octave;
      if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
      if (j >= n_pdesc) continue;                            // pKF->mDescriptors.row(idx) does not exist
      const int d = descriptor_distance(ml_desc + 32 * (size_t)i, kf_point_desc + 32 * (size_t)j);
      if (d < bestDist) { bestDist = d; bestIdx = j; }
    }
    best_idx[i] = bestIdx; best_dist[i] = bestDist;
  }
}
}
