// LSD region growing as ORDERED SPECULATIVE EXECUTION — the per-lane state machine (host/device shared).
//
// What is computed: cv::LineSegmentDetector's seed loop (region_grow -> region2rect -> refine/reduce_region_radius), the
// part of LSDDetector::detect (Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:130-215 -> OpenCV imgproc lsd.cpp)
// that is sequential by definition: seeds are visited in descending gradient-bin order, a pixel consumed by an earlier
// region is unavailable to later ones, and the region angle changes with every added pixel.
//
// How: every seed (index i in the sorted order) is a TASK with priority i.  A task is executed by ONE lane, start to end,
// exactly as the CPU does it (same visiting order, same fp32/fp64 operation order, so the rectangle sums are bit-identical
// to the oracle's sequential sums).  Many tasks run concurrently; the result equals the sequential one because of three
// rules on the per-pixel ownership word `own` (first word of the 16-byte pixel record):
//   * claim     own = min(own, 2i) (atomicMin): the EARLIER task always wins a pixel.  If the previous owner was a later
//               task m > i, m is ABORTed (it used a pixel it would not have had) and is executed again.
//   * read      a pixel owned by an earlier task j < i is USED for task i; if j is not final yet and the pixel was aligned
//               (i.e. i would have taken it), i records a DEPENDENCY on it, re-checked when i becomes final.
//               A pixel that was free but not aligned needs no record: used or not, the outcome is "not added".
//   * finality  tasks become final in index order (the COMMITTER walks the status words): a task is valid when it is DONE,
//               not ABORTed, and its recorded dependencies still hold; the head task can always be re-executed
//               non-speculatively (nothing earlier is in flight), so the scheme cannot livelock.
// refine() releases pixels and grows again: a released pixel keeps the mark 2i+1 FOR GOOD, so that an earlier task taking
// it still aborts i (i's first growth had used it) at any time before i is final.  Once task i is final the odd mark
// reads as "free": everybody tests the parity and the finality of the owner (own >> 1 < frontier), and claims such a
// pixel with a compare-and-swap instead of the atomicMin.
//
// This file holds everything a LANE does (no warp collectives); the warp-level parts (seed scan, task hand-out, commit)
// are in line.cu for the GPU and in tools/grow_sim.cpp for the CPU protocol simulator that checks this very code against
// the oracle under random interleavings (tests/test_grow_protocol.py).
#pragma once

#ifdef __CUDACC__
#define LG_HD __device__ __forceinline__
#define LG_NOINL __device__ __noinline__
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#define LG_HD inline
#define LG_NOINL inline
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
#endif

namespace lg {

constexpr double kPI = 3.14159265358979323846;
constexpr double kDegToRads = kPI / 180;
constexpr int kFree = 0x7fffffff;   // own: nobody
constexpr int kNotDef = -1;         // own: gradient undefined (never available)
constexpr int kRing = 8;            // per-lane ring of recent list entries (shared memory on the GPU)

// per-seed status word: [2:0] state, [3] ABORT, [4] HASX (dependencies and/or a segment follow the list), [31:5] pool offset
constexpr unsigned ST_NONE = 0, ST_NOOP = 1, ST_EATEN = 2, ST_RUN = 3, ST_REDO = 5, ST_DONE = 7, ST_STATE = 7;
constexpr unsigned ST_ABORT = 8, ST_HASX = 16;
constexpr int ST_OFF_SHIFT = 5;
// per-frame control words
enum { C_NXT = 0, C_FIN = 1, C_LOCK = 2, C_NS = 3, C_REDO = 4, C_POOL = 5, C_ERR = 6, C_RQH = 7, C_STAT0 = 8, C_RQT = 15, C_WORDS = 16 };
constexpr int kRedoQ = 1024;       // lossy ring of aborted-but-published tasks (ctl[C_WORDS ..]); what it loses the committer re-executes at the head
constexpr int kCtlStride = C_WORDS + kRedoQ;
enum { ERR_POOL = 1, ERR_SEGCAP = 2, ERR_WATCHDOG = 4 };

struct Params {
  int sw, sh, npx, min_reg_size, seg_cap, lane_cap, pool_cap;
  int window;                // seeds are handed out at most this far ahead of the frontier (0: no limit)
  double prec, prec_hi, density_th;
};
struct Frame {
  int4* rec;                 // {own, angle (float degrees), cos, sin} per scaled pixel
  const float2* seedcs;      // cos/sin a region seeded at the pixel starts from
  const int* sq;             // gx^2 + gy^2
  const unsigned* order;     // seeds: (y << 16) | x, descending bin, row-major inside a bin
  int n;                     // number of seeds (= defined pixels)
  unsigned* st;              // status word per seed
  unsigned* pool;            // published region lists (+ overflow lane buffers), bump allocated from ctl[C_POOL]
  int* ctl;
  const double* wtab;        // sqrt(s / 4.0), s = gx^2 + gy^2
};

enum { P_IDLE = 0, P_START, P_GROW, P_RECT1, P_RECT2, P_RECT3, P_REFSTAT, P_REDUCE, P_FIN2, P_ROLLBACK, P_REDOSTART };

struct Lane {
  int phase, task;
  unsigned* buf; int cap;          // current private list
  unsigned* home; int home_cap;    // the lane's own buffer (buf moves into the pool when a region outgrows it)
  int base, cnt, hi, qi;           // region = buf[base .. base+cnt); hi = entries in use; qi = next entry to expand
  float sumdx, sumdy;
  double reg_angle, prec;
  int fast, stage, fresh;
  int ndep, dep0, dep1;            // dependencies: buf[cap-1-k], k < ndep (dep0/dep1: the last two, to drop immediate repeats)
  int j, m;
  double a0, a1, a2, a3, a4, a5;
  double cx, cy, dx, dy;
  double x1, y1, x2, y2, width;
  double xc, yc, radSq;
  int hasseg;
  unsigned off;                    // FIN2: pool offset; REDOSTART: list being released
  unsigned* ring;                  // the last kRing pushed entries of the list, in fast memory (the queue is popped from it)
  unsigned curp;                   // the entry being expanded
  int pend[8]; int npend;          // claims issued by the previous growing step, settled at the start of the next step
  // snapshot of the neighbourhood (kept in the struct only so that the simulator can split load and use)
  int loaded, F;
  unsigned stw;
  int4 nb[8];
};

// ------------------------------------------------------------------------------------------------ memory operations
#ifdef __CUDACC__
LG_HD int4 ld_rec(const int4* p) {
  int4 v;
  asm volatile("ld.relaxed.gpu.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
LG_HD int ld_i(const int* p) { int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
LG_HD unsigned ld_u(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
LG_HD void st_u(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
LG_HD void st_i(int* p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
LG_HD int a_min(int* p, int v) { return atomicMin(p, v); }
LG_HD int a_cas(int* p, int c, int v) { return atomicCAS(p, c, v); }
LG_HD int a_add(int* p, int v) { return atomicAdd(p, v); }
LG_HD int a_exch(int* p, int v) { return atomicExch(p, v); }
LG_HD int a_max(int* p, int v) { return atomicMax(p, v); }
LG_HD unsigned a_or(unsigned* p, unsigned v) { return atomicOr(p, v); }
LG_HD unsigned a_and(unsigned* p, unsigned v) { return atomicAnd(p, v); }
LG_HD void fence() { __threadfence(); }
LG_HD float as_float(int b) { return __int_as_float(b); }
LG_HD int ldg_i(const int* p) { return __ldg(p); }
LG_HD unsigned ldg_u(const unsigned* p) { return __ldg(p); }
LG_HD float2 ldg_f2(const float2* p) { return __ldg(p); }
LG_HD double ldg_d(const double* p) { return __ldg(p); }
LG_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
LG_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
LG_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
LG_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
#else
LG_HD int4 ld_rec(const int4* p) { return *p; }
LG_HD int ld_i(const int* p) { return *p; }
LG_HD unsigned ld_u(const unsigned* p) { return *p; }
LG_HD void st_u(unsigned* p, unsigned v) { *p = v; }
LG_HD void st_i(int* p, int v) { *p = v; }
LG_HD int a_min(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
LG_HD int a_cas(int* p, int c, int v) { int o = *p; if (o == c) *p = v; return o; }
LG_HD int a_add(int* p, int v) { int o = *p; *p = o + v; return o; }
LG_HD int a_exch(int* p, int v) { int o = *p; *p = v; return o; }
LG_HD int a_max(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
LG_HD unsigned a_or(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
LG_HD unsigned a_and(unsigned* p, unsigned v) { unsigned o = *p; *p = o & v; return o; }
LG_HD void fence() {}
LG_HD float as_float(int b) { float f; memcpy(&f, &b, 4); return f; }
LG_HD int ldg_i(const int* p) { return *p; }
LG_HD unsigned ldg_u(const unsigned* p) { return *p; }
LG_HD float2 ldg_f2(const float2* p) { return *p; }
LG_HD double ldg_d(const double* p) { return *p; }
LG_HD float f_add(float a, float b) { return a + b; }   // host build: -ffp-contract=off
LG_HD float f_sub(float a, float b) { return a - b; }
LG_HD float f_mul(float a, float b) { return a * b; }
LG_HD float f_div(float a, float b) { return a / b; }
#endif

// cv::fastAtan2 (degrees), fp32 without FMA
LG_HD float fast_atan2_deg(float y, float x) {
  const float k = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
  const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
  const float eps = 2.220446049250313e-16f;
  const float ax = fabsf(x), ay = fabsf(y);
  const float c = f_div(fminf(ax, ay), f_add(fmaxf(ax, ay), eps)), c2 = f_mul(c, c);
  float a = f_mul(f_add(f_mul(f_add(f_mul(f_add(f_mul(p7, c2), p5), c2), p3), c2), p1), c);
  if (ax < ay) a = f_sub(90.f, a);
  if (x < 0) a = f_sub(180.f, a);
  if (y < 0) a = f_sub(360.f, a);
  return a;
}
LG_HD double angle_diff_signed(double a, double b) {
  double diff = a - b;
  while (diff <= -kPI) diff += 2 * kPI;
  while (diff > kPI) diff -= 2 * kPI;
  return diff;
}
LG_HD double dist_d(double x1, double y1, double x2, double y2) { return sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1)); }
LG_HD double dist_sq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }

// LineSegmentDetectorImpl::isAligned on a defined pixel.  fast: prec < pi/2, folded to  n <= prec || n >= prec_hi
// (prec_hi = smallest double with 2pi - n <= prec; 2pi - n is exact for n in [pi, 4pi], so both forms agree bit for bit)
LG_HD bool aligned(const Params& P, const Lane& L, double a) {
  const double n1 = fabs(L.reg_angle - a);
  if (L.fast) return (n1 <= P.prec) || (n1 >= P.prec_hi);
  const double n2 = fabs(n1 - 2 * kPI);
  return ((n1 > (3 * kPI) / 2) ? n2 : n1) <= L.prec;
}

LG_HD void lane_reset(Lane& L) {      // (a lane keeps the pool block a large region once made it move to: blocks are never freed)
  L.base = 0; L.cnt = 0; L.hi = 0; L.qi = 0; L.stage = 0; L.ndep = 0; L.dep0 = -1; L.dep1 = -1; L.hasseg = 0; L.loaded = 0; L.npend = 0;
}
LG_HD void frame_error(const Frame& Fm, int code) {
  a_or(reinterpret_cast<unsigned*>(&Fm.ctl[C_ERR]), (unsigned)code);
  a_max(&Fm.ctl[C_FIN], Fm.n);         // give the frame up: every warp of the group leaves its loop (C_FIN only grows)
}
// the private list is full: continue in a pool block of twice the size (scalars only: the Lane stays in registers)
LG_NOINL unsigned* buffer_grow(int pool_cap, unsigned* pool, int* ctl, int n, unsigned* buf, int cap, int used, int ndep) {
  const int ncap = 2 * cap;
  const int off = a_add(&ctl[C_POOL], ncap);
  if (off + ncap > pool_cap) {
    a_or(reinterpret_cast<unsigned*>(&ctl[C_ERR]), (unsigned)ERR_POOL);
    a_max(&ctl[C_FIN], n);
    return nullptr;
  }
  unsigned* nb = pool + off;
  for (int k = 0; k < used; k++) nb[k] = buf[k];
  for (int k = 1; k <= ndep; k++) nb[ncap - k] = buf[cap - k];   // the dependency list lives at the far end
  return nb;
}
LG_HD bool lane_buffer_grow(const Params& P, const Frame& Fm, Lane& L) {
  const int used = (L.base + L.cnt > L.hi) ? L.base + L.cnt : L.hi;   // entries appended in the current step are not in hi yet
  unsigned* nb = buffer_grow(P.pool_cap, Fm.pool, Fm.ctl, Fm.n, L.buf, L.cap, used, L.ndep);
  if (!nb) return false;
  L.buf = nb; L.cap = 2 * L.cap;
  return true;
}
LG_HD void begin_rollback(Lane& L) { L.phase = P_ROLLBACK; L.j = 0; L.loaded = 0; }
LG_HD void to_fin(Lane& L) {
  L.j = 0;
  L.phase = P_FIN2;
  L.off = 0;
}
LG_HD void to_rect1(Lane& L) { L.phase = P_RECT1; L.j = 0; L.a0 = 0; L.a1 = 0; L.a2 = 0; }
LG_HD bool record_dep(const Params& P, const Frame& Fm, Lane& L, int idx) {
  if (idx == L.dep0 || idx == L.dep1) return true;
  const int used = (L.base + L.cnt > L.hi) ? L.base + L.cnt : L.hi;
  if (used + L.ndep + 1 > L.cap) { if (!lane_buffer_grow(P, Fm, L)) return false; }
  L.buf[L.cap - 1 - L.ndep] = (unsigned)idx;
  L.ndep++;
  L.dep1 = L.dep0; L.dep0 = idx;
  return true;
}
// first pixel of a (re)grown region: the seed, already claimed
LG_HD void seed_region(const Params& P, const Frame& Fm, Lane& L, unsigned pix, int sidx, int angbits) {
  L.buf[L.base] = pix;
  L.ring[0] = pix;
  L.cnt = 1; L.qi = 0;
  if (L.base + 1 > L.hi) L.hi = L.base + 1;
  const float2 s0 = ldg_f2(&Fm.seedcs[sidx]);
  L.sumdx = s0.x; L.sumdy = s0.y;
  L.reg_angle = (double)as_float(angbits) * kDegToRads;
  L.phase = P_GROW;
  L.loaded = 0;
}

// ------------------------------------------------------------------------------------------------ phases
// how task i (me = 2i) reads an ownership word o, F = frontier (every task < F is final)
//   candidate: free, a later task's mark, my own dropped pixel, or a pixel dropped by a FINAL earlier task
//   blocked  : held (even) or dropped (odd) by an earlier task that is not known to be final -> dependency if aligned
LG_HD bool own_candidate(int o, int me, int F) { return o > me || (o >= 0 && o < me && (o & 1) && (o >> 1) < F); }
LG_HD bool own_blocked(int o, int me, int F) { return o >= 0 && o < me && (o >> 1) >= F; }
// a later task lost a pixel to an earlier one: it has to run again.  If it is already published, nobody is looking
// at its status word, so it goes on the (lossy) redo ring that idle lanes serve before they take new seeds.
LG_HD void abort_task(const Frame& Fm, int m) {
  const unsigned old = a_or(&Fm.st[m], ST_ABORT);
  if ((old & ST_STATE) == ST_DONE && !(old & ST_ABORT)) {
    const int t = a_add(&Fm.ctl[C_RQT], 1);
    st_i(&Fm.ctl[C_WORDS + (t & (kRedoQ - 1))], m + 1);
  }
}
// claim a candidate seen as o.  Result: kFree = clean; a value below me = somebody earlier was faster (the attempt is
// void); anything else = the mark of the later task the pixel was taken from (settle_claim aborts it)
LG_HD int issue_claim(const Frame& Fm, int idx, int o, int me) {
  if (o > me) return a_min(&Fm.rec[idx].x, me);
  return (a_cas(&Fm.rec[idx].x, o, me) == o) ? kFree : -2;   // dropped by a final task: the word is below me, atomicMin cannot take it
}
LG_HD bool settle_claim(const Frame& Fm, int r, int me) {
  if (r == kFree) return true;
  if (r < me) return false;
  if ((r >> 1) != (me >> 1)) abort_task(Fm, r >> 1);
  return true;
}
LG_HD bool own_claim(const Frame& Fm, int idx, int o, int me) { return settle_claim(Fm, issue_claim(Fm, idx, o, me), me); }

LG_HD void step_start(const Params& P, const Frame& Fm, Lane& L) {
  const int i = L.task, me = 2 * i;
  const unsigned pix = ldg_u(&Fm.order[i]);
  const int sidx = (int)(pix >> 16) * P.sw + (int)(pix & 0xffffu);
  const int4 r = ld_rec(&Fm.rec[sidx]);
  const int F = ld_i(&Fm.ctl[C_FIN]);
  lane_reset(L);
  if (!own_candidate(r.x, me, F)) {     // consumed by an earlier region (or dropped by one that is not final yet)
    st_u(&Fm.st[i], (!(r.x & 1) && (r.x >> 1) < F) ? ST_NOOP : ST_EATEN);
    L.phase = P_IDLE;
    return;
  }
  if (L.fresh) st_u(&Fm.st[i], ST_RUN);
  fence();                              // RUN (and a cleared ABORT) must be in place before the first claim can be stolen
  if (!own_claim(Fm, sidx, r.x, me)) {  // lost the seed between the load and the claim: look again
    L.fresh = 0;                        // (the status word is RUN already)
    return;
  }
  L.prec = P.prec; L.fast = 1;
  seed_region(P, Fm, L, pix, sidx, r.y);
}

LG_HD void grow_done(const Params& P, Lane& L) {
  const int minsz = (L.stage == 0) ? P.min_reg_size : 2;
  if (L.cnt < minsz) { L.hasseg = 0; to_fin(L); }
  else to_rect1(L);
}

template <bool SPLIT>
LG_HD void step_grow(const Params& P, const Frame& Fm, Lane& L) {
  if (!L.loaded) {
    if (L.qi == L.cnt) { grow_done(P, L); return; }
    const unsigned p = (L.cnt - L.qi <= kRing) ? L.ring[L.qi & (kRing - 1)] : L.buf[L.base + L.qi];
    L.curp = p;
    const int x = (int)(p & 0xffffu), y = (int)(p >> 16);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int kk = k + (k >= 4), xx = x + kk % 3 - 1, yy = y + kk / 3 - 1;
      int4 v; v.x = kNotDef; v.y = 0; v.z = 0; v.w = 0;
      if (xx >= 0 && yy >= 0 && xx < P.sw && yy < P.sh) v = ld_rec(&Fm.rec[yy * P.sw + xx]);
      L.nb[k] = v;
    }
    L.stw = ld_u(&Fm.st[L.task]);
    L.F = ld_i(&Fm.ctl[C_FIN]);
    L.loaded = 1;
    if (SPLIT) return;
  }
  L.loaded = 0;
  if (L.stw & ST_ABORT) { begin_rollback(L); return; }
  const unsigned p = L.curp;
  const int x = (int)(p & 0xffffu), y = (int)(p >> 16);
  L.qi++;
  const int me = 2 * L.task;
  // results of the claims: looked at by the NEXT step of this lane (settle_pending), so the atomics are in flight during
  // the rest of this iteration; a stolen-from task learns it one iteration later, a lost race costs one wasted step
#pragma unroll
  for (int k = 0; k < 8; k++) {
    L.pend[k] = kFree;
    const int o = L.nb[k].x;
    const bool cand = own_candidate(o, me, L.F);
    if (cand || own_blocked(o, me, L.F)) {
      const double a = (double)as_float(L.nb[k].y) * kDegToRads;
      if (aligned(P, L, a)) {
        const int kk = k + (k >= 4), xx = x + kk % 3 - 1, yy = y + kk / 3 - 1;
        if (cand) {
          if (L.base + L.cnt + L.ndep >= L.cap) { if (!lane_buffer_grow(P, Fm, L)) { L.phase = P_IDLE; return; } }
          L.pend[k] = issue_claim(Fm, yy * P.sw + xx, o, me);
          L.buf[L.base + L.cnt] = (unsigned)xx | ((unsigned)yy << 16);
          L.ring[L.cnt & (kRing - 1)] = (unsigned)xx | ((unsigned)yy << 16);
          L.cnt++;
          L.sumdx = f_add(L.sumdx, as_float(L.nb[k].z));
          L.sumdy = f_add(L.sumdy, as_float(L.nb[k].w));
          L.reg_angle = (double)fast_atan2_deg(L.sumdy, L.sumdx) * kDegToRads;
        } else {
          if (!record_dep(P, Fm, L, yy * P.sw + xx)) { L.phase = P_IDLE; return; }
        }
      }
    }
  }
  if (L.base + L.cnt > L.hi) L.hi = L.base + L.cnt;
  L.npend = 1;
}

// region2rect, pass 1: weighted centroid (sequential fp64 sums in list order == the oracle's order)
LG_HD void step_rect1(const Params& P, const Frame& Fm, Lane& L) {
#pragma unroll 4
  for (int t = 0; t < 32; t++) {
    if (L.j < L.cnt) {
      const unsigned p = L.buf[L.base + L.j];
      const int px = (int)(p & 0xffffu), py = (int)(p >> 16);
      const double w = ldg_d(&Fm.wtab[ldg_i(&Fm.sq[py * P.sw + px])]);
      L.a0 += (double)px * w; L.a1 += (double)py * w; L.a2 += w;
      L.j++;
    }
  }
  if (L.j == L.cnt) {
    L.cx = L.a0 / L.a2; L.cy = L.a1 / L.a2;
    L.phase = P_RECT2; L.j = 0; L.a3 = 0; L.a4 = 0; L.a5 = 0;
  }
}
// pass 2: inertia (get_theta)
LG_HD void step_rect2(const Params& P, const Frame& Fm, Lane& L) {
#pragma unroll 4
  for (int t = 0; t < 32; t++) {
    if (L.j < L.cnt) {
      const unsigned p = L.buf[L.base + L.j];
      const int px = (int)(p & 0xffffu), py = (int)(p >> 16);
      const double w = ldg_d(&Fm.wtab[ldg_i(&Fm.sq[py * P.sw + px])]);
      const double dx = (double)px - L.cx, dy = (double)py - L.cy;
      L.a3 += dy * dy * w; L.a4 += dx * dx * w; L.a5 -= dx * dy * w;
      L.j++;
    }
  }
  if (L.j == L.cnt) {
    const double Ixx = L.a3, Iyy = L.a4, Ixy = L.a5;
    const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)(lambda - Ixx), (float)Ixy)
                                           : (double)fast_atan2_deg((float)Ixy, (float)(lambda - Iyy));
    theta *= kDegToRads;
    if (fabs(angle_diff_signed(theta, L.reg_angle)) > P.prec) theta += kPI;
    L.dx = cos(theta); L.dy = sin(theta);
    L.phase = P_RECT3; L.j = 0; L.a0 = 0; L.a1 = 0; L.a2 = 0; L.a3 = 0;   // l_min, l_max, w_min, w_max
  }
}
LG_HD void density_decision(const Params& P, const Frame& Fm, Lane& L) {
  const double density = (double)L.cnt / (dist_d(L.x1, L.y1, L.x2, L.y2) * L.width);
  if (density >= P.density_th) { L.hasseg = 1; to_fin(L); return; }
  if (L.stage == 0) {                                   // refine(): release the region, measure the angle spread near the seed
    const unsigned p0 = L.buf[L.base];
    L.xc = (double)(int)(p0 & 0xffffu); L.yc = (double)(int)(p0 >> 16);
    L.phase = P_REFSTAT; L.j = 0; L.m = 0; L.a0 = 0; L.a1 = 0;
    return;
  }
  if (L.stage == 1) {                                   // reduce_region_radius(): initial radius
    const double r1 = dist_sq(L.xc, L.yc, L.x1, L.y1), r2 = dist_sq(L.xc, L.yc, L.x2, L.y2);
    L.radSq = r1 > r2 ? r1 : r2;
    L.stage = 2;
  }
  L.radSq *= 0.75 * 0.75;
  L.phase = P_REDUCE; L.j = 0;
}
// pass 3: extents along / across theta
LG_HD void step_rect3(const Params& P, const Frame& Fm, Lane& L) {
#pragma unroll 4
  for (int t = 0; t < 32; t++) {
    if (L.j < L.cnt) {
      const unsigned p = L.buf[L.base + L.j];
      const double rdx = (double)(int)(p & 0xffffu) - L.cx, rdy = (double)(int)(p >> 16) - L.cy;
      const double l = rdx * L.dx + rdy * L.dy, w = -rdx * L.dy + rdy * L.dx;
      if (l > L.a1) L.a1 = l; else if (l < L.a0) L.a0 = l;
      if (w > L.a3) L.a3 = w; else if (w < L.a2) L.a2 = w;
      L.j++;
    }
  }
  if (L.j == L.cnt) {
    L.x1 = L.cx + L.a0 * L.dx; L.y1 = L.cy + L.a0 * L.dy;
    L.x2 = L.cx + L.a1 * L.dx; L.y2 = L.cy + L.a1 * L.dy;
    L.width = L.a3 - L.a2;
    if (L.width < 1.0) L.width = 1.0;
    density_decision(P, Fm, L);
  }
}
// refine(): every pixel of the region becomes NOTUSED (tentative mark 2i+1: see the header), statistics of the angle
// differences to the seed inside the rectangle width; then the region is grown again with tolerance tau
LG_HD void step_refstat(const Params& P, const Frame& Fm, Lane& L) {
  const int me = 2 * L.task;
  const unsigned p0 = L.buf[L.base];
  const int sidx = (int)(p0 >> 16) * P.sw + (int)(p0 & 0xffffu);
  const double ang_c = (double)as_float(Fm.rec[sidx].y) * kDegToRads;
#pragma unroll 2
  for (int t = 0; t < 8; t++) {
    if (L.j < L.cnt) {
      const unsigned p = L.buf[L.base + L.j];
      const int px = (int)(p & 0xffffu), py = (int)(p >> 16), idx = py * P.sw + px;
      a_cas(&Fm.rec[idx].x, me, me + 1);
      if (dist_d(L.xc, L.yc, (double)px, (double)py) < L.width) {
        const double ang_d = angle_diff_signed((double)as_float(Fm.rec[idx].y) * kDegToRads, ang_c);
        L.a0 += ang_d; L.a1 += ang_d * ang_d; L.m++;
      }
      L.j++;
    }
  }
  if (L.j == L.cnt) {
    const double mean_angle = L.a0 / (double)L.m;
    const double tau = 2.0 * sqrt((L.a1 - 2.0 * mean_angle * L.a0) / (double)L.m + mean_angle * mean_angle);
    // grow again from the same seed, behind the old list (the old pixels are still needed for the final release)
    L.base = L.hi;
    if (L.base + L.ndep >= L.cap) { if (!lane_buffer_grow(P, Fm, L)) { L.phase = P_IDLE; return; } }
    if (!own_claim(Fm, sidx, me + 1, me)) { begin_rollback(L); return; }   // stolen meanwhile (the ABORT flag is on its way)
    L.prec = tau; L.fast = 0; L.stage = 1;
    seed_region(P, Fm, L, p0, sidx, Fm.rec[sidx].y);
  }
}
// reduce_region_radius(): drop the pixels beyond the shrinking radius (swap-remove, as the reference does: the order of
// the survivors decides the order of the next rectangle sums)
LG_HD void step_reduce(const Params& P, const Frame& Fm, Lane& L) {
  const int me = 2 * L.task;
#pragma unroll 2
  for (int t = 0; t < 8; t++) {
    if (L.j < L.cnt) {
      const unsigned p = L.buf[L.base + L.j];
      const int px = (int)(p & 0xffffu), py = (int)(p >> 16);
      if (dist_sq(L.xc, L.yc, (double)px, (double)py) > L.radSq) {
        a_cas(&Fm.rec[py * P.sw + px].x, me, me + 1);
        const unsigned q = L.buf[L.base + L.cnt - 1];
        L.buf[L.base + L.j] = q; L.buf[L.base + L.cnt - 1] = p;   // the removed pixel stays in the buffer (final release)
        L.cnt--;
      } else {
        L.j++;
      }
    }
  }
  if (L.j == L.cnt) {
    if (L.cnt < 2) { L.hasseg = 0; to_fin(L); }
    else to_rect1(L);
  }
}
// end of a task: publish the list (a pending task must be releasable by whoever re-executes it), then DONE
LG_HD void step_fin2(const Params& P, const Frame& Fm, Lane& L) {
  const bool hasx = L.hasseg || L.ndep > 0;
  // layout: [cnt | hasseg << 31] [ndep] [cnt pixels] [ndep dependencies] [4 segment words if hasseg]
  if (L.off == 0) {
    if (ld_u(&Fm.st[L.task]) & ST_ABORT) { begin_rollback(L); return; }
    const int need = 2 + L.cnt + L.ndep + (L.hasseg ? 4 : 0);
    const int off = a_add(&Fm.ctl[C_POOL], need);
    if (off + need > P.pool_cap) { frame_error(Fm, ERR_POOL); L.phase = P_IDLE; return; }
    L.off = (unsigned)off;
    Fm.pool[off] = (unsigned)L.cnt | ((unsigned)L.hasseg << 31);
    Fm.pool[off + 1] = (unsigned)L.ndep;
  }
  unsigned* dst = Fm.pool + L.off + 2;
#pragma unroll 4
  for (int t = 0; t < 16; t++) {
    if (L.j < L.cnt) { dst[L.j] = L.buf[L.base + L.j]; L.j++; }
  }
  if (L.j == L.cnt) {
    unsigned* x = dst + L.cnt;
    for (int k = 0; k < L.ndep; k++) x[k] = L.buf[L.cap - 1 - k];
    if (L.hasseg) {
      x += L.ndep;
      float sg[4] = {(float)((L.x1 + 0.5) / 0.8), (float)((L.y1 + 0.5) / 0.8), (float)((L.x2 + 0.5) / 0.8), (float)((L.y2 + 0.5) / 0.8)};
#ifdef __CUDACC__
      x[0] = __float_as_uint(sg[0]); x[1] = __float_as_uint(sg[1]); x[2] = __float_as_uint(sg[2]); x[3] = __float_as_uint(sg[3]);
#else
      memcpy(x, sg, 16);
#endif
    }
    fence();
    const unsigned old = a_or(&Fm.st[L.task], (L.off << ST_OFF_SHIFT) | (hasx ? ST_HASX : 0u) | 4u);   // RUN (3) -> DONE (7)
    if (old & ST_ABORT) {               // aborted while publishing: nobody else will notice, queue it like a stealer would
      const int t = a_add(&Fm.ctl[C_RQT], 1);
      st_i(&Fm.ctl[C_WORDS + (t & (kRedoQ - 1))], L.task + 1);
    }
    L.phase = P_IDLE;
  }
}
// an aborted attempt gives back everything it holds, then the same task starts again
LG_HD void step_rollback(const Params& P, const Frame& Fm, Lane& L) {
  const int me = 2 * L.task;
#pragma unroll 2
  for (int t = 0; t < 8; t++) {
    if (L.j < L.hi) {
      const unsigned p = L.buf[L.j];
      int* o = &Fm.rec[(int)(p >> 16) * P.sw + (int)(p & 0xffffu)].x;
      if (a_cas(o, me, kFree) == me + 1) a_cas(o, me + 1, kFree);
      L.j++;
    }
  }
  if (L.j >= L.hi) {
    a_and(&Fm.st[L.task], ~ST_ABORT);
    a_add(&Fm.ctl[C_STAT0], 1);
    L.fresh = 0; L.phase = P_START;
  }
}
// the committer found the head task invalid: release its published list, then run it (nothing earlier is in flight)
LG_HD void step_redostart(const Params& P, const Frame& Fm, Lane& L) {
  const int me = 2 * L.task;
  if (L.off != 0) {
    const unsigned* src = Fm.pool + L.off + 2;
#pragma unroll 2
    for (int t = 0; t < 8; t++) {
      if (L.j < L.m) {
        const unsigned p = ld_u(&src[L.j]);
        a_cas(&Fm.rec[(int)(p >> 16) * P.sw + (int)(p & 0xffffu)].x, me, kFree);
        L.j++;
      }
    }
  }
  if (L.off == 0 || L.j >= L.m) {
    st_u(&Fm.st[L.task], ST_RUN);
    a_add(&Fm.ctl[C_STAT0 + 1], 1);
    L.fresh = 0; L.phase = P_START;
  }
}
// a lane takes over the re-execution of task i
LG_HD void lane_take_redo(const Frame& Fm, Lane& L, int i) {
  L.task = i;
  const unsigned w = ld_u(&Fm.st[i]);
  L.off = w >> ST_OFF_SHIFT;
  L.j = 0;
  L.m = L.off ? (int)(ld_u(&Fm.pool[L.off]) & 0x0fffffffu) : 0;
  L.phase = P_REDOSTART;
}
LG_HD void lane_take_seed(Lane& L, int i) { L.task = i; L.fresh = 1; L.phase = P_START; }

LG_HD void settle_pending(const Frame& Fm, Lane& L) {
  const int me = 2 * L.task;
  bool lost = false;
#pragma unroll
  for (int k = 0; k < 8; k++) if (!settle_claim(Fm, L.pend[k], me)) lost = true;
  L.npend = 0;
  if (lost && L.phase != P_ROLLBACK) begin_rollback(L);   // an earlier task took a pixel between my load and my claim
}
template <bool SPLIT>
LG_HD void lane_step(const Params& P, const Frame& Fm, Lane& L) {
  if (L.npend) { settle_pending(Fm, L); if (L.phase == P_ROLLBACK) return; }
  if (L.phase == P_GROW) { step_grow<SPLIT>(P, Fm, L); return; }
  switch (L.phase) {
    case P_START: step_start(P, Fm, L); break;
    case P_RECT1: step_rect1(P, Fm, L); break;
    case P_RECT2: step_rect2(P, Fm, L); break;
    case P_RECT3: step_rect3(P, Fm, L); break;
    case P_REFSTAT: step_refstat(P, Fm, L); break;
    case P_REDUCE: step_reduce(P, Fm, L); break;
    case P_FIN2: step_fin2(P, Fm, L); break;
    case P_ROLLBACK: step_rollback(P, Fm, L); break;
    case P_REDOSTART: step_redostart(P, Fm, L); break;
    default: break;
  }
}

// ------------------------------------------------------------------------------------------------ commit, one slot
// Validity of task i given that every task before it is final (second pass of the committer: the status words were
// seen DONE before the fence, so every steal / claim of an earlier task is visible now).
// at the frontier every earlier task is final: a pixel is USED for task i iff an earlier task HOLDS it (even mark)
LG_HD bool own_used_by_earlier(int o, int me) { return o >= 0 && o < me && !(o & 1); }
LG_HD bool task_valid(const Params& P, const Frame& Fm, int i, unsigned w) {
  const unsigned s = w & ST_STATE;
  if (s == ST_NOOP) return true;
  const int me = 2 * i;
  if (s == ST_EATEN) {
    const unsigned pix = ldg_u(&Fm.order[i]);
    return own_used_by_earlier(ld_i(&Fm.rec[(int)(pix >> 16) * P.sw + (int)(pix & 0xffffu)].x), me);
  }
  if (s != ST_DONE || (w & ST_ABORT)) return false;
  if (!(w & ST_HASX)) return true;
  const unsigned off = w >> ST_OFF_SHIFT;
  const int cnt = (int)(ld_u(&Fm.pool[off]) & 0x0fffffffu), ndep = (int)ld_u(&Fm.pool[off + 1]);
  const unsigned* x = Fm.pool + off + 2 + cnt;
  for (int k = 0; k < ndep; k++)
    if (!own_used_by_earlier(ld_i(&Fm.rec[(int)ld_u(&x[k])].x), me)) return false;
  return true;
}
LG_HD bool task_has_segment(const Frame& Fm, unsigned w, float4& seg) {
  if ((w & ST_STATE) != ST_DONE || !(w & ST_HASX)) return false;
  const unsigned off = w >> ST_OFF_SHIFT;
  const unsigned h = ld_u(&Fm.pool[off]);
  if (!(h >> 31)) return false;
  const unsigned* x = Fm.pool + off + 2 + (h & 0x0fffffffu) + ld_u(&Fm.pool[off + 1]);
  seg.x = as_float((int)ld_u(&x[0])); seg.y = as_float((int)ld_u(&x[1])); seg.z = as_float((int)ld_u(&x[2])); seg.w = as_float((int)ld_u(&x[3]));
  return true;
}

}  // namespace lg
