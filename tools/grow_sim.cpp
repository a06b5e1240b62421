// CPU protocol simulator for the speculative LSD region growing (test infrastructure, never shipped or timed).
//
// It runs the very lane state machine of pl-slam_b200/csrc/lsd_grow_core.cuh (compiled for the host) for W warps x 32
// lanes of ONE frame under a random, adversarial interleaving: lanes advance one micro-step at a time in random order,
// the neighbourhood snapshot of a growing step and its use are separate scheduling units (so every load-then-claim race
// of the GPU happens here too), warps scan / hand out / commit at random times.  The warp-level parts (seed scan, task
// hand-out, in-order commit) mirror k_lsd_grow in line.cu statement by statement, with loops over the 32 lanes in place of
// ballots.  tests/test_grow_protocol.py compares the segments with the oracle's sequential LSD, bit for bit.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -o tools/bin/libgrowsim.so tools/grow_sim.cpp
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <random>
#include <vector>
#include "../pl-slam_b200/csrc/lsd_grow_core.cuh"

using namespace lg;

namespace {

struct Warp {
  Lane lanes[32];
  int wq[64]; int whead = 0, wcount = 0, rsc = 0;
  bool exhausted = false;
};

struct Sim {
  Params P;
  Frame Fm;
  std::vector<int4> rec; std::vector<float2> seedcs; std::vector<int> sq; std::vector<unsigned> order, st, pool;
  std::vector<double> wtab;
  int ctl[kCtlStride];
  std::vector<float4> segs;
  std::vector<std::vector<unsigned>> lanebuf;
  std::vector<unsigned> rings;
  std::vector<Warp> warps;
  bool poll = true; long trace = 0; int boost = 0;
  long commits = 0, redos = 0, steps = 0, iters = 0, busy_hist[33] = {0}, eager = 0, polled = 0, winblock = 0, phase_hist[16] = {0}, redo_abort = 0, redo_dep = 0, redo_eaten = 0;

  void build(const uint8_t* scaled, int sw, int sh) {
    const double ANG_TH = 22.5, QUANT = 2.0;
    P.sw = sw; P.sh = sh; P.npx = sw * sh;
    P.prec = kPI * ANG_TH / 180; P.density_th = 0.7;
    {
      const double twopi = 2 * kPI;
      double c = twopi - P.prec;
      while ((twopi - std::nextafter(c, 0.0)) <= P.prec) c = std::nextafter(c, 0.0);
      while (!((twopi - c) <= P.prec)) c = std::nextafter(c, 10.0);
      P.prec_hi = c;
    }
    const double rho = QUANT / std::sin(P.prec);
    const double LOG_NT = 5 * (std::log10((double)sw) + std::log10((double)sh)) / 2 + std::log10(11.0);
    P.min_reg_size = (int)(size_t)(-LOG_NT / std::log10(ANG_TH / 180));
    const int npx = P.npx;
    rec.assign(npx, int4{kNotDef, 0, 0, 0}); seedcs.assign(npx, float2{0, 0}); sq.assign(npx, 0);
    wtab.resize(2 * 510 * 510 + 1);
    for (size_t s = 0; s < wtab.size(); s++) wtab[s] = std::sqrt((double)s / 4.0);
    struct NP { unsigned pix; int bin; };
    std::vector<NP> ord;
    double max_grad = -1;
    for (int y = 0; y < sh - 1; y++)
      for (int x = 0; x < sw - 1; x++) {
        const uint8_t* r = scaled + (size_t)y * sw; const uint8_t* n = r + sw;
        const int DA = n[x + 1] - r[x], BC = r[x + 1] - n[x], gx = DA + BC, gy = DA - BC, s = gx * gx + gy * gy;
        sq[y * sw + x] = s;
        const double norm = std::sqrt(s / 4.0);
        if (norm <= rho) continue;
        const float deg = fast_atan2_deg((float)gx, (float)(-gy));
        const double ad = (double)deg * kDegToRads;
        const float af = (float)ad;
        float c = cosf(af), sn = sinf(af);
        int4 v; v.x = kFree; memcpy(&v.y, &deg, 4); memcpy(&v.z, &c, 4); memcpy(&v.w, &sn, 4);
        rec[y * sw + x] = v;
        seedcs[y * sw + x] = float2{(float)std::cos(ad), (float)std::sin(ad)};
        if (norm > max_grad) max_grad = norm;
      }
    const double bin_coef = (max_grad > 0) ? 1023.0 / max_grad : 0;
    for (int y = 0; y < sh - 1; y++)
      for (int x = 0; x < sw - 1; x++)
        if (rec[y * sw + x].x == kFree) ord.push_back(NP{(unsigned)x | ((unsigned)y << 16), (int)(std::sqrt(sq[y * sw + x] / 4.0) * bin_coef)});
    std::stable_sort(ord.begin(), ord.end(), [](const NP& a, const NP& b) { return a.bin > b.bin; });
    order.resize(ord.size());
    for (size_t i = 0; i < ord.size(); i++) order[i] = ord[i].pix;
  }

  void setup(int nwarps, int lane_cap, int pool_mult, int window) {
    P.window = window; P.seg_cap = 1 << 20; P.lane_cap = lane_cap; P.pool_cap = pool_mult * P.npx;
    st.assign(order.size() + 64, 0);
    pool.assign(P.pool_cap, 0);
    memset(ctl, 0, sizeof(ctl));
    ctl[C_REDO] = -1; ctl[C_POOL] = 1;
    Fm.rec = rec.data(); Fm.seedcs = seedcs.data(); Fm.sq = sq.data(); Fm.order = order.data(); Fm.n = (int)order.size();
    Fm.st = st.data(); Fm.pool = pool.data(); Fm.ctl = ctl; Fm.wtab = wtab.data();
    warps.assign(nwarps, Warp());
    lanebuf.assign((size_t)nwarps * 32, std::vector<unsigned>(lane_cap));
    rings.assign((size_t)nwarps * 32 * kRing, 0);
    for (int w = 0; w < nwarps; w++)
      for (int l = 0; l < 32; l++) {
        Lane& L = warps[w].lanes[l];
        memset(&L, 0, sizeof(L));
        L.home = lanebuf[(size_t)w * 32 + l].data(); L.home_cap = lane_cap;
        L.ring = rings.data() + ((size_t)w * 32 + l) * kRing;
        L.phase = P_IDLE; L.buf = L.home; L.cap = L.home_cap;
        lane_reset(L);
      }
  }

  // ---- the committer: mirrors warp_commit() of line.cu
  void commit() {
    const int n = Fm.n;
    int F = ctl[C_FIN];
    if (F >= n) return;
    unsigned w1[32];
    int pre1 = 0;
    for (int l = 0; l < 32; l++) {
      const int i = F + l;
      w1[l] = (i < n) ? ld_u(&st[i]) : 0u;
    }
    for (; pre1 < 32; pre1++) {
      const unsigned s = w1[pre1] & ST_STATE;
      if (!(s == ST_NOOP || s == ST_EATEN || s == ST_DONE)) break;
    }
    if (pre1 == 0) return;
    fence();
    unsigned w2[32];
    int pre2 = 0;
    bool ok[32];
    for (int l = 0; l < pre1; l++) { w2[l] = ld_u(&st[F + l]); ok[l] = task_valid(P, Fm, F + l, w2[l]); }
    for (; pre2 < pre1 && ok[pre2]; pre2++) {}
    int ns = ctl[C_NS];
    for (int l = 0; l < pre2; l++) {
      float4 sg;
      if (task_has_segment(Fm, w2[l], sg)) {
        if (ns < P.seg_cap) { if ((int)segs.size() <= ns) segs.resize(ns + 1); segs[ns] = sg; } ns++; }
    }
    ctl[C_NS] = ns;
    commits += pre2;
    a_max(&ctl[C_FIN], F + pre2);
    if (pre2 < pre1) {                                    // the head is DONE but invalid: have it executed again
      const int h = F + pre2;
      // (an idle lane may be taking the same task off the redo ring right now: the compare-and-swap decides)
      const unsigned s2 = w2[pre2] & ST_STATE;
      if ((s2 == ST_DONE || s2 == ST_EATEN) &&
          (unsigned)a_cas(reinterpret_cast<int*>(&st[h]), (int)w2[pre2], (int)((w2[pre2] & ~(ST_STATE | ST_ABORT)) | ST_REDO)) == w2[pre2]) {
        if ((w2[pre2] & ST_STATE) == ST_EATEN) redo_eaten++; else if (w2[pre2] & ST_ABORT) redo_abort++; else redo_dep++;
        fence();
        st_i(&ctl[C_REDO], h);
        redos++;
      }
    }
  }

  // ---- seed scan + hand-out: mirrors warp_feed() of line.cu
  void feed(Warp& W) {
    int idle = 0;
    for (int l = 0; l < 32; l++) idle += (W.lanes[l].phase == P_IDLE);
    if (idle == 0) return;
    int redo = ld_i(&ctl[C_REDO]);
    if (redo >= 0) {
      redo = a_exch(&ctl[C_REDO], -1);
      if (redo >= 0)
        for (int l = 0; l < 32; l++)
          if (W.lanes[l].phase == P_IDLE) { lane_take_redo(Fm, W.lanes[l], redo); idle--; break; }
    }
    // aborted tasks that are already published: re-execute them now rather than when they reach the head
    for (int tries = 0; tries < 2 && idle > 0; tries++) {
      if (ld_i(&ctl[C_RQH]) >= ld_i(&ctl[C_RQT])) break;
      const int hq = a_add(&ctl[C_RQH], 1);
      const int m = a_exch(&ctl[C_WORDS + (hq & (kRedoQ - 1))], 0) - 1;
      if (m < 0) continue;
      const unsigned w = ld_u(&st[m]);
      if ((w & ST_STATE) != ST_DONE || !(w & ST_ABORT)) continue;
      if ((unsigned)a_cas(reinterpret_cast<int*>(&st[m]), (int)w, (int)((w & ~(ST_STATE | ST_ABORT)) | ST_REDO)) != w) continue;
      for (int l = 0; l < 32; l++)
        if (W.lanes[l].phase == P_IDLE) { lane_take_redo(Fm, W.lanes[l], m); idle--; eager++; break; }
    }
    int scans = 0;
    while (idle > W.wcount && !W.exhausted && W.wcount <= 32 && scans < 2) {
      if (P.window > 0 && ld_i(&ctl[C_NXT]) - ld_i(&ctl[C_FIN]) >= P.window) { winblock++; break; }
      scans++;
      const int base = a_add(&ctl[C_NXT], 32);
      if (base >= Fm.n) { W.exhausted = true; break; }
      const int F = ld_i(&ctl[C_FIN]);
      for (int l = 0; l < 32; l++) {
        const int i = base + l;
        if (i >= Fm.n) break;
        const unsigned pix = order[i];
        const int o = ld_i(&rec[(int)(pix >> 16) * P.sw + (int)(pix & 0xffffu)].x);
        if (!own_candidate(o, 2 * i, F)) st_u(&st[i], (!(o & 1) && (o >> 1) < F) ? ST_NOOP : ST_EATEN);
        else { W.wq[(W.whead + W.wcount) & 63] = i; W.wcount++; }
      }
    }
    // lanes still idle: look again at seeds that were found consumed by a task that was not final (it may have let go)
    if (idle > W.wcount && W.wcount <= 32 && poll) {
      const int F = ld_i(&ctl[C_FIN]), hi = std::min(ld_i(&ctl[C_NXT]), Fm.n);
      if (hi > F) {
        if (W.rsc < F || W.rsc >= hi) W.rsc = F;
        const int base = W.rsc;
        W.rsc += 32;
        for (int l = 0; l < 32; l++) {
          const int i = base + l;
          if (i >= hi) break;
          const unsigned w = ld_u(&st[i]);
          if ((w & ST_STATE) != ST_EATEN) continue;
          const unsigned pix = order[i];
          const int o = ld_i(&rec[(int)(pix >> 16) * P.sw + (int)(pix & 0xffffu)].x);
          if (own_candidate(o, 2 * i, F)) {
            if ((unsigned)a_cas(reinterpret_cast<int*>(&st[i]), (int)w, (int)ST_RUN) == w) { W.wq[(W.whead + W.wcount) & 63] = i; W.wcount++; polled++; }
          } else if (!(o & 1) && (o >> 1) < F) st_u(&st[i], ST_NOOP);
        }
      }
    }
    for (int l = 0; l < 32 && W.wcount > 0; l++)
      if (W.lanes[l].phase == P_IDLE) { lane_take_seed(W.lanes[l], W.wq[W.whead & 63]); W.whead++; W.wcount--; }
  }

  int run(unsigned seed, int mode) {
    std::mt19937 rng(seed);
    if (mode == 3) {                                       // strictly serial: one lane of one warp does everything
      for (auto& W : warps) for (int l = 0; l < 32; l++) if (&W != &warps[0] || l > 0) W.lanes[l].phase = -1;
      mode = 0;
    }
    const int nw = (int)warps.size();
    long guard = 0;
    while (ld_i(&ctl[C_FIN]) < Fm.n) {
      if (++guard > 400000000L) return -2;
      Warp& W = warps[rng() % nw];
      if (mode == 0 || (rng() & 3) == 0) {                 // mode 0: lock step like the GPU; else: commits are rare and late
        if (ctl[C_LOCK] == 0) commit();
      }
      feed(W);
      iters++;
      if (trace && iters % (trace * nw) == 0) { int b = 0, g = 0; for (auto& X : warps) for (int l = 0; l < 32; l++) { b += X.lanes[l].phase > P_IDLE; g += X.lanes[l].phase == P_GROW; } int hp = -1, hc = 0, hs = 0; for (auto& X : warps) for (int l = 0; l < 32; l++) if (X.lanes[l].phase > P_IDLE && X.lanes[l].task == ctl[C_FIN]) { hp = X.lanes[l].phase; hc = X.lanes[l].cnt; hs = X.lanes[l].stage; } fprintf(stderr, "t=%ld F=%d NXT=%d busy=%d grow=%d headst=%x head: phase=%d cnt=%d stage=%d\n", iters / nw, ctl[C_FIN], ctl[C_NXT], b, g, st[ctl[C_FIN]], hp, hc, hs); }
      { int b = 0; for (int l = 0; l < 32; l++) { b += (W.lanes[l].phase > P_IDLE); if (W.lanes[l].phase >= 0) phase_hist[W.lanes[l].phase]++; } busy_hist[b]++; }
      int perm[32];
      for (int l = 0; l < 32; l++) perm[l] = l;
      if (mode != 0) std::shuffle(perm, perm + 32, rng);
      if (boost > 0) {                                      // emulation of the cooperative step: the deepest queue gets extra expansions
        int bl = -1, bq = 7;
        for (int l = 0; l < 32; l++) { Lane& L = W.lanes[l]; if (L.phase == P_GROW && L.cnt - L.qi > bq) { bq = L.cnt - L.qi; bl = l; } }
        if (bl >= 0) for (int r = 0; r < boost && W.lanes[bl].phase == P_GROW; r++) lane_step<false>(P, Fm, W.lanes[bl]);
      }
      for (int k = 0; k < 32; k++) {
        Lane& L = W.lanes[perm[k]];
        if (L.phase == P_IDLE) continue;
        if (mode == 2 && (rng() & 1)) continue;            // this lane stalls
        if (mode == 0) { lane_step<false>(P, Fm, L); }
        else lane_step<true>(P, Fm, L);
        steps++;
      }
    }
    return ctl[C_ERR] ? -1 : 0;
  }
};

}  // namespace

// scaled: the 0.8x image LSD works on (oracle_lsd_stages).  Returns the number of segments (or < 0), segments in out.
// stats[0..5]: tasks committed, head re-executions, lane micro-steps, self aborts, pool words used, seeds
extern "C" int grow_sim(const uint8_t* scaled, int sw, int sh, int nwarps, unsigned seed, int mode, int lane_cap, float* out, int cap,
                        long* stats, int window) {
  Sim S;
  S.build(scaled, sw, sh);
  S.setup(nwarps, lane_cap > 0 ? lane_cap : 2048, 4, window);
  S.poll = !getenv("GROWSIM_NOPOLL");
  if (getenv("GROWSIM_BOOST")) S.boost = atoi(getenv("GROWSIM_BOOST"));
  if (getenv("GROWSIM_TRACE")) S.trace = atol(getenv("GROWSIM_TRACE"));
  const int rc = S.run(seed, mode);
  if (stats) {
    stats[0] = S.commits; stats[1] = S.redos; stats[2] = S.steps; stats[3] = S.ctl[C_STAT0]; stats[4] = S.ctl[C_POOL]; stats[5] = S.Fm.n; stats[6] = S.iters; stats[7] = S.redo_abort; stats[8] = S.redo_dep; stats[9] = S.redo_eaten;
    long bs = 0; for (int b = 0; b <= 32; b++) bs += b * S.busy_hist[b]; stats[10] = bs; stats[11] = S.ctl[C_STAT0 + 2]; stats[12] = S.ctl[C_STAT0 + 3]; stats[13] = S.ctl[C_STAT0 + 4]; stats[14] = S.eager; stats[15] = S.winblock; stats[12] = S.polled;
    if (getenv("GROWSIM_VERBOSE")) { fprintf(stderr, "busy:"); for (int b = 0; b <= 32; b++) fprintf(stderr, " %ld", S.busy_hist[b]); fprintf(stderr, "\nphase:"); for (int b = 0; b < 12; b++) fprintf(stderr, " %ld", S.phase_hist[b]); fprintf(stderr, "\n"); }
  }
  if (rc) return rc;
  const int n = S.ctl[C_NS];
  for (int i = 0; i < n && i < cap; i++) memcpy(out + 4 * i, &S.segs[i], 16);
  return n;
}
