// tests/emul/okb_emul.cpp -- TEST INFRASTRUCTURE. Serial host execution of the *parallel formulation* that the CUDA
// kernels implement (same per-element functions from okvis2_b200/csrc/okb_core.h, same phases, same touch-time map),
// so that the formulation can be checked against the oracle on a machine without a GPU. Not part of the product.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../okvis2_b200/csrc/okb_core.h"
#include "../../okvis2_b200/csrc/okb_tables.h"

using namespace okb;

struct Kp { float x, y, size, angle, response; int32_t octave, class_id; };

struct HostLayer { int w, h, pitch; float scale, offset; std::vector<uint8_t> img, score; std::vector<uint32_t> touch; };

static HostTables* g_tables = nullptr;
static int closed_form_mismatches = 0;

static void resize_layer(const HostLayer& s, HostLayer& d)
{
  const double sx = 1. / ((double)d.w / s.w), sy = 1. / ((double)d.h / s.h);
  const bool fast = sx == 2.0 && sy == 2.0;
  if (fast) {
    for (int y = 0; y < d.h; y++) for (int x = 0; x < d.w; x++) d.img[(size_t)y * d.pitch + x] = half_pixel(s.img.data(), s.pitch, x, y);
    return;
  }
  AreaAxis ax, ay; build_area_axis(s.w, d.w, ax); build_area_axis(s.h, d.h, ay);
  for (int y = 0; y < d.h; y++)
    for (int x = 0; x < d.w; x++)
      d.img[(size_t)y * d.pitch + x] = area_pixel(s.img.data(), s.pitch, ax.start[x], ax.count[x], &ax.alpha[(size_t)x * 4],
                                                   ay.start[y], ay.count[y], &ay.alpha[(size_t)y * 4]);
}

struct Cand { uint32_t key; int layer, x, y; bool tie; RefineResult r; int state; };

extern "C" int okb_emul_detect_describe(const uint8_t* img, int W, int H, int threshold, int octaves, int max_kp,
                                        Kp* kp_out, uint8_t* desc_out, int cap, int* stats /*[5]: cands, ties, rounds, raw, skipped emissions*/)
{
  if (!g_tables) { g_tables = new HostTables(); if (!build_host_tables(1.0f, *g_tables)) return -1; }
  const HostTables& T = *g_tables;
  const int n_layers = octaves == 0 ? 1 : 2 * octaves;
  std::vector<HostLayer> HL(n_layers);
  auto init = [&](HostLayer& l, int w, int h, float scale) {
    l.w = w; l.h = h; l.pitch = (w + 15) / 16 * 16; l.scale = scale; l.offset = 0.5f * scale - 0.5f;
    l.img.assign((size_t)l.pitch * h, 0); l.score.assign((size_t)l.pitch * h, 0); l.touch.assign((size_t)l.pitch * h, 0);
  };
  init(HL[0], W, H, 1.0f); HL[0].offset = 0.f;
  for (int y = 0; y < H; y++) memcpy(&HL[0].img[(size_t)y * HL[0].pitch], img + (size_t)y * W, W);
  if (n_layers > 1) { init(HL[1], 2 * (W / 3), 2 * (H / 3), 1.5f); resize_layer(HL[0], HL[1]); }
  for (int i = 2; i < n_layers; i += 2) {
    init(HL[i], HL[i - 2].w / 2, HL[i - 2].h / 2, HL[i - 2].scale * 2); resize_layer(HL[i - 2], HL[i]);
    init(HL[i + 1], HL[i - 1].w / 2, HL[i - 1].h / 2, HL[i - 1].scale * 2); resize_layer(HL[i - 1], HL[i + 1]);
  }
  LayerView L[kMaxLayers];
  for (int i = 0; i < n_layers; i++)
    L[i] = LayerView{HL[i].img.data(), HL[i].score.data(), HL[i].w, HL[i].h, HL[i].pitch, HL[i].pitch, HL[i].scale, HL[i].offset};
  // phase: dense score maps b0 (k_score)
  for (int i = 0; i < n_layers; i++)
    for (int y = 0; y < HL[i].h; y++) for (int x = 0; x < HL[i].w; x++)
      HL[i].score[(size_t)y * HL[i].pitch + x] = (uint8_t)b0_compute(L[i], x, y);
  // phase: candidates (any order on the device; the order here is irrelevant by construction)
  std::vector<Cand> C;
  for (int i = n_layers - 1; i >= 0; i--) {  // deliberately reversed to prove order independence
    const HostLayer& l = HL[i];
    for (int y = l.h - 4; y >= 3; y--) for (int x = 3; x < l.w - 3; x++) {
      const uint8_t* s = &l.score[(size_t)y * l.pitch + x];
      const int c = s[0];
      if (c < threshold) continue;
      bool ok = true, tie = false;
      for (int dy = -1; dy <= 1 && ok; dy++) for (int dx = -1; dx <= 1; dx++) {
        if (!dx && !dy) continue;
        const int v = s[dy * l.pitch + dx];  // values below the threshold can neither exceed nor tie c
        if (v > c) { ok = false; break; }
        if (v == c) tie = true;
      }
      if (!ok) continue;
      Cand cd; cd.key = time_key(i, x, y); cd.layer = i; cd.x = x; cd.y = y; cd.tie = tie; cd.state = tie ? 0 : 1;
      C.push_back(cd);
    }
  }
  // phase (as the device does it): the same candidates from 64x64 score tiles. A strong pixel is tested against its
  // in-tile neighbours only; one on the tile border that they do not beat is "pending" and finished from the complete
  // map (k_refine). The result must be the candidate set above, tie flags included.
  {
    constexpr int TW = 64, TH = 64;
    std::vector<std::pair<uint32_t, bool>> tiled;
    for (int i = 0; i < n_layers; i++) {
      const HostLayer& l = HL[i];
      for (int y0 = 0; y0 < l.h; y0 += TH) for (int x0 = 0; x0 < l.w; x0 += TW)
        for (int y = y0; y < std::min(y0 + TH, l.h); y++) for (int x = x0; x < std::min(x0 + TW, l.w); x++) {
          const int c = l.score[(size_t)y * l.pitch + x];
          if (c < threshold) continue;
          bool is_c = true, tie = false, pending = false;
          for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
            if (!dx && !dy) continue;
            const int xx = x + dx, yy = y + dy;
            if (xx >= x0 && xx < x0 + TW && yy >= y0 && yy < y0 + TH) {
              const int v = l.score[(size_t)yy * l.pitch + xx];
              if (v > c) is_c = false;
              if (v == c) tie = true;
            } else pending = true;
          }
          if (!is_c) continue;
          if (pending) {   // completion from the global map
            tie = false;
            for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
              if (!dx && !dy) continue;
              const int v = l.score[(size_t)(y + dy) * l.pitch + x + dx];
              if (v > c) is_c = false;
              if (v == c) tie = true;
            }
            if (!is_c) continue;
          }
          tiled.push_back({time_key(i, x, y), tie});
        }
    }
    std::vector<std::pair<uint32_t, bool>> dense;
    for (const Cand& c : C) dense.push_back({c.key, c.tie});
    std::sort(tiled.begin(), tiled.end()); std::sort(dense.begin(), dense.end());
    if (tiled != dense) return -2000;
  }
  // tie-cell bitmap: cells around tied candidates and around candidates on a tile border (pending when the score kernel
  // flags them); non-tied maxima whose touch footprint misses every flagged cell do not emit
  std::vector<std::vector<uint32_t>> cells(n_layers, std::vector<uint32_t>(kCellWordsPerLayer, 0u));
  for (const Cand& c : C) {
    const bool on_tile_border = c.x % 64 == 0 || c.x % 64 == 63 || c.y % 64 == 0 || c.y % 64 == 63;
    if (c.tie || on_tile_border)
      for_each_cell(HL[c.layer].w, c.x - 2, c.x + 2, c.y - 2, c.y + 2, [&](int b) { cells[c.layer][b >> 5] |= 1u << (b & 31); return false; });
  }
  auto any_cell = [&](int layer, const TouchBox& t) {
    return for_each_cell(HL[layer].w, t.x_lo, t.x_hi, t.y_lo, t.y_hi, [&](int b) { return ((cells[layer][b >> 5] >> (b & 31)) & 1u) != 0; });
  };
  int emissions_skipped = 0;
  const uint32_t epoch = 1;
  auto touch = [&](int layer, int x, int y, uint32_t time) {
    HostLayer& l = HL[layer];
    if (x < 0 || y < 0 || x >= l.w || y >= l.h) return;
    uint32_t& e = l.touch[(size_t)y * l.pitch + x];
    e = std::max(e, touch_entry(epoch, time));
  };
  auto emit = [&](const Cand& c) {
    if (c.r.own_touch == 1) { for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) touch(c.layer, c.x + dx, c.y + dy, c.key); }
    else if (c.r.own_touch == 2) { for (int dy = -1; dy <= 2; dy++) for (int dx = -1; dx <= 2; dx++) touch(c.layer, c.x + dx, c.y + dy, c.key); }
    if (c.r.has_above) {
      for_each_above_touch(c.layer, c.x, c.y, c.r.above, [&](int x, int y) { touch(c.layer + 1, x, y, c.key); });
      // the device emits the same set through the closed form (one lane per query): check the two enumerations agree
      std::vector<std::pair<int, int>> a, b;
      for_each_above_touch(c.layer, c.x, c.y, c.r.above, [&](int x, int y) { a.push_back({x, y}); });
      ScanIter it; above_window(c.layer, c.x, c.y, it);
      for (int q = 0; q < c.r.above.n_queries; q++) {
        int X, Y; bool blk; above_query_pos(it, q, X, Y, blk);
        b.push_back({X, Y});
        if (blk) { b.push_back({X + 1, Y}); b.push_back({X, Y + 1}); b.push_back({X + 1, Y + 1}); }
      }
      if (!c.r.above.exited) for (int j = 0; j < 9; j++) b.push_back({c.r.above.max_x + j % 3 - 1, c.r.above.max_y + j / 3 - 1});
      std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end());
      std::sort(b.begin(), b.end()); b.erase(std::unique(b.begin(), b.end()), b.end());
      if (a != b) closed_form_mismatches++;
    }
  };
  // phase: refine (pure) + touches of the non-tie maxima
  for (auto& c : C) {
    refine_candidate(L, n_layers, c.layer, c.x, c.y, threshold, c.r);
    if (c.tie || !(c.r.own_touch || c.r.has_above)) continue;
    bool hit = c.r.own_touch && any_cell(c.layer, own_touch_box(c.x, c.y));
    if (!hit && c.r.has_above) hit = any_cell(c.layer + 1, above_touch_box(c.layer, c.x, c.y));
    if (hit) emit(c); else emissions_skipped++;
  }
  // phase: resolve ties in dependency rounds
  std::vector<int> ties;
  for (size_t i = 0; i < C.size(); i++) if (C[i].tie) ties.push_back((int)i);
  auto near_ = [&](const Cand& u, const Cand& t) {  // can events of u reach the 5x5 window of t ?
    if (u.layer == t.layer) return abs(u.x - t.x) <= 4 && abs(u.y - t.y) <= 4;
    if (u.layer + 1 == t.layer) {
      ScanIter it; above_window(u.layer, u.x, u.y, it);
      const int xa = (int)it.x_1 - 1, xb = (int)it.x1 + 2, ya = (int)it.y_1 - 1, yb = (int)it.y1 + 2;
      return !(t.x + 2 < xa || t.x - 2 > xb || t.y + 2 < ya || t.y - 2 > yb);
    }
    return false;
  };
  int rounds = 0, unresolved = (int)ties.size();
  while (unresolved > 0) {
    rounds++;
    std::vector<int> newly;
    for (int ti : ties) {
      Cand& t = C[ti];
      if (t.state != 0) continue;
      bool blocked = false;
      for (int ui : ties) { const Cand& u = C[ui]; if (u.state == 0 && u.key < t.key && near_(u, t)) { blocked = true; break; } }
      if (blocked) continue;
      const HostLayer& l = HL[t.layer];
      int m[5][5];
      for (int dy = -2; dy <= 2; dy++) for (int dx = -2; dx <= 2; dx++) {
        const int x = t.x + dx, y = t.y + dy;
        int v = l.score[(size_t)y * l.pitch + x];  // b0: what the cache holds once touched
        if (v < threshold && !touched_before(l.touch[(size_t)y * l.pitch + x], epoch, t.key)) v = 0;
        m[dy + 2][dx + 2] = v;
      }
      newly.push_back(is_max_2d_5x5(m) ? ti : -ti - 1);
    }
    for (int v : newly) {
      if (v >= 0) { C[v].state = 1; emit(C[v]); } else C[-v - 1].state = 2;
      unresolved--;
    }
  }
  // phase: finalize (order by key, cap, border removal)
  std::vector<int> order;
  for (size_t i = 0; i < C.size(); i++) if (C[i].state == 1 && C[i].r.keep) order.push_back((int)i);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return C[a].key < C[b].key; });
  const int raw = (int)order.size();
  std::vector<char> keep(order.size(), 1);
  if (max_kp > 0 && (int)order.size() > max_kp) {
    std::vector<int> pos(order.size());
    for (size_t i = 0; i < pos.size(); i++) pos[i] = (int)i;
    std::sort(pos.begin(), pos.end(), [&](int a, int b) {
      const float ra = C[order[a]].r.response, rb = C[order[b]].r.response;
      return ra > rb || (ra == rb && a < b);
    });
    std::fill(keep.begin(), keep.end(), 0);
    for (int i = 0; i < max_kp; i++) keep[pos[i]] = 1;
  }
  std::vector<Kp> kps; std::vector<int> kscale;
  for (size_t i = 0; i < order.size(); i++) {
    if (!keep[i]) continue;
    const Cand& c = C[order[i]];
    const int sc = kscale_from_bounds(T.scale_bounds.data(), c.r.size);
    const int border = (int)T.size_list[sc];
    if (c.r.x < (float)border || c.r.x >= (float)(W - border) || c.r.y < (float)border || c.r.y >= (float)(H - border)) continue;
    kps.push_back(Kp{c.r.x, c.r.y, c.r.size, -1.f, c.r.response, c.layer, -1});
    kscale.push_back(sc);
  }
  // phase: integral + describe
  const int ipitch = W + 1;
  std::vector<int32_t> integral((size_t)(W + 1) * (H + 1), 0);
  for (int y = 0; y < H; y++) {
    int32_t rs = 0;
    for (int x = 0; x < W; x++) { rs += HL[0].img[(size_t)y * HL[0].pitch + x]; integral[(size_t)(y + 1) * ipitch + x + 1] = integral[(size_t)y * ipitch + x + 1] + rs; }
  }
  const int n = std::min((int)kps.size(), cap);
  for (int k = 0; k < n; k++) {
    Kp& p = kps[k];
    int values[kPoints];
    const PatternPoint* pat0 = &T.pattern[((size_t)kscale[k] * kRot + 0) * kPoints];
    for (int i = 0; i < kPoints; i++) values[i] = smoothed_intensity(HL[0].img.data(), HL[0].pitch, integral.data(), ipitch, p.x, p.y, pat0[i]);
    int d0 = 0, d1 = 0;
    for (const LongPair& lp : T.long_pairs) { const int dt = values[lp.i] - values[lp.j]; d0 += dt * lp.wdx / 1024; d1 += dt * lp.wdy / 1024; }
    p.angle = (float)(atan2((double)(float)d1, (double)(float)d0) / M_PI * 180.0);
    int theta = (int)(kRot * ((double)p.angle / 360.0) + 0.5);
    if (theta < 0) theta += kRot;
    if (theta >= kRot) theta -= kRot;
    if (p.angle < 0) p.angle += 360.f;
    const PatternPoint* pat = &T.pattern[((size_t)kscale[k] * kRot + theta) * kPoints];
    for (int i = 0; i < kPoints; i++) values[i] = smoothed_intensity(HL[0].img.data(), HL[0].pitch, integral.data(), ipitch, p.x, p.y, pat[i]);
    uint32_t* out = (uint32_t*)(desc_out + (size_t)k * 64);
    for (int w = 0; w < 16; w++) {
      uint32_t word = 0;
      for (int b = 0; b < 32; b++) { const uint32_t pr = T.short_pairs[w * 32 + b]; if (values[pr & 255] > values[pr >> 8]) word |= 1u << b; }
      out[w] = word;
    }
    kp_out[k] = p;
  }
  if (closed_form_mismatches) return -1000 - closed_form_mismatches;
  if (stats) { stats[0] = (int)C.size(); stats[1] = (int)ties.size(); stats[2] = rounds; stats[3] = raw; stats[4] = emissions_skipped; }
  return (int)kps.size();
}

// ---- gate constants: okb::gate_cos (csrc/okb_gatecos.h, the function the matchers use on host and device) against the
//      libm of this machine. Returns the number of arguments on which they differ.
#include <math.h>
#include "../../okvis2_b200/csrc/okb_gatecos.h"
extern "C" double okb_emul_gate_cos(double x) { return okb::gate_cos(x); }
extern "C" long okb_emul_gate_cos_mismatches(long n, unsigned long long seed, double range)
{
  unsigned long long s = seed ? seed : 88172645463325252ull; long bad = 0;
  for (long i = 0; i < n; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    double x = (double)(s >> 11) / 9007199254740992.0 * range;
    if (i % 5 == 0) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x *= (double)(s >> 11) / 9007199254740992.0; }   // more small arguments
    if (i % 7 == 0) x = -x;
    bad += okb::gate_cos(x) != cos(x);
  }
  return bad;
}
// every keypoint size (float) of a binade range at focal length f: sigma = size / f * 0.125 -> cos(2.6 sigma), cos(6 sigma)
extern "C" long okb_emul_gate_cos_sizes(double f, unsigned first_bits, unsigned last_bits, unsigned stride)
{
  long bad = 0;
  for (unsigned b = first_bits; b < last_bits; b += stride) {
    float size; memcpy(&size, &b, 4);
    const double sg = ((double)size / f) * 0.125;
    bad += okb::gate_cos(2.6 * sg) != cos(2.6 * sg);
    bad += okb::gate_cos(6.0 * sg) != cos(6.0 * sg);
  }
  return bad;
}

// ---------------------------------------------------------------------------------------------------------------
// D = 48 mode (okvis2_b200/csrc/okb_harris.cu): the same per-element functions (okb_harris_core.h) and the same parallel
// formulation -- run-parity maxima, occupancy as a per-candidate sum over higher-ranked accepted candidates, decided in push waves --
// executed serially. Returns the number of keypoints; stats[0] = maxima, [1] = rounds, [2] = accepted before the border test.
#include "../../okvis2_b200/csrc/okb_harris_core.h"

extern "C" int okb_emul_harris_brisk2(const uint8_t* img, int W, int H, float radius, int threshold, int max_kp, const float* rays,
                                      const float* jac, float fu, const float* dir, Kp* kp_out, uint8_t* desc_out, int cap, int* stats)
{
  if (!g_tables) { g_tables = new HostTables(); if (!build_host_tables(1.0f, *g_tables)) return -1; }
  const HostTables& T = *g_tables;
  std::vector<int32_t> score((size_t)W * H, 0);
  {
    std::vector<int8_t> gx((size_t)W * H, 0), gy((size_t)W * H, 0);
    for (int y = 1; y <= H - 2; y++)
      for (int x = 1; x <= W - 2; x++) {
        int a, b;
        harris_grad(img + (size_t)(y - 1) * W, img + (size_t)y * W, img + (size_t)(y + 1) * W, x, a, b);
        gx[(size_t)y * W + x] = (int8_t)a; gy[(size_t)y * W + x] = (int8_t)b;
      }
    for (int y = 2; y < H - 2; y++)
      for (int x = 2; x < W - 2; x++) {
        int a = 0, b = 0, c = 0;
        for (int j = 0; j < 3; j++)
          for (int k = 0; k < 3; k++) {
            const int w = (j == 1 ? 2 : 1) * (k == 1 ? 2 : 1);
            const int u = gx[(size_t)(y + j - 1) * W + x + k - 1], v = gy[(size_t)(y + j - 1) * W + x + k - 1];
            a += w * u * u; b += w * v * v; c += w * u * v;
          }
        score[(size_t)y * W + x] = harris_score(a, b, c);
      }
  }
  struct C { uint64_t key; float nsc; int state; };
  std::vector<C> cs;
  for (int y = 2; y < H - 2; y++)
    for (int x = W - 3; x >= 2; x--)   // any visiting order: the test is a pure function of the map
      if (harris_is_maximum(score.data(), W, x, y, threshold))
        cs.push_back(C{((uint64_t)(~(uint32_t)score[(size_t)y * W + x]) << 32) | ((uint32_t)x | ((uint32_t)y << 16)), 0.f, 0});
  std::sort(cs.begin(), cs.end(), [](const C& a, const C& b) { return a.key < b.key; });
  const int n = (int)cs.size();
  stats[0] = n; stats[1] = 0; stats[2] = 0;
  if (n == 0) return 0;
  float lut[kUniLut * kUniLut];
  for (int j = 0; j < kUniLut; j++) for (int i = 0; i < kUniLut; i++) lut[j * kUniLut + i] = uni_lut_host(radius, i - kUniWin, j - kUniWin);
  const float max_score = (float)(int)(~(uint32_t)(cs[0].key >> 32));
  auto sc_of = [&](int i) { return (int)(~(uint32_t)(cs[i].key >> 32)); };
  auto hx_of = [&](int i) { return (int)((uint32_t)cs[i].key & 0xffffu) >> 1; };
  auto hy_of = [&](int i) { return (int)((uint32_t)cs[i].key >> 16) >> 1; };
  for (int i = 0; i < n; i++) cs[i].nsc = uni_nsc(uni_ratio(sc_of(i), max_score));
  // the kernel's formulation (k_uniformity): every candidate counts the higher-ranked candidates whose stamp reaches its cell
  // (word = pending << 20 | stamp sum); the candidates without pending neighbours form the first wave; a wave decides its candidates
  // from their (complete) stamp sums and pushes the decisions -- one add per lower-ranked candidate in reach: the stamp, and one
  // off the pending count -- and whoever drops to zero joins the next wave. After every wave: ranks below the first undecided one
  // are final; stop once they hold max_kp accepted candidates.
  auto reaches = [&](int i, int j, float& l) {   // offset of i seen from j
    const int dx = hx_of(i) - hx_of(j), dy = hy_of(i) - hy_of(j);
    if (dx < -kUniWin || dx > kUniWin || dy < -kUniWin || dy > kUniWin) return false;
    l = lut[(dy + kUniWin) * kUniLut + dx + kUniWin];
    return l != 0.0f;
  };
  std::vector<uint32_t> word(n, 0);
  std::vector<int> queue; queue.reserve(n);
  for (int i = 0; i < n; i++) {
    uint32_t pend = 0; float l;
    for (int j = 0; j < i; j++) if (reaches(i, j, l)) pend++;
    word[i] = pend << 20;
    if (pend == 0) queue.push_back(i);
  }
  int lo = 0, acc = 0;
  size_t head = 0;
  while (head < queue.size()) {
    stats[1]++;
    const size_t tail = queue.size();
    for (size_t q = head; q < tail; q++) {
      const int j = queue[q];
      const bool accepted = !uni_rejected(uni_ratio(sc_of(j), max_score), (int)(word[j] & 0xfffffu));
      cs[j].state = accepted ? 1 : 2;
      for (int i = j + 1; i < n; i++) {
        float l;
        if (!reaches(i, j, l)) continue;
        const uint32_t old = word[i];
        word[i] = old + (accepted ? (uint32_t)uni_stamp(cs[j].nsc, l) : 0u) - (1u << 20);
        if ((old >> 20) == 1u) queue.push_back(i);
      }
    }
    head = tail;
    if (max_kp > 0) {
      int first = lo;
      while (first < n && cs[first].state != 0) first++;
      for (int i = lo; i < first; i++) acc += cs[i].state == 1;
      lo = first;
      if (acc >= max_kp) break;
    }
  }
  if (max_kp <= 0 || acc < max_kp) lo = n;
  const int basic = brisk2_basic_scale_host();
  const PatternPoint* pat0 = &T.pattern[(size_t)basic * kRot * kPoints];
  const int border = (int)T.size_list[basic];
  std::vector<uint32_t> sp;
  {
    const float d_max = (float)(kDmax48 * 1.0), d_min = (float)(8.2 * 1.0);
    for (unsigned i = 1; i < (unsigned)kPoints; i++)
      for (unsigned j = 0; j < i; j++) {
        const float dx = T.pattern[j].x - T.pattern[i].x, dy = T.pattern[j].y - T.pattern[i].y, n2 = dx * dx + dy * dy;
        if (n2 > d_min * d_min) continue;
        if (n2 < d_max * d_max) sp.push_back(i | (j << 8));
      }
    if ((int)sp.size() != kShortPairs48) return -2;
  }
  std::vector<int32_t> integral((size_t)(W + 1) * (H + 1), 0);
  for (int y = 0; y < H; y++) {
    int rs = 0;
    for (int x = 0; x < W; x++) { rs += img[(size_t)y * W + x]; integral[(size_t)(y + 1) * (W + 1) + x + 1] = integral[(size_t)y * (W + 1) + x + 1] + rs; }
  }
  int m = 0, kept = 0;
  for (int i = 0; i < lo; i++) {
    if (cs[i].state != 1) continue;
    if (max_kp > 0 && kept >= max_kp) break;
    kept++;
    const int x = (int)((uint32_t)cs[i].key & 0xffffu), y = (int)((uint32_t)cs[i].key >> 16);
    float dx, dy;
    harris_subpixel(score.data(), W, x, y, dx, dy);
    const float fx = (float)x + dx, fy = (float)y + dy;
    int val[kPoints];
    float angle;
    if (!rays) {
      if ((fx < (float)border) || (fx >= (float)(W - border)) || (fy < (float)border) || (fy >= (float)(H - border))) continue;
      for (int p = 0; p < kPoints; p++) val[p] = smoothed_intensity(img, W, integral.data(), W + 1, fx, fy, pat0[p]);
      int e0 = 0, e1 = 0;
      for (const LongPair& lp : T.long_pairs) { const int dt = val[lp.i] - val[lp.j]; e0 += dt * lp.wdx / 1024; e1 += dt * lp.wdy / 1024; }
      angle = (float)(atan2((double)(float)e1, (double)(float)e0) / 3.14159265358979323846 * 180.0);
      int theta = (int)((double)kRot * ((double)angle / 360.0) + 0.5);
      if (theta < 0) theta += kRot;
      if (theta >= kRot) theta -= kRot;
      if (angle < 0) angle += 360.f;
      for (int p = 0; p < kPoints; p++) val[p] = smoothed_intensity(img, W, integral.data(), W + 1, fx, fy, pat0[(size_t)theta * kPoints + p]);
    } else {
      int u = (int)(fx + 0.5f), v = (int)(fy + 0.5f);
      u = u < 0 ? 0 : (u > W - 1 ? W - 1 : u); v = v < 0 ? 0 : (v > H - 1 ? H - 1 : v);
      float M[4];
      bool ok = brisk2_warp(rays + ((size_t)v * W + u) * 3, jac + ((size_t)v * W + u) * 6, dir, fu, M);
      float xs[kPoints], ys[kPoints];
      for (int p = 0; ok && p < kPoints; p++) { brisk2_sample_pos(M, fx, fy, pat0[p], xs[p], ys[p]); ok = brisk2_sample_inside(xs[p], ys[p], pat0[p].sigma, W, H); }
      if (!ok) continue;
      for (int p = 0; p < kPoints; p++) val[p] = smoothed_intensity_at(img, W, integral.data(), W + 1, xs[p], ys[p], pat0[p].sigma);
      angle = (float)(atan2((double)M[2], (double)M[0]) / 3.14159265358979323846 * 180.0);
      if (angle < 0) angle += 360.f;
    }
    if (m < cap) {
      kp_out[m] = Kp{fx, fy, 12.0f, angle, (float)sc_of(i), 0, -1};
      uint32_t* w = reinterpret_cast<uint32_t*>(desc_out + (size_t)m * 48);
      for (int q = 0; q < 12; q++) w[q] = 0;
      for (int q = 0; q < kShortPairs48; q++) if (val[sp[q] & 255] > val[sp[q] >> 8]) w[q >> 5] |= 1u << (q & 31);
    }
    m++;
  }
  stats[2] = kept;
  return m;
}
