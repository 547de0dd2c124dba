// okb_tables.h -- host-side generation of the static look-up tables (no CUDA types): BRISK sampling pattern
// (60 points x 64 scales x 1024 rotations), short/long pairs, the keypoint-size -> pattern-scale map and the INTER_AREA
// tap tables of the pyramid. Host libm is used on purpose: the values must be bit-identical to what a CPU BRISK builds
// with the same libm; none of this is per-frame work.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "okb_core.h"

namespace okb {

struct AreaAxis {
  std::vector<int> start, count;
  std::vector<float> alpha;  // 4 per destination index
};

struct LongPair { int i, j, wdx, wdy; };

struct HostTables {
  std::vector<PatternPoint> pattern;   // [scale][rot][point]
  std::vector<uint32_t> size_list;     // pattern extent per scale index
  std::vector<uint32_t> short_pairs;   // i | j << 8
  std::vector<LongPair> long_pairs;
  std::vector<float> scale_bounds;     // 63 boundaries of the size -> scale index map
};

inline void build_area_axis(int ssize, int dsize, AreaAxis& ax)
{
  const double inv = (double)dsize / ssize;
  const double scale = 1. / inv;
  ax.start.assign(dsize, 0); ax.count.assign(dsize, 0); ax.alpha.assign((size_t)dsize * 4, 0.f);
  for (int d = 0; d < dsize; d++) {
    const double f1 = d * scale, f2 = f1 + scale;
    const double cell = std::min(scale, ssize - f1);
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = std::min(s2, ssize - 1);
    s1 = std::min(s1, s2);
    int n = 0, first = -1;
    auto push = [&](int si, float a) { if (first < 0) first = si; if (n < 4) ax.alpha[(size_t)d * 4 + n] = a; n++; };
    if (s1 - f1 > 1e-3) push(s1 - 1, (float)((s1 - f1) / cell));
    for (int s = s1; s < s2; s++) push(s, (float)(1.0 / cell));
    if (f2 - s2 > 1e-3) push(s2, (float)(std::min(std::min(f2 - s2, 1.), cell) / cell));
    ax.start[d] = first < 0 ? 0 : first;
    ax.count[d] = std::min(n, 4);
  }
}

inline int kscale_host(float size)
{
  // index of the pattern scale used for a keypoint of this size
  const float ln2 = 0.693147180559945f;
  const float lb_scalerange = (float)(logf(30.f) / ln2);
  const float basic06 = 12.0f * 0.6f;
  int s = (int)(kScales / lb_scalerange * (logf(size / basic06) / ln2) + 0.5);
  return std::min(std::max(s, 0), kScales - 1);
}

inline bool build_host_tables(float pattern_scale, HostTables& T)
{
  const double f = 0.85 * pattern_scale;
  const float radius[5] = {(float)(f * 0.), (float)(f * 2.9), (float)(f * 4.9), (float)(f * 7.4), (float)(f * 10.8)};
  const int number[5] = {1, 10, 14, 15, 20};
  const float d_max = (float)(5.85 * pattern_scale), d_min = (float)(8.2 * pattern_scale);
  std::vector<PatternPoint>& pat = T.pattern; pat.resize((size_t)kPoints * kScales * kRot);
  std::vector<uint32_t>& size_list = T.size_list; size_list.assign(kScales, 0);
  const float lb_scale = (float)(logf(30.f) / log(2.0));
  const float lb_step = lb_scale / (float)kScales;
  size_t idx = 0;
  for (unsigned s = 0; s < (unsigned)kScales; s++) {
    const float sc = (float)pow(2.0, (double)(s * lb_step));
    for (int rot = 0; rot < kRot; rot++) {
      const double theta = (double)rot * 2 * M_PI / (double)kRot;
      for (int ring = 0; ring < 5; ring++)
        for (int num = 0; num < number[ring]; num++) {
          const double alpha = (double)num * 2 * M_PI / (double)number[ring];
          PatternPoint& p = pat[idx++];
          p.x = (float)(sc * radius[ring] * cos(alpha + theta));
          p.y = (float)(sc * radius[ring] * sin(alpha + theta));
          p.sigma = ring == 0 ? 1.3f * sc * 0.5f : (float)(1.3f * sc * (double)radius[ring] * sin(M_PI / number[ring]));
          const uint32_t ext = (uint32_t)(int)ceil((sc * radius[ring]) + p.sigma) + 1;
          size_list[s] = std::max(size_list[s], ext);
        }
    }
  }
  std::vector<uint32_t>& sp = T.short_pairs; std::vector<LongPair>& lp = T.long_pairs; sp.clear(); lp.clear();
  for (unsigned i = 1; i < (unsigned)kPoints; i++)
    for (unsigned j = 0; j < i; j++) {
      const float dx = pat[j].x - pat[i].x, dy = pat[j].y - pat[i].y;
      const float n2 = dx * dx + dy * dy;
      if (n2 > d_min * d_min) {
        LongPair e; e.i = (int)i; e.j = (int)j;
        e.wdx = (int)((dx / n2) * 2048.0 + 0.5); e.wdy = (int)((dy / n2) * 2048.0 + 0.5);
        lp.push_back(e);
      } else if (n2 < d_max * d_max) {
        sp.push_back(i | (j << 8));
      }
    }
  if ((int)sp.size() != kShortPairs || (int)lp.size() != kLongPairs) {
    return false;
  }
  // size -> scale index is monotone: store for k = 1..63 the smallest float whose index is >= k
  std::vector<float>& bounds = T.scale_bounds; bounds.resize(kScales - 1);
  for (int k = 1; k < kScales; k++) {
    uint32_t lo = 0x3f800000u /*1.0f*/, hi = 0x45800000u /*4096.f*/;  // positive floats order like their bits
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      float fm; memcpy(&fm, &mid, 4);
      if (kscale_host(fm) >= k) hi = mid; else lo = mid + 1;
    }
    memcpy(&bounds[k - 1], &lo, 4);
  }
  return true;
}

}  // namespace okb
