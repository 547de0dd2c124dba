// okb_harris_core.h -- per-element arithmetic of the D = 48 mode (the detector / extractor pair OKVIS2 itself constructs,
// reference okvis_frontend/src/Frontend.cpp:2406-2412: brisk::ScaleSpaceFeatureDetector<brisk::HarrisScoreCalculator>(uniformityRadius,
// octaves, absoluteThreshold, maxNumKeypoints) + brisk::BriskDescriptorExtractor(rotationInvariance, scaleInvariance), fed with the
// camera-awareness maps and the extraction direction, Frontend.cpp:232-251). Written once, compiled for the device (okb_harris.cu) and
// for the host (tests/emul), like okb_core.h.
//
// PARITY UNPINNED vs smartroboticslab/brisk@1ef8b42a: that library is an empty directory under the reference tree; the constants the
// reference does not establish are chosen here and listed in DESIGN.md section 2b. The oracle of this mode is oracle/brisk_oracle.c
// section 6.
#pragma once
#include "okb_core.h"

namespace okb {

constexpr int kShortPairs48 = 383;    // short pairs of the 60-point pattern below 5.1 x patternScale: 48 bytes, bit 383 stays zero
constexpr double kDmax48 = 5.1;
constexpr int kUniWin = 15;           // occupancy stamps cover (2 * 15 + 1)^2 half-resolution cells
constexpr int kUniLut = 2 * kUniWin + 1;
constexpr int kHarrisCandCap = 16384; // maxima per frame the uniformity kernel can rank (the score kernel flags an overflow)

// Scharr derivative / 32 at (x, y) of a pitch-linear u8 image (the caller keeps 1 <= x <= W - 2, 1 <= y <= H - 2)
OKB_HD void harris_grad(const uint8_t* r0, const uint8_t* r1, const uint8_t* r2, int x, int& gx, int& gy)
{
  const int sx = 3 * ((int)r0[x + 1] - (int)r0[x - 1]) + 10 * ((int)r1[x + 1] - (int)r1[x - 1]) + 3 * ((int)r2[x + 1] - (int)r2[x - 1]);
  const int sy = 3 * ((int)r2[x - 1] - (int)r0[x - 1]) + 10 * ((int)r2[x] - (int)r0[x]) + 3 * ((int)r2[x + 1] - (int)r0[x + 1]);
  gx = sx >> 5; gy = sy >> 5;
}
// score from the binomially weighted sums of gx gx, gy gy, gx gy over the 3x3 neighbourhood
OKB_HD int harris_score(int a, int b, int c)
{
  a >>= 4; b >>= 4; c >>= 4;
  return a * b - c * c - (((a + b) * (a + b)) >> 4);
}
// is (x, y) a maximum candidate of the row scan: at or above the threshold and no neighbour strictly greater
OKB_HD bool harris_cond(const int32_t* c /*&score[y][x]*/, int pitch, int threshold)
{
  const int v = c[0];
  if (v < threshold) return false;
  return !(c[1] > v || c[-1] > v || c[pitch] > v || c[-pitch] > v || c[pitch + 1] > v || c[pitch - 1] > v || c[-pitch + 1] > v ||
           c[-pitch - 1] > v);
}
// The row scan skips the pixel right of an accepted maximum: inside a run of consecutive candidates every second one is kept.
OKB_HDN bool harris_is_maximum(const int32_t* score, int pitch, int x, int y, int threshold)
{
  const int32_t* c = score + (size_t)y * pitch + x;
  if (!harris_cond(c, pitch, threshold)) return false;
  int run = 0;
  while (x - 1 - run >= 2 && harris_cond(c - 1 - run, pitch, threshold)) run++;
  return (run & 1) == 0;
}
OKB_HDN void harris_subpixel(const int32_t* score, int pitch, int x, int y, float& dx, float& dy)
{
  const int32_t* c = score + (size_t)y * pitch + x;
  const int v = c[0];
  int sh = 0;
  while ((v >> sh) >= 512) sh++;
  int s[9];
OKB_UNROLL
  for (int j = -1; j <= 1; j++)
OKB_UNROLL
    for (int i = -1; i <= 1; i++) {
      int q = c[j * pitch + i];
      if (q < -v) q = -v;
      s[(j + 1) * 3 + (i + 1)] = q >> sh;
    }
  subpixel2D(s[0], s[3], s[6], s[1], s[4], s[7], s[2], s[5], s[8], dx, dy);
}

// ---- uniformity enforcement. The sequential definition keeps a half-resolution occupancy image that accepted candidates stamp
// (saturating add). A candidate only ever reads the occupancy of its OWN cell, and saturating sums of non-negative stamps are
// order-free: occupancy = min(255, sum of the stamps of the accepted candidates of higher rank at that cell). The kernels evaluate
// that sum per candidate instead of building the image.
OKB_HD float uni_ratio(int score, float max_score) { return (float)score / max_score; }
OKB_HD float uni_nsc(float ratio) { return sqrtf(sqrtf(ratio)); }
OKB_HD int uni_stamp(float nsc, float lut) { return (int)((nsc * lut) * 255.0f); }
OKB_HD bool uni_rejected(float ratio, int occupancy)
{
  const float t = (float)(occupancy > 255 ? 255 : occupancy) / 255.0f;
  const float t2 = t * t;
  return ratio < t2 * t2;
}
inline float uni_lut_host(float radius, int dx, int dy)
{
  const double v = 1.0 - sqrt((double)(dx * dx + dy * dy)) / ((double)radius / 2.0);
  return (float)(v > 0.0 ? v : 0.0);
}

// ---- BRISK2 sampling
inline int brisk2_basic_scale_host()
{
  const float ln2 = 0.693147180559945f;
  const float lb_scalerange = (float)(logf(30.f) / ln2);
  const float basic06 = 12.0f * 0.6f;
  const int s = (int)(kScales / lb_scalerange * (logf(1.45f * 12.0f / basic06) / ln2) + 0.5);
  return s < 0 ? 0 : s;
}
// M = J [tx ty] / fu with ty = the extraction direction projected on the tangent plane of the ray e, tx = ty x e
OKB_HDN bool brisk2_warp(const float* e, const float* J, const float* d, float fu, float M[4])
{
  if (e[0] == 0.0f && e[1] == 0.0f && e[2] == 0.0f) return false;
  float de = (d[0] * e[0] + d[1] * e[1]) + d[2] * e[2];
  float y0 = d[0] - de * e[0], y1 = d[1] - de * e[1], y2 = d[2] - de * e[2];
  float n2 = (y0 * y0 + y1 * y1) + y2 * y2;
  if (n2 < 1e-12f) {
    de = e[1];
    y0 = 0.0f - de * e[0]; y1 = 1.0f - de * e[1]; y2 = 0.0f - de * e[2];
    n2 = (y0 * y0 + y1 * y1) + y2 * y2;
  }
  const float inv = 1.0f / sqrtf(n2);
  y0 *= inv; y1 *= inv; y2 *= inv;
  const float x0 = y1 * e[2] - y2 * e[1], x1 = y2 * e[0] - y0 * e[2], x2 = y0 * e[1] - y1 * e[0];
  M[0] = ((J[0] * x0 + J[1] * x1) + J[2] * x2) / fu;
  M[1] = ((J[0] * y0 + J[1] * y1) + J[2] * y2) / fu;
  M[2] = ((J[3] * x0 + J[4] * x1) + J[5] * x2) / fu;
  M[3] = ((J[3] * y0 + J[4] * y1) + J[5] * y2) / fu;
  return true;
}
OKB_HD void brisk2_sample_pos(const float M[4], float kx, float ky, const PatternPoint p, float& xf, float& yf)
{
  xf = kx + (M[0] * p.x + M[1] * p.y);
  yf = ky + (M[2] * p.x + M[3] * p.y);
}
OKB_HD bool brisk2_sample_inside(float xf, float yf, float s, int W, int H)
{
  return (xf - s >= 1.0f) && (xf + s < (float)(W - 2)) && (yf - s >= 1.0f) && (yf + s < (float)(H - 2));
}

}  // namespace okb
