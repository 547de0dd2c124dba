// okb_core.h -- per-element arithmetic of the detect/describe path, written once and compiled both for the
// device (nvcc, sm_100a, -fmad=false) and for the host (tests/emul, g++ -ffp-contract=off) so that the
// parallel formulation can be checked against the oracle without a GPU. Nothing here is a CPU fallback:
// the shipped library only instantiates these functions inside __global__ kernels.
//
// Algorithm: scale-space AGAST + BRISK as specified by SURVEY.md Appendix A (OpenCV-BRISK semantics), behind
// okvis::Frame::detect/describe (reference okvis_cv/include/okvis/implementation/Frame.hpp:140-175).
//
// Parallel restatement of the reference's lazily cached score map: every score query the sequential algorithm
// issues with threshold 1 returns b0(q) = max(B*(q), 0), a pure function of the layer image (B* = largest AGAST
// threshold at which q is still a corner). The only state-dependent reader is the tie-break of the 2-D maximum
// test, handled by the touch-time map (see okb_detect.cu: resolve kernel).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define OKB_HD __device__ __forceinline__
#define OKB_HDN __device__ inline
#else
#define OKB_HD inline
#define OKB_HDN inline
#endif

#if defined(__CUDA_ARCH__)
#define OKB_UNROLL _Pragma("unroll")
#else
#define OKB_UNROLL
#endif

namespace okb {

constexpr int kMaxLayers = 8;
constexpr int kPoints = 60;
constexpr int kScales = 64;
constexpr int kRot = 1024;
constexpr int kShortPairs = 512;
constexpr int kLongPairs = 870;

struct LayerView {
  const uint8_t* img;  // pitch-linear u8
  const uint8_t* b0;   // dense score map b0(q) of this layer (same size, pitch bpitch), zero in the 3-pixel margin
  int w, h, pitch, bpitch;
  float scale, offset;
};

// ---- AGAST ----------------------------------------------------------------------------------------------
OKB_HD int imin(int a, int b) { return a < b ? a : b; }
OKB_HD int imax(int a, int b) { return a > b ? a : b; }

// B*: largest b (>= -1) such that 9 contiguous ring pixels are all > p+b or all < p-b. v[] = ring intensities in
// ring order, p = centre. max over arcs of (min over arc of v) gives the bright bound, min over arcs of max the dark.
OKB_HD int bstar_from_ring16(const int v[16], int p)
{
  int mn2[16], mx2[16];
OKB_UNROLL
  for (int i = 0; i < 16; i++) { mn2[i] = imin(v[i], v[(i + 1) & 15]); mx2[i] = imax(v[i], v[(i + 1) & 15]); }
  int mn4[16], mx4[16];
OKB_UNROLL
  for (int i = 0; i < 16; i++) { mn4[i] = imin(mn2[i], mn2[(i + 2) & 15]); mx4[i] = imax(mx2[i], mx2[(i + 2) & 15]); }
  int best_b = 0, best_d = 255;
OKB_UNROLL
  for (int i = 0; i < 16; i++) {
    int mn8 = imin(mn4[i], mn4[(i + 4) & 15]);
    int mx8 = imax(mx4[i], mx4[(i + 4) & 15]);
    int mn9 = imin(mn8, v[(i + 8) & 15]);
    int mx9 = imax(mx8, v[(i + 8) & 15]);
    best_b = imax(best_b, mn9);
    best_d = imin(best_d, mx9);
  }
  return imax(best_b - p, p - best_d) - 1;
}

OKB_HD int bstar16(const uint8_t* c, int pitch)
{
  int v[16];
  v[0] = c[-3];              v[1] = c[-pitch - 3];      v[2] = c[-2 * pitch - 2];  v[3] = c[-3 * pitch - 1];
  v[4] = c[-3 * pitch];      v[5] = c[-3 * pitch + 1];  v[6] = c[-2 * pitch + 2];  v[7] = c[-pitch + 3];
  v[8] = c[3];               v[9] = c[pitch + 3];       v[10] = c[2 * pitch + 2];  v[11] = c[3 * pitch + 1];
  v[12] = c[3 * pitch];      v[13] = c[3 * pitch - 1];  v[14] = c[2 * pitch - 2];  v[15] = c[pitch - 3];
  return bstar_from_ring16(v, (int)c[0]);
}

// AGAST 5-8 (radius-1 ring), used for the virtual layer below layer 0
OKB_HD int bstar8(const uint8_t* c, int pitch)
{
  int v[8];
  v[0] = c[-1]; v[1] = c[-pitch - 1]; v[2] = c[-pitch]; v[3] = c[-pitch + 1];
  v[4] = c[1];  v[5] = c[pitch + 1];  v[6] = c[pitch];  v[7] = c[pitch - 1];
  int p = c[0];
  int best_b = 0, best_d = 255;
OKB_UNROLL
  for (int i = 0; i < 8; i++) {
    int mn = v[i], mx = v[i];
OKB_UNROLL
    for (int k = 1; k < 5; k++) { mn = imin(mn, v[(i + k) & 7]); mx = imax(mx, v[(i + k) & 7]); }
    best_b = imax(best_b, mn);
    best_d = imin(best_d, mx);
  }
  return imax(best_b - p, p - best_d) - 1;
}

// value every threshold-1 score query of the sequential algorithm returns (0 outside the 3-pixel margin):
// computed from the image ...
OKB_HD int b0_compute(const LayerView& l, int x, int y)
{
  if (x < 3 || y < 3 || x >= l.w - 3 || y >= l.h - 3) return 0;
  int s = bstar16(l.img + (size_t)y * l.pitch + x, l.pitch);
  return s < 1 ? 0 : (s > 254 ? 254 : s);
}
// ... or read from the dense map the score pass has written (what refinement and tie resolution use)
OKB_HD int b0(const LayerView& l, int x, int y)
{
  if ((unsigned)(x - 3) >= (unsigned)(l.w - 6) || (unsigned)(y - 3) >= (unsigned)(l.h - 6)) return 0;   // x < 3 || x >= w - 3 || ... (w, h >= 8)
  return l.b0[(size_t)y * l.bpitch + x];
}
OKB_HD int b0_58(const LayerView& l, int x, int y)
{
  if (x < 2 || y < 2 || x >= l.w - 2 || y >= l.h - 2) return 0;
  int s = bstar8(l.img + (size_t)y * l.pitch + x, l.pitch);
  return s < 1 ? 0 : (s > 254 ? 254 : s);
}
// bilinear score read at float coordinates (scale <= 1 branch of the reference's float getAgastScore)
OKB_HD int b0_f(const LayerView& l, float xf, float yf)
{
  const int x = (int)xf;
  const float rx1 = xf - (float)x;
  const float rx = 1.0f - rx1;
  const int y = (int)yf;
  const float ry1 = yf - (float)y;
  const float ry = 1.0f - ry1;
  const int s00 = b0(l, x, y), s10 = b0(l, x + 1, y), s01 = b0(l, x, y + 1), s11 = b0(l, x + 1, y + 1);
  return (uint8_t)(rx * ry * s00 + rx1 * ry * s10 + rx * ry1 * s01 + rx1 * ry1 * s11);
}

// ---- pyramid ------------------------------------------------------------------------------------------------
// one destination pixel of cv::resize(INTER_AREA), general (non-integer factor) path: float taps, row sums first,
// accumulated in ascending source order, then round-half-even and saturate.
OKB_HD uint8_t area_pixel(const uint8_t* src, int pitch, int xs, int xn, const float* xa, int ys, int yn, const float* ya)
{
  float sum = 0.f;
  for (int j = 0; j < yn; j++) {
    const uint8_t* row = src + (size_t)(ys + j) * pitch + xs;
    float buf = 0.f;
    for (int i = 0; i < xn; i++) buf += (float)row[i] * xa[i];
    if (j == 0) sum = ya[0] * buf; else sum += ya[j] * buf;
  }
#if defined(__CUDACC__)
  int r = __float2int_rn(sum);
#else
  int r = (int)lrintf(sum);
#endif
  return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}
OKB_HD uint8_t half_pixel(const uint8_t* src, int pitch, int x, int y)
{
  const uint8_t* r0 = src + (size_t)(2 * y) * pitch + 2 * x;
  return (uint8_t)(((int)r0[0] + r0[1] + r0[pitch] + r0[pitch + 1] + 2) >> 2);
}

// keypoint size -> pattern scale index through the 63 host-built boundaries (count of boundaries <= size)
OKB_HD int kscale_from_bounds(const float* bounds, float size)
{
  int lo = 0, hi = kScales - 1;  // number of boundaries <= size, in [0, 63]
  while (lo < hi) {
    const int mid = (lo + hi) / 2;
    if (bounds[mid] <= size) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- sub-pixel / scale refinement ---------------------------------------------------------------------------
OKB_HDN float subpixel2D(const int s_0_0, const int s_0_1, const int s_0_2, const int s_1_0, const int s_1_1,
                         const int s_1_2, const int s_2_0, const int s_2_1, const int s_2_2, float& delta_x,
                         float& delta_y)
{
  const int tmp1 = s_0_0 + s_0_2 - 2 * s_1_1 + s_2_0 + s_2_2;
  const int coeff1 = 3 * (tmp1 + s_0_1 - ((s_1_0 + s_1_2) * 2) + s_2_1);
  const int coeff2 = 3 * (tmp1 - ((s_0_1 + s_2_1) * 2) + s_1_0 + s_1_2);
  const int tmp2 = s_0_2 - s_2_0;
  const int tmp3 = (s_0_0 + tmp2 - s_2_2);
  const int tmp4 = tmp3 - 2 * tmp2;
  const int coeff3 = -3 * (tmp3 + s_0_1 - s_2_1);
  const int coeff4 = -3 * (tmp4 + s_1_0 - s_1_2);
  const int coeff5 = (s_0_0 - s_0_2 - s_2_0 + s_2_2) * 4;
  const int coeff6 = -(s_0_0 + s_0_2 - ((s_1_0 + s_0_1 + s_1_2 + s_2_1) * 2) - 5 * s_1_1 + s_2_0 + s_2_2) * 2;
  const int H_det = 4 * coeff1 * coeff2 - coeff5 * coeff5;
  if (H_det == 0) { delta_x = 0.0f; delta_y = 0.0f; return (float)coeff6 / 18.0f; }
  if (!(H_det > 0 && coeff1 < 0)) {
    int tmp_max = coeff3 + coeff4 + coeff5;
    delta_x = 1.0f; delta_y = 1.0f;
    int tmp = -coeff3 + coeff4 - coeff5;
    if (tmp > tmp_max) { tmp_max = tmp; delta_x = -1.0f; delta_y = 1.0f; }
    tmp = coeff3 - coeff4 - coeff5;
    if (tmp > tmp_max) { tmp_max = tmp; delta_x = 1.0f; delta_y = -1.0f; }
    tmp = -coeff3 - coeff4 + coeff5;
    if (tmp > tmp_max) { tmp_max = tmp; delta_x = -1.0f; delta_y = -1.0f; }
    return (float)(tmp_max + coeff1 + coeff2 + coeff6) / 18.0f;
  }
  float dx = (float)(2 * coeff2 * coeff3 - coeff4 * coeff5) / (float)(-H_det);
  float dy = (float)(2 * coeff1 * coeff4 - coeff3 * coeff5) / (float)(-H_det);
  bool tx = false, tx_ = false, ty = false, ty_ = false;
  if (dx > 1.0f) tx = true; else if (dx < -1.0f) tx_ = true;
  if (dy > 1.0f) ty = true;
  if (dy < -1.0f) ty_ = true;
  const float c1 = (float)coeff1, c2 = (float)coeff2, c3 = (float)coeff3, c4 = (float)coeff4, c5 = (float)coeff5,
              c6 = (float)coeff6;
  if (tx || tx_ || ty || ty_) {
    float dx1 = 0.0f, dx2 = 0.0f, dy1 = 0.0f, dy2 = 0.0f;
    if (tx) {
      dx1 = 1.0f; dy1 = -(float)(coeff4 + coeff5) / (float)(2 * coeff2);
      if (dy1 > 1.0f) dy1 = 1.0f; else if (dy1 < -1.0f) dy1 = -1.0f;
    } else if (tx_) {
      dx1 = -1.0f; dy1 = -(float)(coeff4 - coeff5) / (float)(2 * coeff2);
      if (dy1 > 1.0f) dy1 = 1.0f; else if (dy1 < -1.0f) dy1 = -1.0f;
    }
    if (ty) {
      dy2 = 1.0f; dx2 = -(float)(coeff3 + coeff5) / (float)(2 * coeff1);
      if (dx2 > 1.0f) dx2 = 1.0f; else if (dx2 < -1.0f) dx2 = -1.0f;
    } else if (ty_) {
      dy2 = -1.0f; dx2 = -(float)(coeff3 - coeff5) / (float)(2 * coeff1);
      if (dx2 > 1.0f) dx2 = 1.0f; else if (dx2 < -1.0f) dx2 = -1.0f;
    }
    const float max1 = (c1 * dx1 * dx1 + c2 * dy1 * dy1 + c3 * dx1 + c4 * dy1 + c5 * dx1 * dy1 + c6) / 18.0f;
    const float max2 = (c1 * dx2 * dx2 + c2 * dy2 * dy2 + c3 * dx2 + c4 * dy2 + c5 * dx2 * dy2 + c6) / 18.0f;
    if (max1 > max2) { delta_x = dx1; delta_y = dy1; return max1; }
    delta_x = dx2; delta_y = dy2; return max2;
  }
  delta_x = dx; delta_y = dy;
  return (c1 * dx * dx + c2 * dy * dy + c3 * dx + c4 * dy + c5 * dx * dy + c6) / 18.0f;
}

OKB_HD int to_fix1024(float s) { return (int)(1024.0 * (double)s + 0.5); }

OKB_HDN float refine1D(const float s_05, const float s0, const float s05, float& max)
{
  const int i_05 = to_fix1024(s_05), i0 = to_fix1024(s0), i05 = to_fix1024(s05);
  const int three_a = 16 * i_05 - 24 * i0 + 8 * i05;
  if (three_a >= 0) {
    if (s0 >= s_05 && s0 >= s05) { max = s0; return 1.0f; }
    if (s_05 >= s0 && s_05 >= s05) { max = s_05; return 0.75f; }
    if (s05 >= s0 && s05 >= s_05) { max = s05; return 1.5f; }
  }
  const int three_b = -40 * i_05 + 54 * i0 - 14 * i05;
  float ret_val = -(float)three_b / (float)(2 * three_a);
  if (ret_val < 0.75f) ret_val = 0.75f; else if (ret_val > 1.5f) ret_val = 1.5f;
  const int three_c = +24 * i_05 - 27 * i0 + 6 * i05;
  max = (float)three_c + (float)three_a * ret_val * ret_val + (float)three_b * ret_val;
  max /= 3072.0f;
  return ret_val;
}
OKB_HDN float refine1D_1(const float s_05, const float s0, const float s05, float& max)
{
  const int i_05 = to_fix1024(s_05), i0 = to_fix1024(s0), i05 = to_fix1024(s05);
  const int two_a = 9 * i_05 - 18 * i0 + 9 * i05;
  if (two_a >= 0) {
    if (s0 >= s_05 && s0 >= s05) { max = s0; return 1.0f; }
    if (s_05 >= s0 && s_05 >= s05) { max = s_05; return 0.6666666666666666666666666667f; }
    if (s05 >= s0 && s05 >= s_05) { max = s05; return 1.3333333333333333333333333333f; }
  }
  const int two_b = -21 * i_05 + 36 * i0 - 15 * i05;
  float ret_val = -(float)two_b / (float)(2 * two_a);
  if (ret_val < 0.6666666666666666666666666667f) ret_val = 0.666666666666666666666666667f;
  else if (ret_val > 1.33333333333333333333333333f) ret_val = 1.333333333333333333333333333f;
  const int two_c = +12 * i_05 - 16 * i0 + 6 * i05;
  max = (float)two_c + (float)two_a * ret_val * ret_val + (float)two_b * ret_val;
  max /= 2048.0f;
  return ret_val;
}
OKB_HDN float refine1D_2(const float s_05, const float s0, const float s05, float& max)
{
  const int i_05 = to_fix1024(s_05), i0 = to_fix1024(s0), i05 = to_fix1024(s05);
  const int a = 2 * i_05 - 4 * i0 + 2 * i05;
  if (a >= 0) {
    if (s0 >= s_05 && s0 >= s05) { max = s0; return 1.0f; }
    if (s_05 >= s0 && s_05 >= s05) { max = s_05; return 0.7f; }
    if (s05 >= s0 && s05 >= s_05) { max = s05; return 1.5f; }
  }
  const int b = -5 * i_05 + 8 * i0 - 3 * i05;
  float ret_val = -(float)b / (float)(2 * a);
  if (ret_val < 0.7f) ret_val = 0.7f; else if (ret_val > 1.5f) ret_val = 1.5f;
  const int c = +3 * i_05 - 3 * i0 + 1 * i05;
  max = (float)c + (float)a * ret_val * ret_val + (float)b * ret_val;
  max /= 1024.0f;
  return ret_val;
}

OKB_HD float patch_subpixel(const LayerView& l, int x, int y, float& dx, float& dy)
{
  const int s_0_0 = b0(l, x - 1, y - 1), s_1_0 = b0(l, x, y - 1), s_2_0 = b0(l, x + 1, y - 1);
  const int s_0_1 = b0(l, x - 1, y), s_1_1 = b0(l, x, y), s_2_1 = b0(l, x + 1, y);
  const int s_0_2 = b0(l, x - 1, y + 1), s_1_2 = b0(l, x, y + 1), s_2_2 = b0(l, x + 1, y + 1);
  return subpixel2D(s_0_0, s_0_1, s_0_2, s_1_0, s_1_1, s_1_2, s_2_0, s_2_1, s_2_2, dx, dy);
}

// What the scan of the neighbouring layer touched, needed to replay its cache side effects (touch events).
struct ScanTrace {
  int16_t n_queries;  // number of scan queries issued before the early exit (or all of them)
  int16_t exited;     // 1 = left early (no 3x3 patch around the maximum was read)
  int16_t max_x, max_y;
};

// The scan positions of the reference's getScoreMaxAbove/Below in issue order. kind 0 = float (bilinear, touches
// the 2x2 block at (int)xf,(int)yf), kind 1 = integer position. The last row never exits early.
struct ScanIter {
  float x_1, x1, y_1, y1;
  int xa, xb, ya, yb;  // integer interior range [xa, xb], [ya, yb]
  OKB_HD void init(float _x_1, float _x1, float _y_1, float _y1)
  {
    x_1 = _x_1; x1 = _x1; y_1 = _y_1; y1 = _y1;
    xa = (int)x_1 + 1; xb = (int)x1; ya = (int)y_1 + 1; yb = (int)y1;
  }
};

// Generic scan used for both directions, written with a fixed iteration structure (one loop over the at most 16
// grid positions, the early exit kept as a flag) so that the 32 candidates of a warp stay in lockstep on the device.
// Sequential semantics reproduced: positions visited in row-major order of the grid
//   rows  = [y_1, ya..yb, y1],  columns = [x_1, xa..xb, x1]   (first/last are float, bilinear reads)
// every row but the last leaves (ismax = false) at the first value above `threshold`; the running maximum is updated
// on strict >, its position following the reference's per-position formulas. BELOW adds the reference's tie rule for
// interior integer positions. An interior position read bilinearly at integer coordinates returns exactly b0.
template <bool BELOW>
OKB_HDN float scan_neighbour_layer(const LayerView& nl, const ScanIter& it, const int threshold, bool& ismax,
                                   int& max_x, int& max_y, ScanTrace* tr)
{
  const int nx = imax(it.xb - it.xa + 1, 0), ny = imax(it.yb - it.ya + 1, 0);
  const int rowlen = nx + 2, Q = (ny + 2) * rowlen;
  int mx = (int)it.x_1 + 1, my = (int)it.y_1 + 1;
  float maxval = 0.f;
  bool exited = false;
  int nq = Q;
  int r = 0, c = 0;
  for (int q = 0; q < 16; q++) {
    if (q >= Q) break;
    const bool row_first = r == 0, row_last = r == ny + 1, col_first = c == 0, col_last = c == rowlen - 1;
    const int xi = it.xa + c - 1, yi = it.ya + r - 1;
    const float xf = col_first ? it.x_1 : (col_last ? it.x1 : (float)xi);
    const float yf = row_first ? it.y_1 : (row_last ? it.y1 : (float)yi);
    const float tmp = exited ? 0.f : (float)b0_f(nl, xf, yf);
    if (!exited) {
      if (!row_last && tmp > (float)threshold) { exited = true; nq = q + 1; }
      else if (q == 0) maxval = tmp;
      else {
        if (BELOW) {
          if (!row_first && !row_last && !col_first && !col_last && tmp == maxval) {
            const int x = xi, y = yi;
            const int t1 = 2 * (b0(nl, x - 1, y) + b0(nl, x + 1, y) + b0(nl, x, y + 1) + b0(nl, x, y - 1)) +
                           (b0(nl, x + 1, y + 1) + b0(nl, x - 1, y + 1) + b0(nl, x + 1, y - 1) + b0(nl, x - 1, y - 1));
            const int t2 = 2 * (b0(nl, mx - 1, my) + b0(nl, mx + 1, my) + b0(nl, mx, my + 1) + b0(nl, mx, my - 1)) +
                           (b0(nl, mx + 1, my + 1) + b0(nl, mx - 1, my + 1) + b0(nl, mx + 1, my - 1) + b0(nl, mx - 1, my - 1));
            if (t1 > t2) { mx = x; my = y; }
          }
        }
        if (tmp > maxval) {
          maxval = tmp;
          mx = col_first ? (int)(it.x_1 + 1) : (col_last ? (int)it.x1 : xi);
          if (!row_first) my = row_last ? (int)it.y1 : yi;
        }
      }
    }
    if (++c == rowlen) { c = 0; r++; }
  }
  max_x = mx; max_y = my;
  if (tr) { tr->n_queries = (int16_t)nq; tr->exited = exited ? 1 : 0; tr->max_x = exited ? 0 : (int16_t)mx; tr->max_y = exited ? 0 : (int16_t)my; }
  ismax = !exited;
  return exited ? 0.0f : maxval;
}

OKB_HD void above_window(int layer, int x_layer, int y_layer, ScanIter& it)
{
  if (layer % 2 == 0)
    it.init((float)(4 * x_layer - 1 - 2) / 6.0f, (float)(4 * x_layer - 1 + 2) / 6.0f,
            (float)(4 * y_layer - 1 - 2) / 6.0f, (float)(4 * y_layer - 1 + 2) / 6.0f);
  else
    it.init((float)(6 * x_layer - 1 - 3) / 8.0f, (float)(6 * x_layer - 1 + 3) / 8.0f,
            (float)(6 * y_layer - 1 - 3) / 8.0f, (float)(6 * y_layer - 1 + 3) / 8.0f);
}
OKB_HD void below_window(int layer, int x_layer, int y_layer, ScanIter& it)
{
  if (layer % 2 == 0)
    it.init((float)(8 * x_layer + 1 - 4) / 6.0f, (float)(8 * x_layer + 1 + 4) / 6.0f,
            (float)(8 * y_layer + 1 - 4) / 6.0f, (float)(8 * y_layer + 1 + 4) / 6.0f);
  else
    it.init((float)(6 * x_layer + 1 - 3) / 4.0f, (float)(6 * x_layer + 1 + 3) / 4.0f,
            (float)(6 * y_layer + 1 - 3) / 4.0f, (float)(6 * y_layer + 1 + 3) / 4.0f);
}

OKB_HDN float score_max_above(const LayerView* L, const int layer, const int x_layer, const int y_layer,
                              const int threshold, bool& ismax, float& dx, float& dy, ScanTrace* tr)
{
  const LayerView& la = L[layer + 1];
  ScanIter it; above_window(layer, x_layer, y_layer, it);
  int max_x, max_y;
  const float maxval = scan_neighbour_layer<false>(la, it, threshold, ismax, max_x, max_y, tr);
  if (!ismax) return 0.0f;
  float dx_1, dy_1;
  const float refined_max = patch_subpixel(la, max_x, max_y, dx_1, dy_1);
  const float real_x = (float)max_x + dx_1;
  const float real_y = (float)max_y + dy_1;
  bool returnrefined = true;
  if (layer % 2 == 0) {
    dx = (real_x * 6.0f + 1.0f) / 4.0f - (float)x_layer;
    dy = (real_y * 6.0f + 1.0f) / 4.0f - (float)y_layer;
  } else {
    dx = (real_x * 8.0f + 1.0f) / 6.0f - (float)x_layer;
    dy = (real_y * 8.0f + 1.0f) / 6.0f - (float)y_layer;
  }
  if (dx > 1.0f) { dx = 1.0f; returnrefined = false; }
  if (dx < -1.0f) { dx = -1.0f; returnrefined = false; }
  if (dy > 1.0f) { dy = 1.0f; returnrefined = false; }
  if (dy < -1.0f) { dy = -1.0f; returnrefined = false; }
  if (returnrefined) return refined_max > maxval ? refined_max : maxval;
  return maxval;
}

OKB_HDN float score_max_below(const LayerView* L, const int layer, const int x_layer, const int y_layer,
                              const int threshold, bool& ismax, float& dx, float& dy)
{
  const LayerView& lb = L[layer - 1];
  ScanIter it; below_window(layer, x_layer, y_layer, it);
  int max_x, max_y;
  const float maxval = scan_neighbour_layer<true>(lb, it, threshold, ismax, max_x, max_y, nullptr);
  if (!ismax) return 0.0f;
  float dx_1, dy_1;
  const float refined_max = patch_subpixel(lb, max_x, max_y, dx_1, dy_1);
  const float real_x = (float)max_x + dx_1;
  const float real_y = (float)max_y + dy_1;
  bool returnrefined = true;
  if (layer % 2 == 0) {
    dx = (float)(((double)real_x * 6.0 + 1.0) / 8.0) - (float)x_layer;
    dy = (float)(((double)real_y * 6.0 + 1.0) / 8.0) - (float)y_layer;
  } else {
    dx = (float)(((double)real_x * 4.0 - 1.0) / 6.0) - (float)x_layer;
    dy = (float)(((double)real_y * 4.0 - 1.0) / 6.0) - (float)y_layer;
  }
  if (dx > 1.0f) { dx = 1.0f; returnrefined = false; }
  if (dx < -1.0f) { dx = -1.0f; returnrefined = false; }
  if (dy > 1.0f) { dy = 1.0f; returnrefined = false; }
  if (dy < -1.0f) { dy = -1.0f; returnrefined = false; }
  if (returnrefined) return refined_max > maxval ? refined_max : maxval;
  return maxval;
}

// Result of refining one 2-D maximum candidate (pure function of the pyramid).
struct RefineResult {
  float x, y, size, response;
  int8_t keep;       // 1 = a keypoint is emitted (given that the candidate is a 2-D maximum)
  int8_t own_touch;  // cache side effect on the own layer: 0 none, 1 = 3x3 around the point, 2 = 4x4 (x-1..x+2)
  int8_t has_above;  // 1 = above-layer scan was run (trace valid)
  ScanTrace above;
};

// Everything the reference does for one candidate that passed the 2-D maximum test (getKeypoints loop body), in two stages so
// that the device can regroup the candidates in between: stage 1 is the search in the layer above (about half of the candidates
// are not a maximum across scales and end there), stage 2 the layer below, the own patch and the scale / position fit.
struct RefineMid {
  float max_above, dx_above, dy_above;
  int center;
  int mode;   // 0: ended in stage 1 (no keypoint); 1: single-layer pyramid; 2: top layer; 3: between two layers (refine3D)
};

OKB_HDN void refine_stage1(const LayerView* L, int n_layers, int layer, int x_layer, int y_layer, RefineResult& r, RefineMid& m)
{
  r.keep = 0; r.own_touch = 0; r.has_above = 0;
  r.above.n_queries = 0; r.above.exited = 1; r.above.max_x = r.above.max_y = 0;
  r.x = r.y = r.size = r.response = 0.f;
  m.max_above = m.dx_above = m.dy_above = 0.f; m.center = 0;
  if (n_layers == 1) { m.mode = 1; return; }
  if (layer == n_layers - 1) { m.mode = 2; return; }
  // refine3D, first part
  bool ismax = true;
  m.center = b0(L[layer], x_layer, y_layer);
  r.has_above = 1;
  m.max_above = score_max_above(L, layer, x_layer, y_layer, m.center, ismax, m.dx_above, m.dy_above, &r.above);
  m.mode = ismax ? 3 : 0;
}

OKB_HDN void refine_stage2(const LayerView* L, int n_layers, int layer, int x_layer, int y_layer, int threshold, const RefineMid& m,
                           RefineResult& r)
{
  const float basicSize = 12.0f;
  const LayerView& tl = L[layer];
  if (m.mode == 1) {
    float dx, dy;
    const float mx = patch_subpixel(tl, x_layer, y_layer, dx, dy);
    r.x = (float)x_layer + dx; r.y = (float)y_layer + dy; r.size = basicSize; r.response = mx;
    r.keep = 1; r.own_touch = 2;
    return;
  }
  if (m.mode == 2) {
    bool ismax; float dx, dy;
    const int center = (uint8_t)(float)b0(tl, x_layer, y_layer);  // score is >= threshold here
    score_max_below(L, layer, x_layer, y_layer, center, ismax, dx, dy);
    if (!ismax) return;
    float delta_x, delta_y;
    const float mx = patch_subpixel(tl, x_layer, y_layer, delta_x, delta_y);
    r.x = ((float)x_layer + delta_x) * tl.scale + tl.offset;
    r.y = ((float)y_layer + delta_y) * tl.scale + tl.offset;
    r.size = basicSize * tl.scale; r.response = mx; r.keep = 1; r.own_touch = 2;
    return;
  }
  // refine3D, second part
  bool ismax = true;
  const int center = m.center;
  const float max_above = m.max_above, delta_x_above = m.dx_above, delta_y_above = m.dy_above;
  float max, scale, x, y;
  if (layer % 2 == 0) {
    float delta_x_below, delta_y_below;
    float max_below_float;
    if (layer == 0) {
      const int s_0_0 = b0_58(tl, x_layer - 1, y_layer - 1), s_1_0 = b0_58(tl, x_layer, y_layer - 1),
                s_2_0 = b0_58(tl, x_layer + 1, y_layer - 1), s_2_1 = b0_58(tl, x_layer + 1, y_layer),
                s_1_1 = b0_58(tl, x_layer, y_layer), s_0_1 = b0_58(tl, x_layer - 1, y_layer),
                s_0_2 = b0_58(tl, x_layer - 1, y_layer + 1), s_1_2 = b0_58(tl, x_layer, y_layer + 1),
                s_2_2 = b0_58(tl, x_layer + 1, y_layer + 1);
      int max_below = imax(imax(imax(s_0_0, s_1_0), imax(s_2_0, s_2_1)), imax(imax(s_1_1, s_0_1), imax(s_0_2, imax(s_1_2, s_2_2))));
      subpixel2D(s_0_0, s_0_1, s_0_2, s_1_0, s_1_1, s_1_2, s_2_0, s_2_1, s_2_2, delta_x_below, delta_y_below);
      max_below_float = (float)max_below;
    } else {
      max_below_float = score_max_below(L, layer, x_layer, y_layer, center, ismax, delta_x_below, delta_y_below);
      if (!ismax) return;
    }
    float delta_x_layer, delta_y_layer;
    const float max_layer = patch_subpixel(tl, x_layer, y_layer, delta_x_layer, delta_y_layer);
    r.own_touch = 1;
    const float c = (float)center > max_layer ? (float)center : max_layer;
    if (layer == 0) scale = refine1D_2(max_below_float, c, max_above, max);
    else scale = refine1D(max_below_float, c, max_above, max);
    if (scale > 1.0f) {
      const float r0 = (1.5f - scale) / .5f;
      const float r1 = 1.0f - r0;
      x = (r0 * delta_x_layer + r1 * delta_x_above + (float)x_layer) * tl.scale + tl.offset;
      y = (r0 * delta_y_layer + r1 * delta_y_above + (float)y_layer) * tl.scale + tl.offset;
    } else {
      if (layer == 0) {
        const float r0 = (scale - 0.5f) / 0.5f;
        const float r_1 = 1.0f - r0;
        x = r0 * delta_x_layer + r_1 * delta_x_below + (float)x_layer;
        y = r0 * delta_y_layer + r_1 * delta_y_below + (float)y_layer;
      } else {
        const float r0 = (scale - 0.75f) / 0.25f;
        const float r_1 = 1.0f - r0;
        x = (r0 * delta_x_layer + r_1 * delta_x_below + (float)x_layer) * tl.scale + tl.offset;
        y = (r0 * delta_y_layer + r_1 * delta_y_below + (float)y_layer) * tl.scale + tl.offset;
      }
    }
  } else {
    float delta_x_below, delta_y_below;
    const float max_below = score_max_below(L, layer, x_layer, y_layer, center, ismax, delta_x_below, delta_y_below);
    if (!ismax) return;
    float delta_x_layer, delta_y_layer;
    const float max_layer = patch_subpixel(tl, x_layer, y_layer, delta_x_layer, delta_y_layer);
    r.own_touch = 1;
    const float c = (float)center > max_layer ? (float)center : max_layer;
    scale = refine1D_1(max_below, c, max_above, max);
    if (scale > 1.0f) {
      const float r0 = 4.0f - scale * 3.0f;
      const float r1 = 1.0f - r0;
      x = (r0 * delta_x_layer + r1 * delta_x_above + (float)x_layer) * tl.scale + tl.offset;
      y = (r0 * delta_y_layer + r1 * delta_y_above + (float)y_layer) * tl.scale + tl.offset;
    } else {
      const float r0 = scale * 3.0f - 2.0f;
      const float r_1 = 1.0f - r0;
      x = (r0 * delta_x_layer + r_1 * delta_x_below + (float)x_layer) * tl.scale + tl.offset;
      y = (r0 * delta_y_layer + r_1 * delta_y_below + (float)y_layer) * tl.scale + tl.offset;
    }
  }
  scale *= tl.scale;
  if (max > (float)threshold) { r.x = x; r.y = y; r.size = basicSize * scale; r.response = max; r.keep = 1; }
}

OKB_HDN void refine_candidate(const LayerView* L, int n_layers, int layer, int x_layer, int y_layer, int threshold,
                              RefineResult& r)
{
  RefineMid m;
  refine_stage1(L, n_layers, layer, x_layer, y_layer, r, m);
  if (m.mode) refine_stage2(L, n_layers, layer, x_layer, y_layer, threshold, m, r);
}

// ---- tie-cell bitmap -------------------------------------------------------------------------------------------
// One bit per 8x8-pixel cell of a layer, set when the 5x5 window of a tied (or not yet decided, "pending") candidate
// intersects the cell. The touch-time map is only ever read inside those windows, so a maximum whose touch footprint
// misses every flagged cell does not have to emit its touches. for_each_cell visits the bit indices of the cells that the
// pixel box [x_lo, x_hi] x [y_lo, y_hi] covers (clipped to the layer on the low side and in x); f returns true to stop.
constexpr int kCellShift = 3;
constexpr int kCellWordsPerLayer = 2048;                   // 65536 cells: layers up to 2048 x 2048
template <class F>
OKB_HD bool for_each_cell(int layer_w, int x_lo, int x_hi, int y_lo, int y_hi, F f)
{
  const int cw = (layer_w + (1 << kCellShift) - 1) >> kCellShift;
  for (int cy = imax(y_lo, 0) >> kCellShift; cy <= (y_hi >> kCellShift); cy++)
    for (int cx = imax(x_lo, 0) >> kCellShift; cx <= imin(x_hi >> kCellShift, cw - 1); cx++)
      if (f(cy * cw + cx)) return true;
  return false;
}
// footprint boxes of a maximum's cache touches: own layer (x-1..x+2, y-1..y+2 covers both the 3x3 and the 4x4 patch) and
// the window of its above-layer scan with the bilinear / 3x3 margins, in coordinates of layer + 1
struct TouchBox { int x_lo, x_hi, y_lo, y_hi; };
OKB_HD TouchBox own_touch_box(int x, int y) { return TouchBox{x - 1, x + 2, y - 1, y + 2}; }

OKB_HD TouchBox above_touch_box(int layer, int x, int y)
{
  ScanIter it; above_window(layer, x, y, it);
  return TouchBox{(int)it.x_1 - 1, (int)it.x1 + 2, (int)it.y_1 - 1, (int)it.y1 + 2};
}

// ---- touch-time map (state of the reference's lazily filled score cache) --------------------------------------
// time key of a candidate = its position in the sequential processing order (layer, y, x)
OKB_HD uint32_t time_key(int layer, int x, int y) { return ((uint32_t)layer << 22) | ((uint32_t)y << 11) | (uint32_t)x; }
constexpr uint32_t kTimeMask = (1u << 25) - 1u;
// map entry = (epoch << 25) | (kTimeMask - time): larger wins (atomicMax): newer epoch first, then earlier time
OKB_HD uint32_t touch_entry(uint32_t epoch, uint32_t time) { return (epoch << 25) | (kTimeMask - time); }
OKB_HD bool touched_before(uint32_t entry, uint32_t epoch, uint32_t time)
{
  if ((entry >> 25) != epoch) return false;
  return (kTimeMask - (entry & kTimeMask)) < time;
}

// Enumerate the integer pixels the above-layer scan of a refined candidate touched (trace replay). F(x, y) is called
// for every touched pixel of layer+1 (duplicates allowed).
template <class F>
OKB_HDN void for_each_above_touch(int layer, int x_layer, int y_layer, const ScanTrace& tr, F f)
{
  ScanIter it; above_window(layer, x_layer, y_layer, it);
  int q = 0;
  const int n = tr.n_queries;
  auto fl = [&](float xf, float yf) { const int x = (int)xf, y = (int)yf; f(x, y); f(x + 1, y); f(x, y + 1); f(x + 1, y + 1); };
  if (q++ < n) fl(it.x_1, it.y_1);
  for (int x = it.xa; x <= it.xb; x++) if (q++ < n) fl((float)x, it.y_1);
  if (q++ < n) fl(it.x1, it.y_1);
  for (int y = it.ya; y <= it.yb; y++) {
    if (q++ < n) fl(it.x_1, (float)y);
    for (int x = it.xa; x <= it.xb; x++) if (q++ < n) f(x, y);
    if (q++ < n) fl(it.x1, (float)y);
  }
  if (q++ < n) fl(it.x_1, it.y1);
  for (int x = it.xa; x <= it.xb; x++) if (q++ < n) fl((float)x, it.y1);
  if (q++ < n) fl(it.x1, it.y1);
  if (!tr.exited)
    for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) f(tr.max_x + dx, tr.max_y + dy);
}

// Closed form of the scan order above: query q of the above-layer scan of a candidate sits in row q / (nx + 2) and
// column q % (nx + 2) of a grid whose first/last rows and columns are float (bilinear, 2x2 block) positions.
// Returns the integer base pixel (X, Y) and whether the 2x2 block is touched.
OKB_HD void above_query_pos(const ScanIter& it, int q, int& X, int& Y, bool& block2x2)
{
  const int nx = imax(it.xb - it.xa + 1, 0), ny = imax(it.yb - it.ya + 1, 0);
  const int rowlen = nx + 2;   // 2, 3 or 4 for the 4/3x and 3/2x windows; q < 32
  // q / rowlen without the emulated integer division ((q * 11) >> 5 == q / 3 for q < 32)
  const int r = rowlen == 2 ? (q >> 1) : (rowlen == 4 ? (q >> 2) : (rowlen == 3 && q < 32 ? (q * 11) >> 5 : q / rowlen));
  const int c = q - r * rowlen;
  X = c == 0 ? (int)it.x_1 : (c == rowlen - 1 ? (int)it.x1 : it.xa + c - 1);
  Y = r == 0 ? (int)it.y_1 : (r == ny + 1 ? (int)it.y1 : it.ya + r - 1);
  block2x2 = (r == 0) || (r == ny + 1) || (c == 0) || (c == rowlen - 1);
}

// 2-D maximum test with tie-break, on the effective map M(q) the sequential algorithm would see.
// m[5][5] = effective scores around the candidate (m[2][2] = centre). Returns true if it is a maximum.
OKB_HD bool is_max_2d_5x5(const int m[5][5])
{
  // fixed trip counts, no early return: the window stays in registers on the device
  const int center = m[2][2];
  bool ok = true;
OKB_UNROLL
  for (int dy = -1; dy <= 1; dy++)
OKB_UNROLL
    for (int dx = -1; dx <= 1; dx++) if (center < m[2 + dy][2 + dx]) ok = false;
  const int smoothedcenter = 4 * center + 2 * (m[2][1] + m[2][3] + m[1][2] + m[3][2]) + m[1][1] + m[1][3] + m[3][1] + m[3][3];
OKB_UNROLL
  for (int dy = -1; dy <= 1; dy++)
OKB_UNROLL
    for (int dx = -1; dx <= 1; dx++) {
      if (dx == 0 && dy == 0) continue;
      const int cy = 2 + dy, cx = 2 + dx;
      const int other = m[cy - 1][cx - 1] + 2 * m[cy - 1][cx] + m[cy - 1][cx + 1] + 2 * m[cy][cx - 1] + 4 * m[cy][cx] +
                        2 * m[cy][cx + 1] + m[cy + 1][cx - 1] + 2 * m[cy + 1][cx] + m[cy + 1][cx + 1];
      if (m[cy][cx] == center && other > smoothedcenter) ok = false;
    }
  return ok;
}

// ---- descriptor ---------------------------------------------------------------------------------------------
struct PatternPoint { float x, y, sigma; };

// BRISK smoothed intensity of one pattern point. image: pitch-linear u8; integral: (w+1) x (h+1) int32, pitch ipitch.
// (xf, yf) = position of the sample in the image, sigma_half = its smoothing half-width
OKB_HDN int smoothed_intensity_at(const uint8_t* image, int pitch, const int32_t* integral, int ipitch,
                                  const float xf, const float yf, const float sigma_half)
{
  const int x = (int)xf;
  const int y = (int)yf;
  const float area = 4.0f * sigma_half * sigma_half;
  int ret_val;
  if (sigma_half < 0.5f) {
    const int r_x = (int)((xf - (float)x) * 1024.0f);
    const int r_y = (int)((yf - (float)y) * 1024.0f);
    const int r_x_1 = (1024 - r_x);
    const int r_y_1 = (1024 - r_y);
    const uint8_t* ptr = image + x + (size_t)y * pitch;
    ret_val = r_x_1 * r_y_1 * (int)ptr[0] + r_x * r_y_1 * (int)ptr[1] + r_x * r_y * (int)ptr[pitch] +
              r_x_1 * r_y * (int)ptr[pitch + 1];
    return (ret_val + 512) / 1024;
  }
  const int scaling = (int)(4194304.0 / (double)area);
  const int scaling2 = (int)((double)((float)scaling * area) / 1024.0);
  const float x_1 = xf - sigma_half, x1 = xf + sigma_half, y_1 = yf - sigma_half, y1 = yf + sigma_half;
  const int x_left = (int)((double)x_1 + 0.5), y_top = (int)((double)y_1 + 0.5);
  const int x_right = (int)((double)x1 + 0.5), y_bottom = (int)((double)y1 + 0.5);
  const float r_x_1 = (float)x_left - x_1 + 0.5f;
  const float r_y_1 = (float)y_top - y_1 + 0.5f;
  const float r_x1 = x1 - (float)x_right + 0.5f;
  const float r_y1 = y1 - (float)y_bottom + 0.5f;
  const int dx = x_right - x_left - 1;
  const int dy = y_bottom - y_top - 1;
  const float fs = (float)scaling;
  const int A = (int)((r_x_1 * r_y_1) * fs);
  const int B = (int)((r_x1 * r_y_1) * fs);
  const int C = (int)((r_x1 * r_y1) * fs);
  const int D = (int)((r_x_1 * r_y1) * fs);
  const int r_x_1_i = (int)(r_x_1 * fs);
  const int r_y_1_i = (int)(r_y_1 * fs);
  const int r_x1_i = (int)(r_x1 * fs);
  const int r_y1_i = (int)(r_y1 * fs);
  const uint8_t* p00 = image + x_left + (size_t)y_top * pitch;
  // unsigned arithmetic: the reference's int sums wrap modulo 2^32
  uint32_t acc = (uint32_t)A * p00[0] + (uint32_t)B * p00[dx + 1] + (uint32_t)C * p00[(size_t)(dy + 1) * pitch + dx + 1] +
                 (uint32_t)D * p00[(size_t)(dy + 1) * pitch];
  if (dx + dy > 2) {
    const int32_t* i0 = integral + (size_t)y_top * ipitch + x_left;  // I[y_top][x_left]
    const int32_t* i1 = i0 + ipitch;                                  // I[y_top+1][.]
    const int32_t* i2 = i0 + (size_t)(dy + 1) * ipitch;               // I[y_bottom][.]
    const int32_t* i3 = i2 + ipitch;                                  // I[y_bottom+1][.]
    const int xl1 = 1, xr = dx + 1, xr1 = dx + 2;                     // offsets of x_left+1, x_right, x_right+1
    const int upper = (i1[xr] - i0[xr] + i0[xl1] - i1[xl1]);
    const int middle = (i2[xr] - i1[xr] + i1[xl1] - i2[xl1]);
    const int left = (i2[xl1] - i1[xl1] + i1[0] - i2[0]);
    const int right = (i2[xr1] - i1[xr1] + i1[xr] - i2[xr]);
    const int bottom = (i3[xr] - i2[xr] + i2[xl1] - i3[xl1]);
    acc += (uint32_t)upper * (uint32_t)r_y_1_i + (uint32_t)middle * (uint32_t)scaling + (uint32_t)left * (uint32_t)r_x_1_i +
           (uint32_t)right * (uint32_t)r_x1_i + (uint32_t)bottom * (uint32_t)r_y1_i;
  } else {
    // small boxes: direct sums (at most 3 interior pixels in total)
    int upper = 0, bottom = 0, left = 0, right = 0, middle = 0;
    for (int i = 1; i <= dx; i++) { upper += p00[i]; bottom += p00[(size_t)(dy + 1) * pitch + i]; }
    for (int j = 1; j <= dy; j++) {
      const uint8_t* row = p00 + (size_t)j * pitch;
      left += row[0]; right += row[dx + 1];
      for (int i = 1; i <= dx; i++) middle += row[i];
    }
    acc += (uint32_t)upper * (uint32_t)r_y_1_i + (uint32_t)middle * (uint32_t)scaling + (uint32_t)left * (uint32_t)r_x_1_i +
           (uint32_t)right * (uint32_t)r_x1_i + (uint32_t)bottom * (uint32_t)r_y1_i;
  }
  return (int)(acc + (uint32_t)(scaling2 / 2)) / scaling2;
}
OKB_HDN int smoothed_intensity(const uint8_t* image, int pitch, const int32_t* integral, int ipitch,
                               const float key_x, const float key_y, const PatternPoint bp)
{
  return smoothed_intensity_at(image, pitch, integral, ipitch, bp.x + key_x, bp.y + key_y, bp.sigma);
}

}  // namespace okb
