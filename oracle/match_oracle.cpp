/*
 * oracle/match_oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle of the matchers; never on the product path).
 *
 * Plain C++17 transcription of the reference's match loops, loop order and strict-< tie rule preserved:
 *   M1 Frontend::matchToMapByThread            okvis_frontend/src/Frontend.cpp:1515-1590
 *   M2 Frontend::matchToMapByThreadUnitialised okvis_frontend/src/Frontend.cpp:1594-1720
 *   M3 Frontend::matchMotionStereo worker      okvis_frontend/src/Frontend.cpp:1809-1907
 *   M4 Frontend::matchStereo                   okvis_frontend/src/Frontend.cpp:2016-2074
 *   M5 Frontend::verifyRecognisedPlace matcher okvis_frontend/src/Frontend.cpp:329-355
 *   G1 triangulation::triangulateFast          okvis_frontend/src/stereo_triangulation.cpp:50-132
 *   H0 brisk::Hamming::PopcntofXORed           (external/brisk, absent) = popcount(a XOR b) over 16*n bytes
 * Threads: the k-range split of Frontend.cpp:1536-1538 (M1/M2) and the strided k0 split of :1812 (M3).
 *
 * Eigen is not available in the authoring container; 3-vector expressions are written out with the association
 * (x0*y0 + x1*y1) + x2*y2 for dot products and element-wise division for normalized(). Whether the reference's Eigen
 * build associates identically at the last ulp is UNPINNED (it cannot be compiled here); the loops, gates, constants
 * and tie rules are pinned by the source lines cited above. Built with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "camera_project.h"

namespace {

struct V3 { double x, y, z; };
inline V3 ld(const double* p) { return V3{p[0], p[1], p[2]}; }
inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mul(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { double z = dot(a, a); if (z > 0) { double n = sqrt(z); return V3{a.x / n, a.y / n, a.z / n}; } return a; }

inline uint32_t popcnt_xored(const uint8_t* a, const uint8_t* b, int n128)
{
  uint32_t d = 0;
  for (int i = 0; i < n128 * 2; i++) { uint64_t x, y; memcpy(&x, a + 8 * i, 8); memcpy(&y, b + 8 * i, 8); d += (uint32_t)__builtin_popcountll(x ^ y); }
  return d;
}

// stereo_triangulation.cpp:50-132
V3 triangulateFast(V3 p1, V3 e1, V3 p2, V3 e2, double sigma, bool& isValid, bool& isParallel)
{
  isParallel = false; isValid = true;
  V3 t12 = sub(p2, p1);
  double b0 = dot(t12, e1), b1 = dot(t12, e2);
  double A00 = dot(e1, e1), A10 = dot(e1, e2), A01 = -A10, A11 = -dot(e2, e2);
  double det = A00 * A11 - A10 * A01;
  bool invertible = fabs(det) > 1.0e-12;
  double l0 = 0, l1 = 0;
  if (invertible) {
    double invdet = 1.0 / det;
    double i00 = A11 * invdet, i10 = -A10 * invdet, i01 = -A01 * invdet, i11 = A00 * invdet;
    l0 = i00 * b0 + i01 * b1; l1 = i10 * b0 + i11 * b1;
  }
  if (!invertible || l0 < 0.01 || l1 < 0.01) {
    isParallel = true; isValid = true;
    V3 m = add(p1, mul(0.5, t12));
    V3 midpoint = add(m, mul(40.0 * std::max(0.01, norm(t12)), add(e1, e2)));
    if (dot(e1, normalized(sub(midpoint, p1))) < cos(2.6 * sigma)) isValid = false;
    if (dot(e2, normalized(sub(midpoint, p2))) < cos(2.6 * sigma)) isValid = false;
    return midpoint;
  }
  V3 xm = add(mul(l0, e1), p1), xn = add(mul(l1, e2), p2);
  V3 s = add(xm, xn);
  V3 midpoint = V3{s.x / 2.0, s.y / 2.0, s.z / 2.0};
  if (dot(e1, normalized(sub(midpoint, p1))) < cos(2.6 * sigma)) isValid = false;
  if (dot(e2, normalized(sub(midpoint, p2))) < cos(2.6 * sigma)) isValid = false;
  if (dot(normalized(sub(midpoint, p2)), normalized(sub(midpoint, p1))) > cos(6.0 * sigma)) isParallel = true;
  return midpoint;
}

inline double depth(const double* T, V3 p) { return (((T[8] * p.x + T[9] * p.y) + T[10] * p.z) + T[11] * 1.0) / 1.0; }

template <class F> void run_threads(int n_threads, F f)
{
  if (n_threads <= 1) { f(0, 1); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) th.emplace_back(f, t, n_threads);
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

void okvo_triangulate_fast(const double p1[3], const double e1[3], const double p2[3], const double e2[3], double sigma,
                           double hp[4], int* valid, int* parallel)
{
  bool v, p; V3 m = triangulateFast(ld(p1), ld(e1), ld(p2), ld(e2), sigma, v, p);
  hp[0] = m.x; hp[1] = m.y; hp[2] = m.z; hp[3] = 1.0; *valid = v; *parallel = p;
}

void okvo_hamming_matrix(int D, int na, const uint8_t* a, int nb, const uint8_t* b, uint16_t* out)
{
  for (int i = 0; i < na; i++) for (int j = 0; j < nb; j++) out[(size_t)i * nb + j] = (uint16_t)popcnt_xored(a + (size_t)i * D, b + (size_t)j * D, D / 16);
}

// M1 -- landmarks outer (ascending id == ascending slot), keypoints inner, descriptors innermost
void okvo_match_map3d(int D, int n_kp, const uint8_t* kp_desc, const double* kp_xy, const uint8_t* kp_use, int n_cand,
                      const uint8_t* cand_desc, const int32_t* cand_lm, int n_lm, const double* lm_proj,
                      const uint8_t* lm_is3d, double reprojectionThreshold, uint32_t matchThreshold, uint32_t* out_dist,
                      int32_t* out_lm, int n_threads)
{
  (void)n_lm;
  const double thrSq = reprojectionThreshold * reprojectionThreshold;
  std::vector<double> distances(n_kp, (double)matchThreshold);
  for (int k = 0; k < n_kp; k++) out_lm[k] = -1;
  run_threads(n_threads, [&](int t, int T) {
    const int segment = n_kp / T, startK = segment * t, endK = t + 1 == T ? n_kp : startK + segment;
    int c = 0;
    while (c < n_cand) {
      const int lm = cand_lm[c];
      int c_end = c; while (c_end < n_cand && cand_lm[c_end] == lm) c_end++;
      if (lm_is3d[lm]) {
        const double px = lm_proj[2 * lm], py = lm_proj[2 * lm + 1];
        for (int k = startK; k < endK; k++) {
          if (kp_use && !kp_use[k]) continue;
          const double dx = px - kp_xy[2 * k], dy = py - kp_xy[2 * k + 1];
          if (dx * dx + dy * dy > thrSq) continue;
          for (int d = c; d < c_end; d++) {
            const double dist = popcnt_xored(kp_desc + (size_t)k * D, cand_desc + (size_t)d * D, D / 16);
            if (dist < distances[k]) { distances[k] = dist; out_lm[k] = lm; }
          }
        }
      }
      c = c_end;
    }
  });
  for (int k = 0; k < n_kp; k++) out_dist[k] = (uint32_t)distances[k];
}

// M2
void okvo_match_map_uninit(int D, int n_kp, const uint8_t* kp_desc, const double* kp_e_W, const uint8_t* kp_use,
                           const int32_t* kp_prev_lm, int n_cand, const uint8_t* cand_desc, const int32_t* cand_lm,
                           const double* cand_e_W, const double* cand_r_W, int n_lm, const uint8_t* lm_is3d,
                           const double r_WC1[3], double sigma, uint32_t matchThreshold, uint32_t* out_dist,
                           int32_t* out_lm, double* out_hp_W, int32_t* out_ctr, int n_threads)
{
  (void)n_lm;
  std::vector<double> distances(n_kp, (double)matchThreshold);
  for (int k = 0; k < n_kp; k++) { out_lm[k] = -1; out_hp_W[4 * k] = out_hp_W[4 * k + 1] = out_hp_W[4 * k + 2] = out_hp_W[4 * k + 3] = 0.0; }
  const double cos6Sigma = cos(6.0 * sigma);
  const V3 r1 = ld(r_WC1);
  std::vector<int> ctrs(std::max(n_threads, 1), 0);
  run_threads(n_threads, [&](int t, int T) {
    const int segment = n_kp / T, startK = segment * t, endK = t + 1 == T ? n_kp : startK + segment;
    int c = 0;
    while (c < n_cand) {
      const int lm = cand_lm[c];
      int c_end = c; while (c_end < n_cand && cand_lm[c_end] == lm) c_end++;
      if (!lm_is3d[lm]) {
        for (int k = startK; k < endK; k++) {
          if (kp_use && !kp_use[k]) continue;
          const V3 e1_W = ld(kp_e_W + 3 * (size_t)k);
          for (int d = c; d < c_end; d++) {
            const double dist = popcnt_xored(kp_desc + (size_t)k * D, cand_desc + (size_t)d * D, D / 16);
            if (dist < distances[k]) {
              const V3 e0_W = ld(cand_e_W + 3 * (size_t)d), r0_W = ld(cand_r_W + 3 * (size_t)d);
              if (dot(e0_W, e1_W) < cos6Sigma) {
                const V3 et_W = normalized(sub(r1, r0_W));
                const V3 n0_W = normalized(cross(e0_W, et_W));
                const V3 n1_W = normalized(cross(e1_W, et_W));
                if (dot(n0_W, n1_W) < cos6Sigma) continue;
                if (dot(cross(e0_W, e1_W), normalized(add(n0_W, n0_W))) > 0.0) continue;
              }
              bool isValid = false, isParallel = false;
              const V3 hp = triangulateFast(r0_W, e0_W, r1, e1_W, sigma, isValid, isParallel);
              if (!isValid) continue;
              if (!isParallel) {
                if (norm(sub(hp, r0_W)) < 0.2) isValid = false;
                if (norm(sub(hp, r1)) < 0.2) isValid = false;
              }
              if (!isValid) continue;
              if (kp_prev_lm && kp_prev_lm[k] >= 0 && lm == kp_prev_lm[k]) { ctrs[t]++; break; }
              distances[k] = dist; out_lm[k] = lm;
              if (!isParallel) { out_hp_W[4 * k] = hp.x; out_hp_W[4 * k + 1] = hp.y; out_hp_W[4 * k + 2] = hp.z; out_hp_W[4 * k + 3] = 1.0; }
            }
          }
        }
      }
      c = c_end;
    }
  });
  int ctr = 0; for (int v : ctrs) ctr += v;
  if (out_ctr) *out_ctr = ctr;
  for (int k = 0; k < n_kp; k++) out_dist[k] = (uint32_t)distances[k];
}

// M3
void okvo_match_motion_stereo(int D, int n0, const uint8_t* desc0, const uint8_t* use0, const double* e0_W,
                              const double* size_over_f0, int n1, const uint8_t* desc1, const uint8_t* valid1,
                              const double* e1_W, const double r_WC0[3], const double r_WC1[3], const double T_CW0[12],
                              const double T_CW1[12], uint32_t matchThreshold, int32_t* out_k1, uint32_t* out_dist,
                              double* out_hp_W, uint8_t* out_initialisable, int n_threads)
{
  const V3 r0 = ld(r_WC0), r1 = ld(r_WC1);
  run_threads(n_threads, [&](int t, int T) {
    for (int k0 = t; k0 < n0; k0 += T) {
      out_k1[k0] = -1; out_dist[k0] = matchThreshold; out_initialisable[k0] = 0;
      for (int i = 0; i < 4; i++) out_hp_W[4 * k0 + i] = 0.0;
      if (use0 && !use0[k0]) continue;
      uint32_t distances = matchThreshold;
      const V3 e0 = ld(e0_W + 3 * (size_t)k0);
      const double sigma = size_over_f0[k0] * 0.125;
      for (int kk = 0; kk < n1; kk++) {
        const uint32_t dist = popcnt_xored(desc0 + (size_t)k0 * D, desc1 + (size_t)kk * D, D / 16);
        if (dist < distances) {
          bool isValid = false, isParallel = false;
          if (!valid1[kk]) continue;
          const V3 e1 = ld(e1_W + 3 * (size_t)kk);
          if (dot(e0, e1) < 0.5) continue;
          V3 hp = triangulateFast(r0, e0, r1, e1, sigma, isValid, isParallel);
          if (!isValid) continue;
          const double z0 = depth(T_CW0, hp), z1 = depth(T_CW1, hp);
          if (dot(e0, e1) < 0.8) isValid = false;
          if (!isParallel) { if (z0 < 0.2) isValid = false; if (z1 < 0.2) isValid = false; }
          if (isValid) {
            out_k1[k0] = kk; distances = dist; out_initialisable[k0] = !isParallel;
            out_hp_W[4 * k0] = hp.x; out_hp_W[4 * k0 + 1] = hp.y; out_hp_W[4 * k0 + 2] = hp.z; out_hp_W[4 * k0 + 3] = 1.0;
          }
        }
      }
      out_dist[k0] = distances;
    }
  });
}


// M3 as a sequence over the older keyframes of one camera (Frontend::matchMotionStereo, Frontend.cpp:1775-1958): per older
// frame the worker loop (:1809-1907, including the 4 px re-projection check :1897-1904) and then the serial insertion
// (:1915-1954): ascending k0, a match is inserted unless its k1 already carries a landmark -- from before the call or from
// an earlier k0 of this loop -- and the inserted k1 drop out of the candidate set `k1s` of the next older frame (:1789-1801).
// The estimator-state tests (:1813-1821, :1840, :1923-1932) are the caller's `use0` flags; `quality` (acos, :1888) only
// feeds estimator.setLandmark and stays with the caller.
// view i: n0[i] keypoints, arrays at offset off[i] (in keypoints) of desc0 / rays0 / valid0 / size0 / use0; poses T_WC0[i] and
// its inverse T_CW0[i] = T_WC0[i].inverse() as the caller's Transformation gives them (12 doubles each: C row-major, then r);
// outputs at the same offsets. flags: bit 0 matching (matchInfos[k0].matching), bit 1 initialisable, bit 2 inserted.
void okvo_match_motion_stereo_sequence(int D, int n_views, const int32_t* n0, const int32_t* off, const uint8_t* desc0,
                                       const double* rays0, const uint8_t* valid0, const float* size0, const uint8_t* use0,
                                       const double* T_WC0, const double* T_CW0, int n1, const uint8_t* desc1, const double* rays1,
                                       const uint8_t* valid1, const float* xy1 /* n1 x 2 keypoint pt */, const double T_WC1[12],
                                       const double T_CW1_in[12],
                                       int model, const double* intr /* fu fv cu cv k0..k3 */, int width, int height,
                                       uint32_t matchThreshold, uint8_t* matched1 /* n1, in/out */, int32_t* out_k1,
                                       uint32_t* out_dist, double* out_hp_W, uint8_t* out_flags, int n_threads)
{
  const double f0 = 0.5 * (intr[0] + intr[1]);
  auto rot = [](const double* C, V3 v) {
    return V3{(C[0] * v.x + C[1] * v.y) + C[2] * v.z, (C[3] * v.x + C[4] * v.y) + C[5] * v.z, (C[6] * v.x + C[7] * v.y) + C[8] * v.z};
  };
  auto as3x4 = [](const double* T, double* out) {   // (C row-major, r) -> row-major 3x4 [C | r], the layout depth() reads
    for (int i = 0; i < 3; i++) { out[4 * i] = T[3 * i]; out[4 * i + 1] = T[3 * i + 1]; out[4 * i + 2] = T[3 * i + 2]; out[4 * i + 3] = T[9 + i]; }
  };
  double T_CW1[12]; as3x4(T_CW1_in, T_CW1);
  const V3 r1 = V3{T_WC1[9], T_WC1[10], T_WC1[11]};
  std::vector<double> e1_W((size_t)n1 * 3);
  for (int k = 0; k < n1; k++) { const V3 e = normalized(rot(T_WC1, ld(rays1 + 3 * (size_t)k))); e1_W[3 * k] = e.x; e1_W[3 * k + 1] = e.y; e1_W[3 * k + 2] = e.z; }
  for (int v = 0; v < n_views; v++) {
    const int n = n0[v]; const size_t o = (size_t)off[v];
    const double* Tw0 = T_WC0 + 12 * (size_t)v;
    double T_CW0v[12]; as3x4(T_CW0 + 12 * (size_t)v, T_CW0v);
    const V3 r0 = V3{Tw0[9], Tw0[10], Tw0[11]};
    // the compacted unmatched set k1s and its descriptors (Frontend.cpp:1789-1801)
    std::vector<int> k1s; k1s.reserve(n1);
    for (int k1 = 0; k1 < n1; k1++) if (!matched1[k1]) k1s.push_back(k1);
    run_threads(n_threads, [&](int t, int T) {
      for (int k0 = t; k0 < n; k0 += T) {
        out_k1[o + k0] = -1; out_dist[o + k0] = matchThreshold; out_flags[o + k0] = 0;
        for (int i = 0; i < 4; i++) out_hp_W[4 * (o + k0) + i] = 0.0;
        if (use0 && !use0[o + k0]) continue;
        if (!valid0[o + k0]) continue;
        uint32_t distances = matchThreshold;
        bool initialisable = false;
        V3 hps = V3{0, 0, 0};
        int k1_max = -1;
        const V3 e0 = normalized(rot(Tw0, ld(rays0 + 3 * (o + k0))));
        const double sigma = (double)size0[o + k0] / f0 * 0.125;
        for (size_t kk = 0; kk < k1s.size(); kk++) {
          const int k1 = k1s[kk];
          const uint32_t dist = popcnt_xored(desc0 + (o + k0) * D, desc1 + (size_t)k1 * D, D / 16);
          if (dist < distances) {
            bool isValid = false, isParallel = false;
            if (!valid1[k1]) continue;
            const V3 e1 = ld(&e1_W[3 * (size_t)k1]);
            if (dot(e0, e1) < 0.5) continue;
            V3 hp = triangulateFast(r0, e0, r1, e1, sigma, isValid, isParallel);
            if (!isValid) continue;
            const double z0 = depth(T_CW0v, hp), z1 = depth(T_CW1, hp);
            if (dot(e0, e1) < 0.8) isValid = false;
            if (!isParallel) { if (z0 < 0.2) isValid = false; if (z1 < 0.2) isValid = false; }
            if (isValid) { k1_max = k1; distances = dist; hps = hp; initialisable = !isParallel; }
          }
        }
        out_dist[o + k0] = distances;
        if (distances < matchThreshold) {
          out_k1[o + k0] = k1_max;
          out_hp_W[4 * (o + k0)] = hps.x; out_hp_W[4 * (o + k0) + 1] = hps.y; out_hp_W[4 * (o + k0) + 2] = hps.z; out_hp_W[4 * (o + k0) + 3] = 1.0;
          // re-projection into the current view (:1897-1904): T_WC1.inverse() * hps_W, projectHomogeneous (w = 1 > 0)
          const V3 c = rot(T_CW1_in, hps);   // Transformation::operator*(Vector4d): C * head + r * s
          const okvo_cam::P3 pc{c.x + T_CW1_in[9] * 1.0, c.y + T_CW1_in[10] * 1.0, c.z + T_CW1_in[11] * 1.0};
          double pt1p[2];
          const auto st = okvo_cam::project(model, intr, width, height, pc, pt1p);
          const double dx = (double)xy1[2 * k1_max] - pt1p[0], dy = (double)xy1[2 * k1_max + 1] - pt1p[1];
          if (st == okvo_cam::Successful && sqrt(dx * dx + dy * dy) < 4.0) out_flags[o + k0] = (uint8_t)(1 | (initialisable ? 2 : 0));
          else if (initialisable) out_flags[o + k0] = 2;
        }
      }
    });
    for (int k0 = 0; k0 < n; k0++) {   // serial insertion
      if (!(out_flags[o + k0] & 1)) continue;
      const int k1 = out_k1[o + k0];
      if (matched1[k1]) continue;      // already matched
      matched1[k1] = 1;
      out_flags[o + k0] |= 4;
    }
  }
}

// M4 (single-threaded in the reference; n_threads > 1 splits k0 for the all-cores baseline)
void okvo_match_stereo(int D, int n0, const uint8_t* desc0, const uint8_t* valid0, const double* e0_W,
                       const double* size_over_f0, int n1, const uint8_t* desc1, const uint8_t* valid1, const double* e1_W,
                       const double* size_over_f1, const double r_WC0[3], const double r_WC1[3], const double T_CW0[12],
                       const double T_CW1[12], uint32_t matchThreshold, int32_t* out_k1, uint32_t* out_dist,
                       double* out_hp_W, uint8_t* out_initialisable, int n_threads)
{
  const V3 r0 = ld(r_WC0), r1 = ld(r_WC1);
  run_threads(n_threads, [&](int t, int T) {
    for (int k0 = t; k0 < n0; k0 += T) {
      double distances = matchThreshold;
      bool initialisable = false;
      V3 hps = V3{0, 0, 0}; bool have = false;
      int k1_match = -1;
      for (int k1 = 0; k1 < n1; k1++) {
        const uint32_t dist = popcnt_xored(desc0 + (size_t)k0 * D, desc1 + (size_t)k1 * D, D / 16);
        if (dist < distances) {
          const double sigma = std::max(size_over_f0[k0], size_over_f1[k1]) * 0.125;
          bool isValid = false, isParallel = false;
          if (valid0 && !valid0[k0]) continue;
          if (!valid1[k1]) continue;
          const V3 e0 = ld(e0_W + 3 * (size_t)k0), e1 = ld(e1_W + 3 * (size_t)k1);
          V3 hp = triangulateFast(r0, e0, r1, e1, sigma, isValid, isParallel);
          const double z0 = depth(T_CW0, hp), z1 = depth(T_CW1, hp);
          if (!isParallel) {
            if (z0 < 0.05) isValid = false;
            if (z1 < 0.05) isValid = false;
            if (dot(e0, e1) < 0.8) isValid = false;
          }
          if (isValid) { distances = dist; hps = hp; have = true; k1_match = k1; initialisable = !isParallel; }
        }
      }
      out_k1[k0] = k1_match; out_dist[k0] = (uint32_t)distances; out_initialisable[k0] = initialisable;
      out_hp_W[4 * k0] = hps.x; out_hp_W[4 * k0 + 1] = hps.y; out_hp_W[4 * k0 + 2] = hps.z; out_hp_W[4 * k0 + 3] = have ? 1.0 : 0.0;
    }
  });
}

// M5
void okvo_match_place(int D, int n_lm, const int32_t* lm_offsets, const uint8_t* lm_desc, int n_kp, const uint8_t* kp_desc,
                      uint32_t matchThreshold, int32_t* out_k, uint32_t* out_dist)
{
  for (int i = 0; i < n_lm; i++) {
    uint32_t distMin = matchThreshold; int kMin = -1;
    for (int d = lm_offsets[i]; d < lm_offsets[i + 1]; d++)
      for (int k = 0; k < n_kp; k++) {
        const uint32_t dist = popcnt_xored(kp_desc + (size_t)k * D, lm_desc + (size_t)d * D, D / 16);
        if (dist < distMin) { distMin = dist; kMin = k; }
      }
    out_k[i] = kMin; out_dist[i] = distMin;
  }
}

}  // extern "C"
