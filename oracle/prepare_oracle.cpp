/*
 * oracle/prepare_oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle of P1; never on the product path).
 *
 * Restatement of the landmark-candidate preparation that precedes map matching, the serial host loop of
 * Frontend::matchToMap (reference okvis_frontend/src/Frontend.cpp:1196-1360): for one camera of the current frame,
 * every landmark of the map is projected into the view (PinholeCamera::projectHomogeneous -> project,
 * okvis_cv/include/okvis/cameras/implementation/PinholeCamera.hpp:257-292,493-502; RadialTangentialDistortion::distort,
 * implementation/RadialTangentialDistortion.hpp:90-109; EquidistantDistortion::distort, EquidistantDistortion.hpp:86-106),
 * gated by the field of view +- reprThreshold, and its observations are walked NEWEST FIRST (std::set reverse order) to
 *   - decide is3d (Frontend.cpp:1289-1296),
 *   - drop observations with > 0.6 rad viewpoint change or > 50 % scale change (:1298-1310),
 *   - keep the "best 3" descriptors by score (:1312-1343) in a pool of 48(D)-byte rows.
 * The loop is transcribed with its quirks, because downstream code (M1/M2) consumes exactly what it leaves behind:
 *   the accepted descriptor is written to row `o` (not to worstIdx), `o = max(o, worstIdx)` afterwards, and the landmark is
 *   cropped to `o` rows -- so the first accepted observation is overwritten by the second, a landmark needs two accepted
 *   observations to survive, and later landmarks overwrite the stale third row of earlier ones in the shared pool.
 * Eigen expressions are written out with the association (x0*y0 + x1*y1) + x2*y2, C*v row by row, normalized() as
 * element-wise division by sqrt(squaredNorm). T_WC of every (old frame, camera) and T_CW of the current camera are inputs
 * (the caller forms them with okvis::kinematics::Transformation, Frontend.cpp:1214-1216,1276-1280).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {
struct V3 { double x, y, z; };
inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double norm(V3 a) { return sqrt(dot(a, a)); }
inline V3 normalized(V3 a) { const double n = norm(a); return V3{a.x / n, a.y / n, a.z / n}; }
inline V3 scale(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 rot(const double* C, V3 v)
{
  return V3{(C[0] * v.x + C[1] * v.y) + C[2] * v.z, (C[3] * v.x + C[4] * v.y) + C[5] * v.z, (C[6] * v.x + C[7] * v.y) + C[8] * v.z};
}

}  // namespace
#include "camera_project.h"
namespace {
using namespace okvo_cam;
}  // namespace

extern "C" int okvo_prepare_landmarks(
    int n_lm, const double* hp_W, const double* quality, const int32_t* obs_begin, const int32_t* obs /* n_obs x 3 */,
    int n_cams, const double* T_WC_old /* [slot][cam][12] */, const uint8_t* const* desc_tab, const double* const* ray_tab, int D,
    const double* T_WC1 /* C(9) r(3) */, const double* T_CW1, int model, const double* intr, int width, int height,
    double repr_thr, int exclusive,
    int32_t* out_lm, double* out_proj, uint8_t* out_is3d, double* out_p_W, int32_t* out_desc_begin, uint8_t* pool,
    double* out_e_W, double* out_r_W, int32_t* out_kid, int32_t* n_rows)
{
  const double maxU = width + repr_thr, maxV = height + repr_thr;
  const double focalLength = intr[0] + intr[1];
  const V3 r_WC1 = V3{T_WC1[9], T_WC1[10], T_WC1[11]};
  const size_t numDescriptorsToKeep = 3;
  uint8_t* dataPtr = pool;
  int n_out = 0;
  out_desc_begin[0] = 0;
  for (int it = 0; it < n_lm; ++it) {
    bool is3d = false;
    const double* hp = hp_W + 4 * it;
    const V3 p_W = V3{hp[0] / hp[3], hp[1] / hp[3], hp[2] / hp[3]};
    const V3 r_W = sub(p_W, r_WC1);
    const V3 e_W = normalized(r_W);
    const double r = std::max(0.01, norm(r_W));
    // hp_C = T_CW1 * hp_W (Transformation::operator*(Vector4d), implementation/Transformation.hpp:271-278)
    const V3 ch = rot(T_CW1, V3{hp[0], hp[1], hp[2]});
    const double s = hp[3];
    V3 head = V3{ch.x + T_CW1[9] * s, ch.y + T_CW1[10] * s, ch.z + T_CW1[11] * s};
    if (s < 0) head = V3{-head.x, -head.y, -head.z};   // projectHomogeneous (PinholeCamera.hpp:493-502)
    double kp[2] = {0, 0};
    const Status status = project(model, intr, width, height, head, kp);
    if (status == Invalid || status == Behind) continue;
    if (kp[0] < -repr_thr) continue;
    if (kp[1] < -repr_thr) continue;
    if (kp[0] > maxU) continue;
    if (kp[1] > maxV) continue;
    const double q = quality[it];
    std::vector<double> bestScores(numDescriptorsToKeep, 1.0);
    double e_cols[3][3], r_cols[3][3]; int32_t kid_rows[3][3];
    size_t o = 0;
    for (int oi = obs_begin[it + 1] - 1; oi >= obs_begin[it]; --oi) {   // observations.rbegin() .. rend()
      const int32_t* kid = obs + 3 * oi;
      const double* T_old = T_WC_old + ((size_t)kid[0] * n_cams + kid[1]) * 12;
      const V3 r_old_cam = V3{T_old[9], T_old[10], T_old[11]};
      const V3 r_W_old = sub(p_W, r_old_cam);
      if (!is3d) {
        const V3 r_close_W = sub(r_W, scale(0.2 / focalLength / q, r_W_old));
        const double cosA = dot(normalized(r_W), normalized(r_close_W));
        if (cosA > cos(10.0 / focalLength)) is3d = true;
      }
      const double cosViewpointChange = dot(e_W, normalized(r_W_old));
      if (cosViewpointChange < cos(0.6) && !exclusive) continue;
      const double scaleChange = fabs(r - norm(r_W_old)) / r;
      if ((scaleChange > 0.5) && !exclusive) continue;
      const double score = 0.5 * (acos(cosViewpointChange) / 0.6 + scaleChange / 0.5);
      double worstScore = 0.0;
      size_t worstIdx = 0;
      for (size_t n = 0; n < numDescriptorsToKeep; ++n)
        if (bestScores[n] > worstScore) { worstScore = bestScores[n]; worstIdx = n; }
      if (score < bestScores[worstIdx]) {
        const size_t tab = (size_t)kid[0] * n_cams + kid[1];
        memcpy(dataPtr + (size_t)D * o, desc_tab[tab] + (size_t)D * kid[2], D);
        const double* ec = ray_tab[tab] + 3 * (size_t)kid[2];
        const V3 e = rot(T_old, normalized(V3{ec[0], ec[1], ec[2]}));
        e_cols[o][0] = e.x; e_cols[o][1] = e.y; e_cols[o][2] = e.z;
        r_cols[o][0] = r_old_cam.x; r_cols[o][1] = r_old_cam.y; r_cols[o][2] = r_old_cam.z;
        kid_rows[o][0] = kid[0]; kid_rows[o][1] = kid[1]; kid_rows[o][2] = kid[2];
        o = std::max(o, worstIdx);
        bestScores[worstIdx] = score;
      }
    }
    const int row0 = (int)((dataPtr - pool) / D);
    for (size_t j = 0; j < o; ++j) {
      memcpy(out_e_W + 3 * (row0 + j), e_cols[j], 24); memcpy(out_r_W + 3 * (row0 + j), r_cols[j], 24);
      memcpy(out_kid + 3 * (row0 + j), kid_rows[j], 12);
    }
    dataPtr += o * D;
    if (o == 0) continue;
    out_lm[n_out] = it;
    out_proj[2 * n_out] = kp[0]; out_proj[2 * n_out + 1] = kp[1];
    out_is3d[n_out] = is3d ? 1 : 0;
    out_p_W[3 * n_out] = p_W.x; out_p_W[3 * n_out + 1] = p_W.y; out_p_W[3 * n_out + 2] = p_W.z;
    n_out++;
    out_desc_begin[n_out] = (int32_t)((dataPtr - pool) / D);
  }
  *n_rows = (int)((dataPtr - pool) / D);
  return n_out;
}
