/*
 * oracle/backproject_oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle of D4; never on the product path).
 *
 * Restatement of okvis::Frame::computeBackProjections (reference okvis_cv/include/okvis/implementation/Frame.hpp:178-193):
 *   PinholeCamera<D>::backProject       okvis_cv/include/okvis/cameras/implementation/PinholeCamera.hpp:574-592
 *   RadialTangentialDistortion::distort  .../implementation/RadialTangentialDistortion.hpp:110-136 (with Jacobian)
 *   RadialTangentialDistortion::undistort .../RadialTangentialDistortion.hpp:214-253 (Gauss-Newton, 5 iterations)
 *   EquidistantDistortion::distort       .../implementation/EquidistantDistortion.hpp:105-171
 *   EquidistantDistortion::undistort     .../EquidistantDistortion.hpp:319-352 (20 iterations)
 * The 2x2 Eigen expressions (E^T E).inverse() * E^T * e are written out (products as 2-term sums, inverse by the
 * adjugate times 1/det); whether the reference's Eigen build rounds identically at the last ulp is unpinned.
 */
#include <math.h>
#include <stdint.h>

namespace {
struct Cam { int model; double fu, fv, cu, cv, k[4]; };

void distort_rt(const Cam& c, double u0, double u1, double* d, double J[2][2])
{
  const double k1_ = c.k[0], k2_ = c.k[1], p1_ = c.k[2], p2_ = c.k[3];
  const double mx_u = u0 * u0, my_u = u1 * u1, mxy_u = u0 * u1;
  const double rho_u = mx_u + my_u;
  const double rad_dist_u = k1_ * rho_u + k2_ * rho_u * rho_u;
  d[0] = u0 + u0 * rad_dist_u + 2.0 * p1_ * mxy_u + p2_ * (rho_u + 2.0 * mx_u);
  d[1] = u1 + u1 * rad_dist_u + 2.0 * p2_ * mxy_u + p1_ * (rho_u + 2.0 * my_u);
  J[0][0] = 1 + rad_dist_u + k1_ * 2.0 * mx_u + k2_ * rho_u * 4 * mx_u + 2.0 * p1_ * u1 + 6 * p2_ * u0;
  J[1][0] = k1_ * 2.0 * u0 * u1 + k2_ * 4 * rho_u * u0 * u1 + p1_ * 2.0 * u0 + 2.0 * p2_ * u1;
  J[0][1] = J[1][0];
  J[1][1] = 1 + rad_dist_u + k1_ * 2.0 * my_u + k2_ * rho_u * 4 * my_u + 6 * p1_ * u1 + 2.0 * p2_ * u0;
}

void distort_eq(const Cam& c, double u0, double u1, double* d, double J[2][2])
{
  const double k1_ = c.k[0], k2_ = c.k[1], k3_ = c.k[2], k4_ = c.k[3];
  const double r = sqrt(u0 * u0 + u1 * u1);
  const double theta = atan(r);
  const double theta2 = theta * theta;
  const double theta4 = theta2 * theta2;
  const double theta6 = theta4 * theta2;
  const double theta8 = theta4 * theta4;
  const double thetad = theta * (1.0 + k1_ * theta2 + k2_ * theta4 + k3_ * theta6 + k4_ * theta8);
  const double scaling = (r > 1e-8) ? thetad / r : 1.0;
  d[0] = scaling * u0; d[1] = scaling * u1;
  if (r > 1e-8) {
    double t2, t3, t4, t6, t7, t8, t9, t11, t17, t18, t19, t20, t25;
    t2 = u0 * u0; t3 = u1 * u1; t4 = t2 + t3;
    t6 = atan(sqrt(t4)); t7 = t6 * t6; t8 = 1.0 / sqrt(t4); t9 = t7 * t7;
    t11 = 1.0 / ((t2 + t3) + 1.0);
    t17 = (((k1_ * t7 + k2_ * t9) + k3_ * t7 * t9) + k4_ * (t9 * t9)) + 1.0;
    t18 = 1.0 / t4; t19 = 1.0 / sqrt(t4 * t4 * t4); t20 = t6 * t8 * t17;
    t25 = ((k2_ * t6 * t7 * t8 * t11 * u1 * 4.0 + k3_ * t6 * t8 * t9 * t11 * u1 * 6.0) + k4_ * t6 * t7 * t8 * t9 * t11 * u1 * 8.0) +
          k1_ * t6 * t8 * t11 * u1 * 2.0;
    t4 = ((k2_ * t6 * t7 * t8 * t11 * u0 * 4.0 + k3_ * t6 * t8 * t9 * t11 * u0 * 6.0) + k4_ * t6 * t7 * t8 * t9 * t11 * u0 * 8.0) +
         k1_ * t6 * t8 * t11 * u0 * 2.0;
    t7 = t11 * t17 * t18 * u0 * u1;
    J[0][1] = (t7 + t6 * t8 * t25 * u0) - t6 * t17 * t19 * u0 * u1;
    J[1][1] = ((t20 - t3 * t6 * t17 * t19) + t3 * t11 * t17 * t18) + t6 * t8 * t25 * u1;
    J[0][0] = ((t20 - t2 * t6 * t17 * t19) + t2 * t11 * t17 * t18) + t6 * t8 * t4 * u0;
    J[1][0] = (t7 + t6 * t8 * t4 * u1) - t6 * t17 * t19 * u0 * u1;
  } else { J[0][0] = 1; J[0][1] = 0; J[1][0] = 0; J[1][1] = 1; }
}

bool undistort(const Cam& c, const double pd[2], double out[2])
{
  double x_bar[2] = {pd[0], pd[1]};
  const int n = c.model == 1 ? 5 : 20;
  bool success = false;
  for (int i = 0; i < n; i++) {
    double x_tmp[2], E[2][2];
    if (c.model == 1) distort_rt(c, x_bar[0], x_bar[1], x_tmp, E); else distort_eq(c, x_bar[0], x_bar[1], x_tmp, E);
    const double e[2] = {pd[0] - x_tmp[0], pd[1] - x_tmp[1]};
    double E2[2][2];  // E^T * E
    for (int r = 0; r < 2; r++) for (int cc = 0; cc < 2; cc++) E2[r][cc] = E[0][r] * E[0][cc] + E[1][r] * E[1][cc];
    const double invdet = 1.0 / (E2[0][0] * E2[1][1] - E2[1][0] * E2[0][1]);
    const double I[2][2] = {{E2[1][1] * invdet, -E2[0][1] * invdet}, {-E2[1][0] * invdet, E2[0][0] * invdet}};
    double M[2][2];   // I * E^T
    for (int r = 0; r < 2; r++) for (int cc = 0; cc < 2; cc++) M[r][cc] = I[r][0] * E[cc][0] + I[r][1] * E[cc][1];
    x_bar[0] += M[0][0] * e[0] + M[0][1] * e[1];
    x_bar[1] += M[1][0] * e[0] + M[1][1] * e[1];
    const double chi2 = e[0] * e[0] + e[1] * e[1];
    if (chi2 < 1e-6) success = true;
    if (chi2 < 1e-15) { success = true; break; }
  }
  out[0] = x_bar[0]; out[1] = x_bar[1];
  return success;
}
}  // namespace

extern "C" void okvo_back_project(int model, double fu, double fv, double cu, double cv, const double k[4], int n,
                                  const float* kp_xy /* stride 7 floats: cv::KeyPoint records */, int stride_floats,
                                  double* rays, uint8_t* valid)
{
  Cam c; c.model = model; c.fu = fu; c.fv = fv; c.cu = cu; c.cv = cv; for (int i = 0; i < 4; i++) c.k[i] = k[i];
  const double one_over_fu = 1.0 / fu, one_over_fv = 1.0 / fv;
  for (int i = 0; i < n; i++) {
    const double px = (double)kp_xy[(size_t)i * stride_floats], py = (double)kp_xy[(size_t)i * stride_floats + 1];
    const double p2[2] = {(px - cu) * one_over_fu, (py - cv) * one_over_fv};
    double u[2] = {p2[0], p2[1]};
    bool ok = true;
    if (model != 0) ok = undistort(c, p2, u);
    rays[3 * i] = u[0]; rays[3 * i + 1] = u[1]; rays[3 * i + 2] = 1.0;
    valid[i] = ok;
  }
}

// ---- D5 and computeOverlaps: per-pixel uses of the same camera model -------------------------------------------------
namespace {
bool back_project(const Cam& c, double px, double py, double ray[3])
{
  const double one_over_fu = 1.0 / c.fu, one_over_fv = 1.0 / c.fv;   // PinholeCamera keeps the reciprocals as members
  const double p2[2] = {(px - c.cu) * one_over_fu, (py - c.cv) * one_over_fv};
  double u[2] = {p2[0], p2[1]};
  bool ok = true;
  if (c.model != 0) ok = undistort(c, p2, u);
  ray[0] = u[0]; ray[1] = u[1]; ray[2] = 1.0;
  return ok;
}

enum { kSuccessful = 0, kOutside = 1, kBehind = 3, kInvalid = 4 };
// PinholeCamera::project with the point Jacobian (PinholeCamera.hpp:294-372)
int project_jac(const Cam& c, int width, int height, const double p[3], double kp[2], double J[2][3])
{
  if (fabs(p[2]) < 1.0e-12) return kInvalid;
  const double rz = 1.0 / p[2];
  const double rz2 = rz * rz;
  const double u0 = p[0] * rz, u1 = p[1] * rz;
  double d[2], D[2][2];
  if (c.model == 1) distort_rt(c, u0, u1, d, D);
  else if (c.model == 2) distort_eq(c, u0, u1, d, D);
  else { d[0] = u0; d[1] = u1; D[0][0] = 1; D[0][1] = 0; D[1][0] = 0; D[1][1] = 1; }
  J[0][0] = c.fu * D[0][0] * rz;
  J[0][1] = c.fu * D[0][1] * rz;
  J[0][2] = -c.fu * (p[0] * D[0][0] + p[1] * D[0][1]) * rz2;
  J[1][0] = c.fv * D[1][0] * rz;
  J[1][1] = c.fv * D[1][1] * rz;
  J[1][2] = -c.fv * (p[0] * D[1][0] + p[1] * D[1][1]) * rz2;
  kp[0] = c.fu * d[0] + c.cu;
  kp[1] = c.fv * d[1] + c.cv;
  if (kp[0] < 0.0 || kp[1] < 0.0 || kp[0] >= width || kp[1] >= height) return kOutside;
  return p[2] > 0.0 ? kSuccessful : kBehind;
}
Cam make_cam(int model, const double* in) { Cam c; c.model = model; c.fu = in[0]; c.fv = in[1]; c.cu = in[2]; c.cv = in[3]; for (int i = 0; i < 4; i++) c.k[i] = in[4 + i]; return c; }
}  // namespace

// PinholeCamera::initialiseCameraAwarenessMaps (PinholeCamera.hpp:179-208): rays (H x W x 3 floats: the normalised
// back-projection of every pixel, zero where it fails) and imageJacobians (H x W x 6 floats: the 2 x 3 projection Jacobian at
// that ray, row-major; the reference leaves entries of non-Successful projections uninitialised -- here they are zero).
extern "C" void okvo_camera_awareness_maps(int model, const double* intr, int width, int height, float* rays, float* jac)
{
  const Cam c = make_cam(model, intr);
  for (int v = 0; v < height; v++)
    for (int u = 0; u < width; u++) {
      double ray[3];
      if (back_project(c, (double)u, (double)v, ray)) {
        const double n = sqrt((ray[0] * ray[0] + ray[1] * ray[1]) + ray[2] * ray[2]);
        ray[0] /= n; ray[1] /= n; ray[2] /= n;
      } else { ray[0] = ray[1] = ray[2] = 0.0; }
      const size_t i = (size_t)v * width + u;
      rays[3 * i] = (float)ray[0]; rays[3 * i + 1] = (float)ray[1]; rays[3 * i + 2] = (float)ray[2];
      double pt[2], J[2][3];
      for (int k = 0; k < 6; k++) jac[6 * i + k] = 0.f;
      if (project_jac(c, width, height, ray, pt, J) == kSuccessful)
        for (int r = 0; r < 2; r++) for (int k = 0; k < 3; k++) jac[6 * i + 3 * r + k] = (float)J[r][k];
    }
}

// NCameraSystem::computeOverlaps (okvis_cv/src/NCameraSystem.cpp:48-118). models / intr / widths / heights per camera;
// C_rel[(seenBy * n + cam) * 9]: (T_SC[seenBy]->inverse() * *T_SC[cam]).C(), row-major. overlaps: n x n booleans
// (overlaps_[seenBy][cam]); mats (may be NULL): per (seenBy, cam) an offset into `mat_data`, height x width bytes of camera `cam`.
extern "C" void okvo_compute_overlaps(int n, const int32_t* models, const double* intr /* n x 8 */, const int32_t* widths, const int32_t* heights,
                                      const double* C_rel, uint8_t* overlaps, const int64_t* mat_offsets, uint8_t* mat_data)
{
  for (int seenBy = 0; seenBy < n; seenBy++)
    for (int cam = 0; cam < n; cam++) {
      const int W = widths[cam], H = heights[cam];
      uint8_t* mat = mat_offsets ? mat_data + mat_offsets[seenBy * n + cam] : nullptr;
      if (cam == seenBy) { overlaps[seenBy * n + cam] = 1; if (mat) for (size_t i = 0; i < (size_t)W * H; i++) mat[i] = 1; continue; }
      overlaps[seenBy * n + cam] = 0;
      if (mat) for (size_t i = 0; i < (size_t)W * H; i++) mat[i] = 0;
      const Cam c = make_cam(models[cam], intr + 8 * cam), o = make_cam(models[seenBy], intr + 8 * seenBy);
      const double* C = C_rel + 9 * (size_t)(seenBy * n + cam);
      for (int u = 0; u < W; u++)
        for (int v = 0; v < H; v++) {
          double ray[3];
          back_project(c, (double)u, (double)v, ray);   // the success flag is ignored, as in the reference
          const double ro[3] = {(C[0] * ray[0] + C[1] * ray[1]) + C[2] * ray[2], (C[3] * ray[0] + C[4] * ray[1]) + C[5] * ray[2],
                                (C[6] * ray[0] + C[7] * ray[1]) + C[8] * ray[2]};
          double pt[2], J[2][3];
          if (project_jac(o, widths[seenBy], heights[seenBy], ro, pt, J) != kSuccessful) continue;
          double ver[3];
          back_project(o, pt[0], pt[1], ver);
          const double n0 = sqrt((ro[0] * ro[0] + ro[1] * ro[1]) + ro[2] * ro[2]), n1 = sqrt((ver[0] * ver[0] + ver[1] * ver[1]) + ver[2] * ver[2]);
          const double a[3] = {ro[0] / n0, ro[1] / n0, ro[2] / n0}, b[3] = {ver[0] / n1, ver[1] / n1, ver[2] / n1};
          if (fabs(((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) - 1.0) < 1.0e-10) {
            if (mat) mat[(size_t)v * W + u] = 1;
            overlaps[seenBy * n + cam] = 1;
          }
        }
    }
}
