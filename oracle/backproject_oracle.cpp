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
