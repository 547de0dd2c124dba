// okb_camdev.h -- pinhole camera model on the device: distortion, back-projection (D4), projection.
//
// Restates okvis::cameras::PinholeCamera<D> (reference okvis_cv/include/okvis/cameras/implementation/PinholeCamera.hpp:
// project :257-292, projectHomogeneous :493-502, backProject :574-592) with RadialTangentialDistortion
// (implementation/RadialTangentialDistortion.hpp:90-136 distort + Jacobian, :214-253 undistort, 5 Gauss-Newton iterations),
// EquidistantDistortion (implementation/EquidistantDistortion.hpp:86-171, :319-352, 20 iterations) and NoDistortion.
// Pure fp64 with the expression order of the reference, no FMA contraction (-fmad=false). The equidistant model needs atan:
// the device atan is within 1 ulp of libm, so its results can differ in the last bits (documented in DESIGN.md).
#pragma once
#include <math.h>
#include <string.h>

#include "../../include/okvis_b200.h"

namespace okb {

struct Model { int model; double fu, fv, cu, cv, one_over_fu, one_over_fv, k[4]; };

__device__ __forceinline__ void distort_radtan(const Model& m, double u0, double u1, double& d0, double& d1, double (&J)[2][2])
{
  const double k1 = m.k[0], k2 = m.k[1], p1 = m.k[2], p2 = m.k[3];
  const double mx_u = u0 * u0, my_u = u1 * u1, mxy_u = u0 * u1;
  const double rho_u = mx_u + my_u;
  const double rad_dist_u = k1 * rho_u + k2 * rho_u * rho_u;
  d0 = u0 + u0 * rad_dist_u + 2.0 * p1 * mxy_u + p2 * (rho_u + 2.0 * mx_u);
  d1 = u1 + u1 * rad_dist_u + 2.0 * p2 * mxy_u + p1 * (rho_u + 2.0 * my_u);
  J[0][0] = 1 + rad_dist_u + k1 * 2.0 * mx_u + k2 * rho_u * 4 * mx_u + 2.0 * p1 * u1 + 6 * p2 * u0;
  J[1][0] = k1 * 2.0 * u0 * u1 + k2 * 4 * rho_u * u0 * u1 + p1 * 2.0 * u0 + 2.0 * p2 * u1;
  J[0][1] = J[1][0];
  J[1][1] = 1 + rad_dist_u + k1 * 2.0 * my_u + k2 * rho_u * 4 * my_u + 6 * p1 * u1 + 2.0 * p2 * u0;
}

__device__ __forceinline__ void distort_equi(const Model& m, double u0, double u1, double& d0, double& d1, double (&J)[2][2])
{
  const double k1 = m.k[0], k2 = m.k[1], k3 = m.k[2], k4 = m.k[3];
  const double r = sqrt(u0 * u0 + u1 * u1);
  const double theta = atan(r);
  const double theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
  const double thetad = theta * (1.0 + k1 * theta2 + k2 * theta4 + k3 * theta6 + k4 * theta8);
  const double scaling = (r > 1e-8) ? thetad / r : 1.0;
  d0 = scaling * u0; d1 = scaling * u1;
  if (r > 1e-8) {
    double t2 = u0 * u0, t3 = u1 * u1, t4 = t2 + t3;
    const double t6 = atan(sqrt(t4));
    double t7 = t6 * t6;
    const double t8 = 1.0 / sqrt(t4);
    const double t9 = t7 * t7;
    const double t11 = 1.0 / ((t2 + t3) + 1.0);
    const double t17 = (((k1 * t7 + k2 * t9) + k3 * t7 * t9) + k4 * (t9 * t9)) + 1.0;
    const double t18 = 1.0 / t4;
    const double t19 = 1.0 / sqrt(t4 * t4 * t4);
    const double t20 = t6 * t8 * t17;
    const double t25 = ((k2 * t6 * t7 * t8 * t11 * u1 * 4.0 + k3 * t6 * t8 * t9 * t11 * u1 * 6.0) + k4 * t6 * t7 * t8 * t9 * t11 * u1 * 8.0) +
                       k1 * t6 * t8 * t11 * u1 * 2.0;
    t4 = ((k2 * t6 * t7 * t8 * t11 * u0 * 4.0 + k3 * t6 * t8 * t9 * t11 * u0 * 6.0) + k4 * t6 * t7 * t8 * t9 * t11 * u0 * 8.0) +
         k1 * t6 * t8 * t11 * u0 * 2.0;
    t7 = t11 * t17 * t18 * u0 * u1;
    J[0][1] = (t7 + t6 * t8 * t25 * u0) - t6 * t17 * t19 * u0 * u1;
    J[1][1] = ((t20 - t3 * t6 * t17 * t19) + t3 * t11 * t17 * t18) + t6 * t8 * t25 * u1;
    J[0][0] = ((t20 - t2 * t6 * t17 * t19) + t2 * t11 * t17 * t18) + t6 * t8 * t4 * u0;
    J[1][0] = (t7 + t6 * t8 * t4 * u1) - t6 * t17 * t19 * u0 * u1;
  } else {
    J[0][0] = 1.0; J[0][1] = 0.0; J[1][0] = 0.0; J[1][1] = 1.0;
  }
}

// PinholeCamera::backProject: returns success, ray = (x, y, 1)
__device__ inline bool back_project(const Model& m, double px, double py, double& rx, double& ry)
{
  const double q0 = (px - m.cu) * m.one_over_fu, q1 = (py - m.cv) * m.one_over_fv;
  if (m.model == 0) { rx = q0; ry = q1; return true; }
  double x0 = q0, x1 = q1;
  const int n = m.model == 1 ? 5 : 20;
  bool success = false;
  for (int i = 0; i < n; i++) {
    double t0, t1, E[2][2];
    if (m.model == 1) distort_radtan(m, x0, x1, t0, t1, E); else distort_equi(m, x0, x1, t0, t1, E);
    const double e0 = q0 - t0, e1 = q1 - t1;
    // du = (E^T E)^-1 * E^T * e, evaluated as ((E2^-1 * E^T) * e) with 2-term sums left to right
    const double a = E[0][0] * E[0][0] + E[1][0] * E[1][0], b = E[0][0] * E[0][1] + E[1][0] * E[1][1];
    const double c = E[0][1] * E[0][0] + E[1][1] * E[1][0], d = E[0][1] * E[0][1] + E[1][1] * E[1][1];
    const double invdet = 1.0 / (a * d - c * b);
    const double i00 = d * invdet, i10 = -c * invdet, i01 = -b * invdet, i11 = a * invdet;
    const double m00 = i00 * E[0][0] + i01 * E[0][1], m01 = i00 * E[1][0] + i01 * E[1][1];
    const double m10 = i10 * E[0][0] + i11 * E[0][1], m11 = i10 * E[1][0] + i11 * E[1][1];
    x0 += m00 * e0 + m01 * e1;
    x1 += m10 * e0 + m11 * e1;
    const double chi2 = e0 * e0 + e1 * e1;
    if (chi2 < 1e-6) success = true;
    if (chi2 < 1e-15) { success = true; break; }
  }
  rx = x0; ry = x1;
  return success;
}


// PinholeCamera::project: 0 Successful, 1 OutsideImage, 3 Behind, 4 Invalid (no mask support: 2 Masked never occurs)
enum { kProjSuccessful = 0, kProjOutside = 1, kProjMasked = 2, kProjBehind = 3, kProjInvalid = 4 };
__device__ __forceinline__ int project(const Model& m, int width, int height, double px, double py, double pz, double& kx, double& ky)
{
  if (fabs(pz) < 1.0e-12) return kProjInvalid;
  const double rz = 1.0 / pz;
  const double u0 = px * rz, u1 = py * rz;
  double d0, d1;
  if (m.model == 1) {
    const double k1 = m.k[0], k2 = m.k[1], p1 = m.k[2], p2 = m.k[3];
    const double mx_u = u0 * u0, my_u = u1 * u1, mxy_u = u0 * u1;
    const double rho_u = mx_u + my_u;
    const double rad_dist_u = k1 * rho_u + k2 * rho_u * rho_u;
    d0 = u0 + u0 * rad_dist_u + 2.0 * p1 * mxy_u + p2 * (rho_u + 2.0 * mx_u);
    d1 = u1 + u1 * rad_dist_u + 2.0 * p2 * mxy_u + p1 * (rho_u + 2.0 * my_u);
  } else if (m.model == 2) {
    const double rr = sqrt(u0 * u0 + u1 * u1);
    const double theta = atan(rr);
    const double theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
    const double thetad = theta * (1.0 + m.k[0] * theta2 + m.k[1] * theta4 + m.k[2] * theta6 + m.k[3] * theta8);
    const double scaling = (rr > 1e-8) ? thetad / rr : 1.0;
    d0 = scaling * u0; d1 = scaling * u1;
  } else { d0 = u0; d1 = u1; }
  kx = m.fu * d0 + m.cu; ky = m.fv * d1 + m.cv;
  if (kx < 0.0 || ky < 0.0 || kx >= width || ky >= height) return kProjOutside;
  return pz > 0.0 ? kProjSuccessful : kProjBehind;
}

inline Model to_model(const okb_camera_model_t& c)
{
  Model m; memset(&m, 0, sizeof(m));
  m.model = c.model; m.fu = c.fu; m.fv = c.fv; m.cu = c.cu; m.cv = c.cv;
  m.one_over_fu = 1.0 / c.fu; m.one_over_fv = 1.0 / c.fv;   // PinholeCamera keeps the reciprocals as members
  for (int i = 0; i < 4; i++) m.k[i] = c.k[i];
  return m;
}

}  // namespace okb
