/* oracle/camera_project.h -- TEST INFRASTRUCTURE ONLY. PinholeCamera<D>::project restated once for the oracles that need it
 * (P1: prepare_oracle.cpp, M3 re-projection check: match_oracle.cpp). Reference: okvis_cv/include/okvis/cameras/implementation/
 * PinholeCamera.hpp:257-292 (project), :493-502 (projectHomogeneous); RadialTangentialDistortion.hpp:90-109,
 * EquidistantDistortion.hpp:86-106 (distort). */
#pragma once
#include <math.h>
namespace okvo_cam {
struct P3 { double x, y, z; };
enum Status { Successful, OutsideImage, Masked, Behind, Invalid };

// PinholeCamera<D>::project (PinholeCamera.hpp:257-292); no mask
template <class V> inline Status project(int model, const double* in /*fu fv cu cv k0..k3*/, int width, int height, V p, double* kp)
{
  if (fabs(p.z) < 1.0e-12) return Invalid;
  const double rz = 1.0 / p.z;
  const double u0 = p.x * rz, u1 = p.y * rz;
  double d0, d1;
  if (model == 1) {
    const double k1_ = in[4], k2_ = in[5], p1_ = in[6], p2_ = in[7];
    const double mx_u = u0 * u0, my_u = u1 * u1, mxy_u = u0 * u1;
    const double rho_u = mx_u + my_u;
    const double rad_dist_u = k1_ * rho_u + k2_ * rho_u * rho_u;
    d0 = u0 + u0 * rad_dist_u + 2.0 * p1_ * mxy_u + p2_ * (rho_u + 2.0 * mx_u);
    d1 = u1 + u1 * rad_dist_u + 2.0 * p2_ * mxy_u + p1_ * (rho_u + 2.0 * my_u);
  } else if (model == 2) {
    const double k1_ = in[4], k2_ = in[5], k3_ = in[6], k4_ = in[7];
    const double r = sqrt(u0 * u0 + u1 * u1);
    const double theta = atan(r);
    const double theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
    const double thetad = theta * (1.0 + k1_ * theta2 + k2_ * theta4 + k3_ * theta6 + k4_ * theta8);
    const double scaling = (r > 1e-8) ? thetad / r : 1.0;
    d0 = scaling * u0; d1 = scaling * u1;
  } else { d0 = u0; d1 = u1; }
  kp[0] = in[0] * d0 + in[2];
  kp[1] = in[1] * d1 + in[3];
  if (kp[0] < 0.0 || kp[1] < 0.0 || kp[0] >= width || kp[1] >= height) return OutsideImage;
  return p.z > 0.0 ? Successful : Behind;
}
}  // namespace okvo_cam
