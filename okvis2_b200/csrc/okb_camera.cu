// okb_camera.cu -- D4: back-projection of keypoints to unit-z rays, on the device.
//
// Replaces okvis::Frame::computeBackProjections (reference okvis_cv/include/okvis/implementation/Frame.hpp:178-193) ->
// PinholeCamera<D>::backProject (okvis_cv/include/okvis/cameras/implementation/PinholeCamera.hpp:574-592) ->
// Distortion::undistort, a Gauss-Newton solve on the distortion model:
//   RadialTangentialDistortion (implementation/RadialTangentialDistortion.hpp:90-136 distort+Jacobian, :214-253 undistort,
//                               5 iterations) -- pure fp64 arithmetic, reproduced exactly (no FMA contraction);
//   EquidistantDistortion      (implementation/EquidistantDistortion.hpp:105-171, :319-352, 20 iterations) -- needs atan:
//                               the device atan is within 1 ulp of libm, so rays can differ in the last bits (documented);
//   no distortion.
// Also the world-frame preparation of the stereo matcher inputs for the device-resident pipeline.
#include <math.h>
#include <string.h>

#include "okb_internal.h"
#include "okb_gatecos.h"

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
__device__ bool back_project(const Model& m, double px, double py, double& rx, double& ry)
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

// kp: [frames][cap] records, count: [frames]; rays: [frames][cap][3] doubles, valid: [frames][cap]
__global__ void __launch_bounds__(128) k_backproject(Model m, const okb_keypoint_t* kp, const int32_t* count, int cap, int n_fixed,
                                                     double* rays, uint8_t* valid)
{
  const int frame = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = count ? min(count[frame], cap) : n_fixed;
  if (k >= n) return;
  const size_t i = (size_t)frame * cap + k;
  double rx, ry;
  const bool ok = back_project(m, (double)kp[i].x, (double)kp[i].y, rx, ry);
  rays[3 * i] = rx; rays[3 * i + 1] = ry; rays[3 * i + 2] = 1.0;
  valid[i] = ok ? 1 : 0;
}

// world-frame inputs of the stereo matcher: e_W = (C_WC * e_C).normalized(), size/f, cos(2.6 sigma), cos(6 sigma)
struct PrepArgs { double C[9]; double f; };
__global__ void __launch_bounds__(128) k_stereo_prep(PrepArgs p, const okb_keypoint_t* kp, const double* rays, const int32_t* count,
                                                     int cap, double* e_W, double* sof, double* c26, double* c6)
{
  const int frame = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= min(count[frame], cap)) return;
  const size_t i = (size_t)frame * cap + k;
  const double x = rays[3 * i], y = rays[3 * i + 1], z = rays[3 * i + 2];
  const double wx = (p.C[0] * x + p.C[1] * y) + p.C[2] * z;
  const double wy = (p.C[3] * x + p.C[4] * y) + p.C[5] * z;
  const double wz = (p.C[6] * x + p.C[7] * y) + p.C[8] * z;
  const double n2 = (wx * wx + wy * wy) + wz * wz;
  double ex = wx, ey = wy, ez = wz;
  if (n2 > 0.0) { const double n = sqrt(n2); ex = wx / n; ey = wy / n; ez = wz / n; }
  e_W[3 * i] = ex; e_W[3 * i + 1] = ey; e_W[3 * i + 2] = ez;
  const double s = (double)kp[i].size / p.f;
  sof[i] = s;
  const double sigma = s * 0.125;
  c26[i] = gate_cos(2.6 * sigma); c6[i] = gate_cos(6.0 * sigma);   // == host libm, okb_gatecos.h
}

static Model to_model(const okb_camera_model_t& c)
{
  Model m; memset(&m, 0, sizeof(m));
  m.model = c.model; m.fu = c.fu; m.fv = c.fv; m.cu = c.cu; m.cv = c.cv;
  m.one_over_fu = 1.0 / c.fu; m.one_over_fv = 1.0 / c.fv;   // PinholeCamera keeps the reciprocals as members
  for (int i = 0; i < 4; i++) m.k[i] = c.k[i];
  return m;
}

int camera_backproject_batch(okb_context* ctx, int cam, int n_frames)
{
  CamWorkspace& ws = ctx->cams[cam];
  if (!ws.has_model) return OKB_OK;
  k_backproject<<<dim3((ws.kp_cap + 127) / 128, n_frames), 128, 0, ws.stream>>>(to_model(ws.model), ws.d_kp, ws.d_count, ws.kp_cap, 0,
                                                                               ws.d_rays, ws.d_rays_valid);
  ctx->launches++;
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

int camera_stereo_prep(okb_context* ctx, const okb_camera_model_t& model, const double C_WC[9], const okb_keypoint_t* d_kp,
                       const int32_t* d_count, int cap, int n_frames, double* d_rays, uint8_t* d_valid, double* d_eW, double* d_sof,
                       double* d_c26, double* d_c6, cudaStream_t st)
{
  k_backproject<<<dim3((cap + 127) / 128, n_frames), 128, 0, st>>>(to_model(model), d_kp, d_count, cap, 0, d_rays, d_valid);
  PrepArgs p; for (int i = 0; i < 9; i++) p.C[i] = C_WC[i];
  p.f = 0.5 * (model.fu + model.fv);
  k_stereo_prep<<<dim3((cap + 127) / 128, n_frames), 128, 0, st>>>(p, d_kp, d_rays, d_count, cap, d_eW, d_sof, d_c26, d_c6);
  ctx->launches += 2;
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

}  // namespace okb

using namespace okb;

extern "C" {

int okb_set_camera_model(okb_context_t* ctx, int cam, const okb_camera_model_t* model)
{
  if (!ctx || cam < 0 || cam >= ctx->n_cams || !model || model->model < 0 || model->model > 2 || !(model->fu > 0) || !(model->fv > 0)) {
    set_error("okb_set_camera_model: bad arguments"); return OKB_ERR_ARGUMENT;
  }
  ctx->cams[cam].model = *model; ctx->cams[cam].has_model = 1;
  return OKB_OK;
}

int okb_back_project(okb_context_t* ctx, int cam, int n, const okb_keypoint_t* kp, double* rays_out, uint8_t* valid_out)
{
  if (!ctx || cam < 0 || cam >= ctx->n_cams || n < 0 || (n > 0 && (!kp || !rays_out || !valid_out))) {
    set_error("okb_back_project: bad arguments"); return OKB_ERR_ARGUMENT;
  }
  CamWorkspace& ws = ctx->cams[cam];
  if (!ws.has_model) { set_error("okb_back_project: camera %d has no model (okb_set_camera_model)", cam); return OKB_ERR_ARGUMENT; }
  if (n == 0) return OKB_OK;
  OKB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ws.stream;
  okb_keypoint_t* d_kp = nullptr; double* d_r = nullptr; uint8_t* d_v = nullptr;
  OKB_CUDA(cudaMallocAsync(&d_kp, (size_t)n * sizeof(okb_keypoint_t), st));
  OKB_CUDA(cudaMallocAsync(&d_r, (size_t)n * 24, st));
  OKB_CUDA(cudaMallocAsync(&d_v, (size_t)n, st));
  OKB_CUDA(cudaMemcpyAsync(d_kp, kp, (size_t)n * sizeof(okb_keypoint_t), cudaMemcpyHostToDevice, st));
  k_backproject<<<dim3((n + 127) / 128, 1), 128, 0, st>>>(to_model(ws.model), d_kp, nullptr, n, n, d_r, d_v);
  ctx->launches++;
  OKB_CUDA(cudaGetLastError());
  OKB_CUDA(cudaMemcpyAsync(rays_out, d_r, (size_t)n * 24, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(cudaMemcpyAsync(valid_out, d_v, (size_t)n, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(cudaFreeAsync(d_kp, st)); OKB_CUDA(cudaFreeAsync(d_r, st)); OKB_CUDA(cudaFreeAsync(d_v, st));
  OKB_CUDA(cudaStreamSynchronize(st));
  return OKB_OK;
}

}  // extern "C"
