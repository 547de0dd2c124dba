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
#include "okb_camdev.h"

namespace okb {

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

void k_backproject_ext(const Model& m, const okb_keypoint_t* d_kp, const int32_t* d_count, int cap, int n_frames, double* d_rays,
                       uint8_t* d_valid, cudaStream_t st)
{
  k_backproject<<<dim3((cap + 127) / 128, n_frames), 128, 0, st>>>(m, d_kp, d_count, cap, 0, d_rays, d_valid);
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
