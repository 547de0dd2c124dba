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

#include <algorithm>

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


// D5: PinholeCamera::initialiseCameraAwarenessMaps (cameras/implementation/PinholeCamera.hpp:179-208), one thread per pixel
__device__ __forceinline__ int project_jac(const Model& m, int width, int height, double px, double py, double pz, double& kx, double& ky, double (&J)[2][3])
{
  if (fabs(pz) < 1.0e-12) return kProjInvalid;
  const double rz = 1.0 / pz;
  const double rz2 = rz * rz;
  const double u0 = px * rz, u1 = py * rz;
  double d0, d1, D[2][2];
  if (m.model == 1) distort_radtan(m, u0, u1, d0, d1, D);
  else if (m.model == 2) distort_equi(m, u0, u1, d0, d1, D);
  else { d0 = u0; d1 = u1; D[0][0] = 1; D[0][1] = 0; D[1][0] = 0; D[1][1] = 1; }
  J[0][0] = m.fu * D[0][0] * rz;
  J[0][1] = m.fu * D[0][1] * rz;
  J[0][2] = -m.fu * (px * D[0][0] + py * D[0][1]) * rz2;
  J[1][0] = m.fv * D[1][0] * rz;
  J[1][1] = m.fv * D[1][1] * rz;
  J[1][2] = -m.fv * (px * D[1][0] + py * D[1][1]) * rz2;
  kx = m.fu * d0 + m.cu; ky = m.fv * d1 + m.cv;
  if (kx < 0.0 || ky < 0.0 || kx >= width || ky >= height) return kProjOutside;
  return pz > 0.0 ? kProjSuccessful : kProjBehind;
}

__global__ void __launch_bounds__(128) k_awareness_maps(Model m, int width, int height, float* rays, float* jac)
{
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= width) return;
  double rx, ry, rz = 1.0;
  if (back_project(m, (double)u, (double)v, rx, ry)) {
    const double n = sqrt((rx * rx + ry * ry) + rz * rz);
    rx /= n; ry /= n; rz /= n;
  } else { rx = ry = rz = 0.0; }
  const size_t i = (size_t)v * width + u;
  rays[3 * i] = (float)rx; rays[3 * i + 1] = (float)ry; rays[3 * i + 2] = (float)rz;
  double kx, ky, J[2][3];
  const bool ok = project_jac(m, width, height, rx, ry, rz, kx, ky, J) == kProjSuccessful;
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) jac[6 * i + 3 * r + k] = ok ? (float)J[r][k] : 0.f;   // the reference leaves non-Successful entries uninitialised
}

// NCameraSystem::computeOverlaps (okvis_cv/src/NCameraSystem.cpp:48-118) for one (seenBy, cam) pair, one thread per pixel of `cam`
struct OverlapPair { Model cam, other; int w, h, ow, oh; double C[9]; uint8_t* mat; uint8_t* flag; };
__global__ void __launch_bounds__(128) k_compute_overlaps(const __grid_constant__ OverlapPair p)
{
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= p.w) return;
  double rx, ry;
  back_project(p.cam, (double)u, (double)v, rx, ry);   // the success flag is ignored, as in the reference
  const double ox = (p.C[0] * rx + p.C[1] * ry) + p.C[2] * 1.0, oy = (p.C[3] * rx + p.C[4] * ry) + p.C[5] * 1.0,
               oz = (p.C[6] * rx + p.C[7] * ry) + p.C[8] * 1.0;
  double kx, ky;
  uint8_t hit = 0;
  if (project(p.other, p.ow, p.oh, ox, oy, oz, kx, ky) == kProjSuccessful) {
    double vx, vy;
    back_project(p.other, kx, ky, vx, vy);
    const double n0 = sqrt((ox * ox + oy * oy) + oz * oz), n1 = sqrt((vx * vx + vy * vy) + 1.0 * 1.0);
    const double d = ((ox / n0) * (vx / n1) + (oy / n0) * (vy / n1)) + (oz / n0) * (1.0 / n1);
    if (fabs(d - 1.0) < 1.0e-10) hit = 1;
  }
  if (p.mat) p.mat[(size_t)v * p.w + u] = hit;
  if (hit) *p.flag = 1;
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

// both sides of a stereo pair in one launch (grid: keypoint tiles x frames x 2): D4 + the preparation above, and the reset of the
// matcher's reduction arrays (per-query best key, per-frame hit counter), which would otherwise be two more nodes in front of the scan
struct PairPrep {
  Model m[2]; PrepArgs p[2];
  const okb_keypoint_t* kp[2]; const int32_t* count[2]; int cap[2];
  double* rays[2]; uint8_t* valid[2]; double* e_W[2]; double* sof[2]; double* c26[2]; double* c6[2];
  unsigned long long* best; int32_t* hit_cnt;
};
__global__ void __launch_bounds__(128) k_stereo_prep_pair(const __grid_constant__ PairPrep a)
{
  const int frame = blockIdx.y, s = blockIdx.z;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int cap = a.cap[s];
  if (s == 0) {
    if (k == 0) a.hit_cnt[frame] = 0;
    if (k < cap) a.best[(size_t)frame * cap + k] = ~0ull;
  }
  if (k >= min(a.count[s][frame], cap)) return;
  const size_t i = (size_t)frame * cap + k;
  const okb_keypoint_t kp = a.kp[s][i];
  double x, y;
  const bool ok = back_project(a.m[s], (double)kp.x, (double)kp.y, x, y);
  const double z = 1.0;
  a.rays[s][3 * i] = x; a.rays[s][3 * i + 1] = y; a.rays[s][3 * i + 2] = z;
  a.valid[s][i] = ok ? 1 : 0;
  const PrepArgs& p = a.p[s];
  const double wx = (p.C[0] * x + p.C[1] * y) + p.C[2] * z;
  const double wy = (p.C[3] * x + p.C[4] * y) + p.C[5] * z;
  const double wz = (p.C[6] * x + p.C[7] * y) + p.C[8] * z;
  const double n2 = (wx * wx + wy * wy) + wz * wz;
  double ex = wx, ey = wy, ez = wz;
  if (n2 > 0.0) { const double n = sqrt(n2); ex = wx / n; ey = wy / n; ez = wz / n; }
  a.e_W[s][3 * i] = ex; a.e_W[s][3 * i + 1] = ey; a.e_W[s][3 * i + 2] = ez;
  const double sz = (double)kp.size / p.f;
  a.sof[s][i] = sz;
  const double sigma = sz * 0.125;
  a.c26[s][i] = gate_cos(2.6 * sigma); a.c6[s][i] = gate_cos(6.0 * sigma);   // == host libm, okb_gatecos.h
}

int camera_stereo_prep_pair(okb_context* ctx, const okb_camera_model_t* const model[2], const double* const C_WC[2], const okb_keypoint_t* const d_kp[2],
                            const int32_t* const d_count[2], const int cap[2], int n_frames, double* const d_rays[2], uint8_t* const d_valid[2],
                            double* const d_eW[2], double* const d_sof[2], double* const d_c26[2], double* const d_c6[2],
                            unsigned long long* d_best, int32_t* d_hit_cnt, cudaStream_t st)
{
  PairPrep a;
  for (int s = 0; s < 2; s++) {
    a.m[s] = to_model(*model[s]);
    for (int i = 0; i < 9; i++) a.p[s].C[i] = C_WC[s][i];
    a.p[s].f = 0.5 * (model[s]->fu + model[s]->fv);
    a.kp[s] = d_kp[s]; a.count[s] = d_count[s]; a.cap[s] = cap[s]; a.rays[s] = d_rays[s]; a.valid[s] = d_valid[s];
    a.e_W[s] = d_eW[s]; a.sof[s] = d_sof[s]; a.c26[s] = d_c26[s]; a.c6[s] = d_c6[s];
  }
  a.best = d_best; a.hit_cnt = d_hit_cnt;
  k_stereo_prep_pair<<<dim3((std::max(cap[0], cap[1]) + 127) / 128, n_frames, 2), 128, 0, st>>>(a);
  ctx->launches++;
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
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
  ctx->cams[cam].maps_ready = 0;   // camera-awareness maps of an earlier model are stale (okb_camera_awareness_maps recomputes them)
  return OKB_OK;
}


int okb_set_extraction_direction(okb_context_t* ctx, int cam, const double C_WC[9])
{
  if (!ctx || cam < 0 || cam >= ctx->n_cams || !C_WC) { set_error("okb_set_extraction_direction: bad arguments"); return OKB_ERR_ARGUMENT; }
  // T_WC.inverse().C() = C_WC^T; times g_W = (0, 0, -1): minus the third row of C_WC, evaluated as the 3x3 product does
  // (0 * a + 0 * b + (-1) * c per component), then narrowed to float as the reference's cv::Vec3f
  for (int i = 0; i < 3; i++) ctx->cams[cam].extraction_dir[i] = (float)((C_WC[0 + i] * 0.0 + C_WC[3 + i] * 0.0) + C_WC[6 + i] * -1.0);
  return OKB_OK;
}

int okb_get_extraction_direction(okb_context_t* ctx, int cam, float dir_out[3])
{
  if (!ctx || cam < 0 || cam >= ctx->n_cams || !dir_out) { set_error("okb_get_extraction_direction: bad arguments"); return OKB_ERR_ARGUMENT; }
  for (int i = 0; i < 3; i++) dir_out[i] = ctx->cams[cam].extraction_dir[i];
  return OKB_OK;
}

int okb_camera_awareness_maps(okb_context_t* ctx, int cam, float* rays_out, float* jac_out)
{
  if (!ctx || cam < 0 || cam >= ctx->n_cams) { set_error("okb_camera_awareness_maps: bad camera"); return OKB_ERR_ARGUMENT; }
  CamWorkspace& ws = ctx->cams[cam];
  if (!ws.has_model) { set_error("okb_camera_awareness_maps: camera %d has no model (okb_set_camera_model)", cam); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  const int W = ws.cfg.width, H = ws.cfg.height;
  const size_t n = (size_t)W * H;
  if (!ws.d_ray_map) { OKB_CUDA(cudaMalloc(&ws.d_ray_map, n * 12)); OKB_CUDA(cudaMalloc(&ws.d_jac_map, n * 24)); }
  k_awareness_maps<<<dim3((W + 127) / 128, H), 128, 0, ws.stream>>>(to_model(ws.model), W, H, ws.d_ray_map, ws.d_jac_map);
  ctx->launches++;
  OKB_CUDA(cudaGetLastError());
  if (rays_out) OKB_CUDA(cudaMemcpyAsync(rays_out, ws.d_ray_map, n * 12, cudaMemcpyDeviceToHost, ws.stream));
  if (jac_out) OKB_CUDA(cudaMemcpyAsync(jac_out, ws.d_jac_map, n * 24, cudaMemcpyDeviceToHost, ws.stream));
  OKB_CUDA(cudaStreamSynchronize(ws.stream));
  ws.maps_ready = 1;
  return OKB_OK;
}

int okb_compute_overlaps(okb_context_t* ctx, int n_cams, const okb_camera_model_t* models, const int32_t* widths, const int32_t* heights,
                         const double* C_rel, uint8_t* overlaps_out, uint8_t* const* mats_out)
{
  if (!ctx || n_cams < 1 || n_cams > 64 || !models || !widths || !heights || !C_rel || !overlaps_out) { set_error("okb_compute_overlaps: bad arguments"); return OKB_ERR_ARGUMENT; }
  for (int i = 0; i < n_cams; i++)
    if (widths[i] < 1 || heights[i] < 1 || models[i].model < 0 || models[i].model > 2 || !(models[i].fu > 0) || !(models[i].fv > 0)) {
      set_error("okb_compute_overlaps: camera %d", i); return OKB_ERR_ARGUMENT;
    }
  OKB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->match_slots[0].stream;
  size_t max_px = 0; for (int i = 0; i < n_cams; i++) max_px = std::max(max_px, (size_t)widths[i] * heights[i]);
  uint8_t* d_flags = nullptr; uint8_t* d_mat = nullptr;
  OKB_CUDA(cudaMalloc(&d_flags, (size_t)n_cams * n_cams));
  OKB_CUDA(cudaMemsetAsync(d_flags, 0, (size_t)n_cams * n_cams, st));
  if (mats_out) OKB_CUDA(cudaMalloc(&d_mat, max_px));
  int rc = OKB_OK;
  for (int s = 0; s < n_cams && rc == OKB_OK; s++)
    for (int c = 0; c < n_cams && rc == OKB_OK; c++) {
      uint8_t* host_mat = mats_out ? mats_out[s * n_cams + c] : nullptr;
      if (s == c) { if (host_mat) memset(host_mat, 1, (size_t)widths[c] * heights[c]); continue; }   // self-visibility is trivial
      OverlapPair p; memset(&p, 0, sizeof(p));
      p.cam = to_model(models[c]); p.other = to_model(models[s]); p.w = widths[c]; p.h = heights[c]; p.ow = widths[s]; p.oh = heights[s];
      for (int i = 0; i < 9; i++) p.C[i] = C_rel[9 * (size_t)(s * n_cams + c) + i];
      p.mat = host_mat ? d_mat : nullptr; p.flag = d_flags + s * n_cams + c;
      k_compute_overlaps<<<dim3((p.w + 127) / 128, p.h), 128, 0, st>>>(p);
      ctx->launches++;
      if (cudaGetLastError() != cudaSuccess) { set_error("okb_compute_overlaps: launch failed"); rc = OKB_ERR_CUDA; break; }
      if (host_mat && cudaMemcpyAsync(host_mat, d_mat, (size_t)p.w * p.h, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = OKB_ERR_CUDA;
      if (host_mat && cudaStreamSynchronize(st) != cudaSuccess) rc = OKB_ERR_CUDA;
    }
  if (rc == OKB_OK && cudaMemcpyAsync(overlaps_out, d_flags, (size_t)n_cams * n_cams, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = OKB_ERR_CUDA;
  if (cudaStreamSynchronize(st) != cudaSuccess) rc = OKB_ERR_CUDA;
  cudaFree(d_flags); cudaFree(d_mat);
  if (rc == OKB_OK) for (int i = 0; i < n_cams; i++) overlaps_out[i * n_cams + i] = 1;
  if (rc == OKB_ERR_CUDA) set_error("okb_compute_overlaps: CUDA error");
  return rc;
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
