// okb_prepare.cu -- P1: landmark-candidate preparation on the device (the step immediately before M1/M2).
//
// Replaces the serial host loop of Frontend::matchToMap (reference okvis_frontend/src/Frontend.cpp:1196-1360): project
// every landmark into the current camera (PinholeCamera::projectHomogeneous, cameras/implementation/PinholeCamera.hpp:
// 257-292,493-502), gate by the field of view +- reprThreshold, walk its observations newest first to decide is3d, drop
// observations with too much viewpoint / scale change and keep the "best 3" descriptors -- with the loop's exact
// bookkeeping (write row `o`, o = max(o, worstIdx), crop to `o` rows: see oracle/prepare_oracle.cpp for what that leaves
// behind). The result is the packed pool M1/M2 consume: cand_desc / cand_lm / lm_proj / lm_is3d (+ e_W, r_W, kid rows).
//
// Parallel formulation: the loop body is independent per landmark (the shared pool only couples landmarks through the
// running row offset), so
//   k_p1_select   one thread per landmark: projection, gates, the observation walk; leaves (keep, o, the observations
//                 that end up in rows 0..1, projection, is3d, p_W),
//   k_p1_scan     one CTA: exclusive scans of `keep` and of `o` in landmark order (= LandmarkId order of the map),
//   k_p1_scatter  16 threads per kept landmark: pool rows (descriptor bytes from the keyframe feature store, e_W, r_W).
// Descriptors and back-projections of the frames in the window live in a device-resident store (okb_store_frame).
// fp64 throughout, -fmad=false; cos(0.6) and cos(10/f) are host libm constants; acos runs on the device (it only feeds
// the comparison between the scores of one landmark's observations).
#include <math.h>
#include <string.h>

#include <vector>

#include "okb_internal.h"

namespace okb {

struct PV3 { double x, y, z; };
__device__ __forceinline__ PV3 psub(PV3 a, PV3 b) { return PV3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double pdot(PV3 a, PV3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ double pnorm(PV3 a) { return sqrt(pdot(a, a)); }
__device__ __forceinline__ PV3 pnormalized(PV3 a) { const double n = pnorm(a); return PV3{a.x / n, a.y / n, a.z / n}; }
__device__ __forceinline__ PV3 prot(const double* C, PV3 v)
{
  return PV3{(C[0] * v.x + C[1] * v.y) + C[2] * v.z, (C[3] * v.x + C[4] * v.y) + C[5] * v.z, (C[6] * v.x + C[7] * v.y) + C[8] * v.z};
}

struct P1Args {
  int n_lm, n_cams, D;
  const double* hp_W; const double* quality; const int32_t* obs_begin; const int32_t* obs; const double* T_WC_old;
  const uint8_t* const* desc_tab; const double* const* ray_tab;
  double T_WC1[12], T_CW1[12];
  int model; double fu, fv, cu, cv, k[4];
  int width, height; double thr; int exclusive;
  double cos06, cos10f, focal;
  // per landmark
  uint8_t* keep; uint8_t* o; int32_t* row_obs; double* proj; uint8_t* is3d; double* p_W;
  int32_t* lm_off; int32_t* row_off; int32_t* totals;
  // outputs
  int32_t* out_lm; double* out_proj; uint8_t* out_is3d; double* out_p_W; int32_t* out_desc_begin; uint8_t* pool; int32_t* cand_lm;
  double* out_e_W; double* out_r_W; int32_t* out_kid;
};

__global__ void __launch_bounds__(128) k_p1_select(const __grid_constant__ P1Args a)
{
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= a.n_lm) return;
  a.keep[it] = 0; a.o[it] = 0;
  const double* hp = a.hp_W + 4 * (size_t)it;
  const double h0 = hp[0], h1 = hp[1], h2 = hp[2], h3 = hp[3];
  const PV3 p_W = PV3{h0 / h3, h1 / h3, h2 / h3};
  const PV3 r_W = psub(p_W, PV3{a.T_WC1[9], a.T_WC1[10], a.T_WC1[11]});
  const PV3 e_W = pnormalized(r_W);
  const double r = fmax(0.01, pnorm(r_W));
  const PV3 ch = prot(a.T_CW1, PV3{h0, h1, h2});
  PV3 head = PV3{ch.x + a.T_CW1[9] * h3, ch.y + a.T_CW1[10] * h3, ch.z + a.T_CW1[11] * h3};
  if (h3 < 0) head = PV3{-head.x, -head.y, -head.z};
  // PinholeCamera::project
  if (fabs(head.z) < 1.0e-12) return;   // Invalid
  const double rz = 1.0 / head.z;
  const double u0 = head.x * rz, u1 = head.y * rz;
  double d0, d1;
  if (a.model == 1) {
    const double k1 = a.k[0], k2 = a.k[1], p1 = a.k[2], p2 = a.k[3];
    const double mx_u = u0 * u0, my_u = u1 * u1, mxy_u = u0 * u1;
    const double rho_u = mx_u + my_u;
    const double rad_dist_u = k1 * rho_u + k2 * rho_u * rho_u;
    d0 = u0 + u0 * rad_dist_u + 2.0 * p1 * mxy_u + p2 * (rho_u + 2.0 * mx_u);
    d1 = u1 + u1 * rad_dist_u + 2.0 * p2 * mxy_u + p1 * (rho_u + 2.0 * my_u);
  } else if (a.model == 2) {
    const double rr = sqrt(u0 * u0 + u1 * u1);
    const double theta = atan(rr);
    const double theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
    const double thetad = theta * (1.0 + a.k[0] * theta2 + a.k[1] * theta4 + a.k[2] * theta6 + a.k[3] * theta8);
    const double scaling = (rr > 1e-8) ? thetad / rr : 1.0;
    d0 = scaling * u0; d1 = scaling * u1;
  } else { d0 = u0; d1 = u1; }
  const double kx = a.fu * d0 + a.cu, ky = a.fv * d1 + a.cv;
  const bool outside = kx < 0.0 || ky < 0.0 || kx >= a.width || ky >= a.height;
  if (!outside && !(head.z > 0.0)) return;   // Behind (only reported for points that land inside the image)
  if (kx < -a.thr) return;
  if (ky < -a.thr) return;
  if (kx > a.width + a.thr) return;
  if (ky > a.height + a.thr) return;
  const double q = a.quality[it];
  double best[3] = {1.0, 1.0, 1.0};
  int rows[3] = {-1, -1, -1};
  int o = 0;
  bool is3d = false;
  for (int oi = a.obs_begin[it + 1] - 1; oi >= a.obs_begin[it]; --oi) {
    const int32_t* kid = a.obs + 3 * (size_t)oi;
    const double* T_old = a.T_WC_old + ((size_t)kid[0] * a.n_cams + kid[1]) * 12;
    const PV3 r_W_old = psub(p_W, PV3{T_old[9], T_old[10], T_old[11]});
    if (!is3d) {
      const double f = 0.2 / a.focal / q;
      const PV3 r_close_W = psub(r_W, PV3{f * r_W_old.x, f * r_W_old.y, f * r_W_old.z});
      const double cosA = pdot(pnormalized(r_W), pnormalized(r_close_W));
      if (cosA > a.cos10f) is3d = true;
    }
    const double cosVC = pdot(e_W, pnormalized(r_W_old));
    if (cosVC < a.cos06 && !a.exclusive) continue;
    const double scaleChange = fabs(r - pnorm(r_W_old)) / r;
    if ((scaleChange > 0.5) && !a.exclusive) continue;
    const double score = 0.5 * (acos(cosVC) / 0.6 + scaleChange / 0.5);
    double worstScore = 0.0;
    int worstIdx = 0;
#pragma unroll
    for (int n = 0; n < 3; ++n) if (best[n] > worstScore) { worstScore = best[n]; worstIdx = n; }
    if (score < best[worstIdx]) {
      rows[o] = oi;
      o = max(o, worstIdx);
      best[worstIdx] = score;
    }
  }
  if (o == 0) return;
  a.keep[it] = 1; a.o[it] = (uint8_t)o;
  a.row_obs[2 * (size_t)it] = rows[0]; a.row_obs[2 * (size_t)it + 1] = rows[1];
  a.proj[2 * (size_t)it] = kx; a.proj[2 * (size_t)it + 1] = ky;
  a.is3d[it] = is3d ? 1 : 0;
  a.p_W[3 * (size_t)it] = p_W.x; a.p_W[3 * (size_t)it + 1] = p_W.y; a.p_W[3 * (size_t)it + 2] = p_W.z;
}

// exclusive scans in landmark order: lm_off (kept landmarks before this one) and row_off (pool rows before this one)
__global__ void __launch_bounds__(1024) k_p1_scan(const __grid_constant__ P1Args a)
{
  __shared__ int sh_l[33], sh_r[33];
  __shared__ int base_l, base_r;
  if (threadIdx.x == 0) { base_l = 0; base_r = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int start = 0; start < a.n_lm; start += 1024) {
    const int i = start + threadIdx.x;
    const int kl = (i < a.n_lm && a.keep[i]) ? 1 : 0;
    const int kr = kl ? (int)a.o[i] : 0;
    int il = kl, ir = kr;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int tl = __shfl_up_sync(0xffffffffu, il, s), tr = __shfl_up_sync(0xffffffffu, ir, s);
      if (lane >= s) { il += tl; ir += tr; }
    }
    if (lane == 31) { sh_l[warp] = il; sh_r[warp] = ir; }
    __syncthreads();
    if (warp == 0) {
      const int wl = sh_l[lane], wr = sh_r[lane];
      int xl = wl, xr = wr;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const int tl = __shfl_up_sync(0xffffffffu, xl, s), tr = __shfl_up_sync(0xffffffffu, xr, s);
        if (lane >= s) { xl += tl; xr += tr; }
      }
      sh_l[lane] = xl - wl; sh_r[lane] = xr - wr;
      if (lane == 31) { sh_l[32] = xl; sh_r[32] = xr; }
    }
    __syncthreads();
    if (i < a.n_lm) { a.lm_off[i] = base_l + sh_l[warp] + il - kl; a.row_off[i] = base_r + sh_r[warp] + ir - kr; }
    __syncthreads();
    if (threadIdx.x == 0) { base_l += sh_l[32]; base_r += sh_r[32]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { a.totals[0] = base_l; a.totals[1] = base_r; a.out_desc_begin[base_l] = base_r; }
}

__global__ void __launch_bounds__(256) k_p1_scatter(const __grid_constant__ P1Args a)
{
  const int it = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, t = threadIdx.x & 15;
  if (it >= a.n_lm || !a.keep[it]) return;
  const int j = a.lm_off[it], row0 = a.row_off[it], o = a.o[it];
  if (t == 0) {
    a.out_lm[j] = it;
    a.out_proj[2 * (size_t)j] = a.proj[2 * (size_t)it]; a.out_proj[2 * (size_t)j + 1] = a.proj[2 * (size_t)it + 1];
    a.out_is3d[j] = a.is3d[it];
    a.out_p_W[3 * (size_t)j] = a.p_W[3 * (size_t)it]; a.out_p_W[3 * (size_t)j + 1] = a.p_W[3 * (size_t)it + 1];
    a.out_p_W[3 * (size_t)j + 2] = a.p_W[3 * (size_t)it + 2];
    a.out_desc_begin[j] = row0;
  }
  const int row = t >> 3, part = t & 7;   // 8 threads per pool row
  if (row >= o) return;
  const int oi = a.row_obs[2 * (size_t)it + row];
  const int32_t* kid = a.obs + 3 * (size_t)oi;
  const size_t tab = (size_t)kid[0] * a.n_cams + kid[1];
  const uint8_t* src = a.desc_tab[tab] + (size_t)a.D * kid[2];
  uint8_t* dst = a.pool + (size_t)a.D * (row0 + row);
  for (int b = part * 8; b < a.D; b += 64) *reinterpret_cast<uint2*>(dst + b) = *reinterpret_cast<const uint2*>(src + b);
  if (part == 0) {
    const double* T_old = a.T_WC_old + tab * 12;
    const double* ec = a.ray_tab[tab] + 3 * (size_t)kid[2];
    const PV3 e = prot(T_old, pnormalized(PV3{ec[0], ec[1], ec[2]}));
    double* eo = a.out_e_W + 3 * (size_t)(row0 + row);
    eo[0] = e.x; eo[1] = e.y; eo[2] = e.z;
    double* ro = a.out_r_W + 3 * (size_t)(row0 + row);
    ro[0] = T_old[9]; ro[1] = T_old[10]; ro[2] = T_old[11];
    int32_t* ko = a.out_kid + 3 * (size_t)(row0 + row);
    ko[0] = kid[0]; ko[1] = kid[1]; ko[2] = kid[2];
    a.cand_lm[row0 + row] = j;
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
struct StoreEntry { uint8_t* d_desc = nullptr; double* d_rays = nullptr; int n = 0, cap = 0; };
struct PrepareState {
  int n_slots = 0, n_cams = 0, D = 0;
  std::vector<StoreEntry> store;
  const uint8_t** d_desc_tab = nullptr; const double** d_ray_tab = nullptr; bool tab_dirty = true;
  uint8_t* d_buf = nullptr; size_t d_cap = 0;
  cudaStream_t stream = nullptr;
  // device results of the last okb_prepare_landmarks
  const uint8_t* r_pool = nullptr; const int32_t* r_cand_lm = nullptr; const double* r_proj = nullptr; const uint8_t* r_is3d = nullptr;
  int r_rows = 0, r_lm = 0;
};

static PrepareState* state(okb_context* ctx)
{
  if (!ctx->prepare) ctx->prepare = new PrepareState();
  return static_cast<PrepareState*>(ctx->prepare);
}

void prepare_free(okb_context* ctx)
{
  if (!ctx->prepare) return;
  PrepareState* s = static_cast<PrepareState*>(ctx->prepare);
  for (auto& e : s->store) { cudaFree(e.d_desc); cudaFree(e.d_rays); }
  cudaFree(s->d_desc_tab); cudaFree(s->d_ray_tab); cudaFree(s->d_buf);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  ctx->prepare = nullptr;
}

}  // namespace okb

using namespace okb;

extern "C" {

int okb_store_configure(okb_context_t* ctx, int n_slots, int n_cams, int D)
{
  if (!ctx || n_slots < 1 || n_cams < 1 || (D != 48 && D != 64)) { set_error("okb_store_configure: bad arguments"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  prepare_free(ctx);
  PrepareState* s = state(ctx);
  s->n_slots = n_slots; s->n_cams = n_cams; s->D = D;
  s->store.resize((size_t)n_slots * n_cams);
  OKB_CUDA(cudaMalloc(&s->d_desc_tab, sizeof(void*) * s->store.size()));
  OKB_CUDA(cudaMalloc(&s->d_ray_tab, sizeof(void*) * s->store.size()));
  OKB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  return OKB_OK;
}

static int store_slot(okb_context_t* ctx, int slot, int cam, int n, StoreEntry** out)
{
  PrepareState* s = ctx ? static_cast<PrepareState*>(ctx->prepare) : nullptr;
  if (!s || slot < 0 || slot >= s->n_slots || cam < 0 || cam >= s->n_cams || n < 0) { set_error("okb_store_frame: bad arguments (configure the store first)"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  StoreEntry& e = s->store[(size_t)slot * s->n_cams + cam];
  if (n > e.cap) {
    OKB_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(e.d_desc); cudaFree(e.d_rays); e.d_desc = nullptr; e.d_rays = nullptr;
    e.cap = (n + 255) / 256 * 256;
    OKB_CUDA(cudaMalloc(&e.d_desc, (size_t)e.cap * s->D));
    OKB_CUDA(cudaMalloc(&e.d_rays, (size_t)e.cap * 24));
    s->tab_dirty = true;
  }
  e.n = n;
  *out = &e;
  return OKB_OK;
}

int okb_store_frame(okb_context_t* ctx, int slot, int cam, int n, const uint8_t* desc, const double* rays)
{
  StoreEntry* e = nullptr;
  const int rc = store_slot(ctx, slot, cam, n, &e);
  if (rc) return rc;
  PrepareState* s = static_cast<PrepareState*>(ctx->prepare);
  if (n > 0 && (!desc || !rays)) { set_error("okb_store_frame: null input"); return OKB_ERR_ARGUMENT; }
  if (n > 0) {
    OKB_CUDA(cudaMemcpyAsync(e->d_desc, desc, (size_t)n * s->D, cudaMemcpyHostToDevice, s->stream));
    OKB_CUDA(cudaMemcpyAsync(e->d_rays, rays, (size_t)n * 24, cudaMemcpyHostToDevice, s->stream));
    OKB_CUDA(cudaStreamSynchronize(s->stream));   // the caller's buffers are only valid for the call
  }
  return OKB_OK;
}

int okb_store_frame_from_last(okb_context_t* ctx, int slot, int cam, int batch_index)
{
  if (!ctx || cam < 0 || cam >= ctx->n_cams) { set_error("okb_store_frame_from_last: bad camera"); return OKB_ERR_ARGUMENT; }
  CamWorkspace& ws = ctx->cams[cam];
  if (batch_index < 0 || batch_index >= ws.cfg.max_batch || !ws.d_rays) { set_error("okb_store_frame_from_last: bad batch index / no camera model"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  OKB_CUDA(cudaStreamSynchronize(ws.stream));
  int n = 0;
  OKB_CUDA(cudaMemcpy(&n, ws.d_count + batch_index, 4, cudaMemcpyDeviceToHost));
  StoreEntry* e = nullptr;
  const int rc = store_slot(ctx, slot, cam, n, &e);
  if (rc) return rc;
  PrepareState* s = static_cast<PrepareState*>(ctx->prepare);
  const size_t D = (size_t)ws.cfg.descriptor_bytes;
  if (s->D != (int)D) { set_error("okb_store_frame_from_last: the store holds %d-byte descriptors, camera %d makes %d-byte ones", s->D, cam, (int)D); return OKB_ERR_ARGUMENT; }
  if (n > 0) {
    OKB_CUDA(cudaMemcpyAsync(e->d_desc, ws.d_desc + (size_t)batch_index * ws.kp_cap * D, (size_t)n * D, cudaMemcpyDeviceToDevice, s->stream));
    OKB_CUDA(cudaMemcpyAsync(e->d_rays, ws.d_rays + (size_t)batch_index * ws.kp_cap * 3, (size_t)n * 24, cudaMemcpyDeviceToDevice, s->stream));
    OKB_CUDA(cudaStreamSynchronize(s->stream));
  }
  return OKB_OK;
}

int okb_prepare_landmarks(okb_context_t* ctx, const okb_prepare_view_t* view, int n_lm, const double* hp_W, const double* quality,
                          const int32_t* obs_begin, int n_obs, const int32_t* obs, const double* T_WC_old,
                          int32_t* n_out, int32_t* n_rows, int32_t* out_lm, double* out_proj, uint8_t* out_is3d, double* out_p_W,
                          int32_t* out_desc_begin, uint8_t* out_pool, int32_t* out_cand_lm, double* out_e_W, double* out_r_W,
                          int32_t* out_kid)
{
  PrepareState* s = ctx ? static_cast<PrepareState*>(ctx->prepare) : nullptr;
  if (!s || !view || n_lm < 0 || n_obs < 0 || !n_out || !n_rows) { set_error("okb_prepare_landmarks: bad arguments (configure the store first)"); return OKB_ERR_ARGUMENT; }
  if (n_lm > 0 && (!hp_W || !quality || !obs_begin || !T_WC_old || (n_obs > 0 && !obs))) { set_error("okb_prepare_landmarks: null input"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  *n_out = 0; *n_rows = 0;
  s->r_rows = s->r_lm = 0;
  if (n_lm == 0) return OKB_OK;
  for (int i = 0; i < n_lm; i++)
    if (obs_begin[i] < 0 || obs_begin[i + 1] < obs_begin[i] || obs_begin[i + 1] > n_obs) { set_error("okb_prepare_landmarks: obs_begin not monotone at %d", i); return OKB_ERR_ARGUMENT; }
  for (int i = 0; i < n_obs; i++) {
    const int32_t* k = obs + 3 * (size_t)i;
    if (k[0] < 0 || k[0] >= s->n_slots || k[1] < 0 || k[1] >= s->n_cams || k[2] < 0 || k[2] >= s->store[(size_t)k[0] * s->n_cams + k[1]].n) {
      set_error("okb_prepare_landmarks: observation %d (frame slot %d, camera %d, keypoint %d) is not in the store", i, k[0], k[1], k[2]);
      return OKB_ERR_ARGUMENT;
    }
  }
  cudaStream_t st = s->stream;
  if (s->tab_dirty) {
    std::vector<const uint8_t*> dt(s->store.size()); std::vector<const double*> rt(s->store.size());
    for (size_t i = 0; i < s->store.size(); i++) { dt[i] = s->store[i].d_desc; rt[i] = s->store[i].d_rays; }
    OKB_CUDA(cudaMemcpyAsync(s->d_desc_tab, dt.data(), sizeof(void*) * dt.size(), cudaMemcpyHostToDevice, st));
    OKB_CUDA(cudaMemcpyAsync(s->d_ray_tab, rt.data(), sizeof(void*) * rt.size(), cudaMemcpyHostToDevice, st));
    OKB_CUDA(cudaStreamSynchronize(st));
    s->tab_dirty = false;
  }
  // device arena: inputs | per-landmark scratch | outputs
  const size_t L = (size_t)n_lm, O = (size_t)n_obs, R = 2 * L, D = (size_t)s->D;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t i_hp = take(L * 32), i_q = take(L * 8), i_ob = take((L + 1) * 4), i_obs = take(O * 12 + 16),
               i_T = take((size_t)s->n_slots * s->n_cams * 96);
  const size_t s_keep = take(L), s_o = take(L), s_rows = take(L * 8), s_proj = take(L * 16), s_3d = take(L), s_pw = take(L * 24),
               s_lo = take(L * 4), s_ro = take(L * 4), s_tot = take(8);
  const size_t o_lm = take(L * 4), o_proj = take(L * 16), o_3d = take(L), o_pw = take(L * 24), o_db = take((L + 1) * 4),
               o_pool = take(R * D), o_cl = take(R * 4), o_e = take(R * 24), o_r = take(R * 24), o_kid = take(R * 12);
  if (off > s->d_cap) {
    OKB_CUDA(cudaStreamSynchronize(st));
    cudaFree(s->d_buf); s->d_buf = nullptr; s->d_cap = 0;
    OKB_CUDA(cudaMalloc(&s->d_buf, off + off / 4));
    s->d_cap = off + off / 4;
  }
  uint8_t* d = s->d_buf;
  OKB_CUDA(cudaMemcpyAsync(d + i_hp, hp_W, L * 32, cudaMemcpyHostToDevice, st));
  OKB_CUDA(cudaMemcpyAsync(d + i_q, quality, L * 8, cudaMemcpyHostToDevice, st));
  OKB_CUDA(cudaMemcpyAsync(d + i_ob, obs_begin, (L + 1) * 4, cudaMemcpyHostToDevice, st));
  if (O) OKB_CUDA(cudaMemcpyAsync(d + i_obs, obs, O * 12, cudaMemcpyHostToDevice, st));
  OKB_CUDA(cudaMemcpyAsync(d + i_T, T_WC_old, (size_t)s->n_slots * s->n_cams * 96, cudaMemcpyHostToDevice, st));
  P1Args a;
  memset(&a, 0, sizeof(a));
  a.n_lm = n_lm; a.n_cams = s->n_cams; a.D = s->D;
  a.hp_W = (const double*)(d + i_hp); a.quality = (const double*)(d + i_q); a.obs_begin = (const int32_t*)(d + i_ob);
  a.obs = (const int32_t*)(d + i_obs); a.T_WC_old = (const double*)(d + i_T);
  a.desc_tab = s->d_desc_tab; a.ray_tab = s->d_ray_tab;
  memcpy(a.T_WC1, view->T_WC1, 96); memcpy(a.T_CW1, view->T_CW1, 96);
  a.model = view->model.model; a.fu = view->model.fu; a.fv = view->model.fv; a.cu = view->model.cu; a.cv = view->model.cv;
  for (int i = 0; i < 4; i++) a.k[i] = view->model.k[i];
  a.width = view->width; a.height = view->height; a.thr = view->repr_threshold; a.exclusive = view->exclusive;
  a.focal = view->model.fu + view->model.fv;
  a.cos06 = cos(0.6); a.cos10f = cos(10.0 / a.focal);
  a.keep = d + s_keep; a.o = d + s_o; a.row_obs = (int32_t*)(d + s_rows); a.proj = (double*)(d + s_proj); a.is3d = d + s_3d;
  a.p_W = (double*)(d + s_pw); a.lm_off = (int32_t*)(d + s_lo); a.row_off = (int32_t*)(d + s_ro); a.totals = (int32_t*)(d + s_tot);
  a.out_lm = (int32_t*)(d + o_lm); a.out_proj = (double*)(d + o_proj); a.out_is3d = d + o_3d; a.out_p_W = (double*)(d + o_pw);
  a.out_desc_begin = (int32_t*)(d + o_db); a.pool = d + o_pool; a.cand_lm = (int32_t*)(d + o_cl); a.out_e_W = (double*)(d + o_e);
  a.out_r_W = (double*)(d + o_r); a.out_kid = (int32_t*)(d + o_kid);
  k_p1_select<<<(n_lm + 127) / 128, 128, 0, st>>>(a);
  k_p1_scan<<<1, 1024, 0, st>>>(a);
  k_p1_scatter<<<(n_lm * 16 + 255) / 256, 256, 0, st>>>(a);
  ctx->launches += 3;
  int32_t tot[2] = {0, 0};
  OKB_CUDA(cudaMemcpyAsync(tot, d + s_tot, 8, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(cudaStreamSynchronize(st));
  OKB_CUDA(cudaGetLastError());
  const size_t nl = (size_t)tot[0], nr = (size_t)tot[1];
  *n_out = tot[0]; *n_rows = tot[1];
  s->r_pool = d + o_pool; s->r_cand_lm = (const int32_t*)(d + o_cl); s->r_proj = (const double*)(d + o_proj); s->r_is3d = d + o_3d;
  s->r_rows = tot[1]; s->r_lm = tot[0];
  auto back = [&](void* dst, size_t src, size_t bytes) -> int {
    if (dst && bytes) OKB_CUDA(cudaMemcpyAsync(dst, d + src, bytes, cudaMemcpyDeviceToHost, st));
    return OKB_OK;
  };
  int rc = OKB_OK;
  if (!rc) rc = back(out_lm, o_lm, nl * 4);
  if (!rc) rc = back(out_proj, o_proj, nl * 16);
  if (!rc) rc = back(out_is3d, o_3d, nl);
  if (!rc) rc = back(out_p_W, o_pw, nl * 24);
  if (!rc) rc = back(out_desc_begin, o_db, (nl + 1) * 4);
  if (!rc) rc = back(out_pool, o_pool, nr * D);
  if (!rc) rc = back(out_cand_lm, o_cl, nr * 4);
  if (!rc) rc = back(out_e_W, o_e, nr * 24);
  if (!rc) rc = back(out_r_W, o_r, nr * 24);
  if (!rc) rc = back(out_kid, o_kid, nr * 12);
  if (rc) return rc;
  OKB_CUDA(cudaStreamSynchronize(st));
  return OKB_OK;
}

int okb_prepared_device(okb_context_t* ctx, const uint8_t** d_cand_desc, const int32_t** d_cand_lm, const double** d_lm_proj,
                        const uint8_t** d_lm_is3d, int32_t* n_cand, int32_t* n_lm)
{
  PrepareState* s = ctx ? static_cast<PrepareState*>(ctx->prepare) : nullptr;
  if (!s) { set_error("okb_prepared_device: nothing prepared"); return OKB_ERR_ARGUMENT; }
  if (d_cand_desc) *d_cand_desc = s->r_pool;
  if (d_cand_lm) *d_cand_lm = s->r_cand_lm;
  if (d_lm_proj) *d_lm_proj = s->r_proj;
  if (d_lm_is3d) *d_lm_is3d = s->r_is3d;
  if (n_cand) *n_cand = s->r_rows;
  if (n_lm) *n_lm = s->r_lm;
  return OKB_OK;
}

}  // extern "C"
