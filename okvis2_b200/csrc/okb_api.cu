// okb_api.cu -- extern "C" shim of libokvis_b200.so (declared in include/okvis_b200.h): context life cycle,
// host<->device staging around the detect/describe kernels, inspection hooks. No CPU fallback anywhere.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "okb_internal.h"
#include "okb_gatecos.h"

namespace okb {
static thread_local char g_err[512] = "";
static char g_err_global[512] = "";
static std::mutex g_err_mutex;
void set_error(const char* fmt, ...)
{
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(g_err_mutex);
  memcpy(g_err_global, g_err, sizeof(g_err));
}
void detect_collect_timing(okb_context* ctx, int cam);
int match_init(okb_context* ctx);
void match_free(okb_context* ctx);
}  // namespace okb

using namespace okb;

extern "C" {

const char* okb_last_error(void) { return g_err[0] ? g_err : g_err_global; }
const char* okb_version(void) { return "okvis2_b200 0.1 (sm_100a)"; }

int okb_create(int device, int n_cams, const okb_camera_config_t* cfgs, okb_context_t** out)
{
  if (!out || n_cams < 0 || (n_cams > 0 && !cfgs)) { set_error("okb_create: bad arguments"); return OKB_ERR_ARGUMENT; }
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    set_error("okb_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
    return OKB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n_dev) { set_error("okb_create: device %d of %d", device, n_dev); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(device));
  okb_context* ctx = new okb_context();
  ctx->device = device; ctx->n_cams = n_cams; ctx->cams.resize(n_cams);
  float ps = n_cams > 0 ? cfgs[0].pattern_scale : 1.0f;
  if (ps <= 0.f) ps = 1.0f;
  for (int i = 1; i < n_cams; i++) {   // the sampling-pattern tables are per context
    const float pi = cfgs[i].pattern_scale <= 0.f ? 1.0f : cfgs[i].pattern_scale;
    if (pi != ps) { set_error("okb_create: camera %d has pattern_scale %g, camera 0 has %g (one pattern table per context)", i, pi, ps); delete ctx; return OKB_ERR_ARGUMENT; }
  }
  int rc = tables_init(ctx, ps);
  if (rc != OKB_OK) { delete ctx; return rc; }
  for (int i = 0; i < n_cams; i++) {
    ctx->cams[i].cfg = cfgs[i];
    okb_camera_config_t& c = ctx->cams[i].cfg;
    if (c.max_batch < 1) c.max_batch = 1;
    if (c.descriptor_bytes == 0) c.descriptor_bytes = 64;
    if (c.descriptor_bytes != 64 && c.descriptor_bytes != 48) {
      set_error("camera %d: descriptor_bytes=%d (64 = AGAST + BRISK-512, 48 = Harris + BRISK2)", i, c.descriptor_bytes);
      okb_destroy(ctx); return OKB_ERR_ARGUMENT;
    }
    if (c.octaves < 0 || c.max_keypoints < 0 || c.max_keypoints >= (1 << 20)) {
      set_error("camera %d: octaves %d / max_keypoints %d out of range", i, c.octaves, c.max_keypoints); okb_destroy(ctx); return OKB_ERR_ARGUMENT;
    }
    if (c.descriptor_bytes == 48) {
      if (c.octaves != 0) {
        set_error("camera %d: the Harris + BRISK2 mode (descriptor_bytes = 48) is single-scale: octaves must be 0 (every shipped okvis configuration)", i);
        okb_destroy(ctx); return OKB_ERR_UNSUPPORTED;
      }
      if (!(c.uniformity_radius > 0.f) || c.threshold < 1) {
        set_error("camera %d: uniformity_radius %g / absolute threshold %d", i, c.uniformity_radius, c.threshold); okb_destroy(ctx); return OKB_ERR_ARGUMENT;
      }
    } else if (c.threshold < 1 || c.threshold > 254) { set_error("camera %d: threshold %d", i, c.threshold); okb_destroy(ctx); return OKB_ERR_ARGUMENT; }
    rc = detect_init_camera(ctx, i);
    if (rc != OKB_OK) { okb_destroy(ctx); return rc; }
  }
  rc = match_init(ctx);
  if (rc != OKB_OK) { okb_destroy(ctx); return rc; }
  if (cudaMalloc(&ctx->d_scan_status, 4) != cudaSuccess || cudaMemset(ctx->d_scan_status, 0, 4) != cudaSuccess) {
    set_error("okb_create: scan status word"); okb_destroy(ctx); return OKB_ERR_CUDA;
  }
  {
    // self-check of the gate constants: gate_cos must return what THIS machine's libm returns (it restates glibc's
    // algorithm, okb_gatecos.h); 65 536 arguments over the gate's range
    uint64_t s = 88172645463325252ull; int bad = 0;
    for (int i = 0; i < 65536; i++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const double x = (double)(s >> 11) / 9007199254740992.0 * ((i & 3) ? 0.5 : 0.855);
      bad += gate_cos(x) != cos(x);
    }
    ctx->gate_cos_exact = bad == 0;
  }
  { const cudaError_t e2 = cudaDeviceSynchronize();
    if (e2 != cudaSuccess) { set_error("okb_create: %s", cudaGetErrorString(e2)); okb_destroy(ctx); return OKB_ERR_CUDA; } }
  *out = ctx;
  return OKB_OK;
}

void okb_destroy(okb_context_t* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < ctx->n_cams; i++) detect_free_camera(ctx, i);
  match_free(ctx);
  stream_free(ctx);
  prepare_free(ctx);
  aux_free(ctx);
  if (ctx->stereo_scratch) cudaFree(ctx->stereo_scratch);
  if (ctx->d_scan_status) cudaFree(ctx->d_scan_status);
  if (ctx->motion.d) cudaFree(ctx->motion.d);
  if (ctx->motion.h) cudaFreeHost(ctx->motion.h);
  tables_free(ctx);
  delete ctx;
}

int okb_gate_cos_exact(const okb_context_t* ctx) { return ctx ? ctx->gate_cos_exact : 0; }
int64_t okb_launch_count(const okb_context_t* ctx) { return ctx ? ctx->launches : 0; }
int okb_set_blocking_sync(okb_context_t* ctx, int on) { if (!ctx) return OKB_ERR_ARGUMENT; ctx->blocking_sync = on ? 1 : 0; return OKB_OK; }
void* okb_stream(okb_context_t* ctx, int cam)
{
  if (!ctx) return nullptr;
  if (cam < 0 || cam >= ctx->n_cams) return (void*)ctx->match_slots[0].stream;
  return (void*)ctx->cams[cam].stream;
}
int okb_sync(okb_context_t* ctx)
{
  if (!ctx) return OKB_ERR_ARGUMENT;
  OKB_CUDA(cudaSetDevice(ctx->device));
  for (int i = 0; i < ctx->n_cams; i++) OKB_CUDA(cudaStreamSynchronize(ctx->cams[i].stream));
  for (int i = 0; i < kMatchSlots; i++) OKB_CUDA(cudaStreamSynchronize(ctx->match_slots[i].stream));
  if (ctx->d_scan_status) {
    int32_t st = 0;
    OKB_CUDA(cudaMemcpy(&st, ctx->d_scan_status, 4, cudaMemcpyDeviceToHost));
    if (st) { set_error("okb_sync: the tensor-core Hamming scan reported a timed-out barrier wait (flags 0x%x)", st); return OKB_ERR_CUDA; }
  }
  return OKB_OK;
}

static int check_cam(okb_context_t* ctx, int cam, const char* who)
{
  if (!ctx || cam < 0 || cam >= ctx->n_cams) { set_error("%s: bad context/camera %d", who, cam); return OKB_ERR_ARGUMENT; }
  return OKB_OK;
}

static int status_to_error(const CamWorkspace& ws, int n_frames)
{
  for (int b = 0; b < n_frames; b++)
    if (ws.h_status[b]) {
      set_error("detect: device capacity exceeded on frame %d (flags 0x%x: 1=a layer overflowed its share of the %d candidate slots, 2=ties, 4=keypoints>sort cap, "
                "8=keypoints>output cap %d, 16=TMA tile load timed out)", b, ws.h_status[b], ws.cand_cap, ws.kp_cap);
      return OKB_ERR_CAPACITY;
    }
  return OKB_OK;
}

int okb_detect_describe_batch(okb_context_t* ctx, int cam, int n_frames, const uint8_t* images, size_t stride_bytes,
                              okb_keypoint_t* kp_out, uint8_t* desc_out, int cap, int* n_out)
{
  int rc = check_cam(ctx, cam, "okb_detect_describe");
  if (rc) return rc;
  CamWorkspace& ws = ctx->cams[cam];
  const int W = ws.cfg.width, H = ws.cfg.height;
  const size_t D = (size_t)ws.cfg.descriptor_bytes;
  if (!images || !kp_out || !desc_out || !n_out || cap < 0 || n_frames < 1 || n_frames > ws.cfg.max_batch || stride_bytes < (size_t)W) {
    set_error("okb_detect_describe: bad arguments (n_frames %d of max_batch %d)", n_frames, ws.cfg.max_batch);
    return OKB_ERR_ARGUMENT;
  }
  OKB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ws.stream;
  static_assert(sizeof(okb_keypoint_t) == 28, "cv::KeyPoint layout");
  // host -> device: pageable caller memory (cv::Mat) is staged through the pinned buffer; page-locked caller memory
  // (cudaHostAlloc / cudaHostRegister) is read by the copy engine directly
  const bool in_pinned = host_pinned(images);
  if (in_pinned && stride_bytes == (size_t)W) {
    OKB_CUDA(cudaMemcpyAsync(ws.d_in, images, (size_t)W * H * n_frames, cudaMemcpyHostToDevice, st));
  } else if (in_pinned) {   // row-by-row DMA: correct for padded images, much slower than the dense copy
    OKB_CUDA(cudaMemcpy2DAsync(ws.d_in, (size_t)W, images, stride_bytes, (size_t)W, (size_t)H * n_frames, cudaMemcpyHostToDevice, st));
  } else {
    for (int b = 0; b < n_frames; b++)
      for (int y = 0; y < H; y++) memcpy(ws.h_img + ((size_t)b * H + y) * W, images + ((size_t)b * H + y) * stride_bytes, (size_t)W);
    OKB_CUDA(cudaMemcpyAsync(ws.d_in, ws.h_img, (size_t)W * H * n_frames, cudaMemcpyHostToDevice, st));
  }
  rc = detect_run_device(ctx, cam, n_frames, ws.d_in, W);
  if (rc) return rc;
  OKB_CUDA(cudaMemcpyAsync(ws.h_count, ws.d_count, 4 * n_frames, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(cudaMemcpyAsync(ws.h_status, ws.d_status, 4 * n_frames, cudaMemcpyDeviceToHost, st));
  // keypoint / descriptor payload. The counts come back first (one extra synchronisation of a few microseconds): only the rows that
  // are filled in some frame of the batch cross PCIe (the device arrays are sized by the detector's capacity, ~30 % more than a
  // capped frame fills). Page-locked caller buffers receive them directly (rows k >= n_out[b] are then unspecified); otherwise they
  // land in the pinned staging and are trimmed on the host.
  OKB_CUDA(wait_stream(ctx, st));
  int max_n = 0;
  for (int b = 0; b < n_frames; b++) max_n = ws.h_count[b] > max_n ? ws.h_count[b] : max_n;
  const int rows_cap = cap < ws.kp_cap ? cap : ws.kp_cap;
  const int rows = max_n < rows_cap ? max_n : rows_cap;
  const bool out_pinned = rows_cap > 0 && host_pinned(kp_out) && host_pinned(desc_out);
  if (rows > 0) {
    if (out_pinned) {
      OKB_CUDA(cudaMemcpy2DAsync(kp_out, (size_t)cap * sizeof(okb_keypoint_t), ws.d_kp, (size_t)ws.kp_cap * sizeof(okb_keypoint_t),
                                 (size_t)rows * sizeof(okb_keypoint_t), n_frames, cudaMemcpyDeviceToHost, st));
      OKB_CUDA(cudaMemcpy2DAsync(desc_out, (size_t)cap * D, ws.d_desc, (size_t)ws.kp_cap * D, (size_t)rows * D, n_frames,
                                 cudaMemcpyDeviceToHost, st));
    } else {
      OKB_CUDA(cudaMemcpy2DAsync(ws.h_kp, (size_t)ws.kp_cap * sizeof(okb_keypoint_t), ws.d_kp, (size_t)ws.kp_cap * sizeof(okb_keypoint_t),
                                 (size_t)rows * sizeof(okb_keypoint_t), n_frames, cudaMemcpyDeviceToHost, st));
      OKB_CUDA(cudaMemcpy2DAsync(ws.h_desc, (size_t)ws.kp_cap * D, ws.d_desc, (size_t)ws.kp_cap * D, (size_t)rows * D, n_frames,
                                 cudaMemcpyDeviceToHost, st));
    }
  }
  ws.h_rays_frames = 0;
  if (ws.has_model) {   // the rays of D4 ride along (okb_last_back_projections): no second round trip for computeBackProjections
    if (rows > 0) {
      OKB_CUDA(cudaMemcpy2DAsync(ws.h_rays, (size_t)ws.kp_cap * 24, ws.d_rays, (size_t)ws.kp_cap * 24, (size_t)rows * 24, n_frames, cudaMemcpyDeviceToHost, st));
      OKB_CUDA(cudaMemcpy2DAsync(ws.h_rays_valid, (size_t)ws.kp_cap, ws.d_rays_valid, (size_t)ws.kp_cap, (size_t)rows, n_frames, cudaMemcpyDeviceToHost, st));
    }
    ws.h_rays_frames = n_frames;
  }
  OKB_CUDA(wait_stream(ctx, st));
  rc = status_to_error(ws, n_frames);
  if (rc) return rc;
  for (int b = 0; b < n_frames; b++) {
    int n = ws.h_count[b];
    if (n > cap) { set_error("okb_detect_describe: %d keypoints do not fit the caller's capacity %d", n, cap); return OKB_ERR_CAPACITY; }
    n_out[b] = n;
    if (out_pinned) continue;
    memcpy(kp_out + (size_t)b * cap, ws.h_kp + (size_t)b * ws.kp_cap, (size_t)n * sizeof(okb_keypoint_t));
    memcpy(desc_out + (size_t)b * cap * D, ws.h_desc + (size_t)b * ws.kp_cap * D, (size_t)n * D);
  }
  return OKB_OK;
}

int okb_detect_describe(okb_context_t* ctx, int cam, const uint8_t* image, size_t stride_bytes, okb_keypoint_t* kp_out,
                        uint8_t* desc_out, int cap, int* n_out)
{
  return okb_detect_describe_batch(ctx, cam, 1, image, stride_bytes, kp_out, desc_out, cap, n_out);
}

int okb_detect_describe_batch_device(okb_context_t* ctx, int cam, int n_frames, const uint8_t* d_images)
{
  int rc = check_cam(ctx, cam, "okb_detect_describe_batch_device");
  if (rc) return rc;
  CamWorkspace& ws = ctx->cams[cam];
  if (!d_images || n_frames < 1 || n_frames > ws.cfg.max_batch) { set_error("okb_detect_describe_batch_device: bad arguments"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  return detect_run_device(ctx, cam, n_frames, d_images, ws.cfg.width);
}

int okb_last_back_projections(okb_context_t* ctx, int cam, int frame, int cap, double* rays_out, uint8_t* valid_out, int* n_out)
{
  int rc = check_cam(ctx, cam, "okb_last_back_projections");
  if (rc) return rc;
  CamWorkspace& ws = ctx->cams[cam];
  if (!rays_out || !valid_out || !n_out || frame < 0 || frame >= ws.h_rays_frames) {
    set_error("okb_last_back_projections: no back-projections for frame %d (camera model set before okb_detect_describe?)", frame);
    return OKB_ERR_ARGUMENT;
  }
  const int n = ws.h_count[frame];
  if (n > cap) { set_error("okb_last_back_projections: %d rays > capacity %d", n, cap); return OKB_ERR_CAPACITY; }
  memcpy(rays_out, ws.h_rays + (size_t)frame * ws.kp_cap * 3, (size_t)n * 24);
  memcpy(valid_out, ws.h_rays_valid + (size_t)frame * ws.kp_cap, (size_t)n);
  *n_out = n;
  return OKB_OK;
}

int okb_fetch_features(okb_context_t* ctx, int cam, int frame, okb_keypoint_t* kp_out, uint8_t* desc_out, int cap, int* n_out)
{
  int rc = check_cam(ctx, cam, "okb_fetch_features");
  if (rc) return rc;
  CamWorkspace& ws = ctx->cams[cam];
  if (frame < 0 || frame >= ws.cfg.max_batch || !n_out) { set_error("okb_fetch_features: bad arguments"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  OKB_CUDA(cudaStreamSynchronize(ws.stream));
  int n = 0, status = 0;
  OKB_CUDA(cudaMemcpy(&n, ws.d_count + frame, 4, cudaMemcpyDeviceToHost));
  OKB_CUDA(cudaMemcpy(&status, ws.d_status + frame, 4, cudaMemcpyDeviceToHost));
  if (status) { ws.h_status[0] = status; return status_to_error(ws, 1); }
  if (n > cap) { set_error("okb_fetch_features: %d keypoints > capacity %d", n, cap); return OKB_ERR_CAPACITY; }
  *n_out = n;
  if (kp_out) OKB_CUDA(cudaMemcpy(kp_out, ws.d_kp + (size_t)frame * ws.kp_cap, (size_t)n * sizeof(okb_keypoint_t), cudaMemcpyDeviceToHost));
  if (desc_out) OKB_CUDA(cudaMemcpy(desc_out, ws.d_desc + (size_t)frame * ws.kp_cap * ws.cfg.descriptor_bytes, (size_t)n * ws.cfg.descriptor_bytes, cudaMemcpyDeviceToHost));
  return OKB_OK;
}

int okb_device_features(okb_context_t* ctx, int cam, const okb_keypoint_t** d_kp, const uint8_t** d_desc,
                        const int32_t** d_count, int* capacity)
{
  int rc = check_cam(ctx, cam, "okb_device_features");
  if (rc) return rc;
  CamWorkspace& ws = ctx->cams[cam];
  if (d_kp) *d_kp = ws.d_kp;
  if (d_desc) *d_desc = ws.d_desc;
  if (d_count) *d_count = ws.d_count;
  if (capacity) *capacity = ws.kp_cap;
  return OKB_OK;
}

int okb_device_back_projections(okb_context_t* ctx, int cam, const double** d_rays, const uint8_t** d_valid)
{
  int rc = check_cam(ctx, cam, "okb_device_back_projections");
  if (rc) return rc;
  CamWorkspace& ws = ctx->cams[cam];
  if (!ws.has_model) { set_error("okb_device_back_projections: camera %d has no model (okb_set_camera_model)", cam); return OKB_ERR_ARGUMENT; }
  if (d_rays) *d_rays = ws.d_rays;
  if (d_valid) *d_valid = ws.d_rays_valid;
  return OKB_OK;
}

size_t okb_feature_block_bytes(int n_frames, int capacity)
{
  const size_t counts = ((size_t)n_frames * 4 + 255) & ~(size_t)255;
  return counts + (size_t)n_frames * capacity * (sizeof(okb_keypoint_t) + 64);
}

int okb_export_features(okb_context_t* ctx, int cam, int n_frames, void* d_block)
{
  int rc = check_cam(ctx, cam, "okb_export_features");
  if (rc) return rc;
  CamWorkspace& ws = ctx->cams[cam];
  if (!d_block || n_frames < 1 || n_frames > ws.cfg.max_batch) { set_error("okb_export_features: bad arguments"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  uint8_t* p = (uint8_t*)d_block;
  const size_t counts = ((size_t)n_frames * 4 + 255) & ~(size_t)255;
  const size_t kp_bytes = (size_t)n_frames * ws.kp_cap * sizeof(okb_keypoint_t);
  OKB_CUDA(cudaMemcpyAsync(p, ws.d_count, (size_t)n_frames * 4, cudaMemcpyDeviceToDevice, ws.stream));
  OKB_CUDA(cudaMemcpyAsync(p + counts, ws.d_kp, kp_bytes, cudaMemcpyDeviceToDevice, ws.stream));
  OKB_CUDA(cudaMemcpyAsync(p + counts + kp_bytes, desc_slots(ws), (size_t)n_frames * ws.kp_cap * 64, cudaMemcpyDeviceToDevice, ws.stream));
  return OKB_OK;
}

int okb_num_layers(okb_context_t* ctx, int cam) { return check_cam(ctx, cam, "okb_num_layers") ? -1 : ctx->cams[cam].n_layers; }

int okb_layer_info(okb_context_t* ctx, int cam, int layer, int* width, int* height, float* scale, float* offset)
{
  int rc = check_cam(ctx, cam, "okb_layer_info");
  if (rc) return rc;
  CamWorkspace& ws = ctx->cams[cam];
  if (layer < 0 || layer >= ws.n_layers) { set_error("okb_layer_info: layer %d", layer); return OKB_ERR_ARGUMENT; }
  const LayerGeom& g = ws.geom[layer];
  if (width) *width = g.w; if (height) *height = g.h; if (scale) *scale = g.scale; if (offset) *offset = g.offset_px;
  return OKB_OK;
}

int okb_fetch_layer(okb_context_t* ctx, int cam, int frame, int layer, uint8_t* image_out, uint8_t* score_out)
{
  int rc = check_cam(ctx, cam, "okb_fetch_layer");
  if (rc) return rc;
  if (ctx->cams[cam].cfg.descriptor_bytes == 48) {   // the D = 48 mode has no pyramid and no AGAST score maps
    set_error("okb_fetch_layer: camera %d runs the Harris + BRISK2 mode (no scale-space layers)", cam); return OKB_ERR_UNSUPPORTED;
  }
  CamWorkspace& ws = ctx->cams[cam];
  if (layer < 0 || layer >= ws.n_layers || frame < 0 || frame >= ws.cfg.max_batch) { set_error("okb_fetch_layer: bad index"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  OKB_CUDA(cudaStreamSynchronize(ws.stream));
  const LayerGeom& g = ws.geom[layer];
  const size_t fo = (size_t)frame * ws.dl.frame_stride + g.offset;
  if (image_out) {
    if (layer == 0) OKB_CUDA(cudaMemcpy(image_out, ws.d_in + (size_t)frame * g.w * g.h, (size_t)g.w * g.h, cudaMemcpyDeviceToHost));
    else OKB_CUDA(cudaMemcpy2D(image_out, g.w, ws.d_img + fo, g.pitch, g.w, g.h, cudaMemcpyDeviceToHost));
  }
  if (score_out) OKB_CUDA(cudaMemcpy2D(score_out, g.w, ws.d_score + fo, g.pitch, g.w, g.h, cudaMemcpyDeviceToHost));
  return OKB_OK;
}

int okb_debug_stamps(okb_context_t* ctx, int cam, int frame, long long* out16)
{
  int rc = check_cam(ctx, cam, "okb_debug_stamps");
  if (rc) return rc;
  CamWorkspace& ws = ctx->cams[cam];
  if (frame < 0 || frame >= ws.cfg.max_batch || !out16) { set_error("okb_debug_stamps: bad arguments"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  OKB_CUDA(cudaStreamSynchronize(ws.stream));
  OKB_CUDA(cudaMemcpy(out16, ws.d_dbg + (size_t)frame * 16, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return OKB_OK;
}

int64_t okb_pyramid_score_bytes(okb_context_t* ctx, int cam) { return check_cam(ctx, cam, "okb_pyramid_score_bytes") ? -1 : ctx->cams[cam].ps_bytes; }

int okb_enable_timers(okb_context_t* ctx, int on) { if (!ctx) return OKB_ERR_ARGUMENT; ctx->timers_on = on; return OKB_OK; }
int okb_reset_timers(okb_context_t* ctx)
{
  if (!ctx) return OKB_ERR_ARGUMENT;
  for (auto& ws : ctx->cams) { if (ws.pending_timing) { cudaEventSynchronize(ws.ev[3]); ws.pending_timing = 0; } ws.ps_ms = ws.total_ms = ws.score_ms = 0; ws.ps_launches = 0; }
  return OKB_OK;
}
int okb_get_score_kernel_ms(okb_context_t* ctx, int cam, double* score_ms)
{
  int rc = check_cam(ctx, cam, "okb_get_score_kernel_ms");
  if (rc) return rc;
  detect_collect_timing(ctx, cam);
  if (score_ms) *score_ms = ctx->cams[cam].score_ms;
  return OKB_OK;
}

int okb_get_timers(okb_context_t* ctx, int cam, double* pyramid_score_ms, int64_t* pyramid_score_launches, double* total_ms)
{
  int rc = check_cam(ctx, cam, "okb_get_timers");
  if (rc) return rc;
  detect_collect_timing(ctx, cam);
  CamWorkspace& ws = ctx->cams[cam];
  if (pyramid_score_ms) *pyramid_score_ms = ws.ps_ms;
  if (pyramid_score_launches) *pyramid_score_launches = ws.ps_launches;
  if (total_ms) *total_ms = ws.total_ms;
  return OKB_OK;
}

}  // extern "C"
