// okb_internal.h -- context / workspace definitions shared by the translation units of libokvis_b200.so
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/okvis_b200.h"
#include "okb_core.h"
#include "okb_tables.h"

namespace okb {

void set_error(const char* fmt, ...);

#define OKB_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      okb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return OKB_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

// the device-resident matcher forms and the live path read the camera's descriptor block with 64-byte rows
#define OKB_REQUIRE_D64(ws, who)                                                                                              \
  do {                                                                                                                        \
    if ((ws).cfg.descriptor_bytes != 64) {                                                                                    \
      okb::set_error("%s: device-resident form for 64-byte descriptors only (this camera has %d; use the host-buffer matcher " \
                     "forms, which take D = 48)", who, (ws).cfg.descriptor_bytes);                                           \
      return OKB_ERR_UNSUPPORTED;                                                                                             \
    }                                                                                                                         \
  } while (0)

struct LayerGeom {
  int w, h, pitch;
  size_t offset;  // byte offset of this layer inside one frame's layer block
  float scale, offset_px;
  int parent;     // layer it is sampled from (-1 for layer 0)
  int fast2;      // 1 = exact 2x2 mean, 0 = general area tables
  int max_taps = 0;   // general: largest tap count per axis (3 for the 2/3-sample layer)
  // device copies of the area tables (general path)
  int *d_xs = nullptr, *d_xn = nullptr, *d_ys = nullptr, *d_yn = nullptr;
  float *d_xa = nullptr, *d_ya = nullptr;
};

struct DeviceLayer {  // POD passed to kernels
  int w, h, pitch;
  uint32_t offset;
  float scale, offset_px;
};
struct DeviceLayers {
  int n;
  DeviceLayer l[kMaxLayers];
  uint32_t frame_stride;  // bytes of one frame's layer block
};

// refined candidate record kept on the device between the refine, resolve and finalize kernels

// a candidate whose 2-D maximum test ties with a neighbour, appended by k_refine for k_resolve: what its cache touches look like
// if it turns out to be a maximum (8 bytes) + its time key and record index
struct TieInfo {
  int8_t own_touch, has_above, exited, n_queries;
  int16_t max_x, max_y;
};
struct TieEntry { uint32_t key, rec; TieInfo info; };

// device scratch + page-locked table mirror of one M3 sequence (okb_match_motion_stereo_device*)
struct MotionScratch { void* d = nullptr; size_t cap = 0; void* h = nullptr; size_t h_cap = 0; int pinned_staging = 0; };

struct CamWorkspace {
  okb_camera_config_t cfg;
  uint8_t* m2_d = nullptr; size_t m2_cap = 0;   // scratch of okb_match_map_uninit_device (poses, world rays, use mask)
  void* harris = nullptr;                     // okb::HarrisState (okb_harris.cu): workspace of the D = 48 mode
  uint8_t* d_desc64 = nullptr;                // D = 48 only: the same rows in 64-byte slots with a zero tail (desc_slots)
  float extraction_dir[3] = {0.f, 0.f, -1.f};   // D1: gravity in the camera frame (okb_set_extraction_direction)
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;            // side stream (integral image)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_pyr = nullptr;
  int n_layers = 0;
  LayerGeom geom[kMaxLayers];
  DeviceLayers dl;
  int cand_cap = 0;   // candidate capacity per frame (sum of the per-layer regions)
  int cand_off[kMaxLayers + 1] = {0};
  int kp_cap = 0;     // keypoint capacity per frame (output rows)
  uint32_t* d_epoch = nullptr;   // [0] touch-map epoch (1..126), [1] wrap flag of the current call; advanced on the device
  // device buffers ([max_batch] leading dimension)
  uint8_t* d_in = nullptr;      // input frames (pitch = width) of the host-buffer entry points
  uint8_t* d_img = nullptr;     // layer images (layers >= 1)
  uint8_t* d_score = nullptr;   // thresholded score maps
  uint32_t* d_touch = nullptr;  // touch-time maps
  TieEntry* d_ties = nullptr;       // per frame: kMaxTies entries; d_tie_count (in the zero block) holds the fill
  int32_t* d_tie_count = nullptr;
  uint32_t* d_tie_sorted = nullptr; // per frame: the ties in time order, field by field (k_tie_gather -> k_resolve)
  uint32_t* d_tie_cells = nullptr;  // per frame: one bit per 16x16 cell of every layer, set around tied candidates
  int32_t* d_integral = nullptr;  // (w+1) x (h+1) per frame
  uint32_t* d_cand = nullptr;   // candidate keys
  int32_t* d_cand_count = nullptr;   // start of the per-call zero block: counts | status | tie cells (zero_bytes in total)
  size_t zero_bytes = 0;
  uint32_t* d_fkey = nullptr;   // per candidate: time key | keep << 30 | decided-maximum << 31 (k_refine, k_resolve -> k_finalize)
  void* d_fval = nullptr;       // per candidate: float4 x, y, size, response
  uint32_t* d_fslot = nullptr;  // k_finalize scratch for frames with more than 16384 candidates
  okb_keypoint_t* d_kp = nullptr;
  int32_t* d_kscale = nullptr;
  uint8_t* d_desc = nullptr;
  int32_t* d_count = nullptr;
  int32_t* d_status = nullptr;  // per frame overflow flags
  // M1 (device-resident form): keypoint grid and merge slots
  // D4: camera model and the rays of the last batch
  okb_camera_model_t model; int has_model = 0;
  double* d_rays = nullptr; uint8_t* d_rays_valid = nullptr;
  float* d_ray_map = nullptr; float* d_jac_map = nullptr; int maps_ready = 0;   // D5: camera-awareness maps (H x W x 3, H x W x 6)
  cudaEvent_t ev_done = nullptr;
  // TMA tensor maps of the internal layers (layer 0 is encoded per call)
  CUtensorMap tma[kMaxLayers]; int tma_use[kMaxLayers] = {0}; int tma_ready = 0;
  int score_tile_h = 32;          // rows per score tile: 32 (128-thread CTAs, default) or 64 (256-thread CTAs)
  void* d_tiles = nullptr; int n_tiles = 0, n_tiles0 = 0;   // (layer, x0, y0) of every score tile of a frame, layer-major; tiles of layer 0
  long long* d_dbg = nullptr;   // per-frame cycle stamps of the single-CTA kernels (okb_debug_stamps)
  uint8_t* d_m1_rows = nullptr; size_t m1_rows_cap = 0;   // M1 row bins of the device-resident form (grown on demand)
  // staging of the host-buffer batch matchers (okb_match_map3d_batch / okb_match_stereo_batch), grown on demand
  uint8_t* m_d = nullptr; uint8_t* m_h = nullptr; size_t m_cap = 0;
  uint8_t* m3_d = nullptr; uint8_t* m3_h = nullptr; size_t m3_cap = 0;   // staging of okb_match_motion_stereo_batch
  MotionScratch motion;
  // pinned staging
  uint8_t* h_img = nullptr;
  okb_keypoint_t* h_kp = nullptr;
  uint8_t* h_desc = nullptr;
  int32_t* h_count = nullptr;
  int32_t* h_status = nullptr;
  double* h_rays = nullptr; uint8_t* h_rays_valid = nullptr; int h_rays_frames = 0;
  // timers
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_mid = nullptr;
  double ps_ms = 0, total_ms = 0, score_ms = 0;
  int64_t ps_launches = 0;
  int pending_timing = 0;
  int64_t ps_bytes = 0;
};

// staging arena + stream of the host-buffer matchers. There are kMatchSlots of them; a calling thread is bound to one
// slot (round robin at its first call) and holds the slot's mutex during a call, so that e.g. the map matching of two
// cameras can be issued concurrently from two host threads (the reference spawns num_matching_threads workers).
constexpr int kMatchSlots = 4;
struct MatchWorkspace {
  cudaStream_t stream = nullptr;
  void* d_buf = nullptr; size_t d_cap = 0;
  void* h_buf = nullptr; size_t h_cap = 0;
  std::mutex* mtx = nullptr;
};

}  // namespace okb

namespace okb {
// The device-resident matcher forms (tensor-core Hamming scans of M3 / M4, the feature block of the camera-sharded exchange) read
// descriptor rows in 64-byte slots. A 48-byte row in such a slot with a zero tail has the same Hamming distances (the tails cancel),
// so a D = 48 camera keeps a second copy of its rows in that layout and those forms work on it unchanged.
inline const uint8_t* desc_slots(const CamWorkspace& ws) { return ws.cfg.descriptor_bytes == 48 ? ws.d_desc64 : ws.d_desc; }
}  // namespace okb

struct okb_context {
  int device = 0;
  int n_cams = 0;
  std::vector<okb::CamWorkspace> cams;
  okb::MatchWorkspace match_slots[okb::kMatchSlots];
  std::atomic<int> next_slot{0};
  // static tables
  okb::PatternPoint* d_pattern = nullptr;  // [scale][rot][point]
  uint32_t* d_short_pairs = nullptr;       // packed (i | j << 8)
  int4* d_long_pairs = nullptr;            // (i, j, wdx, wdy)
  int n_short = 0, n_long = 0;
  float* d_scale_bounds = nullptr;         // 63 float boundaries of the keypoint-size -> scale-index map
  uint32_t* d_size_list = nullptr;         // pattern extent per scale index
  float pattern_scale = 1.0f;
  okb::PatternPoint h_pat0[okb::kPoints];  // host copy: the pattern at scale 0, rotation 0 (pair classification)
  uint32_t h_size_list[okb::kScales];      // host copy of the pattern extents
  int timers_on = 0;
  void* stereo_scratch = nullptr; size_t stereo_cap = 0;   // device scratch of okb_match_stereo_device*
  okb::MotionScratch motion;   // scratch of okb_match_motion_stereo_device_ptr (one per context)
  int64_t launches = 0;
  int gate_cos_exact = 0;    // okb_create's self-check: gate_cos == this machine's libm cos on 65 536 arguments
  int blocking_sync = 0;     // 1: host-buffer entry points wait on a cudaEventBlockingSync event (the thread sleeps) instead of spinning
  void* prepare = nullptr;   // okb::PrepareState (okb_prepare.cu): keyframe feature store + P1 workspace
  void* aux = nullptr;       // okb::AuxState (okb_aux.cu): keyframe-overlap / BoW workspaces
  int32_t* d_scan_status = nullptr;   // one word: error flags of k_scan_umma (a timed-out barrier wait: never non-zero in a correct build)
  void* stream_state = nullptr;   // okb::StreamState (okb_stream.cu): arenas + CUDA graph of okb_process_multiframe
};

namespace okb {
int detect_init_camera(okb_context* ctx, int cam);
void detect_free_camera(okb_context* ctx, int cam);
int detect_run_device(okb_context* ctx, int cam, int n_frames, const uint8_t* d_images, int src_pitch, cudaEvent_t input_ready = nullptr);
int camera_backproject_batch(okb_context* ctx, int cam, int n_frames);
// D = 48 mode (okb_harris.cu)
int harris_init_camera(okb_context* ctx, int cam);
void harris_free_camera(okb_context* ctx, int cam);
int harris_run_device(okb_context* ctx, int cam, int n_frames, const uint8_t* d_images, int src_pitch, cudaEvent_t input_ready);
void harris_stage_direction(okb_context* ctx, int cam);   // live path: refresh the page-locked mirror of the extraction direction
struct Model;
int camera_stereo_prep_pair(okb_context* ctx, const okb_camera_model_t* const model[2], const double* const C_WC[2], const okb_keypoint_t* const d_kp[2],
                            const int32_t* const d_count[2], const int cap[2], int n_frames, double* const d_rays[2], uint8_t* const d_valid[2],
                            double* const d_eW[2], double* const d_sof[2], double* const d_c26[2], double* const d_c6[2],
                            unsigned long long* d_best, int32_t* d_hit_cnt, cudaStream_t st);
void k_backproject_ext(const Model& m, const okb_keypoint_t* d_kp, const int32_t* d_count, int cap, int n_frames, double* d_rays,
                       uint8_t* d_valid, cudaStream_t st);   // D4 on explicit device blocks (okb_camera.cu)
int camera_stereo_prep(okb_context* ctx, const okb_camera_model_t& model, const double C_WC[9], const okb_keypoint_t* d_kp,
                       const int32_t* d_count, int cap, int n_frames, double* d_rays, uint8_t* d_valid, double* d_eW, double* d_sof,
                       double* d_c26, double* d_c6, cudaStream_t st);
MatchWorkspace& match_ws(okb_context* ctx);   // the calling thread's slot
// true when `p` is page-locked host memory the copy engines can address (cudaHostAlloc / cudaHostRegister)
inline bool host_pinned(const void* p)
{
  cudaPointerAttributes at;
  if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}
// wait for `st` from a host-buffer entry point: spin (lowest latency) or sleep on a blocking event (okb_set_blocking_sync: for
// hosts with more waiting threads than cores, e.g. 8 ranks x 3 sequences x 3 threads)
inline cudaError_t wait_stream(okb_context* ctx, cudaStream_t st)
{
  if (!ctx->blocking_sync) return cudaStreamSynchronize(st);
  static thread_local cudaEvent_t ev = nullptr;
  static thread_local int ev_device = -1;
  if (ev_device != ctx->device) {
    if (ev) cudaEventDestroy(ev);
    ev = nullptr;
    const cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    ev_device = ctx->device;
  }
  const cudaError_t e = cudaEventRecord(ev, st);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ev);
}
// M3 sequence over the older keyframes with an explicit scratch area (okb_match.cu); the compaction of its matching entries
int match_map3d_enqueue(okb_context* ctx, int cam, int D, int n_frames, int n_cand, const uint8_t* d_cand_desc, const int32_t* d_cand_lm, int n_lm,
                        const double* d_lm_proj, const uint8_t* d_lm_is3d, double reprojection_threshold, uint32_t match_threshold,
                        uint32_t* d_out_dist, int32_t* d_out_lm, cudaStream_t bin_stream, cudaEvent_t bin_done);
int motion_restage(MotionScratch& ms, int n_frames, int n_older, const okb_older_view_t* older, int cap0, const double* T_WC1, const double* T_CW1);
int motion_sequence(okb_context* ctx, MotionScratch& ms, int n_frames, int cap1, const okb_keypoint_t* d_kp1, const uint8_t* d_desc1,
                    const int32_t* d_count1, const okb_camera_model_t* model, int width, int height, const double* T_WC1, const double* T_CW1,
                    int n_older, const okb_older_view_t* older, int cap0, uint32_t match_threshold, cudaStream_t st, uint8_t* d_matched1,
                    int32_t* d_out_k1, uint32_t* d_out_dist, double* d_out_hp_W, uint8_t* d_out_flags, const double* d_rays1, const uint8_t* d_valid1,
                    const int32_t* d_m1_lm = nullptr);
void m3_compact_launch(int n_frames, int cap0, int n_older, int cap_m, const int32_t* k1, const double* hp, const uint8_t* flags, int32_t* n_match,
                       int32_t* m_k0, int32_t* m_k1, uint8_t* m_flags, double* m_hp, cudaStream_t st);
void prepare_free(okb_context* ctx);
void stream_free(okb_context* ctx);
void aux_free(okb_context* ctx);
int tables_init(okb_context* ctx, float pattern_scale);
void tables_free(okb_context* ctx);
}  // namespace okb
