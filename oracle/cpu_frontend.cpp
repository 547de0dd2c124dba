/*
 * oracle/cpu_frontend.cpp -- TEST / BASELINE INFRASTRUCTURE ONLY (never on the product path).
 *
 * The CPU arm of the benchmark as BASELINE.md describes the reference's own front-end: plain C++ with std::thread, no
 * Python in the loop. Per multiframe (one synchronized set of camera images), like ThreadedSlam::processFrame
 * (reference okvis_multisensor_processing/src/ThreadedSlam.cpp:429-463,512-533):
 *   1. detect + describe, ONE std::thread PER CAMERA (ThreadedSlam.cpp:432-448)          -> oracle BRISK restatement
 *   2. Frame::computeBackProjections per camera (implementation/Frame.hpp:178-193)
 *   3. Frontend::matchToMap: M1 per camera with num_matching_threads workers (Frontend.cpp:1370-1385, k-range split)
 *   4. Frontend::matchMotionStereo: M3 against the older keyframes, num_matching_threads workers per view (:1807-1913),
 *      serial insertion between the views (:1915-1954)
 *   5. Frontend::matchStereo: M4, single-threaded (:2016-2074)
 * Two schedules: `workers == 1` = the reference configuration (multiframes one after the other, steps 1 and 3/4 threaded
 * as above; config/euroc.yaml:71 num_matching_threads = 4); `workers > 1` = all host cores: that many multiframes in flight,
 * each processed by one thread. Times with std::chrono::steady_clock per multiframe (reference timers "1 DetectAndDescribe",
 * "2.01 / 2.02 / 2.10").
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "oracle.h"

extern "C" {
void okvo_match_map3d(int D, int n_kp, const uint8_t* kp_desc, const double* kp_xy, const uint8_t* kp_use, int n_cand,
                      const uint8_t* cand_desc, const int32_t* cand_lm, int n_lm, const double* lm_proj, const uint8_t* lm_is3d,
                      double reprThreshold, uint32_t matchThreshold, uint32_t* out_dist, int32_t* out_lm, int n_threads);
void okvo_match_motion_stereo_sequence(int D, int n_views, const int32_t* n0, const int32_t* off, const uint8_t* desc0, const double* rays0,
                                       const uint8_t* valid0, const float* size0, const uint8_t* use0, const double* T_WC0,
                                       const double* T_CW0, int n1, const uint8_t* desc1, const double* rays1, const uint8_t* valid1,
                                       const float* xy1, const double T_WC1[12], const double T_CW1_in[12], int model, const double* intr,
                                       int width, int height, uint32_t matchThreshold, uint8_t* matched1, int32_t* out_k1,
                                       uint32_t* out_dist, double* out_hp_W, uint8_t* out_flags, int n_threads);
void okvo_match_stereo(int D, int n0, const uint8_t* desc0, const uint8_t* valid0, const double* e0_W, const double* size_over_f0, int n1,
                       const uint8_t* desc1, const uint8_t* valid1, const double* e1_W, const double* size_over_f1, const double r_WC0[3],
                       const double r_WC1[3], const double T_CW0[12], const double T_CW1[12], uint32_t matchThreshold, int32_t* out_k1,
                       uint32_t* out_dist, double* out_hp_W, uint8_t* out_initialisable, int n_threads);
void okvo_back_project(int model, double fu, double fv, double cu, double cv, const double k[4], int n, const float* kp_xy, int stride_floats,
                       double* rays, uint8_t* valid);
}

/* one camera of the workload */
typedef struct {
  const uint8_t* images;                 /* n_frames x H x W */
  int32_t model; int32_t pad_;
  double intr[8];                        /* fu fv cu cv k0..k3 */
  double T_WC[12], T_CW[12];             /* pose of the camera in the current frame (C row-major, r) and its inverse */
  /* landmark pool of M1 */
  int32_t n_cand, n_lm; const uint8_t* cand_desc; const int32_t* cand_lm; const double* lm_proj; const uint8_t* lm_is3d;
  /* older keyframe views of M3 (concatenated) */
  int32_t n_views, pad2_; const int32_t* v_n; const int32_t* v_off; const uint8_t* v_desc; const double* v_rays; const uint8_t* v_valid;
  const float* v_size; const uint8_t* v_use; const double* v_T_WC; const double* v_T_CW;
} okvo_cam_work_t;

typedef struct {
  int32_t W, H, threshold, octaves, max_kp, n_cams, n_frames, warmup;
  int32_t detect_threads_per_frame;   /* 1 = cameras one after the other, n_cams = one std::thread per camera */
  int32_t match_threads;              /* num_matching_threads */
  int32_t workers;                    /* multiframes in flight */
  int32_t stereo;                     /* 1 = M4 camera 0 -> camera 1 */
} okvo_frontend_cfg_t;

namespace {
struct CamOut {
  std::vector<okvo_keypoint_t> kp; std::vector<uint8_t> desc; int n = 0;
  std::vector<double> rays, eW, sof, xy; std::vector<float> xyf; std::vector<uint8_t> valid, matched;
  std::vector<uint32_t> m1_dist; std::vector<int32_t> m1_lm;
};

struct Worker {
  std::vector<okvo_brisk_t*> brisk;   // one detector object per camera (they hold the pyramid of the last image)
  std::vector<CamOut> out;
  std::vector<int32_t> k1; std::vector<uint32_t> dist; std::vector<double> hp; std::vector<uint8_t> fl;
};

long process(const okvo_frontend_cfg_t& c, const okvo_cam_work_t* cams, Worker& w, int frame, long* n_kp)
{
  const int cap = 1 << 16;   // raw detections before the cap
  const size_t img_bytes = (size_t)c.W * c.H;
  auto detect = [&](int cam) {
    CamOut& o = w.out[cam];
    if ((int)o.kp.size() < cap) { o.kp.resize(cap); o.desc.resize((size_t)cap * 64); }
    o.n = okvo_brisk_detect_and_compute(w.brisk[cam], cams[cam].images + (size_t)frame * img_bytes, c.W, c.H, c.W, c.max_kp, o.kp.data(), cap, o.desc.data());
    const int n = o.n;
    o.rays.resize((size_t)n * 3); o.valid.resize(n); o.eW.resize((size_t)n * 3); o.sof.resize(n); o.xy.resize((size_t)n * 2); o.xyf.resize((size_t)n * 2);
    const double* in = cams[cam].intr;
    okvo_back_project(cams[cam].model, in[0], in[1], in[2], in[3], in + 4, n, reinterpret_cast<const float*>(o.kp.data()), 7, o.rays.data(), o.valid.data());
    const double* C = cams[cam].T_WC; const double f = 0.5 * (in[0] + in[1]);
    for (int k = 0; k < n; k++) {
      const double x = o.rays[3 * k], y = o.rays[3 * k + 1], z = o.rays[3 * k + 2];
      const double wx = (C[0] * x + C[1] * y) + C[2] * z, wy = (C[3] * x + C[4] * y) + C[5] * z, wz = (C[6] * x + C[7] * y) + C[8] * z;
      const double nn = sqrt((wx * wx + wy * wy) + wz * wz);
      o.eW[3 * k] = wx / nn; o.eW[3 * k + 1] = wy / nn; o.eW[3 * k + 2] = wz / nn;
      o.sof[k] = (double)o.kp[k].size / f;
      o.xy[2 * k] = o.kp[k].x; o.xy[2 * k + 1] = o.kp[k].y; o.xyf[2 * k] = o.kp[k].x; o.xyf[2 * k + 1] = o.kp[k].y;
    }
  };
  if (c.detect_threads_per_frame > 1) {
    std::vector<std::thread> th;
    for (int cam = 0; cam < c.n_cams; cam++) th.emplace_back(detect, cam);
    for (auto& t : th) t.join();
  } else {
    for (int cam = 0; cam < c.n_cams; cam++) detect(cam);
  }
  long matches = 0;
  for (int cam = 0; cam < c.n_cams; cam++) {
    CamOut& o = w.out[cam]; const okvo_cam_work_t& cw = cams[cam];
    const int n = o.n;
    *n_kp += n;
    o.m1_dist.resize(n); o.m1_lm.resize(n); o.matched.assign(n, 0);
    okvo_match_map3d(64, n, o.desc.data(), o.xy.data(), nullptr, cw.n_cand, cw.cand_desc, cw.cand_lm, cw.n_lm, cw.lm_proj, cw.lm_is3d, 20.0, 60,
                     o.m1_dist.data(), o.m1_lm.data(), c.match_threads);
    for (int k = 0; k < n; k++) if (o.m1_lm[k] >= 0) { o.matched[k] = 1; matches++; }
    if (cw.n_views > 0) {
      size_t tot = 0; for (int v = 0; v < cw.n_views; v++) tot += cw.v_n[v];
      w.k1.resize(tot); w.dist.resize(tot); w.hp.resize(tot * 4); w.fl.resize(tot);
      okvo_match_motion_stereo_sequence(64, cw.n_views, cw.v_n, cw.v_off, cw.v_desc, cw.v_rays, cw.v_valid, cw.v_size, cw.v_use, cw.v_T_WC, cw.v_T_CW,
                                        n, o.desc.data(), o.rays.data(), o.valid.data(), o.xyf.data(), cw.T_WC, cw.T_CW, cw.model, cw.intr, c.W, c.H, 60,
                                        o.matched.data(), w.k1.data(), w.dist.data(), w.hp.data(), w.fl.data(), c.match_threads);
      for (size_t i = 0; i < tot; i++) matches += (w.fl[i] & 4) != 0;
    }
  }
  if (c.stereo && c.n_cams >= 2) {
    CamOut& a = w.out[0]; CamOut& b = w.out[1];
    auto t3x4 = [](const double* T, double* o) { for (int i = 0; i < 3; i++) { o[4 * i] = T[3 * i]; o[4 * i + 1] = T[3 * i + 1]; o[4 * i + 2] = T[3 * i + 2]; o[4 * i + 3] = T[9 + i]; } };
    double T0[12], T1[12]; t3x4(cams[0].T_CW, T0); t3x4(cams[1].T_CW, T1);
    w.k1.resize(a.n); w.dist.resize(a.n); w.hp.resize((size_t)a.n * 4); w.fl.resize(a.n);
    okvo_match_stereo(64, a.n, a.desc.data(), a.valid.data(), a.eW.data(), a.sof.data(), b.n, b.desc.data(), b.valid.data(), b.eW.data(), b.sof.data(),
                      cams[0].T_WC + 9, cams[1].T_WC + 9, T0, T1, 60, w.k1.data(), w.dist.data(), w.hp.data(), w.fl.data(), 1);
    for (int k = 0; k < a.n; k++) matches += w.k1[k] >= 0;
  }
  return matches;
}
}  // namespace

/* Processes frames [0, n_frames) after `warmup` untimed ones (frames are reused cyclically for the warm-up).
 * per_frame_ms: n_frames entries (wall time of each multiframe inside its worker). Returns 0. */
extern "C" int okvo_frontend_run(const okvo_frontend_cfg_t* cfg, const okvo_cam_work_t* cams, double* per_frame_ms, double* total_s,
                                 long* n_kp, long* n_matches)
{
  const okvo_frontend_cfg_t c = *cfg;
  const int W = c.workers < 1 ? 1 : c.workers;
  std::vector<Worker> ws(W);
  for (auto& w : ws) {
    w.brisk.resize(c.n_cams); w.out.resize(c.n_cams);
    for (int cam = 0; cam < c.n_cams; cam++) w.brisk[cam] = okvo_brisk_create(c.threshold, c.octaves, 1.0f);
  }
  std::atomic<int> next{0}; std::atomic<long> kp{0}, mt{0};
  auto run = [&](int wi, int first, int count, bool timed) {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= count) return;
      const auto t0 = std::chrono::steady_clock::now();
      long k = 0;
      const long m = process(c, cams, ws[wi], (first + i) % c.n_frames, &k);
      const auto t1 = std::chrono::steady_clock::now();
      if (timed) { per_frame_ms[i] = std::chrono::duration<double, std::milli>(t1 - t0).count(); kp += k; mt += m; }
    }
  };
  auto phase = [&](int count, bool timed) {
    next = 0;
    std::vector<std::thread> th;
    for (int wi = 1; wi < W; wi++) th.emplace_back(run, wi, 0, count, timed);
    run(0, 0, count, timed);
    for (auto& t : th) t.join();
  };
  phase(c.warmup, false);
  const auto t0 = std::chrono::steady_clock::now();
  phase(c.n_frames, true);
  const auto t1 = std::chrono::steady_clock::now();
  *total_s = std::chrono::duration<double>(t1 - t0).count();
  *n_kp = kp; *n_matches = mt;
  for (auto& w : ws) for (auto* b : w.brisk) okvo_brisk_destroy(b);
  return 0;
}
