// bench/e2e_driver.cpp -- the host loop a C++ integrator of libokvis_b200.so runs per stereo frame, used by bench.py for
// the `e2e` figure (HOST buffers in, HOST buffers out, every H2D/D2H copy inside the timed region). It mirrors the
// reference's per-frame driver: one host thread per camera for detect+describe
// (okvis_multisensor_processing/src/ThreadedSlam.cpp:432-448), then stereo matching (Frontend::matchStereo,
// Frontend.cpp:1982) and map matching (Frontend::matchToMap, Frontend.cpp:1171) from the main thread.
// Only the C ABI of include/okvis_b200.h is used.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../include/okvis_b200.h"

namespace {
struct Worker {  // persistent detection thread of one camera
  std::thread th; std::mutex m; std::condition_variable cv;
  bool has_job = false, done = false, quit = false;
  std::function<void()> job;
  void start() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(m);
      for (;;) {
        cv.wait(lk, [this] { return has_job || quit; });
        if (quit) return;
        lk.unlock(); job(); lk.lock();
        has_job = false; done = true; cv.notify_all();
      }
    });
  }
  void submit(std::function<void()> j) { std::lock_guard<std::mutex> lk(m); job = std::move(j); has_job = true; done = false; cv.notify_all(); }
  void wait() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [this] { return done; }); }
  void stop() { { std::lock_guard<std::mutex> lk(m); quit = true; cv.notify_all(); } th.join(); }
};
}  // namespace

// M3 inputs of the streaming loop (one camera): views + poses of ONE frame, list capacity
struct okb_stream_m3 { int32_t n_older, cap0, cap_m, pad_; const okb_older_view_t* older[2]; const double* T_WC1[2]; const double* T_CW1[2]; };

extern "C" int okb_e2e_run(okb_context_t* ctx, int n_frames, int warmup, int W, int H, const uint8_t* left, const uint8_t* right,
                           int cap, double f, const int* n_cand, const uint8_t* const* cand_desc, const int32_t* const* cand_lm,
                           const int* n_lm, const double* const* lm_proj, const uint8_t* const* lm_is3d, const okb_stream_m3* m3,
                           double* seconds, long long* h2d_bytes, long long* d2h_bytes, long long* total_kp, long long* total_matches)
{
  std::vector<uint8_t> matched[2]; std::vector<int32_t> m_n[2], m_k0[2], m_k1[2]; std::vector<uint8_t> m_fl[2]; std::vector<double> m_hp[2];
  const int n_older = m3 ? m3->n_older : 0;
  for (int c = 0; c < 2 && n_older > 0; c++) {
    matched[c].resize(cap); m_n[c].resize(n_older); m_k0[c].resize((size_t)n_older * m3->cap_m); m_k1[c].resize((size_t)n_older * m3->cap_m);
    m_fl[c].resize((size_t)n_older * m3->cap_m); m_hp[c].resize((size_t)n_older * m3->cap_m * 4);
  }
  std::vector<okb_keypoint_t> kp[2]; std::vector<uint8_t> desc[2]; int n[2] = {0, 0}; int rc2[2] = {0, 0};
  for (int c = 0; c < 2; c++) { kp[c].resize(cap); desc[c].resize((size_t)cap * 64); }
  std::vector<double> e[2], sof[2], xy[2]; std::vector<uint8_t> valid[2];
  std::vector<int32_t> k1(cap), lm(cap), lm_b(cap); std::vector<uint32_t> dist_b(cap); std::vector<uint32_t> dist(cap); std::vector<double> hp((size_t)cap * 4); std::vector<uint8_t> init(cap);
  const double r0[3] = {0, 0, 0}, r1[3] = {0.11, 0, 0};
  const double T0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, T1[12] = {1, 0, 0, -0.11, 0, 1, 0, 0, 0, 0, 1, 0};
  Worker w; w.start();
  long long h2d = 0, d2h = 0, nkp = 0, nm = 0;
  const uint8_t* imgs[2] = {left, right};
  auto frame = [&](int i) -> int {
    w.submit([&, i] { rc2[1] = okb_detect_describe(ctx, 1, imgs[1] + (size_t)i * W * H, W, kp[1].data(), desc[1].data(), cap, &n[1]); });
    rc2[0] = okb_detect_describe(ctx, 0, imgs[0] + (size_t)i * W * H, W, kp[0].data(), desc[0].data(), cap, &n[0]);
    w.wait();
    if (rc2[0] || rc2[1]) return rc2[0] ? rc2[0] : rc2[1];
    for (int c = 0; c < 2; c++) {  // Frame::computeBackProjections on the device, then the packing the matchers need
      e[c].resize((size_t)n[c] * 3); sof[c].resize(n[c]); xy[c].resize((size_t)n[c] * 2); valid[c].resize(n[c]);
      int nr = 0;   // rays were computed on the device during okb_detect_describe (camera model set)
      const int rcb = okb_last_back_projections(ctx, c, 0, n[c], e[c].data(), valid[c].data(), &nr);
      if (rcb) return rcb;
      for (int k = 0; k < n[c]; k++) {   // e_W = (C_WC * e_C).normalized() with C_WC = I
        const double x = e[c][3 * k], y = e[c][3 * k + 1], z = e[c][3 * k + 2];
        const double nn = sqrt((x * x + y * y) + z * z);
        e[c][3 * k] = x / nn; e[c][3 * k + 1] = y / nn; e[c][3 * k + 2] = z / nn;
        sof[c][k] = (double)kp[c][k].size / f; xy[c][2 * k] = kp[c][k].x; xy[c][2 * k + 1] = kp[c][k].y;
      }
      d2h += (long long)n[c] * 25;
    }
    // map matching of camera 1 on the worker thread (its own matcher slot/stream), stereo + camera 0 on this thread
    std::vector<int32_t>& lm1 = lm_b; std::vector<uint32_t>& dist1 = dist_b;
    int rc_m1 = 0;
    w.submit([&] {
      rc_m1 = okb_match_map3d(ctx, 64, n[1], desc[1].data(), xy[1].data(), nullptr, n_cand[1], cand_desc[1], cand_lm[1], n_lm[1],
                              lm_proj[1], lm_is3d[1], 20.0, 60, dist1.data(), lm1.data());
    });
    int rc = okb_match_stereo(ctx, 64, n[0], desc[0].data(), valid[0].data(), e[0].data(), sof[0].data(), n[1], desc[1].data(),
                              valid[1].data(), e[1].data(), sof[1].data(), r0, r1, T0, T1, 60, k1.data(), dist.data(), hp.data(), init.data());
    if (!rc) {
      for (int k = 0; k < n[0]; k++) nm += k1[k] >= 0;
      rc = okb_match_map3d(ctx, 64, n[0], desc[0].data(), xy[0].data(), nullptr, n_cand[0], cand_desc[0], cand_lm[0], n_lm[0],
                           lm_proj[0], lm_is3d[0], 20.0, 60, dist.data(), lm.data());
    }
    w.wait();
    if (rc || rc_m1) return rc ? rc : rc_m1;
    for (int k = 0; k < n[0]; k++) nm += lm[k] >= 0;
    for (int k = 0; k < n[1]; k++) nm += lm1[k] >= 0;
    if (n_older > 0) {   // M3 of both cameras (camera 1 on the worker thread), candidate set = keypoints without a landmark from M1
      auto motion = [&](int c, const int32_t* lmv) -> int {
        for (int k = 0; k < cap; k++) matched[c][k] = (k < n[c] && lmv[k] >= 0) ? 1 : 0;
        return okb_match_motion_stereo_batch(ctx, c, 1, m3->T_WC1[c], m3->T_CW1[c], n_older, m3->older[c], m3->cap0, 60, cap, matched[c].data(),
                                             m3->cap_m, m_n[c].data(), m_k0[c].data(), m_k1[c].data(), m_fl[c].data(), m_hp[c].data());
      };
      int rc_b = 0;
      w.submit([&] { rc_b = motion(1, lm1.data()); });
      const int rc_a = motion(0, lm.data());
      w.wait();
      if (rc_a || rc_b) return rc_a ? rc_a : rc_b;
      for (int c = 0; c < 2; c++) {
        for (int v = 0; v < n_older; v++) for (int j = 0; j < m_n[c][v]; j++) nm += (m_fl[c][(size_t)v * m3->cap_m + j] & 4) != 0;
        h2d += 192 + cap + (long long)n_older * (long long)sizeof(okb_older_view_t); d2h += (long long)n_older * (4 + (long long)m3->cap_m * 41) + cap;
      }
    }
    h2d += (long long)(n[0] + n[1]) * (64 + 24 + 8 + 2 * 8 + 1); d2h += (long long)n[0] * (4 + 4 + 32 + 1);
    for (int c = 0; c < 2; c++) {
      h2d += (long long)W * H + (long long)n[c] * (64 + 16) + (long long)n_cand[c] * 68 + (long long)n_lm[c] * 17;
      d2h += (long long)n[c] * (28 + 64 + 8);
      nkp += n[c];
    }
    return 0;
  };
  int rc = 0;
  for (int i = 0; i < warmup && !rc; i++) rc = frame(i % n_frames);
  h2d = d2h = nkp = nm = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n_frames && !rc; i++) rc = frame(i);
  if (!rc) rc = okb_sync(ctx);
  const auto t1 = std::chrono::steady_clock::now();
  w.stop();
  *seconds = std::chrono::duration<double>(t1 - t0).count();
  *h2d_bytes = h2d; *d2h_bytes = d2h; *total_kp = nkp; *total_matches = nm;
  return rc;
}

// ---- replay step (the same step as bench.py's device-resident `value` leg, but HOST buffers in and out): a batch of stereo
// frames per call, one host thread per camera: okb_detect_describe_batch + okb_match_map3d_batch, then
// okb_match_stereo_batch. All buffers are caller-provided (bench.py allocates them page-locked).
struct okb_replay_io {
  int32_t n_steps, warmup, batch, ring, W, H, cap, pad_;
  const uint8_t* img[2];                                  // ring * batch frames per camera
  okb_keypoint_t* kp[2]; uint8_t* desc[2]; int32_t* n[2];  // batch x cap
  int32_t n_cand[2]; int32_t n_lm[2];
  const uint8_t* cand_desc[2]; const int32_t* cand_lm[2]; const double* lm_proj[2]; const uint8_t* lm_is3d[2];
  uint32_t* m1_dist[2]; int32_t* m1_lm[2];
  int32_t* k1; uint32_t* sdist; double* hp; uint8_t* init;
  double seconds; long long h2d, d2h, nkp, nm;             // results
  int32_t lane, lanes;                                     // this replay is sequence `lane` of `lanes` concurrent ones (0, 0 = alone)
  // M3 (Frontend::matchMotionStereo): per camera the older keyframe views (device blocks: the keyframe feature store), poses,
  // the matched mask (host, derived from the M1 result) and the compact match lists
  int32_t n_older, cap0, cap_m, pad2_;
  const okb_older_view_t* older[2];                        // batch * n_older entries
  const double* T_WC1[2]; const double* T_CW1[2];          // batch x 12
  uint8_t* matched[2];                                     // batch x cap
  int32_t* n_match[2]; int32_t* m_k0[2]; int32_t* m_k1[2]; uint8_t* m_flags[2]; double* m_hp[2];
  long long n_m3;                                          // inserted M3 matches of the last step
};

namespace {
struct StartGate {   // all lanes finish their warm-up, then start the timed region together
  std::atomic<int> arrived{0}; int lanes = 1;
  std::chrono::steady_clock::time_point t0;
  std::mutex m; std::condition_variable cv;
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    if (++arrived == lanes) { t0 = std::chrono::steady_clock::now(); cv.notify_all(); }
    else cv.wait(lk, [this] { return arrived.load() >= lanes; });
  }
};
}  // namespace

static int replay_one(okb_context_t* ctx, okb_replay_io* io, StartGate* gate, std::chrono::steady_clock::time_point* t_end);

extern "C" int okb_e2e_replay(okb_context_t* ctx, okb_replay_io* io)
{
  StartGate gate; std::chrono::steady_clock::time_point t1;
  const int rc = replay_one(ctx, io, &gate, &t1);
  io->seconds = std::chrono::duration<double>(t1 - gate.t0).count();
  return rc;
}

// `lanes` independent sequences replayed concurrently on ONE GPU, each through its own library handle (own streams, own
// workspaces) and its own pair of host threads: while one lane's batch is in its kernels, the other lane's images are on
// the copy engine and its results on the way back. Wall clock from the common start to the last lane's okb_sync.
extern "C" int okb_e2e_replay_lanes(okb_context_t** ctx, okb_replay_io** io, int lanes, double* seconds)
{
  StartGate gate; gate.lanes = lanes;
  std::vector<std::thread> th; std::vector<int> rc(lanes, 0);
  std::vector<std::chrono::steady_clock::time_point> t1(lanes);
  for (int l = 0; l < lanes; l++) th.emplace_back([&, l] { rc[l] = replay_one(ctx[l], io[l], &gate, &t1[l]); });
  for (auto& t : th) t.join();
  auto last = t1[0];
  for (int l = 1; l < lanes; l++) if (t1[l] > last) last = t1[l];
  *seconds = std::chrono::duration<double>(last - gate.t0).count();
  for (int l = 0; l < lanes; l++) if (rc[l]) return rc[l];
  return 0;
}

static int replay_one(okb_context_t* ctx, okb_replay_io* io, StartGate* gate, std::chrono::steady_clock::time_point* t_end)
{
  const double C0[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, r0[3] = {0, 0, 0}, r1[3] = {0.11, 0, 0};
  const int B = io->batch, cap = io->cap;
  const size_t frame = (size_t)io->W * io->H;
  int rcs[2] = {0, 0};
  const bool trace = getenv("OKB_E2E_TRACE") != nullptr;
  double t_det[2] = {0, 0}, t_m1[2] = {0, 0}, t_m3[2] = {0, 0}, t_m4 = 0;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  auto camera = [&](int c, int s) {
    const int g = io->lanes > 1 ? s * io->lanes + io->lane : s;   // lanes replay interleaved parts of the ring
    const uint8_t* imgs = io->img[c] + (size_t)(g % io->ring) * B * frame;
    const auto ta = now();
    int rc = okb_detect_describe_batch(ctx, c, B, imgs, (size_t)io->W, io->kp[c], io->desc[c], cap, io->n[c]);
    const auto tb = now();
    t_det[c] += ms(ta, tb);
    if (!rc) rc = okb_match_map3d_batch(ctx, c, 64, B, io->n_cand[c], io->cand_desc[c], io->cand_lm[c], io->n_lm[c], io->lm_proj[c],
                                        io->lm_is3d[c], 20.0, 60, cap, io->m1_dist[c], io->m1_lm[c]);
    const auto tc = now();
    t_m1[c] += ms(tb, tc);
    if (!rc && io->n_older > 0) {
      // the keypoints that M1 gave a landmark leave the candidate set of M3 (Frontend.cpp:1792-1795)
      for (int b = 0; b < B; b++)
        for (int k = 0; k < cap; k++) io->matched[c][(size_t)b * cap + k] = (k < io->n[c][b] && io->m1_lm[c][(size_t)b * cap + k] >= 0) ? 1 : 0;
      rc = okb_match_motion_stereo_batch(ctx, c, B, io->T_WC1[c], io->T_CW1[c], io->n_older, io->older[c], io->cap0, 60, cap, io->matched[c],
                                         io->cap_m, io->n_match[c], io->m_k0[c], io->m_k1[c], io->m_flags[c], io->m_hp[c]);
      t_m3[c] += ms(tc, now());
    }
    rcs[c] = rc;
  };
  Worker w; w.start();
  auto step = [&](int s) -> int {
    w.submit([&, s] { camera(1, s); });
    camera(0, s);
    w.wait();
    if (rcs[0] || rcs[1]) return rcs[0] ? rcs[0] : rcs[1];
    const auto ta = now();
    const int rc4 = okb_match_stereo_batch(ctx, 0, 1, B, C0, r0, C0, r1, 60, cap, io->k1, io->sdist, io->hp, io->init);
    t_m4 += ms(ta, now());
    return rc4;
  };
  int rc = 0;
  for (int s = 0; s < io->warmup && !rc; s++) rc = step(s);
  if (!rc) rc = okb_sync(ctx);
  gate->wait();
  const auto t0 = gate->t0;
  for (int s = 0; s < io->n_steps && !rc; s++) rc = step(io->warmup + s);
  if (!rc) rc = okb_sync(ctx);
  const auto t1 = std::chrono::steady_clock::now();
  *t_end = t1;
  w.stop();
  io->seconds = std::chrono::duration<double>(t1 - t0).count();
  if (trace) {
    const int n = io->n_steps + io->warmup;
    fprintf(stderr, "[okb_e2e_replay] per step (ms): detect %.3f / %.3f  map3d %.3f / %.3f  motion %.3f / %.3f  stereo %.3f\n", t_det[0] / n, t_det[1] / n,
            t_m1[0] / n, t_m1[1] / n, t_m3[0] / n, t_m3[1] / n, t_m4 / n);
  }
  // bytes per step, counted from the copies the calls issue (the feature arrays: the rows filled in some frame of the batch, taken
  // from the last step; the matcher results: row stride = cap)
  long long h2d = 0, d2h = 0;
  for (int c = 0; c < 2; c++) {
    int rows = 0;
    for (int b = 0; b < B; b++) rows = io->n[c][b] > rows ? io->n[c][b] : rows;
    h2d += (long long)B * frame + (long long)io->n_cand[c] * 68 + (long long)B * io->n_lm[c] * 16 + io->n_lm[c];
    d2h += (long long)B * rows * (28 + 64 + 25) + 8LL * B + (long long)B * cap * 8;
  }
  d2h += (long long)B * cap * 41;
  if (io->n_older > 0)
    for (int c = 0; c < 2; c++) {
      h2d += (long long)B * 192 + (long long)B * cap + (long long)B * io->n_older * (long long)sizeof(okb_older_view_t);
      int max_cnt = 0;   // the compact lists come back trimmed to the longest one of the batch
      for (int i = 0; i < B * io->n_older; i++) max_cnt = io->n_match[c][i] > max_cnt ? io->n_match[c][i] : max_cnt;
      d2h += (long long)B * io->n_older * (4 + (long long)max_cnt * 41) + (long long)B * cap;
    }
  io->h2d = h2d; io->d2h = d2h;
  long long nkp = 0, nm = 0;   // of the last step (sanity numbers for the bench line)
  for (int c = 0; c < 2; c++)
    for (int b = 0; b < B; b++) {
      nkp += io->n[c][b];
      for (int k = 0; k < io->n[c][b]; k++) nm += io->m1_lm[c][(size_t)b * cap + k] >= 0;
    }
  for (int b = 0; b < B; b++)
    for (int k = 0; k < io->n[0][b]; k++) nm += io->k1[(size_t)b * cap + k] >= 0;
  long long n_m3 = 0;
  if (io->n_older > 0)
    for (int c = 0; c < 2; c++)
      for (int i = 0; i < B * io->n_older; i++)
        for (int j = 0; j < io->n_match[c][i] && j < io->cap_m; j++) n_m3 += (io->m_flags[c][(size_t)i * io->cap_m + j] & 4) != 0;
  io->nkp = nkp; io->nm = nm + n_m3; io->n_m3 = n_m3;
  return rc;
}

// ---- live use through okb_process_multiframe: one call per stereo frame (detect both cameras, M1, M3 sequence, M4), replayed
// as a CUDA graph from the second frame on. Host buffers in and out.
extern "C" int okb_e2e_multiframe(okb_context_t* ctx, int n_frames, int warmup, int W, int H, const uint8_t* left, const uint8_t* right, int cap,
                                  const int* n_cand, const uint8_t* const* cand_desc, const int32_t* const* cand_lm, const int* n_lm,
                                  const double* const* lm_proj, const uint8_t* const* lm_is3d, const okb_stream_m3* m3, double* seconds,
                                  double* worst_ms, long long* h2d_bytes, long long* d2h_bytes, long long* total_kp, long long* total_matches)
{
  const int n_older = m3 ? m3->n_older : 0, cap_m = m3 ? m3->cap_m : 1;
  std::vector<okb_keypoint_t> kp[2]; std::vector<uint8_t> desc[2], valid[2], m3f[2], init(cap);
  std::vector<double> rays[2], m3hp[2], hp((size_t)cap * 4);
  std::vector<uint32_t> m1d[2], sdist(cap); std::vector<int32_t> m1l[2], m3n[2], m3k0[2], m3k1[2], k1(cap);
  okb_multiframe_cam_t io[2]; memset(io, 0, sizeof(io));
  const uint8_t* imgs[2] = {left, right};
  for (int c = 0; c < 2; c++) {
    kp[c].resize(cap); desc[c].resize((size_t)cap * 64); rays[c].resize((size_t)cap * 3); valid[c].resize(cap); m1d[c].resize(cap); m1l[c].resize(cap);
    m3n[c].resize(n_older + 1); m3k0[c].resize((size_t)(n_older + 1) * cap_m); m3k1[c].resize((size_t)(n_older + 1) * cap_m);
    m3f[c].resize((size_t)(n_older + 1) * cap_m); m3hp[c].resize((size_t)(n_older + 1) * cap_m * 4);
    okb_multiframe_cam_t& q = io[c];
    q.stride_bytes = W; q.n_cand = n_cand[c]; q.n_lm = n_lm[c]; q.cand_desc = cand_desc[c]; q.cand_lm = cand_lm[c]; q.lm_is3d = lm_is3d[c]; q.lm_proj = lm_proj[c];
    if (n_older > 0) { q.T_WC1 = m3->T_WC1[c]; q.T_CW1 = m3->T_CW1[c]; q.n_older = n_older; q.cap0 = m3->cap0; q.older = m3->older[c]; }
    q.cap = cap; q.kp = kp[c].data(); q.desc = desc[c].data(); q.rays = rays[c].data(); q.rays_valid = valid[c].data();
    q.m1_dist = m1d[c].data(); q.m1_lm = m1l[c].data(); q.cap_m = cap_m; q.m3_n = m3n[c].data(); q.m3_k0 = m3k0[c].data(); q.m3_k1 = m3k1[c].data();
    q.m3_flags = m3f[c].data(); q.m3_hp_W = m3hp[c].data();
  }
  okb_multiframe_stereo_t st; memset(&st, 0, sizeof(st));
  st.cam0 = 0; st.cam1 = 1; st.C_WC0[0] = st.C_WC0[4] = st.C_WC0[8] = 1.0; st.C_WC1[0] = st.C_WC1[4] = st.C_WC1[8] = 1.0; st.r_WC1[0] = 0.11;
  st.k1 = k1.data(); st.dist = sdist.data(); st.hp_W = hp.data(); st.initialisable = init.data();
  long long nkp = 0, nm = 0; double worst = 0;
  auto frame = [&](int i, bool first) -> int {
    for (int c = 0; c < 2; c++) { io[c].image = imgs[c] + (size_t)i * W * H; io[c].pool_changed = first ? 1 : 0; }
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = okb_process_multiframe(ctx, 2, io, 1, &st, 20.0, 60);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ms > worst) worst = ms;
    if (rc) return rc;
    for (int c = 0; c < 2; c++) {
      nkp += io[c].n;
      for (int k = 0; k < io[c].n; k++) nm += m1l[c][k] >= 0;
      for (int v = 0; v < n_older; v++) for (int j = 0; j < m3n[c][v]; j++) nm += (m3f[c][(size_t)v * cap_m + j] & 4) != 0;
    }
    for (int k = 0; k < io[0].n; k++) nm += k1[k] >= 0;
    return 0;
  };
  int rc = 0;
  for (int i = 0; i < warmup && !rc; i++) rc = frame(i % n_frames, i == 0);
  nkp = nm = 0; worst = 0;
  { double tmp[4]; okb_stream_timing(ctx, tmp, 1); }   // host phase times of the timed frames only
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n_frames && !rc; i++) rc = frame(i, false);
  const auto t1 = std::chrono::steady_clock::now();
  *seconds = std::chrono::duration<double>(t1 - t0).count(); *worst_ms = worst;
  // bytes per frame of the copies inside okb_process_multiframe (row strides = device capacity `cap`)
  long long h2d = 0, d2h = 0;
  for (int c = 0; c < 2; c++) {
    h2d += (long long)W * H + (long long)n_lm[c] * 16 + (n_older > 0 ? 192 + (long long)n_older * (long long)sizeof(okb_older_view_t) : 0);
    d2h += 8 + (long long)cap * (28 + 64 + 25 + 8) + (n_older > 0 ? (long long)n_older * (4 + (long long)cap_m * 41) : 0);
  }
  d2h += (long long)cap * 41;
  *h2d_bytes = h2d * n_frames; *d2h_bytes = d2h * n_frames; *total_kp = nkp; *total_matches = nm;
  return rc;
}
