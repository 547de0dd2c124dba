// okb_harris.cu -- the D = 48 mode of detect + describe: Harris score + uniformity-enforcement detector and the 48-byte BRISK2
// extractor (camera-aware and aligned with the extraction direction when the camera-awareness maps are on the device), i.e. the pair
// Frontend::initialiseBriskFeatureDetectors constructs (reference okvis_frontend/src/Frontend.cpp:2406-2412) and
// Frontend::detectAndDescribe feeds (Frontend.cpp:232-251). octaves = 0 (every shipped okvis configuration, config/*.yaml).
//
// PARITY UNPINNED vs smartroboticslab/brisk@1ef8b42a (okb_harris_core.h). Bit-exact against oracle/brisk_oracle.c section 6.
//
// Kernels (one batch of frames per launch):
//   k_harris_score   64x16 tiles: u8 tile + 2-pixel ring in shared memory -> Scharr/32 gradients as char2 -> 3x3 binomial sums of the
//                    three products -> int32 score map (written once, 4 B per pixel)
//   k_harris_maxima  one thread per pixel: 8-neighbour test against the score map (L2), the row scan's skip rule resolved by the
//                    parity of the run of candidates to the left, append (score, x | y << 16)
//   k_uniformity     one CTA per frame: bitonic sort of the maxima by (score desc, y, x) in shared memory, 32-pixel cell lists, then the
//                    greedy uniformity enforcement as ROUNDS: a candidate is decided once every higher-ranked candidate whose stamp
//                    can reach its cell is decided (occupancy of its own cell = min(255, sum of accepted stamps)); stops as soon as
//                    max_keypoints are accepted in the decided prefix; sub-pixel refinement, border / warp validity, ordered
//                    compaction into cv::KeyPoint records
//   k_describe48     one warp per keypoint: 60 smoothed samples at the one pattern scale, placed by the per-keypoint 2x2 warp
//                    (camera-aware) or by the rotation table after the long-pair orientation (plain), 383 comparisons -> 12 words
#include "okb_harris_core.h"
#include "okb_internal.h"

namespace okb {

struct HarrisState {
  int32_t* d_score = nullptr;         // [B][H][W]
  uint2* d_cand = nullptr;            // [B][kHarrisCandCap]
  int32_t* d_sorted_score = nullptr;  // [B][kHarrisCandCap]
  float* d_lut = nullptr;             // 31 x 31 stamp weights
  uint32_t* d_short48 = nullptr;      // 384 packed pairs (i | j << 8), the last one (0, 0)
  int basic_scale = 0;
  int border = 0;
  float radius = 0.f;
};

void integral_run(CamWorkspace& ws, const uint8_t* d_images, int src_pitch, size_t in_stride, int W, int H, int B, cudaStream_t st);   // okb_detect.cu

// ---------------------------------------------------------------------------------------------------------------
constexpr int kHT_W = 64, kHT_H = 16;
__global__ void __launch_bounds__(256) k_harris_score(const uint8_t* in0, int pitch, size_t frame_stride, int W, int H, int32_t* score)
{
  __shared__ uint8_t img[kHT_H + 4][kHT_W + 8];
  __shared__ char2 grad[kHT_H + 2][kHT_W + 2];
  const int frame = blockIdx.z, x0 = blockIdx.x * kHT_W, y0 = blockIdx.y * kHT_H;
  const uint8_t* in = in0 + (size_t)frame * frame_stride;
  for (int i = threadIdx.x; i < (kHT_H + 4) * (kHT_W + 4); i += 256) {
    const int r = i / (kHT_W + 4), c = i % (kHT_W + 4);
    const int y = y0 - 2 + r, x = x0 - 2 + c;
    img[r][c] = (x >= 0 && x < W && y >= 0 && y < H) ? in[(size_t)y * pitch + x] : (uint8_t)0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (kHT_H + 2) * (kHT_W + 2); i += 256) {
    const int r = i / (kHT_W + 2), c = i % (kHT_W + 2);
    const int y = y0 - 1 + r, x = x0 - 1 + c;
    int gx = 0, gy = 0;
    if (x >= 1 && x <= W - 2 && y >= 1 && y <= H - 2) harris_grad(img[r], img[r + 1], img[r + 2], c + 1, gx, gy);
    grad[r][c] = make_char2((signed char)gx, (signed char)gy);
  }
  __syncthreads();
  int32_t* out = score + (size_t)frame * W * H;
  for (int i = threadIdx.x; i < kHT_H * kHT_W; i += 256) {
    const int oy = i / kHT_W, ox = i % kHT_W;
    const int y = y0 + oy, x = x0 + ox;
    if (x >= W || y >= H) continue;
    int s = 0;
    if (x >= 2 && x < W - 2 && y >= 2 && y < H - 2) {
      int a = 0, b = 0, c = 0;
#pragma unroll
      for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const char2 g = grad[oy + j][ox + k];
          const int w = (j == 1 ? 2 : 1) * (k == 1 ? 2 : 1);
          const int u = g.x, v = g.y;
          a += w * u * u; b += w * v * v; c += w * u * v;
        }
      s = harris_score(a, b, c);
    }
    out[(size_t)y * W + x] = s;
  }
}

__global__ void __launch_bounds__(128) k_harris_maxima(const int32_t* score, int W, int H, int threshold, uint2* cand, int32_t* count,
                                                       int count_stride, int32_t* status)
{
  const int frame = blockIdx.z;
  const int x = blockIdx.x * 128 + threadIdx.x + 2, y = blockIdx.y + 2;
  if (x >= W - 2 || y >= H - 2) return;
  const int32_t* sc = score + (size_t)frame * W * H;
  if (!harris_is_maximum(sc, W, x, y, threshold)) return;
  const int slot = atomicAdd(&count[frame * count_stride], 1);
  if (slot < kHarrisCandCap) cand[(size_t)frame * kHarrisCandCap + slot] = make_uint2((uint32_t)sc[(size_t)y * W + x], (uint32_t)x | ((uint32_t)y << 16));
  else atomicOr(&status[frame], 1);
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int kUniThreads = 1024;
constexpr int kUniMaxCells = 64 * 64;
constexpr int kUniWindow = 4096;   // ranks beyond the first undecided one that a round looks at
constexpr size_t kUniSmem = (size_t)kHarrisCandCap * 8 + kHarrisCandCap + (size_t)kHarrisCandCap * 2 + (size_t)(kUniMaxCells + 1) * 4 +
                            (size_t)kUniLut * kUniLut * 4 + (size_t)kPoints * sizeof(PatternPoint) + 64 * 4;

// exclusive prefix sum of one value per thread over the block; returns the thread's offset, *total = the block's sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums /*32 ints of shared memory*/, int* total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  __syncthreads();
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    warp_sums[lane] = w;
  }
  __syncthreads();
  const int base = warp ? warp_sums[warp - 1] : 0;
  *total = warp_sums[31];
  return base + incl - v;
}

__global__ void __launch_bounds__(kUniThreads) k_uniformity(const int32_t* score_maps, int W, int H, const uint2* cand, const int32_t* cand_count,
                                                            int count_stride, int32_t* sorted_score, const float* lut_g, const PatternPoint* pat0_g,
                                                            int max_kp, int kp_cap, int border, const float* ray_map, const float* jac_map,
                                                            float fu, float d0, float d1, float d2, okb_keypoint_t* kp_out, int32_t* count_out,
                                                            int32_t* status)
{
  extern __shared__ unsigned long long keys[];                                   // kHarrisCandCap
  uint8_t* state = reinterpret_cast<uint8_t*>(keys + kHarrisCandCap);            // 0 undecided, 1 accepted, 2 rejected, 3 kept + valid
  uint16_t* list = reinterpret_cast<uint16_t*>(state + kHarrisCandCap);          // ranks grouped by cell
  int* cell = reinterpret_cast<int*>(list + kHarrisCandCap);                     // kUniMaxCells + 1
  float* lut = reinterpret_cast<float*>(cell + kUniMaxCells + 1);                // 31 x 31
  PatternPoint* pat0 = reinterpret_cast<PatternPoint*>(lut + kUniLut * kUniLut); // 60
  int* sh = reinterpret_cast<int*>(pat0 + kPoints);                              // 32 scan words + scalars
  int* s_lo = sh + 32; int* s_acc = sh + 33; int* s_first = sh + 34;
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int n = min(cand_count[frame * count_stride], kHarrisCandCap);
  const uint2* cd = cand + (size_t)frame * kHarrisCandCap;
  const int32_t* smap = score_maps + (size_t)frame * W * H;
  int32_t* ss = sorted_score + (size_t)frame * kHarrisCandCap;
  okb_keypoint_t* kps = kp_out + (size_t)frame * kp_cap;
  for (int i = tid; i < kUniLut * kUniLut; i += kUniThreads) lut[i] = lut_g[i];
  for (int i = tid; i < kPoints * 3; i += kUniThreads) reinterpret_cast<float*>(pat0)[i] = reinterpret_cast<const float*>(pat0_g)[i];
  if (n == 0) { if (tid == 0) count_out[frame] = 0; return; }
  // ---- rank: (score desc, y asc, x asc)
  int P = 1024; while (P < n) P <<= 1;
  for (int i = tid; i < P; i += kUniThreads) {
    unsigned long long k = ~0ull;
    if (i < n) { const uint2 c = cd[i]; k = ((unsigned long long)(~c.x) << 32) | c.y; }
    keys[i] = k;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += kUniThreads) {
        const int q = i ^ j;
        if (q > i) {
          const unsigned long long a = keys[i], b = keys[q];
          if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[q] = a; }
        }
      }
      __syncthreads();
    }
  const float max_score = (float)(int)(~(uint32_t)(keys[0] >> 32));
  // ---- cells of 16 x 16 half-resolution positions (a stamp reaches +-15): counting sort of the ranks by cell
  const int cw = ((W - 1) / 2) / 16 + 1, ch = ((H - 1) / 2) / 16 + 1, n_cells = cw * ch;
  for (int i = tid; i <= n_cells; i += kUniThreads) cell[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += kUniThreads) {
    const unsigned long long k = keys[i];
    const int sc = (int)(~(uint32_t)(k >> 32));
    const uint32_t xy = (uint32_t)k;
    ss[i] = sc;
    keys[i] = ((unsigned long long)__float_as_uint(uni_nsc(uni_ratio(sc, max_score))) << 32) | xy;
    state[i] = 0;
    const int hx = (int)(xy & 0xffffu) >> 1, hy = (int)(xy >> 16) >> 1;
    atomicAdd(&cell[(hy >> 4) * cw + (hx >> 4)], 1);
  }
  __syncthreads();
  {
    const int per = (n_cells + kUniThreads - 1) / kUniThreads;   // <= 4
    int local[4], sum = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) { const int c = tid * per + q; local[q] = (q < per && c < n_cells) ? cell[c] : 0; sum += local[q]; }
    int total;
    int off = block_exclusive_scan(sum, sh, &total);
#pragma unroll
    for (int q = 0; q < 4; q++) { const int c = tid * per + q; if (q < per && c < n_cells) { cell[c] = off; off += local[q]; } }
  }
  __syncthreads();
  for (int i = tid; i < n; i += kUniThreads) {   // cell[c] advances to the END of cell c: afterwards cell c = [c ? cell[c - 1] : 0, cell[c])
    const uint32_t xy = (uint32_t)keys[i];
    const int hx = (int)(xy & 0xffffu) >> 1, hy = (int)(xy >> 16) >> 1;
    list[atomicAdd(&cell[(hy >> 4) * cw + (hx >> 4)], 1)] = (uint16_t)i;
  }
  if (tid == 0) { *s_lo = 0; *s_acc = 0; *s_first = n; }
  __syncthreads();
  // ---- rounds
  volatile uint8_t* vstate = state;
  int lo = 0, acc = 0;
  while (true) {
    const int hi = min(n, lo + kUniWindow);
    for (int i = lo + tid; i < hi; i += kUniThreads) {
      if (vstate[i]) continue;
      const unsigned long long ki = keys[i];
      const uint32_t xy = (uint32_t)ki;
      const int hx = (int)(xy & 0xffffu) >> 1, hy = (int)(xy >> 16) >> 1;
      const int cx = hx >> 4, cy = hy >> 4;
      int sum = 0; bool blocked = false;
      for (int yy = max(cy - 1, 0); yy <= min(cy + 1, ch - 1) && !blocked; yy++)
        for (int xx = max(cx - 1, 0); xx <= min(cx + 1, cw - 1) && !blocked; xx++) {
          const int c = yy * cw + xx;
          for (int t = c ? cell[c - 1] : 0, te = cell[c]; t < te; t++) {
            const int j = list[t];
            if (j >= i) continue;
            const unsigned long long kj = keys[j];
            const int dx = hx - ((int)((uint32_t)kj & 0xffffu) >> 1), dy = hy - ((int)((uint32_t)kj >> 16) >> 1);
            if (dx < -kUniWin || dx > kUniWin || dy < -kUniWin || dy > kUniWin) continue;
            const float l = lut[(dy + kUniWin) * kUniLut + dx + kUniWin];
            if (l == 0.0f) continue;
            const int sj = vstate[j];
            if (sj == 0) { blocked = true; break; }
            if (sj == 1) sum += uni_stamp(__uint_as_float((uint32_t)(kj >> 32)), l);
          }
        }
      if (blocked) { atomicMin(s_first, i); continue; }
      const float nsc = __uint_as_float((uint32_t)(ki >> 32));
      // ratio = score / max is recomputed from the sorted score (nsc is its fourth root)
      const float ratio = uni_ratio(ss[i], max_score);
      (void)nsc;
      vstate[i] = uni_rejected(ratio, sum) ? 2 : 1;
    }
    __syncthreads();
    const int first = min(*s_first, hi);   // every rank below `first` is decided
    int mine = 0;
    for (int i = lo + tid; i < first; i += kUniThreads) mine += state[i] == 1;
    if (mine) atomicAdd(s_acc, mine);
    __syncthreads();
    acc = *s_acc; lo = first;
    __syncthreads();
    if (tid == 0) *s_first = n;
    __syncthreads();
    if (lo >= n || (max_kp > 0 && acc >= max_kp)) break;
  }
  // ---- ranks [0, lo) are decided. Keep the first max_kp accepted ones, refine, test the pattern against the image border, compact.
  const int per = (lo + kUniThreads - 1) / kUniThreads;
  const int r0 = min(tid * per, lo), r1 = min(r0 + per, lo);
  int cnt = 0;
  for (int i = r0; i < r1; i++) cnt += state[i] == 1;
  int total_acc;
  int a = block_exclusive_scan(cnt, sh, &total_acc);
  const bool aware = ray_map != nullptr;
  const float d[3] = {d0, d1, d2};
  int valid = 0;
  for (int i = r0; i < r1; i++) {
    if (state[i] != 1) continue;
    const int idx = a++;
    if (max_kp > 0 && idx >= max_kp) { state[i] = 2; continue; }
    const uint32_t xy = (uint32_t)keys[i];
    const int x = (int)(xy & 0xffffu), y = (int)(xy >> 16);
    float dx, dy;
    harris_subpixel(smap, W, x, y, dx, dy);
    const float fx = (float)x + dx, fy = (float)y + dy;
    bool ok;
    if (!aware) {
      ok = !((fx < (float)border) || (fx >= (float)(W - border)) || (fy < (float)border) || (fy >= (float)(H - border)));
    } else {
      int u = (int)(fx + 0.5f), v = (int)(fy + 0.5f);
      u = u < 0 ? 0 : (u > W - 1 ? W - 1 : u); v = v < 0 ? 0 : (v > H - 1 ? H - 1 : v);
      float M[4];
      ok = brisk2_warp(ray_map + ((size_t)v * W + u) * 3, jac_map + ((size_t)v * W + u) * 6, d, fu, M);
      if (ok)
        for (int p = 0; p < kPoints; p++) {
          float xf, yf;
          brisk2_sample_pos(M, fx, fy, pat0[p], xf, yf);
          ok = ok && brisk2_sample_inside(xf, yf, pat0[p].sigma, W, H);
        }
    }
    if (ok) {
      state[i] = 3; valid++;
      keys[i] = ((unsigned long long)__float_as_uint(fy) << 32) | __float_as_uint(fx);
    } else {
      state[i] = 2;
    }
  }
  int total;
  int f = block_exclusive_scan(valid, sh, &total);
  for (int i = r0; i < r1; i++) {
    if (state[i] != 3) continue;
    const int idx = f++;
    if (idx >= kp_cap) continue;
    okb_keypoint_t k;
    k.x = __uint_as_float((uint32_t)keys[i]); k.y = __uint_as_float((uint32_t)(keys[i] >> 32));
    k.size = 12.0f; k.angle = -1.0f; k.response = (float)ss[i]; k.octave = 0; k.class_id = -1;
    kps[idx] = k;
  }
  if (tid == 0) {
    if (total > kp_cap) atomicOr(&status[frame], 8);
    count_out[frame] = min(total, kp_cap);
  }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_describe48(const uint8_t* in0, int in_pitch, size_t in_frame_stride, int W, int H, const int32_t* integral,
                                                    int ipitch, const PatternPoint* pattern /*[rot][point] of the one scale*/, const uint32_t* short48,
                                                    const int4* long_pairs, const float* ray_map, const float* jac_map, float fu, float d0, float d1,
                                                    float d2, okb_keypoint_t* kp, const int32_t* count, int kp_cap, uint8_t* desc)
{
  __shared__ int values[4][64];
  const int frame = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 4 + warp;
  if (k >= count[frame]) return;
  const uint8_t* img = in0 + (size_t)frame * in_frame_stride;
  const int32_t* I = integral + (size_t)frame * ipitch * (H + 1);
  okb_keypoint_t* kpp = kp + (size_t)frame * kp_cap + k;
  const float kx = kpp->x, ky = kpp->y;
  int* val = values[warp];
  float angle;
  if (ray_map) {
    int u = (int)(kx + 0.5f), v = (int)(ky + 0.5f);
    u = u < 0 ? 0 : (u > W - 1 ? W - 1 : u); v = v < 0 ? 0 : (v > H - 1 ? H - 1 : v);
    const float d[3] = {d0, d1, d2};
    float M[4];
    brisk2_warp(ray_map + ((size_t)v * W + u) * 3, jac_map + ((size_t)v * W + u) * 6, d, fu, M);
    for (int i = lane; i < kPoints; i += 32) {
      const PatternPoint p = pattern[i];
      float xf, yf;
      brisk2_sample_pos(M, kx, ky, p, xf, yf);
      val[i] = smoothed_intensity_at(img, in_pitch, I, ipitch, xf, yf, p.sigma);
    }
    angle = (float)(atan2((double)M[2], (double)M[0]) / 3.14159265358979323846 * 180.0);
    if (angle < 0) angle += 360.f;
  } else {
    for (int i = lane; i < kPoints; i += 32) val[i] = smoothed_intensity(img, in_pitch, I, ipitch, kx, ky, pattern[i]);
    __syncwarp();
    int e0 = 0, e1 = 0;
    for (int q = lane; q < kLongPairs; q += 32) {
      const int4 lp = __ldg(&long_pairs[q]);
      const int dt = val[lp.x] - val[lp.y];
      e0 += dt * lp.z / 1024;
      e1 += dt * lp.w / 1024;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { e0 += __shfl_xor_sync(0xffffffffu, e0, o); e1 += __shfl_xor_sync(0xffffffffu, e1, o); }
    angle = (float)(atan2((double)(float)e1, (double)(float)e0) / 3.14159265358979323846 * 180.0);
    int theta = (int)((double)kRot * ((double)angle / 360.0) + 0.5);
    if (theta < 0) theta += kRot;
    if (theta >= kRot) theta -= kRot;
    if (angle < 0) angle += 360.f;
    __syncwarp();
    const PatternPoint* pat = pattern + (size_t)theta * kPoints;
    for (int i = lane; i < kPoints; i += 32) val[i] = smoothed_intensity(img, in_pitch, I, ipitch, kx, ky, pat[i]);
  }
  __syncwarp();
  uint32_t mine = 0;
#pragma unroll
  for (int w = 0; w < 12; w++) {
    const uint32_t pr = __ldg(&short48[w * 32 + lane]);
    const uint32_t word = __ballot_sync(0xffffffffu, val[pr & 255] > val[pr >> 8]);
    if (lane == w) mine = word;
  }
  if (lane < 12) reinterpret_cast<uint32_t*>(desc + ((size_t)frame * kp_cap + k) * 48)[lane] = mine;
  if (lane == 0) kpp->angle = angle;
}

// ---------------------------------------------------------------------------------------------------------------
int harris_init_camera(okb_context* ctx, int cam)
{
  CamWorkspace& ws = ctx->cams[cam];
  const okb_camera_config_t& c = ws.cfg;
  HarrisState* hs = new HarrisState();
  ws.harris = hs;
  const int B = c.max_batch;
  hs->radius = c.uniformity_radius;
  hs->basic_scale = brisk2_basic_scale_host();
  if (hs->basic_scale >= kScales) { set_error("BRISK2 basic scale %d", hs->basic_scale); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaMalloc(&hs->d_score, (size_t)c.width * c.height * 4 * B));
  OKB_CUDA(cudaMalloc(&hs->d_cand, (size_t)kHarrisCandCap * sizeof(uint2) * B));
  OKB_CUDA(cudaMalloc(&hs->d_sorted_score, (size_t)kHarrisCandCap * 4 * B));
  float lut[kUniLut * kUniLut];
  for (int j = 0; j < kUniLut; j++) for (int i = 0; i < kUniLut; i++) lut[j * kUniLut + i] = uni_lut_host(hs->radius, i - kUniWin, j - kUniWin);
  OKB_CUDA(cudaMalloc(&hs->d_lut, sizeof(lut)));
  OKB_CUDA(cudaMemcpy(hs->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
  // short pairs below 5.1 x patternScale, in the (i, j < i) enumeration order of the pattern generator
  std::vector<uint32_t> sp;
  const float d_max = (float)(kDmax48 * ctx->pattern_scale);
  for (unsigned i = 1; i < (unsigned)kPoints; i++)
    for (unsigned j = 0; j < i; j++) {
      const float dx = ctx->h_pat0[j].x - ctx->h_pat0[i].x, dy = ctx->h_pat0[j].y - ctx->h_pat0[i].y;
      const float n2 = dx * dx + dy * dy;
      const float d_min = (float)(8.2 * ctx->pattern_scale);
      if (n2 > d_min * d_min) continue;
      if (n2 < d_max * d_max) sp.push_back(i | (j << 8));
    }
  if ((int)sp.size() != kShortPairs48) {
    set_error("pattern_scale %.3f gives %d short pairs below 5.1 x patternScale (383 expected)", ctx->pattern_scale, (int)sp.size());
    return OKB_ERR_UNSUPPORTED;
  }
  sp.resize(384, 0u);
  OKB_CUDA(cudaMalloc(&hs->d_short48, 384 * 4));
  OKB_CUDA(cudaMemcpy(hs->d_short48, sp.data(), 384 * 4, cudaMemcpyHostToDevice));
  hs->border = (int)ctx->h_size_list[hs->basic_scale];
  OKB_CUDA(cudaFuncSetAttribute(k_uniformity, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUniSmem));
  return OKB_OK;
}

void harris_free_camera(okb_context* ctx, int cam)
{
  CamWorkspace& ws = ctx->cams[cam];
  HarrisState* hs = (HarrisState*)ws.harris;
  if (!hs) return;
  cudaFree(hs->d_score); cudaFree(hs->d_cand); cudaFree(hs->d_sorted_score); cudaFree(hs->d_lut); cudaFree(hs->d_short48);
  delete hs;
  ws.harris = nullptr;
}

int harris_run_device(okb_context* ctx, int cam, int n_frames, const uint8_t* d_images, int src_pitch, cudaEvent_t input_ready)
{
  CamWorkspace& ws = ctx->cams[cam];
  HarrisState* hs = (HarrisState*)ws.harris;
  const okb_camera_config_t& c = ws.cfg;
  const int W = c.width, H = c.height, B = n_frames;
  cudaStream_t st = ws.stream;
  OKB_CUDA(cudaMemsetAsync(ws.d_cand_count, 0, ws.zero_bytes, st));
  if (input_ready) OKB_CUDA(cudaStreamWaitEvent(st, input_ready, 0));
  if (ctx->timers_on) cudaEventRecord(ws.ev[0], st);
  const size_t in_stride = (size_t)src_pitch * H;
  const int ipitch = W + 1;
  // the integral image only needs the frames: side stream, underneath the detector
  OKB_CUDA(cudaEventRecord(ws.ev_fork, st));
  OKB_CUDA(cudaStreamWaitEvent(ws.stream2, ws.ev_fork, 0));
  integral_run(ws, d_images, src_pitch, in_stride, W, H, B, ws.stream2);
  OKB_CUDA(cudaEventRecord(ws.ev_join, ws.stream2));
  k_harris_score<<<dim3((W + kHT_W - 1) / kHT_W, (H + kHT_H - 1) / kHT_H, B), 256, 0, st>>>(d_images, src_pitch, in_stride, W, H, hs->d_score);
  k_harris_maxima<<<dim3((W - 4 + 127) / 128, H - 4, B), 128, 0, st>>>(hs->d_score, W, H, c.threshold, hs->d_cand, ws.d_cand_count, kMaxLayers,
                                                                       ws.d_status);
  if (ctx->timers_on) { cudaEventRecord(ws.ev_mid, st); cudaEventRecord(ws.ev[1], st); }
  const bool aware = ws.maps_ready && ws.has_model;
  const float* rays = aware ? ws.d_ray_map : nullptr;
  const float* jac = aware ? ws.d_jac_map : nullptr;
  const float fu = aware ? (float)ws.model.fu : 1.0f;
  const PatternPoint* pat = ctx->d_pattern + (size_t)hs->basic_scale * kRot * kPoints;
  k_uniformity<<<B, kUniThreads, kUniSmem, st>>>(hs->d_score, W, H, hs->d_cand, ws.d_cand_count, kMaxLayers, hs->d_sorted_score, hs->d_lut, pat,
                                                 c.max_keypoints, ws.kp_cap, hs->border, rays, jac, fu, ws.extraction_dir[0], ws.extraction_dir[1],
                                                 ws.extraction_dir[2], ws.d_kp, ws.d_count, ws.d_status);
  if (ctx->timers_on) cudaEventRecord(ws.ev[2], st);
  OKB_CUDA(cudaStreamWaitEvent(st, ws.ev_join, 0));
  k_describe48<<<dim3((ws.kp_cap + 3) / 4, B), 128, 0, st>>>(d_images, src_pitch, in_stride, W, H, ws.d_integral, ipitch, pat, hs->d_short48,
                                                             ctx->d_long_pairs, rays, jac, fu, ws.extraction_dir[0], ws.extraction_dir[1],
                                                             ws.extraction_dir[2], ws.d_kp, ws.d_count, ws.kp_cap, ws.d_desc);
  ctx->launches += 6;
  { int rc = camera_backproject_batch(ctx, cam, B); if (rc) return rc; }
  if (ctx->timers_on) { cudaEventRecord(ws.ev[3], st); ws.pending_timing = 1; }
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

}  // namespace okb
