// okb_harris.cu -- the D = 48 mode of detect + describe: Harris score + uniformity-enforcement detector and the 48-byte BRISK2
// extractor (camera-aware and aligned with the extraction direction when the camera-awareness maps are on the device), i.e. the pair
// Frontend::initialiseBriskFeatureDetectors constructs (reference okvis_frontend/src/Frontend.cpp:2406-2412) and
// Frontend::detectAndDescribe feeds (Frontend.cpp:232-251). octaves = 0 (every shipped okvis configuration, config/*.yaml).
//
// PARITY UNPINNED vs smartroboticslab/brisk@1ef8b42a (okb_harris_core.h). Bit-exact against oracle/brisk_oracle.c section 6.
//
// Kernels (one batch of frames per launch):
//   k_harris_score   a warp per 32-column strip marching down a band of rows, everything in registers + warp shuffles: pixel row ->
//                    Scharr/32 gradients -> 3x3 binomial sums of the three products -> int32 score row (stored once) -> flags of the
//                    8-neighbour test one row behind (stored as bytes)
//   k_harris_maxima  one thread per 16 flags: the row scan's skip rule resolved by the parity of the run of candidates to the left,
//                    append (score, x | y << 16)
//   k_uniformity     one CTA per frame: bitonic sort of the maxima by (score desc, y, x) in shared memory, 32-pixel cell lists, then the
//                    greedy uniformity enforcement as WAVES: every candidate counts the higher-ranked candidates whose stamp can
//                    reach its cell; a candidate whose count is zero is decided (occupancy of its own cell = min(255, sum of accepted
//                    stamps)) and pushes its decision to the lower-ranked ones in reach with one shared-memory atomic each (stamp
//                    added, count decremented); stops as soon as max_keypoints are accepted in the decided prefix; sub-pixel
//                    refinement, border / warp validity, ordered compaction into cv::KeyPoint records
//   k_describe48     one warp per keypoint: 60 smoothed samples at the one pattern scale, placed by the per-keypoint 2x2 warp
//                    (camera-aware) or by the rotation table after the long-pair orientation (plain), 383 comparisons -> 12 words
#include <type_traits>

#include "okb_harris_core.h"
#include "okb_internal.h"

namespace okb {

struct HarrisState {
  int32_t* d_score = nullptr;         // [B][H][W]
  uint8_t* d_cond = nullptr;          // [B][H][cpitch] maximum-candidate flags
  int cpitch = 0;
  uint2* d_cand = nullptr;            // [B][kHarrisCandCap]
  int32_t* d_sorted_score = nullptr;  // [B][kHarrisCandCap]
  uint32_t* d_sorted_xy = nullptr;    // [B][kHarrisCandCap]
  float* d_lut = nullptr;             // 31 x 31 stamp weights
  uint32_t* d_short48 = nullptr;      // 384 packed pairs (i | j << 8), the last one (0, 0)
  int basic_scale = 0;
  int border = 0;
  float radius = 0.f;
  float* d_dir = nullptr;             // live path: the extraction direction on the device (a captured graph cannot carry it as an argument)
  float* h_dir = nullptr;             //            and its page-locked mirror
  int dir_from_device = 0;
  int cshift = 4;                     // log2 of the uniformity kernel's cell size in half-resolution positions
};

void integral_run(CamWorkspace& ws, const uint8_t* d_images, int src_pitch, size_t in_stride, int W, int H, int B, cudaStream_t st);   // okb_detect.cu

// ---------------------------------------------------------------------------------------------------------------
// Score + maximum-candidate pass. A warp owns a strip of 32 columns (26 of them produce output: the derivative, the horizontal
// binomial sum and the neighbour test each cost one column on either side) and marches down a band of rows; everything between the
// pixel load and the two stores lives in registers, horizontal neighbours come from warp shuffles, vertical ones from the previous
// rows' registers. Per row step: pixel row r -> gradient row r-1 -> horizontal sums of the three products -> score row r-2
// (stored) -> candidate flags of row r-3 (stored as bytes: at or above the threshold and no 8-neighbour strictly greater).
constexpr int kHsCols = 26, kHsWarps = 4;   // rows per band: a launch parameter (60 for batches; 20 for a single frame, whose ~230 warps of
                                            // 66 dependent row steps would otherwise be one long latency chain)
__global__ void __launch_bounds__(32 * kHsWarps) k_harris_score(const uint8_t* __restrict__ in0, int pitch, size_t frame_stride, int W, int H,
                                                                 int threshold, int32_t* __restrict__ score, uint8_t* __restrict__ cond, int cpitch, int band)
{
  const int lane = threadIdx.x & 31;
  const int strip = blockIdx.x * kHsWarps + (threadIdx.x >> 5);
  const int x = strip * kHsCols - 3 + lane;
  if (strip * kHsCols >= W) return;
  const int yb = blockIdx.y * band, ye = min(yb + band, H);
  const int frame = blockIdx.z;
  const uint8_t* in = in0 + (size_t)frame * frame_stride;
  int32_t* out = score + (size_t)frame * W * H;
  uint8_t* cnd = cond + (size_t)frame * cpitch * H;
  const bool x_in = x >= 0 && x < W;
  const bool x_grad = x >= 1 && x <= W - 2, x_score = x >= 2 && x < W - 2;
  const bool x_store = lane >= 3 && lane < 3 + kHsCols && x < W;
  int I0 = 0, I1 = 0;                       // pixel rows r-2, r-1
  int hA0 = 0, hA1 = 0, hB0 = 0, hB1 = 0, hC0 = 0, hC1 = 0;   // horizontal sums of gradient rows r-3, r-2
  int S1 = 0;                               // score row r-3
  int M0 = 0, side1 = 0;                    // row maximum over (x-1, x, x+1) of score row r-4; max(x-1, x+1) of score row r-3
  // rows yb-3 .. ye+2 are loaded; the first steps only fill the pipeline. kFull: every row condition of the step holds (rows inside the
  // image with their margins, both stores inside the band), which is the case for all but the first six and the last one to three steps of
  // a band. The candidate test needs no position check at all: scores are zero where they are not defined and
  // the threshold is at least 1.
  auto step = [&](const int r, auto full) {
    constexpr bool kFull = decltype(full)::value;
    const int I2 = (x_in && (kFull || (r >= 0 && r < H))) ? (int)in[(size_t)r * pitch + x] : 0;
    // gradient row g = r - 1
    const int g = r - 1;
    const int v = 3 * (I0 + I2) + 10 * I1, dv = I2 - I0;
    const int pk = (v << 16) | (dv & 0xffff);
    const int pkL = __shfl_up_sync(0xffffffffu, pk, 1), pkR = __shfl_down_sync(0xffffffffu, pk, 1);
    const int sx = (pkR >> 16) - (pkL >> 16);
    const int sy = 3 * ((int)(short)(pkL & 0xffff) + (int)(short)(pkR & 0xffff)) + 10 * dv;
    int gpk = 0;
    if (x_grad && (kFull || (g >= 1 && g <= H - 2))) gpk = ((sx >> 5) & 0xff) | (((sy >> 5) & 0xff) << 8);
    // horizontal binomial sums of the three products as int8 dot products: (gL, g, g, gR) . (gL, g, g, gR) = gL^2 + 2 g^2 + gR^2
    const int gL = __shfl_up_sync(0xffffffffu, gpk, 1), gR = __shfl_down_sync(0xffffffffu, gpk, 1);
    const int P = (int)__byte_perm(__byte_perm(gL, gpk, 0x0440), gR, 0x4210);
    const int Q = (int)__byte_perm(__byte_perm(gL, gpk, 0x1551), gR, 0x5210);
    const int hA2 = __dp4a(P, P, 0), hB2 = __dp4a(Q, Q, 0), hC2 = __dp4a(P, Q, 0);
    // score row s = r - 2
    const int srow = r - 2;
    int S2 = 0;
    if (x_score && (kFull || (srow >= 2 && srow < H - 2))) S2 = harris_score(hA0 + 2 * hA1 + hA2, hB0 + 2 * hB1 + hB2, hC0 + 2 * hC1 + hC2);
    if (x_store && (kFull || (srow >= yb && srow < ye))) out[(size_t)srow * W + x] = S2;
    const int S2L = __shfl_up_sync(0xffffffffu, S2, 1), S2R = __shfl_down_sync(0xffffffffu, S2, 1);
    const int side2 = max(S2L, S2R), M2 = max(side2, S2);
    // candidate flags of row m = r - 3
    const int m = r - 3;
    if (x_store && (kFull || (m >= yb && m < ye))) cnd[(size_t)m * cpitch + x] = (S1 >= threshold && max(max(M0, M2), side1) <= S1) ? 1 : 0;
    I0 = I1; I1 = I2; hA0 = hA1; hA1 = hA2; hB0 = hB1; hB1 = hB2; hC0 = hC1; hC1 = hC2;
    M0 = max(side1, S1); S1 = S2; side1 = side2;
  };
  using Yes = std::true_type; using No = std::false_type;
  const int r_lo = max(yb + 3, 4), r_hi = min(ye + 1, H - 1);   // the steps in [r_lo, r_hi] need no row checks
  for (int r = yb - 3; r < r_lo; r++) step(r, No());
#pragma unroll 2
  for (int r = r_lo; r <= r_hi; r++) step(r, Yes());
  for (int r = max(r_hi + 1, r_lo); r <= ye + 2; r++) step(r, No());
}

// maxima from the candidate flags: inside a run of consecutive candidates of a row every second one is kept (the row scan skips the
// pixel right of an accepted maximum). One thread per 16 flags.
__global__ void __launch_bounds__(256) k_harris_maxima(const int32_t* score, const uint8_t* cond, int cpitch, int W, int H, uint2* cand,
                                                       int32_t* count, int count_stride, int32_t* status)
{
  __shared__ int s_n, s_base;
  __shared__ uint2 s_c[256 * 8];   // 16 flags hold at most 8 maxima (every second pixel of a run)
  const int frame = blockIdx.y;
  const int chunks = cpitch >> 4;
  const int id = blockIdx.x * 256 + threadIdx.x;
  const int y = id / chunks, x0 = (id - y * chunks) * 16;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  if (y < H) {
    const uint8_t* row = cond + ((size_t)frame * H + y) * cpitch;
    const uint4 f = *reinterpret_cast<const uint4*>(row + x0);
    if ((f.x | f.y | f.z | f.w) != 0u) {
      const uint32_t wds[4] = {f.x, f.y, f.z, f.w};
      const int32_t* sc = score + (size_t)frame * W * H + (size_t)y * W;
      for (int i = 0; i < 16; i++) {
        if (!((wds[i >> 2] >> (8 * (i & 3))) & 1u)) continue;
        const int x = x0 + i;
        int run = 0;
        while (x - 1 - run >= 2 && row[x - 1 - run]) run++;
        if (run & 1) continue;
        s_c[atomicAdd(&s_n, 1)] = make_uint2((uint32_t)sc[x], (uint32_t)x | ((uint32_t)y << 16));
      }
    }
  }
  __syncthreads();
  const int n = s_n;
  if (n == 0) return;
  if (threadIdx.x == 0) s_base = atomicAdd(&count[frame * count_stride], n);   // one global atomic per CTA: the per-frame counter is one address
  __syncthreads();
  const int base = s_base;
  for (int i = threadIdx.x; i < n; i += 256) {
    if (base + i < kHarrisCandCap) cand[(size_t)frame * kHarrisCandCap + base + i] = s_c[i];
    else atomicOr(&status[frame], 1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int kUniThreads = 1024;
constexpr int kUniMaxCells = 64 * 64;
// shared memory: [sort keys (8 B per candidate) | afterwards: cell-ordered entries + per-candidate words (4 + 4 B)] state, ready queue,
// cell ends, stamp weights, the pattern at rotation 0, scalars
constexpr size_t kUniSmem = (size_t)kHarrisCandCap * 8 + kHarrisCandCap + (size_t)kHarrisCandCap * 2 + (size_t)(kUniMaxCells + 1) * 4 +
                            (size_t)kUniLut * kUniLut * 4 + (size_t)kPoints * sizeof(PatternPoint) + 64 * 4;
constexpr int kPendShift = 20;   // word = pending higher-ranked neighbours << 20 | sum of the accepted neighbours' stamps
// at most 31 x 31 maxima share a 62 x 62 pixel window (no two are 8-neighbours): 961 < 2^12 pending, 961 x 255 < 2^20 of stamps

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

// The neighbours of a candidate at half-resolution position (hx, hy) whose stamps can reach it (and which it can reach) lie in the
// 3 x 3 cells of (1 << cshift)^2 positions around its cell (cshift = 3 when no stamp reaches further than 8 positions, else 4); the cells of one cell row are contiguous in the cell-ordered entry list.
// entry = rank << 14 | hx << 4 | (hy & 15). f(rank, dx, dy) is called by the kUniGroup lanes that share the candidate (sub = lane
// within the group) for every entry in the window. A full warp per candidate leaves the kernel latency-bound (a dependent chain of
// shared-memory loads and one atomic per entry, ~40 entries per candidate): groups of 8 keep 128 candidates in flight per CTA.
constexpr int kUniGroup = 8;                          // lanes that share one candidate
constexpr int kUniGroups = kUniThreads / kUniGroup;   // candidates in flight per CTA
template <typename F>
__device__ __forceinline__ void for_each_neighbour(const uint32_t* entries, const int* cell_end, int cw, int ch, int cshift, int hx, int hy, int sub, F f)
{
  const int cx = hx >> cshift, cy = hy >> cshift;
  const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, cw - 1);
  for (int yy = max(cy - 1, 0); yy <= min(cy + 1, ch - 1); yy++) {
    const int c0 = yy * cw + x_lo, c1 = yy * cw + x_hi;
    const int t0 = c0 ? cell_end[c0 - 1] : 0, t1 = cell_end[c1];
    for (int t = t0 + sub; t < t1; t += kUniGroup) {
      const uint32_t e = entries[t];
      const int dx = (int)((e >> 4) & 1023u) - hx, dy = (yy << cshift) + (int)(e & 15u & ((1u << cshift) - 1u)) - hy;
      if (dx < -kUniWin || dx > kUniWin || dy < -kUniWin || dy > kUniWin) continue;
      f((int)(e >> 14), dx, dy);
    }
  }
}

__global__ void __launch_bounds__(kUniThreads) k_uniformity(const int32_t* score_maps, int W, int H, const uint2* cand, const int32_t* cand_count,
                                                            int count_stride, int32_t* sorted_score, uint32_t* sorted_xy, const float* lut_g, int cshift,
                                                            const PatternPoint* pat0_g, int max_kp, int kp_cap, int border, const float* ray_map,
                                                            const float* jac_map, float fu, float d0, float d1, float d2, const float* dir_dev,
                                                            okb_keypoint_t* kp_out, int32_t* count_out, int32_t* status, long long* dbg)
{
  if (dir_dev) { d0 = dir_dev[0]; d1 = dir_dev[1]; d2 = dir_dev[2]; }
#define OKB_STAMP(i) if (dbg && threadIdx.x == 0) dbg[blockIdx.x * 16 + (i)] = clock64()
  OKB_STAMP(0);
  extern __shared__ unsigned long long keys[];                                   // kHarrisCandCap sort keys ...
  uint32_t* entries = reinterpret_cast<uint32_t*>(keys);                         // ... then: kHarrisCandCap cell-ordered entries
  uint32_t* word = entries + kHarrisCandCap;                                     //           kHarrisCandCap pending / stamp-sum words
  uint8_t* state = reinterpret_cast<uint8_t*>(keys + kHarrisCandCap);            // 0 undecided, 1 accepted, 2 rejected, 3 kept + valid
  uint16_t* queue = reinterpret_cast<uint16_t*>(state + kHarrisCandCap);         // ranks in the order they became decidable
  int* cell = reinterpret_cast<int*>(queue + kHarrisCandCap);                    // kUniMaxCells + 1 cell ends
  float* lut = reinterpret_cast<float*>(cell + kUniMaxCells + 1);                // 31 x 31
  PatternPoint* pat0 = reinterpret_cast<PatternPoint*>(lut + kUniLut * kUniLut); // 60
  int* sh = reinterpret_cast<int*>(pat0 + kPoints);                              // 32 scan words + scalars
  int* s_tail = sh + 32; int* s_acc = sh + 33; int* s_first = sh + 34; int* s_total = sh + 35; int* s_ext = sh + 36;   // s_ext: 2 words
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int n = min(cand_count[frame * count_stride], kHarrisCandCap);
  const uint2* cd = cand + (size_t)frame * kHarrisCandCap;
  const int32_t* smap = score_maps + (size_t)frame * W * H;
  int32_t* ss = sorted_score + (size_t)frame * kHarrisCandCap;
  uint32_t* sxy = sorted_xy + (size_t)frame * kHarrisCandCap;
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
  OKB_STAMP(1);
  const float max_score = (float)(int)(~(uint32_t)(keys[0] >> 32));
  // ---- the ranked list goes to global memory (the key array is reused below); cells of 16 x 16 half-resolution positions
  const int cw = (((W - 1) / 2) >> cshift) + 1, ch = (((H - 1) / 2) >> cshift) + 1, n_cells = cw * ch;
  for (int i = tid; i <= n_cells; i += kUniThreads) cell[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += kUniThreads) {
    const unsigned long long k = keys[i];
    const uint32_t xy = (uint32_t)k;
    ss[i] = (int)(~(uint32_t)(k >> 32));
    sxy[i] = xy;
    state[i] = 0;
    const int hx = (int)(xy & 0xffffu) >> 1, hy = (int)(xy >> 16) >> 1;
    atomicAdd(&cell[(hy >> cshift) * cw + (hx >> cshift)], 1);
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
    const uint32_t xy = sxy[i];
    const int hx = (int)(xy & 0xffffu) >> 1, hy = (int)(xy >> 16) >> 1;
    entries[atomicAdd(&cell[(hy >> cshift) * cw + (hx >> cshift)], 1)] = ((uint32_t)i << 14) | ((uint32_t)hx << 4) | (uint32_t)(hy & 15);
  }
  if (tid == 0) { *s_tail = 0; *s_acc = 0; *s_first = n; *s_total = 0; s_ext[0] = 0; s_ext[1] = 0; }
  __syncthreads();
  // ---- pending counts: higher-ranked candidates whose stamp reaches this one's cell (a warp per candidate, lanes over the entries)
  const int group = tid / kUniGroup, sub = tid % kUniGroup;
  {
    uint32_t xy_next = group < n ? sxy[group] : 0u;
    for (int base = 0; base < n; base += kUniGroups) {   // the trip count is uniform over the warp (shuffles inside)
      const int i = base + group;
      const uint32_t xy = xy_next;
      if (i + kUniGroups < n) xy_next = sxy[i + kUniGroups];
      int cnt = 0;
      if (i < n) {
        const int hx = (int)(xy & 0xffffu) >> 1, hy = (int)(xy >> 16) >> 1;
        for_each_neighbour(entries, cell, cw, ch, cshift, hx, hy, sub, [&](int j, int dx, int dy) {
          if (j < i && lut[(dy + kUniWin) * kUniLut + dx + kUniWin] != 0.0f) cnt++;
        });
      }
#pragma unroll
      for (int o = kUniGroup / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (sub == 0 && i < n) {
        word[i] = (uint32_t)cnt << kPendShift;
        if (cnt == 0) queue[atomicAdd(s_tail, 1)] = (uint16_t)i;
      }
    }
  }
  __syncthreads();
  OKB_STAMP(2);
  // ---- waves: the candidates that became decidable in the previous wave are decided (their stamp sum is complete) and PUSH their
  //      decision to the lower-ranked candidates in reach: one atomic adds the stamp and takes one off the pending count
  int head = 0, tail = *s_tail, lo = 0, acc = 0, waves = 0;
  while (head < tail) {
    waves++;
    for (int q = head + group; q < tail; q += kUniGroups) {
      const int j = queue[q];
      const uint32_t xy = sxy[j];
      const int hx = (int)(xy & 0xffffu) >> 1, hy = (int)(xy >> 16) >> 1;
      const float ratio = uni_ratio(ss[j], max_score);
      const bool accepted = !uni_rejected(ratio, (int)(word[j] & ((1u << kPendShift) - 1u)));
      const float nsc = uni_nsc(ratio);
      if (sub == 0) { state[j] = accepted ? 1 : 2; if (accepted && max_kp > 0) atomicAdd(s_total, 1); }
      for_each_neighbour(entries, cell, cw, ch, cshift, hx, hy, sub, [&](int i, int dx, int dy) {
        if (i <= j) return;
        const float l = lut[(-dy + kUniWin) * kUniLut - dx + kUniWin];   // the stamp of j at the cell of i: offset (i - j)
        if (l == 0.0f) return;
        const uint32_t delta = (accepted ? (uint32_t)uni_stamp(nsc, l) : 0u) - (1u << kPendShift);
        const uint32_t old = atomicAdd(&word[i], delta);
        if ((old >> kPendShift) == 1u) queue[atomicAdd(s_tail, 1)] = (uint16_t)i;
      });
    }
    __syncthreads();
    head = tail; tail = *s_tail;
    if (max_kp > 0 && *s_total >= max_kp) {   // ranks below the first undecided one are final: stop once they hold max_kp accepted candidates
      // (looked at only once max_kp candidates have been accepted anywhere)
      int mine = n;
      for (int i = lo + tid; i < n; i += kUniThreads) if (state[i] == 0) { mine = i; break; }
      if (mine < n) atomicMin(s_first, mine);
      __syncthreads();
      const int first = *s_first;
      int c = 0;
      for (int i = lo + tid; i < first; i += kUniThreads) c += state[i] == 1;
      if (c) atomicAdd(s_acc, c);
      __syncthreads();
      acc = *s_acc; lo = first;
      __syncthreads();
      if (tid == 0) *s_first = n;
      if (acc >= max_kp) break;
    } else {
      __syncthreads();
    }
  }
  if (max_kp <= 0 || acc < max_kp) lo = n;   // every candidate was decided
  __syncthreads();
  OKB_STAMP(3);
  if (dbg && tid == 0) { dbg[blockIdx.x * 16 + 5] = waves; dbg[blockIdx.x * 16 + 6] = n; dbg[blockIdx.x * 16 + 7] = lo; }
  // ---- ranks [0, lo) are decided. The first max_kp accepted ones, in rank order, go into a compact list (one thread each from here
  //      on): sub-pixel refinement, the pattern against the image border, ordered compaction into keypoint records.
  const int per = (lo + kUniThreads - 1) / kUniThreads;
  const int r0 = min(tid * per, lo), r1 = min(r0 + per, lo);
  int cnt = 0;
  for (int i = r0; i < r1; i++) cnt += state[i] == 1;
  int total_acc;
  int a = block_exclusive_scan(cnt, sh, &total_acc);
  const int n_keep = max_kp > 0 ? min(total_acc, max_kp) : total_acc;
  uint16_t* kept = queue;                            // the ready queue is dead
  for (int i = r0; i < r1; i++) if (state[i] == 1) { if (a < n_keep) kept[a] = (uint16_t)i; a++; }
  const bool aware = ray_map != nullptr;
  const float d[3] = {d0, d1, d2};
  // the largest |coordinate| and the largest sigma of the pattern: a keypoint further than their warped bound from the border needs no
  // per-sample test (non-negative floats order like their bits)
  if (tid < kPoints) {
    atomicMax(&s_ext[0], __float_as_int(fmaxf(fabsf(pat0[tid].x), fabsf(pat0[tid].y))));
    atomicMax(&s_ext[1], __float_as_int(pat0[tid].sigma));
  }
  float2* pos = reinterpret_cast<float2*>(keys);     // the entry / word arrays are dead too
  uint8_t* okf = state;                              // and so are the states once the list is made
  __syncthreads();
  for (int t = tid; t < n_keep; t += kUniThreads) {
    const uint32_t xy = sxy[kept[t]];
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
      const float ext = fmaxf(fabsf(M[0]) + fabsf(M[1]), fabsf(M[2]) + fabsf(M[3])) * __int_as_float(s_ext[0]) + __int_as_float(s_ext[1]) + 0.05f;
      const bool surely_inside = (fx - ext >= 1.0f) && (fx + ext < (float)(W - 2)) && (fy - ext >= 1.0f) && (fy + ext < (float)(H - 2));
      if (ok && !surely_inside)
        for (int p = 0; p < kPoints; p++) {
          float xf, yf;
          brisk2_sample_pos(M, fx, fy, pat0[p], xf, yf);
          ok = ok && brisk2_sample_inside(xf, yf, pat0[p].sigma, W, H);
        }
    }
    pos[t] = make_float2(fx, fy);
    okf[t] = ok ? 1 : 0;
  }
  __syncthreads();
  const int per2 = (n_keep + kUniThreads - 1) / kUniThreads;
  const int t0 = min(tid * per2, n_keep), t1 = min(t0 + per2, n_keep);
  int valid = 0;
  for (int t = t0; t < t1; t++) valid += okf[t];
  int total;
  int f = block_exclusive_scan(valid, sh, &total);
  for (int t = t0; t < t1; t++) {
    if (!okf[t]) continue;
    const int idx = f++;
    if (idx >= kp_cap) continue;
    okb_keypoint_t k;
    k.x = pos[t].x; k.y = pos[t].y;
    k.size = 12.0f; k.angle = -1.0f; k.response = (float)ss[kept[t]]; k.octave = 0; k.class_id = -1;
    kps[idx] = k;
  }
  if (tid == 0) {
    if (total > kp_cap) atomicOr(&status[frame], 8);
    count_out[frame] = min(total, kp_cap);
  }
  OKB_STAMP(4);
#undef OKB_STAMP
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_describe48(const uint8_t* in0, int in_pitch, size_t in_frame_stride, int W, int H, const int32_t* integral,
                                                    int ipitch, const PatternPoint* pattern /*[rot][point] of the one scale*/, const uint32_t* short48,
                                                    const int4* long_pairs, const float* ray_map, const float* jac_map, float fu, float d0, float d1,
                                                    float d2, const float* dir_dev, okb_keypoint_t* kp, const int32_t* count, int kp_cap, uint8_t* desc, uint8_t* desc64)
{
  __shared__ int values[4][64];
  if (dir_dev) { d0 = dir_dev[0]; d1 = dir_dev[1]; d2 = dir_dev[2]; }
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
  if (lane < 16) reinterpret_cast<uint32_t*>(desc64 + ((size_t)frame * kp_cap + k) * 64)[lane] = lane < 12 ? mine : 0u;   // 64-byte slot, zero tail
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
  hs->cpitch = (c.width + 15) / 16 * 16;
  OKB_CUDA(cudaMalloc(&hs->d_cond, (size_t)hs->cpitch * c.height * B));
  OKB_CUDA(cudaMemset(hs->d_cond, 0, (size_t)hs->cpitch * c.height * B));
  OKB_CUDA(cudaMalloc(&hs->d_cand, (size_t)kHarrisCandCap * sizeof(uint2) * B));
  OKB_CUDA(cudaMalloc(&hs->d_sorted_score, (size_t)kHarrisCandCap * 4 * B));
  OKB_CUDA(cudaMalloc(&hs->d_sorted_xy, (size_t)kHarrisCandCap * 4 * B));
  float lut[kUniLut * kUniLut];
  for (int j = 0; j < kUniLut; j++) for (int i = 0; i < kUniLut; i++) lut[j * kUniLut + i] = uni_lut_host(hs->radius, i - kUniWin, j - kUniWin);
  {
    // cells of 8 x 8 positions when every stamp weight beyond 8 positions is zero (small radii: 4 x fewer entries per 3 x 3 block) and
    // the image has no more than kUniMaxCells of them
    int reach = 0;
    for (int i = 0; i <= kUniWin; i++) if (lut[kUniWin * kUniLut + kUniWin + i] != 0.0f) reach = i;
    const long long cells8 = (long long)((((c.width - 1) / 2) >> 3) + 1) * ((((c.height - 1) / 2) >> 3) + 1);
    hs->cshift = (reach <= 8 && cells8 <= kUniMaxCells) ? 3 : 4;
  }
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
  OKB_CUDA(cudaMalloc(&hs->d_dir, 16)); OKB_CUDA(cudaMemset(hs->d_dir, 0, 16)); OKB_CUDA(cudaMallocHost(&hs->h_dir, 16));
  OKB_CUDA(cudaMalloc(&ws.d_desc64, (size_t)ws.kp_cap * 64 * B));
  OKB_CUDA(cudaMemset(ws.d_desc64, 0, (size_t)ws.kp_cap * 64 * B));
  OKB_CUDA(cudaFuncSetAttribute(k_uniformity, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUniSmem));
  return OKB_OK;
}

void harris_free_camera(okb_context* ctx, int cam)
{
  CamWorkspace& ws = ctx->cams[cam];
  HarrisState* hs = (HarrisState*)ws.harris;
  if (!hs) return;
  cudaFree(hs->d_score); cudaFree(hs->d_cond); cudaFree(hs->d_cand); cudaFree(hs->d_sorted_score); cudaFree(hs->d_sorted_xy); cudaFree(hs->d_dir); if (hs->h_dir) cudaFreeHost(hs->h_dir); cudaFree(hs->d_lut); cudaFree(hs->d_short48);
  delete hs;
  ws.harris = nullptr;
  cudaFree(ws.d_desc64); ws.d_desc64 = nullptr;
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
  {
    const int strips = (W + kHsCols - 1) / kHsCols;
    const int band = B >= 8 ? 60 : (B >= 2 ? 30 : 20);
    k_harris_score<<<dim3((strips + kHsWarps - 1) / kHsWarps, (H + band - 1) / band, B), 32 * kHsWarps, 0, st>>>(
        d_images, src_pitch, in_stride, W, H, c.threshold, hs->d_score, hs->d_cond, hs->cpitch, band);
    k_harris_maxima<<<dim3((hs->cpitch / 16 * H + 255) / 256, B), 256, 0, st>>>(hs->d_score, hs->d_cond, hs->cpitch, W, H, hs->d_cand,
                                                                               ws.d_cand_count, kMaxLayers, ws.d_status);
  }
  if (ctx->timers_on) { cudaEventRecord(ws.ev_mid, st); cudaEventRecord(ws.ev[1], st); }
  const bool aware = ws.maps_ready && ws.has_model;
  const float* rays = aware ? ws.d_ray_map : nullptr;
  const float* jac = aware ? ws.d_jac_map : nullptr;
  const float fu = aware ? (float)ws.model.fu : 1.0f;
  const PatternPoint* pat = ctx->d_pattern + (size_t)hs->basic_scale * kRot * kPoints;
  // live path (okb_process_multiframe): the direction comes from the page-locked mirror through a copy node, so that a replayed graph
  // sees this frame's value; everywhere else it is a launch argument
  const float* dir_dev = nullptr;
  if (hs->dir_from_device) { memcpy(hs->h_dir, ws.extraction_dir, 12); OKB_CUDA(cudaMemcpyAsync(hs->d_dir, hs->h_dir, 12, cudaMemcpyHostToDevice, st)); dir_dev = hs->d_dir; hs->dir_from_device = 0; }
  k_uniformity<<<B, kUniThreads, kUniSmem, st>>>(hs->d_score, W, H, hs->d_cand, ws.d_cand_count, kMaxLayers, hs->d_sorted_score, hs->d_sorted_xy, hs->d_lut, hs->cshift, pat,
                                                 c.max_keypoints, ws.kp_cap, hs->border, rays, jac, fu, ws.extraction_dir[0], ws.extraction_dir[1],
                                                 ws.extraction_dir[2], dir_dev, ws.d_kp, ws.d_count, ws.d_status, ws.d_dbg);
  if (ctx->timers_on) cudaEventRecord(ws.ev[2], st);
  OKB_CUDA(cudaStreamWaitEvent(st, ws.ev_join, 0));
  k_describe48<<<dim3((ws.kp_cap + 3) / 4, B), 128, 0, st>>>(d_images, src_pitch, in_stride, W, H, ws.d_integral, ipitch, pat, hs->d_short48,
                                                             ctx->d_long_pairs, rays, jac, fu, ws.extraction_dir[0], ws.extraction_dir[1],
                                                             ws.extraction_dir[2], dir_dev, ws.d_kp, ws.d_count, ws.kp_cap, ws.d_desc, ws.d_desc64);
  ctx->launches += 6;
  { int rc = camera_backproject_batch(ctx, cam, B); if (rc) return rc; }
  if (ctx->timers_on) { cudaEventRecord(ws.ev[3], st); ws.pending_timing = 1; }
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

void harris_stage_direction(okb_context* ctx, int cam)
{
  CamWorkspace& ws = ctx->cams[cam];
  HarrisState* hs = (HarrisState*)ws.harris;
  if (!hs) return;
  memcpy(hs->h_dir, ws.extraction_dir, 12);
  hs->dir_from_device = 1;   // consumed by the next harris_run_device (direct submission or capture); a replay only needs the mirror
}

}  // namespace okb
