// okb_detect.cu -- detect + describe kernels (sm_100a) and their host-side sequencing.
//
// Replaces, for one camera, detector_->detect(image_, keypoints_) and extractor_->compute(image_, keypoints_,
// descriptors_) (reference okvis_cv/include/okvis/implementation/Frame.hpp:152,167) as driven by
// Frontend::detectAndDescribe (reference okvis_frontend/src/Frontend.cpp:221-269).
//
// Pipeline per batch of frames (blockIdx.z / blockIdx.y = frame):
//   k_resize        INTER_AREA pyramid layers (2/3-sample and half-sample), table driven, bit-exact rounding
//   k_score_nms     dense AGAST 9-16 score map b0 of every layer (u8; tiles staged in shared memory, 16x2 SIMD min/max)
//                   fused with the 3x3 non-max candidates + tie flag (warp-ballot compaction)
//   k_refine        sub-pixel / scale refinement of every candidate (pure), cache-touch events of non-tie maxima
//   k_resolve       order-exact resolution of tied maxima (touch-time map)
//   k_finalize      order by (layer, y, x), keep the N strongest, drop border keypoints, pattern scale index
//   k_integral_*    int32 integral image
//   k_describe      one warp per keypoint: 2 x 60 smoothed samples, orientation, 512 bits via ballot
#include <cuda.h>   // CUtensorMap type and enums only; the encoder is resolved through cudaGetDriverEntryPoint
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "okb_internal.h"

namespace okb {

// ---------------------------------------------------------------------------------------------------------------
struct FrameViews {
  LayerView L[kMaxLayers];
  uint8_t* score[kMaxLayers];
  uint32_t* touch[kMaxLayers];
  int n;
};

__device__ __forceinline__ void make_views(const DeviceLayers& dl, const uint8_t* in0, int in_pitch, size_t in_frame_stride,
                                           uint8_t* img_block, uint8_t* score_block, uint32_t* touch_block, int frame,
                                           FrameViews& v)
{
  v.n = dl.n;
  uint8_t* ib = img_block + (size_t)frame * dl.frame_stride;
  uint8_t* sb = score_block + (size_t)frame * dl.frame_stride;
  uint32_t* tb = touch_block ? touch_block + (size_t)frame * dl.frame_stride : nullptr;
#pragma unroll
  for (int i = 0; i < kMaxLayers; i++) {
    if (i < dl.n) {
      const DeviceLayer& d = dl.l[i];
      v.L[i].w = d.w; v.L[i].h = d.h; v.L[i].scale = d.scale; v.L[i].offset = d.offset_px;
      if (i == 0) { v.L[i].img = in0 + (size_t)frame * in_frame_stride; v.L[i].pitch = in_pitch; }
      else { v.L[i].img = ib + d.offset; v.L[i].pitch = d.pitch; }
      v.L[i].b0 = sb + d.offset; v.L[i].bpitch = d.pitch;
      v.score[i] = sb + d.offset;
      v.touch[i] = tb ? tb + d.offset : nullptr;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pyramid
struct ResizeJob {
  const uint8_t* src; int src_pitch; size_t src_frame_stride;
  uint8_t* dst; int dst_pitch; size_t dst_frame_stride;
  int dw, dh, fast2;
  const int *xs, *xn, *ys, *yn; const float *xa, *ya;
};
struct ResizeJobs { ResizeJob j[2]; int n; int tiles_x[2], tiles_y[2]; };

__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ ResizeJobs jobs)
{
  int t = blockIdx.x, ji = 0;
  const int n0 = jobs.tiles_x[0] * jobs.tiles_y[0];
  if (t >= n0) { t -= n0; ji = 1; }
  const ResizeJob& J = jobs.j[ji];
  const int tx = t % jobs.tiles_x[ji], ty = t / jobs.tiles_x[ji];
  const int frame = blockIdx.y;
  const uint8_t* src = J.src + (size_t)frame * J.src_frame_stride;
  uint8_t* dst = J.dst + (size_t)frame * J.dst_frame_stride;
  if (J.fast2) {
    // tile = 128 x 8 destination pixels; a thread makes 4 of them from two 8-byte source loads
    const int x = tx * 128 + (threadIdx.x & 31) * 4, y = ty * 8 + (threadIdx.x >> 5);
    if (x >= J.dw || y >= J.dh) return;
    const uint8_t* r0 = src + (size_t)(2 * y) * J.src_pitch + 2 * x;
    const bool vec = ((J.src_pitch & 7) == 0) && ((((uintptr_t)src) & 7) == 0) && (x + 3 < J.dw) && ((J.dst_pitch & 3) == 0);
    if (vec) {
      const uint2 a = *reinterpret_cast<const uint2*>(r0), b = *reinterpret_cast<const uint2*>(r0 + J.src_pitch);
      auto px = [](uint32_t u, uint32_t v, int sh) {
        return (((u >> sh) & 255u) + ((u >> (sh + 8)) & 255u) + ((v >> sh) & 255u) + ((v >> (sh + 8)) & 255u) + 2u) >> 2;
      };
      const uint32_t out = px(a.x, b.x, 0) | (px(a.x, b.x, 16) << 8) | (px(a.y, b.y, 0) << 16) | (px(a.y, b.y, 16) << 24);
      *reinterpret_cast<uint32_t*>(dst + (size_t)y * J.dst_pitch + x) = out;
    } else {
      for (int i = 0; i < 4 && x + i < J.dw; i++) dst[(size_t)y * J.dst_pitch + x + i] = half_pixel(src, J.src_pitch, x + i, y);
    }
  } else {
    // tile = 32 x 32 destination pixels, 4 rows per thread
    const int x = tx * 32 + (threadIdx.x & 31);
    const int y0 = ty * 32 + (threadIdx.x >> 5) * 4;
    if (x >= J.dw) return;
    const int xs = J.xs[x], xn = J.xn[x];
    float xa[4];
#pragma unroll
    for (int i = 0; i < 4; i++) xa[i] = J.xa[x * 4 + i];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int y = y0 + r;
      if (y < J.dh) {
        float ya[4];
#pragma unroll
        for (int i = 0; i < 4; i++) ya[i] = J.ya[y * 4 + i];
        dst[(size_t)y * J.dst_pitch + x] = area_pixel(src, J.src_pitch, xs, xn, xa, J.ys[y], J.yn[y], ya);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// score map tiling
constexpr int kTileW = 64, kTileH = 64;   // one warp per tile row (8 rows per warp), one lane per horizontal pixel pair

// per-layer regions of the candidate list (so that a warp of the refinement kernel sees candidates of ONE layer and
// follows one code path); cand_count is [frames][kMaxLayers]
struct CandRegions { int off[kMaxLayers + 1]; };
__device__ __forceinline__ int cand_total(const CandRegions& cr, const int32_t* count, int n_layers, int* prefix /*kMaxLayers+1*/)
{
  int acc = 0;
#pragma unroll
  for (int l = 0; l < kMaxLayers; l++) { prefix[l] = acc; if (l < n_layers) acc += min(count[l], cr.off[l + 1] - cr.off[l]); }
  prefix[kMaxLayers] = acc;
  return acc;
}

struct TileMap { int n_layers; int tile_prefix[kMaxLayers + 1]; int tiles_x[kMaxLayers]; };

__device__ __forceinline__ int find_layer(const TileMap& tm, int tile)
{
  int l = 0;
#pragma unroll
  for (int i = 1; i < kMaxLayers; i++) if (i < tm.n_layers && tile >= tm.tile_prefix[i]) l = i;
  return l;
}

// Dense AGAST 9-16 score b0 = clamp(B*, 0, 254) of four horizontally adjacent pixels, two at a time in 16x2 SIMD
// (VIMNMX3.U16x2): B* = max(max_arcs min_arc(ring) - p, p - min_arcs max_arc(ring)) - 1 over the 16 arcs of 9 pixels.
__device__ __forceinline__ uint32_t u16x2_lo(uint32_t w) { return __byte_perm(w, 0, 0x4140); }  // (b0, b1)
__device__ __forceinline__ uint32_t u16x2_hi(uint32_t w) { return __byte_perm(w, 0, 0x4342); }  // (b2, b3)

__device__ __forceinline__ uint32_t b0_pair(const uint32_t (&v)[16], uint32_t p)
{
  uint32_t m3[16], M3[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    m3[i] = __vimin3_u16x2(v[i], v[(i + 1) & 15], v[(i + 2) & 15]);
    M3[i] = __vimax3_u16x2(v[i], v[(i + 1) & 15], v[(i + 2) & 15]);
  }
  uint32_t m9[16], M9[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    m9[i] = __vimin3_u16x2(m3[i], m3[(i + 3) & 15], m3[(i + 6) & 15]);
    M9[i] = __vimax3_u16x2(M3[i], M3[(i + 3) & 15], M3[(i + 6) & 15]);
  }
  uint32_t bb = __vimax3_u16x2(m9[0], m9[1], m9[2]), bd = __vimin3_u16x2(M9[0], M9[1], M9[2]);
#pragma unroll
  for (int i = 3; i < 15; i += 2) { bb = __vimax3_u16x2(bb, m9[i], m9[i + 1]); bd = __vimin3_u16x2(bd, M9[i], M9[i + 1]); }
  bb = __vmaxu2(bb, m9[15]); bd = __vminu2(bd, M9[15]);
  // per half: t = max(bb - p, p - bd, 0) computed without borrows; b0 = min(max(t, 1) - 1, 254)
  const uint32_t tb = __vmaxu2(bb, p) - p;
  const uint32_t td = p - __vminu2(bd, p);
  const uint32_t t = __vmaxu2(tb, td);
  return __vminu2(__vmaxu2(t, 0x00010001u) - 0x00010001u, 0x00FE00FEu);
}

// ---- TMA (cp.async.bulk.tensor) staging of the image tiles -----------------------------------------------------
// One 3-D tensor map per layer: (x: w bytes, y: h rows of `pitch` bytes, frame). A single elected thread issues one
// bulk tensor copy of the kImgW x kImgH box per CTA; out-of-image elements are zero-filled by the hardware, the
// completion is signalled on a shared-memory mbarrier.
struct alignas(64) TmaMaps { CUtensorMap m[kMaxLayers]; int use[kMaxLayers]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// Tie-cell bitmap (okb_core.h: for_each_cell): flagged by the score kernel around tied / pending candidates, tested by
// k_refine before it emits the touches of a maximum.
constexpr int kCellWordsPerFrame = kMaxLayers * kCellWordsPerLayer;
__device__ __forceinline__ void cells_flag(uint32_t* cells /*layer*/, int layer_w, int x_lo, int x_hi, int y_lo, int y_hi)
{
  for_each_cell(layer_w, x_lo, x_hi, y_lo, y_hi, [&](int b) {
    if (b < kCellWordsPerLayer * 32) atomicOr(&cells[b >> 5], 1u << (b & 31));
    return false;
  });
}
__device__ __forceinline__ bool cells_any(const uint32_t* cells /*layer*/, int layer_w, const TouchBox& t)
{
  return for_each_cell(layer_w, t.x_lo, t.x_hi, t.y_lo, t.y_hi,
                       [&](int b) { return b >= kCellWordsPerLayer * 32 || ((__ldg(&cells[b >> 5]) >> (b & 31)) & 1u); });
}

// Fused score + non-max suppression. Tile = kTileW x kTileH (64 x 64) pixels of one layer of one frame, 256 threads.
//   1. the image tile with its ring halo (3 rows above/below, 16 bytes left/right: TMA wants 16-byte granular boxes
//      AND box origins) arrives in shared memory by ONE TMA bulk tensor copy, zero outside the image;
//   2. the bytes are expanded once into two 16x2 planes, E[r][k] = (p[2k], p[2k+1]) and O[r][k] = (p[2k+1], p[2k+2]):
//      every ring sample of a horizontal pixel PAIR is then a single conflict-free LDS.32 (E for even dx, O for odd dx)
//      and the main loop is nothing but the 81 VIMNMX3.U16x2 of b0_pair (the ALU pipe is this kernel's limiter);
//   3. a warp owns a tile row per iteration (lane = pixel pair, 8 rows per warp): dense scores b0 go to the global score
//      map (64 contiguous bytes per warp) and to a shared score tile; the pixels with score >= threshold ("strong", a few
//      per cent) of a row leave the loop as two ballots (even / odd pixels), no atomics inside the loop;
//   4. the ballots are compacted into a dense list and the strong pixels get the 3x3 non-max test from the shared score
//      tile, one per thread. A strong pixel ON the tile border whose
//      in-tile neighbours do not already beat it is emitted with the PENDING flag (bit 30): its out-of-tile neighbours
//      are scores of another CTA, so k_refine finishes the test from the global map (complete by then). There is no
//      score halo, i.e. no pixel is scored twice.
//   Candidate word: time key | tie << 31 | pending << 30; the tile's candidates are appended with one atomicAdd.
constexpr int kHaloX = 16;           // measured on B200: the innermost TMA coordinate must be a multiple of 16 bytes
                                     // (bench/tma_probe.cu: x = -8, 8, 376 raise "illegal instruction", -16, 0, 384 work)
constexpr int kImgW = kTileW + 2 * kHaloX;   // bytes per staged image row: x0-16 .. x0+79
constexpr int kImgH = kTileH + 6;    // rows y0-3 .. y0+66
constexpr int kExp0 = 3, kExp1 = 21; // staged words (4 bytes) that are expanded: bytes 12 .. 83 cover x0-4 .. x0+67
constexpr int kPlaneW = 2 * (kExp1 - kExp0);   // 16x2 words per plane row; plane word j holds staged bytes 2j+12 (E) / 2j+13 (O)
constexpr int kScoreThreads = 256;
constexpr uint32_t kCandTie = 0x80000000u, kCandPending = 0x40000000u, kCandKeyMask = 0x3fffffffu;

__global__ void __launch_bounds__(kScoreThreads) k_score_nms(const __grid_constant__ TmaMaps maps, const __grid_constant__ DeviceLayers dl,
                                                             const __grid_constant__ TileMap tm,
                                                             const uint8_t* in0, int in_pitch, size_t in_frame_stride,
                                                             uint8_t* img_block, uint8_t* score_block, uint32_t* cand,
                                                             int32_t* cand_count, int cand_cap, const __grid_constant__ CandRegions cr,
                                                             int threshold, int32_t* status, uint32_t* tie_cells)
{
  __shared__ __align__(128) uint8_t tile[kImgH][kImgW];
  __shared__ __align__(16) uint32_t planes[2][kImgH][kPlaneW];   // E and O; reused as the candidate list after the scoring
  __shared__ __align__(16) uint8_t sc[kTileH][kTileW];
  uint32_t (*pe)[kPlaneW] = planes[0];
  uint32_t (*po)[kPlaneW] = planes[1];
  uint32_t* out_list = &planes[0][0][0];   // kTileH * kTileW entries <= 2 * kImgH * kPlaneW; first written after the
                                           // __syncthreads() that ends the scoring loop (the planes are dead by then)
  static_assert(kTileH * kTileW <= 2 * kImgH * kPlaneW, "candidate list must fit into the planes");
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint16_t strong[kTileH * kTileW];
  __shared__ uint32_t smask[kTileH][2];            // per tile row: ballots of the strong even / odd pixels
  __shared__ int n_strong, n_out, out_base, tma_failed;
  const int frame = blockIdx.y;
  const int layer = find_layer(tm, blockIdx.x);
  const int t = blockIdx.x - tm.tile_prefix[layer];
  const int tx = t % tm.tiles_x[layer], ty = t / tm.tiles_x[layer];
  const DeviceLayer d = dl.l[layer];
  const uint8_t* img; int pitch;
  if (layer == 0) { img = in0 + (size_t)frame * in_frame_stride; pitch = in_pitch; }
  else { img = img_block + (size_t)frame * dl.frame_stride + d.offset; pitch = d.pitch; }
  uint8_t* score = score_block + (size_t)frame * dl.frame_stride + d.offset;
  const int x0 = tx * kTileW, y0 = ty * kTileH;
  if (threadIdx.x == 0) { n_strong = 0; n_out = 0; tma_failed = 0; }
  if (threadIdx.x < 2 * kTileH) (&smask[0][0])[threadIdx.x] = 0u;
  if (maps.use[layer]) {
    // TMA: one bulk tensor copy per CTA; out-of-image bytes arrive as zeros
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, kImgH * kImgW);
      tma_load_3d(&tile[0][0], &maps.m[layer], &bar, x0 - kHaloX, y0 - 3, frame);
    }
    // one warp polls the mbarrier (try_wait suspends in hardware), the others park at the CTA barrier and leave the
    // issue slots to the resident CTAs that are computing
    if (threadIdx.x < 32) {
      const long long t_start = clock64();
      while (!mbar_try_wait(&bar, 0)) {
        if (clock64() - t_start > 400000000ll) {   // ~0.2 s: never spin forever on a broken descriptor
          if (threadIdx.x == 0) { atomicOr(&status[frame], 16); tma_failed = 1; }
          break;
        }
      }
    }
    __syncthreads();
    if (tma_failed) return;
  } else {
    const bool word_ok = ((pitch & 3) == 0) && ((((uintptr_t)img) & 3) == 0);
    for (int i = threadIdx.x; i < kImgH * (kImgW / 4); i += kScoreThreads) {
      const int r = i / (kImgW / 4), c = i % (kImgW / 4);
      const int y = y0 - 3 + r, x = x0 - kHaloX + c * 4;
      uint32_t w = 0;
      if (y >= 0 && y < d.h) {
        const uint8_t* row = img + (size_t)y * pitch;
        if (word_ok && x >= 0 && x + 3 < d.w) w = *reinterpret_cast<const uint32_t*>(row + x);
        else {
#pragma unroll
          for (int b = 0; b < 4; b++) { const int xx = x + b; if (xx >= 0 && xx < d.w) w |= (uint32_t)row[xx] << (8 * b); }
        }
      }
      *reinterpret_cast<uint32_t*>(&tile[r][c * 4]) = w;
    }
    __syncthreads();
  }
  // ---- expansion into the two 16x2 planes
  for (int i = threadIdx.x; i < kImgH * (kExp1 - kExp0); i += kScoreThreads) {
    const int r = i / (kExp1 - kExp0), m = i % (kExp1 - kExp0);
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&tile[r][4 * (m + kExp0)]);
    const uint32_t nx = *reinterpret_cast<const uint32_t*>(&tile[r][4 * (m + kExp0) + 4]);
    uint2 e, o;
    e.x = __byte_perm(w, 0, 0x4140); e.y = __byte_perm(w, 0, 0x4342);
    o.x = __byte_perm(w, 0, 0x4241); o.y = __byte_perm(w, nx, 0x7473) & 0x00ff00ffu;   // (b3, nx.b0)
    *reinterpret_cast<uint2*>(&pe[r][2 * m]) = e;
    *reinterpret_cast<uint2*>(&po[r][2 * m]) = o;
  }
  __syncthreads();
  // ---- dense scores: warp = tile row, lane = pixel pair (x0 + 2*lane, +1). Every ring sample is one LDS.32 at a
  //      compile-time offset from a single per-lane pointer (the two planes are one array).
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int PW = kPlaneW, OO = kImgH * kPlaneW;          // row pitch of a plane, offset of the O plane (words)
  const int xg = x0 + 2 * lane;
  // scores are zero in the 3-pixel margin of the layer: per-lane column mask, per-row on/off
  const uint32_t colmask = ((xg >= 3 && xg < d.w - 3) ? 0x0000ffffu : 0u) | ((xg + 1 >= 3 && xg + 1 < d.w - 3) ? 0xffff0000u : 0u);
  const int rows_here = min(kTileH, d.h - y0);
  const uint32_t thr2 = (uint32_t)(0x8000 - min(max(threshold, 1), 0x7fff)) * 0x00010001u;
  {
    const uint32_t* p = &planes[0][warp][2 + lane];          // plane word of this pair (staged byte 16 + 2*lane), tile row `warp`
    uint8_t* gout = score + (size_t)(y0 + warp) * d.pitch + xg;   // pitch is a multiple of 64
    uint8_t* sout = &sc[warp][2 * lane];
#pragma unroll 1
    for (int row = warp; row < rows_here; row += kScoreThreads / 32) {
      uint32_t v[16];
      v[0] = p[OO + 3 * PW - 2];  v[1] = p[OO + 2 * PW - 2];  v[2] = p[1 * PW - 1];       v[3] = p[OO - 1];
      v[4] = p[0];                v[5] = p[OO];               v[6] = p[1 * PW + 1];       v[7] = p[OO + 2 * PW + 1];
      v[8] = p[OO + 3 * PW + 1];  v[9] = p[OO + 4 * PW + 1];  v[10] = p[5 * PW + 1];      v[11] = p[OO + 6 * PW];
      v[12] = p[6 * PW];          v[13] = p[OO + 6 * PW - 1]; v[14] = p[5 * PW - 1];      v[15] = p[OO + 4 * PW - 2];
      uint32_t s = b0_pair(v, p[3 * PW]);
      const int y = y0 + row;
      s &= (y >= 3 && y < d.h - 3) ? colmask : 0u;
      const uint16_t packed = (uint16_t)__byte_perm(s, 0, 0x4420);
      *reinterpret_cast<uint16_t*>(sout) = packed;
      *reinterpret_cast<uint16_t*>(gout) = packed;
      // strong pixels (score >= threshold; scores <= 254 so the 16-bit adds cannot carry): two ballots per row
      const uint32_t hit = s + thr2;
      const unsigned m_even = __ballot_sync(0xffffffffu, hit & 0x8000u), m_odd = __ballot_sync(0xffffffffu, hit & 0x80000000u);
      if (lane == 0) { smask[row][0] = m_even; smask[row][1] = m_odd; }
      p += (kScoreThreads / 32) * PW; gout += (size_t)(kScoreThreads / 32) * d.pitch; sout += (kScoreThreads / 32) * kTileW;
    }
  }
  __syncthreads();
  // ---- the strong pixels (a few per cent) are compacted into a dense list, then get the 3x3 non-max test one per thread
  if (threadIdx.x < 2 * kTileH) {
    uint32_t m = (&smask[0][0])[threadIdx.x];
    if (m) {
      int pos = atomicAdd(&n_strong, __popc(m));
      const int r = threadIdx.x >> 1, odd = threadIdx.x & 1;
      while (m) { const int b = __ffs(m) - 1; m &= m - 1; strong[pos++] = (uint16_t)(r * kTileW + 2 * b + odd); }
    }
  }
  __syncthreads();
  const int ns = n_strong;
  for (int i = threadIdx.x; i < ns; i += kScoreThreads) {
    const int r = strong[i] / kTileW, x = strong[i] % kTileW;
    const int c = sc[r][x];
    bool is_c = true, tie = false, pending = false;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        if (dx == 0 && dy == 0) continue;
        const int rr = r + dy, xx = x + dx;
        if ((unsigned)rr < (unsigned)kTileH && (unsigned)xx < (unsigned)kTileW) {
          const int v = sc[rr][xx];   // rows at / below the image end are never neighbours of a strong pixel (y < h - 3)
          if (v > c) is_c = false;
          if (v == c) tie = true;
        } else pending = true;
      }
    if (is_c) {
      const int pos = atomicAdd(&n_out, 1);
      out_list[pos] = time_key(layer, x0 + x, y0 + r) | (pending ? kCandPending : (tie ? kCandTie : 0u));
      if (pending || tie)
        cells_flag(tie_cells + (size_t)frame * kCellWordsPerFrame + layer * kCellWordsPerLayer, d.w, x0 + x - 2, x0 + x + 2,
                   y0 + r - 2, y0 + r + 2);
    }
  }
  __syncthreads();
  const int no = n_out;
  if (no == 0) return;
  if (threadIdx.x == 0) out_base = atomicAdd(&cand_count[frame * kMaxLayers + layer], no);
  __syncthreads();
  const int base = out_base, room = cr.off[layer + 1] - cr.off[layer];
  for (int i = threadIdx.x; i < no; i += kScoreThreads)
    if (base + i < room) cand[(size_t)frame * cand_cap + cr.off[layer] + base + i] = out_list[i];
}

// Cache-touch events of one maximum, emitted by a full warp (all arguments warp-uniform): lanes take the positions of
// the above-layer scan (closed form of the scan order in okb_core.h: rows of [x_1, xa..xb, x1]) and of the 3x3 / 4x4
// patches, so the divergent, sequential replay of for_each_above_touch never runs on the device.
__device__ __forceinline__ void emit_touches_warp(const DeviceLayers& dl, uint32_t* touch_frame, uint32_t key, int own_touch,
                                                  int has_above, ScanTrace tr, uint32_t entry, int lane)
{
  const int layer = (int)(key >> 22), y = (int)((key >> 11) & 2047), x = (int)(key & 2047);
  if (own_touch && lane < 16) {
    const DeviceLayer d = dl.l[layer];
    const int hi = own_touch == 2 ? 2 : 1;
    const int dx = (lane & 3) - 1, dy = (lane >> 2) - 1;
    const int xx = x + dx, yy = y + dy;
    if (dx <= hi && dy <= hi && xx >= 0 && yy >= 0 && xx < d.w && yy < d.h)
      atomicMax(&touch_frame[d.offset + (size_t)yy * d.pitch + xx], entry);
  }
  if (has_above) {
    const DeviceLayer d = dl.l[layer + 1];
    uint32_t* tm = touch_frame + d.offset;
    ScanIter it; above_window(layer, x, y, it);
    auto put = [&](int xx, int yy) { if (xx >= 0 && yy >= 0 && xx < d.w && yy < d.h) atomicMax(&tm[(size_t)yy * d.pitch + xx], entry); };
    for (int q = lane; q < tr.n_queries; q += 32) {
      int X, Y; bool blk;
      above_query_pos(it, q, X, Y, blk);
      put(X, Y);
      if (blk) { put(X + 1, Y); put(X, Y + 1); put(X + 1, Y + 1); }  // bilinear read: 2x2 block
    }
    if (!tr.exited && lane < 9) put(tr.max_x + lane % 3 - 1, tr.max_y + lane / 3 - 1);
  }
}

__global__ void __launch_bounds__(128) k_refine(const __grid_constant__ DeviceLayers dl, const uint8_t* in0, int in_pitch, size_t in_frame_stride,
                                                uint8_t* img_block, uint8_t* score_block, uint32_t* touch_block,
                                                const uint32_t* cand, const int32_t* cand_count, int cand_cap,
                                                const __grid_constant__ CandRegions cr, CandRecord* rec, int threshold, uint32_t epoch,
                                                const uint32_t* tie_cells)
{
  const int frame = blockIdx.y;
  int prefix[kMaxLayers + 1];
  const int n = cand_total(cr, cand_count + frame * kMaxLayers, dl.n, prefix);
  if (blockIdx.x * blockDim.x >= n) return;
  __shared__ FrameViews v;  // dynamically indexed by layer: keep it out of local memory
  if (threadIdx.x == 0) make_views(dl, in0, in_pitch, in_frame_stride, img_block, score_block, touch_block, frame, v);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool have = i < n;
  uint32_t key = 0; int tie = 0;
  RefineResult r;
  r.keep = 0; r.own_touch = 0; r.has_above = 0; r.above.n_queries = 0; r.above.exited = 1; r.above.max_x = r.above.max_y = 0;
  if (have) {
    int l = 0;   // records are dense and layer-major: record i is candidate i - prefix[l] of layer l
#pragma unroll
    for (int j = 1; j < kMaxLayers; j++) if (i >= prefix[j]) l = j;
    const uint32_t c = cand[(size_t)frame * cand_cap + cr.off[l] + (i - prefix[l])];
    key = c & kCandKeyMask; tie = (int)(c >> 31);
    bool alive = true;
    if (c & kCandPending) {
      // a strong pixel on the border of its score tile: finish the 3x3 non-max test from the (now complete) dense map
      const LayerView& lv = v.L[l];
      const uint8_t* s = lv.b0 + (size_t)((key >> 11) & 2047) * lv.bpitch + (key & 2047);
      const int cc = s[0];
      tie = 0;
#pragma unroll
      for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
          if (dx == 0 && dy == 0) continue;
          const int val = s[dy * lv.bpitch + dx];
          if (val > cc) alive = false;
          if (val == cc) tie = 1;
        }
      if (!alive) tie = 0;
    }
    if (alive) refine_candidate(v.L, v.n, (int)(key >> 22), (int)(key & 2047), (int)((key >> 11) & 2047), threshold, r);
    CandRecord out;
    out.x = r.x; out.y = r.y; out.size = r.size; out.response = r.response; out.key = key;
    out.keep = r.keep; out.own_touch = r.own_touch; out.has_above = r.has_above; out.tie = (int8_t)tie;
    out.above = r.above; out.state = tie ? 0 : 1; out.pad[0] = out.pad[1] = out.pad[2] = 0;
    rec[(size_t)frame * cand_cap + i] = out;
  }
  // cache-touch events of the non-tied maxima, one maximum at a time by the whole warp
  uint32_t* touch_frame = touch_block + (size_t)frame * dl.frame_stride;
  bool emits = have && !tie && (r.own_touch || r.has_above);
  if (emits) {
    // the touches are only read inside the 5x5 windows of tied candidates: skip the emission when the footprint (own
    // 4x4 patch; window of the above-layer scan with its bilinear / 3x3 margins) misses every flagged cell
    const int layer = (int)(key >> 22), y = (int)((key >> 11) & 2047), x = (int)(key & 2047);
    const uint32_t* cells = tie_cells + (size_t)frame * kCellWordsPerFrame;
    bool hit = false;
    if (r.own_touch) hit = cells_any(cells + layer * kCellWordsPerLayer, dl.l[layer].w, own_touch_box(x, y));
    if (!hit && r.has_above) hit = cells_any(cells + (layer + 1) * kCellWordsPerLayer, dl.l[layer + 1].w, above_touch_box(layer, x, y));
    emits = hit;
  }
  unsigned m = __ballot_sync(0xffffffffu, emits);
  const unsigned tr_a = (uint32_t)(uint16_t)r.above.n_queries | ((uint32_t)(uint16_t)r.above.exited << 16);
  const unsigned tr_b = (uint32_t)(uint16_t)r.above.max_x | ((uint32_t)(uint16_t)r.above.max_y << 16);
  const unsigned flags = (uint32_t)r.own_touch | ((uint32_t)r.has_above << 8);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const uint32_t k2 = __shfl_sync(0xffffffffu, key, src);
    const unsigned a2 = __shfl_sync(0xffffffffu, tr_a, src), b2 = __shfl_sync(0xffffffffu, tr_b, src);
    const unsigned f2 = __shfl_sync(0xffffffffu, flags, src);
    ScanTrace t2; t2.n_queries = (int16_t)(a2 & 0xffff); t2.exited = (int16_t)(a2 >> 16);
    t2.max_x = (int16_t)(b2 & 0xffff); t2.max_y = (int16_t)(b2 >> 16);
    emit_touches_warp(dl, touch_frame, k2, (int)(f2 & 0xff), (int)(f2 >> 8), t2, touch_entry(epoch, k2), lane);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// block-wide bitonic sort of n (power of two) 64-bit keys in shared memory, ascending
__device__ void bitonic_sort_u64(unsigned long long* a, int n)
{
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
}

constexpr int kMaxTies = 4096;
constexpr int kMaxBlockers = 12;
constexpr int kResolveSmem = kMaxTies * (8 + 2 * kMaxBlockers + 2 + 8 + 4);

struct TieInfo {  // what a tie's cache touches look like if it turns out to be a maximum (8 bytes)
  int8_t own_touch, has_above, exited, n_queries;
  int16_t max_x, max_y;
};

// One CTA per frame. Resolves, in dependency rounds, the candidates whose 2-D maximum test ties with a neighbour:
// their outcome depends on which sub-threshold scores the sequential algorithm had already cached when it reached them.
// Touches by NON-tied maxima are already in the touch-time map (k_refine); touches by EARLIER TIES are applied here
// from shared memory: a tie waits only for the earlier ties whose touches can reach its 5x5 window (same layer within
// 4 px, or the layer below through the window of its above-scan), listed once; when they are all decided it ORs the
// footprint of the winners among them into its 25-bit "touched" mask and evaluates the reference's isMax2D.
__global__ void __launch_bounds__(512) k_resolve(const __grid_constant__ DeviceLayers dl, const uint8_t* score_block, const uint32_t* touch_block,
                                                 const int32_t* cand_count, int cand_cap, const __grid_constant__ CandRegions cr,
                                                 CandRecord* rec, uint32_t epoch, int threshold, int32_t* status, long long* dbg)
{
#define OKB_STAMP(i) if (dbg && threadIdx.x == 0) dbg[blockIdx.x * 16 + (i)] = clock64()
  OKB_STAMP(0);
  extern __shared__ unsigned long long resolve_smem[];
  unsigned long long* ties = resolve_smem;                                   // key << 32 | record index, sorted
  uint16_t (*blockers)[kMaxBlockers] = reinterpret_cast<uint16_t (*)[kMaxBlockers]>(ties + kMaxTies);
  short4* box = reinterpret_cast<short4*>(blockers + kMaxTies);              // above-scan window (layer+1 coords) ...
  TieInfo* info = reinterpret_cast<TieInfo*>(box);                           // ... replaced by the touch info after the lists are built
  uint32_t* tmask = reinterpret_cast<uint32_t*>(box + kMaxTies);             // window pixels touched before this tie
  int8_t* state = reinterpret_cast<int8_t*>(tmask + kMaxTies);               // 0 unresolved, 1 maximum, 2 rejected
  int8_t* n_block = state + kMaxTies;                                        // -1: list overflowed, scan instead
  __shared__ int n_ties, n_unresolved, layer_start[kMaxLayers + 1];
  const int frame = blockIdx.x;
  int prefix_[kMaxLayers + 1];
  const int n = cand_total(cr, cand_count + frame * kMaxLayers, dl.n, prefix_);
  CandRecord* R = rec + (size_t)frame * cand_cap;
  if (threadIdx.x == 0) n_ties = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (R[i].tie) {
      const int p = atomicAdd(&n_ties, 1);
      if (p < kMaxTies) ties[p] = ((unsigned long long)R[i].key << 32) | (unsigned)i;
    }
  __syncthreads();
  int T = n_ties;
  if (T > kMaxTies) { if (threadIdx.x == 0) atomicOr(&status[frame], 2); T = kMaxTies; }
  if (T == 0) return;
  int P = 1; while (P < T) P <<= 1;
  for (int i = T + threadIdx.x; i < P; i += blockDim.x) ties[i] = ~0ull;
  __syncthreads();
  OKB_STAMP(1);
  bitonic_sort_u64(ties, P);
  OKB_STAMP(2);
  if (threadIdx.x <= kMaxLayers) {
    // first tie index of every layer (ties are sorted by key, layer is the top field)
    const int l = threadIdx.x;
    int lo = 0, hi = T;
    const unsigned long long target = (unsigned long long)time_key(l, 0, 0) << 32;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (ties[mid] < target) lo = mid + 1; else hi = mid; }
    layer_start[l] = (l >= kMaxLayers) ? T : lo;
  }
  if (threadIdx.x == 0) n_unresolved = T;
  const uint8_t* score_frame = score_block + (size_t)frame * dl.frame_stride;
  const uint32_t* touch_frame = touch_block + (size_t)frame * dl.frame_stride;
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const uint32_t key = (uint32_t)(ties[i] >> 32);
    const int layer = (int)(key >> 22), y = (int)((key >> 11) & 2047), x = (int)(key & 2047);
    state[i] = 0;
    ScanIter it; above_window(layer, x, y, it);
    box[i] = make_short4((short)((int)it.x_1 - 1), (short)((int)it.x1 + 2), (short)((int)it.y_1 - 1), (short)((int)it.y1 + 2));
    // touches by the non-tied maxima (all emitted before this kernel started)
    const DeviceLayer d = dl.l[layer];
    const uint32_t* tm = touch_frame + d.offset;
    uint32_t mask = 0;
#pragma unroll
    for (int j = 0; j < 25; j++) {
      const uint32_t e = __ldcg(&tm[(size_t)(y + j / 5 - 2) * d.pitch + (x + j % 5 - 2)]);
      if (touched_before(e, epoch, key)) mask |= 1u << j;
    }
    tmask[i] = mask;
  }
  __syncthreads();
  OKB_STAMP(7);
  // enumerate the possible blockers of tie ti (earlier ties whose touches can reach its window); F(u) true = stop
  auto for_each_blocker = [&](int ti, auto F) {
    const uint32_t key = (uint32_t)(ties[ti] >> 32);
    const int layer = (int)(key >> 22), y = (int)((key >> 11) & 2047), x = (int)(key & 2047);
    {
      const unsigned long long lo_key = (unsigned long long)time_key(layer, 0, max(y - 4, 0)) << 32;
      int lo = layer_start[layer], hi = ti;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (ties[mid] < lo_key) lo = mid + 1; else hi = mid; }
      for (int u = lo; u < ti; u++) {
        const int ux = (int)((uint32_t)(ties[u] >> 32) & 2047);
        if (abs(ux - x) <= 4) if (F(u)) return;   // |dy| <= 4 by the key range
      }
    }
    if (layer > 0) {
      // rows of the layer below that can map into [y-2, y+2] (+ scan margins); the ratio is 3/2 or 4/3
      const int uy_lo = max((y - 5) * 4 / 3 - 3, 0), uy_hi = (y + 5) * 3 / 2 + 4;
      const unsigned long long lo_key = (unsigned long long)time_key(layer - 1, 0, min(uy_lo, 2047)) << 32;
      int lo = layer_start[layer - 1], hi = layer_start[layer];
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (ties[mid] < lo_key) lo = mid + 1; else hi = mid; }
      for (int u = lo; u < layer_start[layer]; u++) {
        const int uy = (int)(((uint32_t)(ties[u] >> 32) >> 11) & 2047);
        if (uy > uy_hi) break;
        const short4 bx = box[u];
        if (!(x + 2 < bx.x || x - 2 > bx.y || y + 2 < bx.z || y - 2 > bx.w)) if (F(u)) return;
      }
    }
  };
  for (int ti = threadIdx.x; ti < T; ti += blockDim.x) {
    int nb = 0;
    for_each_blocker(ti, [&](int u) {
      if (nb < kMaxBlockers) { blockers[ti][nb++] = (uint16_t)u; return false; }
      nb = -1; return true;
    });
    n_block[ti] = (int8_t)nb;
  }
  __syncthreads();
  OKB_STAMP(15);
  // the boxes are no longer needed unless a list overflowed (then the scan above is reused every round, with boxes):
  // keep them in that (rare) case and read the touch info from the records instead
  __shared__ int any_overflow;
  if (threadIdx.x == 0) any_overflow = 0;
  __syncthreads();
  for (int ti = threadIdx.x; ti < T; ti += blockDim.x) if (n_block[ti] < 0) any_overflow = 1;
  __syncthreads();
  const bool use_info = !any_overflow;
  if (use_info)
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
      const CandRecord& c = R[(int)(ties[i] & 0xffffffffu)];
      TieInfo ti; ti.own_touch = c.own_touch; ti.has_above = c.has_above; ti.exited = (int8_t)c.above.exited;
      ti.n_queries = (int8_t)c.above.n_queries; ti.max_x = c.above.max_x; ti.max_y = c.above.max_y;
      info[i] = ti;
    }
  __syncthreads();
  OKB_STAMP(3);
  auto get_info = [&](int u) {
    if (use_info) return info[u];
    const CandRecord& c = R[(int)(ties[u] & 0xffffffffu)];
    TieInfo ti; ti.own_touch = c.own_touch; ti.has_above = c.has_above; ti.exited = (int8_t)c.above.exited;
    ti.n_queries = (int8_t)c.above.n_queries; ti.max_x = c.above.max_x; ti.max_y = c.above.max_y;
    return ti;
  };
  // window pixels of tie (layer, x, y) that winner u would have touched
  auto footprint = [&](int u, int layer, int x, int y) {
    const uint32_t uk = (uint32_t)(ties[u] >> 32);
    const int ul = (int)(uk >> 22), uy = (int)((uk >> 11) & 2047), ux = (int)(uk & 2047);
    const TieInfo f = get_info(u);
    uint32_t mask = 0;
    auto mark = [&](int px, int py) {
      const int dx = px - x + 2, dy = py - y + 2;
      if ((unsigned)dx < 5u && (unsigned)dy < 5u) mask |= 1u << (dy * 5 + dx);
    };
    if (ul == layer) {
      if (f.own_touch) {
        const int hi = f.own_touch == 2 ? 2 : 1;
        for (int dy = -1; dy <= hi; dy++) for (int dx = -1; dx <= hi; dx++) mark(ux + dx, uy + dy);
      }
    } else if (f.has_above) {
      ScanIter it; above_window(ul, ux, uy, it);
      for (int q = 0; q < f.n_queries; q++) {
        int X, Y; bool blk; above_query_pos(it, q, X, Y, blk);
        mark(X, Y);
        if (blk) { mark(X + 1, Y); mark(X, Y + 1); mark(X + 1, Y + 1); }
      }
      if (!f.exited) for (int j = 0; j < 9; j++) mark(f.max_x + j % 3 - 1, f.max_y + j / 3 - 1);
    }
    return mask;
  };
  int rounds = 0;
  while (true) {
    int8_t decided[(kMaxTies + 511) / 512];
    int nd = 0;
    for (int ti = threadIdx.x; ti < T; ti += blockDim.x, nd++) {
      decided[nd] = 0;
      if (state[ti] != 0) continue;
      const uint32_t key = (uint32_t)(ties[ti] >> 32);
      const int layer = (int)(key >> 22), y = (int)((key >> 11) & 2047), x = (int)(key & 2047);
      bool blocked = false;
      uint32_t mask = tmask[ti];
      const int nb = n_block[ti];
      if (nb >= 0) {
        for (int i = 0; i < nb; i++) if (state[blockers[ti][i]] == 0) { blocked = true; break; }
        if (!blocked) for (int i = 0; i < nb; i++) { const int u = blockers[ti][i]; if (state[u] == 1) mask |= footprint(u, layer, x, y); }
      } else {
        for_each_blocker(ti, [&](int u) { if (state[u] == 0) { blocked = true; return true; } return false; });
        if (!blocked) for_each_blocker(ti, [&](int u) { if (state[u] == 1) mask |= footprint(u, layer, x, y); return false; });
      }
      if (blocked) continue;
      const DeviceLayer d = dl.l[layer];
      const uint8_t* sc = score_frame + d.offset;
      int m[5][5];
#pragma unroll
      for (int dy = -2; dy <= 2; dy++)
#pragma unroll
        for (int dx = -2; dx <= 2; dx++) {
          int val = sc[(size_t)(y + dy) * d.pitch + (x + dx)];  // dense b0 = what the cache holds once the pixel was touched
          if (val < threshold && !((mask >> ((dy + 2) * 5 + dx + 2)) & 1u)) val = 0;
          m[dy + 2][dx + 2] = val;
        }
      decided[nd] = is_max_2d_5x5(m) ? 1 : 2;
    }
    __syncthreads();   // every thread has read the states of this round
    nd = 0;
    for (int ti = threadIdx.x; ti < T; ti += blockDim.x, nd++)
      if (decided[nd]) { state[ti] = decided[nd]; atomicSub(&n_unresolved, 1); }
    __syncthreads();
    rounds++;
    if (n_unresolved <= 0) break;
  }
  OKB_STAMP(4);
  for (int ti = threadIdx.x; ti < T; ti += blockDim.x) R[(int)(ties[ti] & 0xffffffffu)].state = state[ti];
  if (dbg && threadIdx.x == 0) { dbg[blockIdx.x * 16 + 5] = rounds; dbg[blockIdx.x * 16 + 6] = T; }
#undef OKB_STAMP
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int kSortCap = 16384;

__device__ __forceinline__ uint32_t float_order_bits(float f)
{ // monotone map float -> uint32
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// exclusive scan of one int per thread across the block (blockDim.x == 1024); returns the total in `total`
__device__ int block_exclusive_scan_1024(int val, int* sh /*33 ints*/, int& total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = val;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) sh[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = sh[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    sh[lane] = wi - w;
    if (lane == 31) sh[32] = wi;
  }
  __syncthreads();
  const int res = sh[warp] + incl - val;
  total = sh[32];
  __syncthreads();
  return res;
}

constexpr int kMaxRows = 8192;   // sum of the layer heights
constexpr int kFinalizeSmem = kSortCap * 8 + kSortCap * 4 + (kMaxRows + 2) * 2 + kSortCap;

// One CTA (1024 threads) per frame: order the surviving keypoints by (layer, y, x), keep the max_kp strongest
// (ties: earlier first), drop the ones whose sampling pattern leaves the image, write cv::KeyPoint records.
// Ordering is a counting sort by (layer, row) followed by a tiny in-place insertion sort inside every row group.
__global__ void __launch_bounds__(1024) k_finalize(const __grid_constant__ DeviceLayers dl, const int32_t* cand_count, int cand_cap,
                                                   const __grid_constant__ CandRegions cr, const CandRecord* rec,
                                                   const float* scale_bounds, const uint32_t* size_list, int W, int H,
                                                   int max_kp, int kp_cap, okb_keypoint_t* kp_out, int32_t* kscale_out,
                                                   int32_t* count_out, int32_t* status, long long* dbg)
{
#define OKB_STAMP(i) if (dbg && threadIdx.x == 0) dbg[blockIdx.x * 16 + (i)] = clock64()
  OKB_STAMP(8);
  extern __shared__ unsigned long long keys[];  // kSortCap (key << 32 | record index), ordered
  uint32_t* resp = reinterpret_cast<uint32_t*>(keys + kSortCap);          // kSortCap response bits
  uint16_t* rowoff = reinterpret_cast<uint16_t*>(resp + kSortCap);        // kMaxRows + 1 group offsets
  uint8_t* flag = reinterpret_cast<uint8_t*>(rowoff + kMaxRows + 2);      // kSortCap keep flags
  __shared__ int n_valid, hist[256], sh_scan[33], rowbase[kMaxLayers + 1];
  __shared__ uint32_t sel_prefix;
  __shared__ int sel_remaining;
  const int frame = blockIdx.x;
  int prefix_[kMaxLayers + 1];
  const int n = cand_total(cr, cand_count + frame * kMaxLayers, dl.n, prefix_);
  if (threadIdx.x < dl.n && cand_count[frame * kMaxLayers + threadIdx.x] > cr.off[threadIdx.x + 1] - cr.off[threadIdx.x])
    atomicOr(&status[frame], 1);   // a layer's region of the candidate list overflowed
  const CandRecord* R = rec + (size_t)frame * cand_cap;
  if (threadIdx.x == 0) {
    n_valid = 0;
    int acc = 0;
    for (int l = 0; l < kMaxLayers; l++) { rowbase[l] = acc; if (l < dl.n) acc += dl.l[l].h; }
    rowbase[kMaxLayers] = acc;
  }
  __syncthreads();
  const int n_rows = rowbase[kMaxLayers];
  uint32_t* rowcnt = reinterpret_cast<uint32_t*>(keys);   // counts live in the (still unused) key array during pass 1
  for (int i = threadIdx.x; i <= n_rows; i += blockDim.x) rowcnt[i] = 0;
  __syncthreads();
  auto row_of = [&](uint32_t key) { return rowbase[key >> 22] + (int)((key >> 11) & 2047); };
  // pass 1: count the valid keypoints per (layer, row)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const CandRecord& c = R[i];
    if (c.state == 1 && c.keep) { atomicAdd(&rowcnt[row_of(c.key)], 1u); atomicAdd(&n_valid, 1); }
  }
  __syncthreads();
  int V = n_valid;
  const bool overflow = V > kSortCap || n_rows > kMaxRows;
  if (overflow) { if (threadIdx.x == 0) { atomicOr(&status[frame], 4); count_out[frame] = 0; } return; }
  // exclusive scan of the row counts -> group offsets (16 bit: V <= 16384)
  {
    const int chunk = (n_rows + (int)blockDim.x - 1) / (int)blockDim.x;
    const int beg = min((int)threadIdx.x * chunk, n_rows), end = min(beg + chunk, n_rows);
    int sum = 0;
    for (int i = beg; i < end; i++) sum += (int)rowcnt[i];
    int total = 0;
    int run = block_exclusive_scan_1024(sum, sh_scan, total);
    // the counts must be consumed before the offsets overwrite anything: stage this thread's counts in registers
    // (chunk <= 8 for kMaxRows = 8192)
    int cnts[8];
    for (int i = beg, j = 0; i < end; i++, j++) cnts[j] = (int)rowcnt[i];
    __syncthreads();
    for (int i = beg, j = 0; i < end; i++, j++) { rowoff[i] = (uint16_t)run; run += cnts[j]; }
    if (threadIdx.x == blockDim.x - 1) rowoff[n_rows] = (uint16_t)V;
    if (end == n_rows && beg < end) rowoff[n_rows] = (uint16_t)run;
  }
  __syncthreads();
  // pass 2: scatter into the row groups (fill counters reuse flag[] as 8-bit counters is too small: use resp[] as 32-bit)
  for (int i = threadIdx.x; i < n_rows; i += blockDim.x) resp[i] = 0;   // n_rows <= kMaxRows <= kSortCap
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const CandRecord& c = R[i];
    if (c.state == 1 && c.keep) {
      const int row = row_of(c.key);
      const int slot = (int)rowoff[row] + (int)atomicAdd(&resp[row], 1u);
      keys[slot] = ((unsigned long long)c.key << 32) | (unsigned)i;
    }
  }
  __syncthreads();
  // pass 3: order inside every row group (a handful of entries) by insertion sort, one thread per group
  for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
    const int beg = rowoff[r], end = rowoff[r + 1];
    for (int i = beg + 1; i < end; i++) {
      const unsigned long long k = keys[i];
      int j = i - 1;
      while (j >= beg && keys[j] > k) { keys[j + 1] = keys[j]; j--; }
      keys[j + 1] = k;
    }
  }
  __syncthreads();
  OKB_STAMP(10);
  for (int i = threadIdx.x; i < V; i += blockDim.x) resp[i] = float_order_bits(R[(int)(keys[i] & 0xffffffffu)].response);
  __syncthreads();
  // ---- strongest max_kp: radix select of the max_kp-th largest response
  const bool capped = max_kp > 0 && V > max_kp;
  uint32_t thr_bits = 0; int n_equal_keep = 0;
  if (capped) {
    if (threadIdx.x == 0) { sel_prefix = 0; sel_remaining = max_kp; }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = sel_prefix;
      const int rem = sel_remaining;   // read here, a barrier away from the owner lane's update below (racecheck-clean)
      const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
      for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const uint32_t rb = resp[i];
        if ((rb & himask) == (prefix & himask)) atomicAdd(&hist[(rb >> shift) & 255], 1);
      }
      __syncthreads();
      if (threadIdx.x < 32) {
        // find the bin b (from the top) where the running count reaches `remaining`: lane l owns bins 8l .. 8l+7
        const int lane = threadIdx.x;
        int mine = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) mine += hist[lane * 8 + j];
        int suffix = mine;   // inclusive suffix sum over the lanes (higher lanes own higher bins)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, suffix, o); if (lane + o < 32) suffix += t; }
        const int above = suffix - mine;
        const bool owner = above < rem && above + mine >= rem;   // exactly one lane (total >= rem)
        if (owner) {
          int r2 = rem - above, b = lane * 8 + 7;
          for (; b > lane * 8; b--) { if (hist[b] >= r2) break; r2 -= hist[b]; }
          sel_prefix = prefix | ((uint32_t)b << shift);
          sel_remaining = r2;  // how many of the elements matching the prefix so far are still to be kept
        }
      }
      __syncthreads();
    }
    thr_bits = sel_prefix; n_equal_keep = sel_remaining;
  }
  OKB_STAMP(11);
  // each thread owns a contiguous chunk so that the scans preserve the (layer, y, x) order
  const int chunk = (V + (int)blockDim.x - 1) / (int)blockDim.x;
  const int beg = min((int)threadIdx.x * chunk, V), end = min(beg + chunk, V);
  int total = 0;
  if (capped) {
    int my_eq = 0;
    for (int i = beg; i < end; i++) my_eq += (resp[i] == thr_bits);
    int eq_before = block_exclusive_scan_1024(my_eq, sh_scan, total);
    for (int i = beg; i < end; i++) {
      const uint32_t rb = resp[i];
      bool keep = rb > thr_bits;
      if (rb == thr_bits) { keep = eq_before < n_equal_keep; eq_before++; }
      flag[i] = keep;
    }
  } else {
    for (int i = beg; i < end; i++) flag[i] = 1;
  }
  int my_keep = 0;
  for (int i = beg; i < end; i++) {
    if (!flag[i]) continue;
    const CandRecord& c = R[(int)(keys[i] & 0xffffffffu)];
    const int sc = kscale_from_bounds(scale_bounds, c.size);
    const int bd = (int)size_list[sc];
    const bool out = c.x < (float)bd || c.x >= (float)(W - bd) || c.y < (float)bd || c.y >= (float)(H - bd);
    flag[i] = out ? 0 : (uint8_t)(sc + 1);
    my_keep += !out;
  }
  int pos = block_exclusive_scan_1024(my_keep, sh_scan, total);
  for (int i = beg; i < end; i++) {
    if (!flag[i]) continue;
    if (pos < kp_cap) {
      const CandRecord& c = R[(int)(keys[i] & 0xffffffffu)];
      okb_keypoint_t k;
      k.x = c.x; k.y = c.y; k.size = c.size; k.angle = -1.f; k.response = c.response;
      k.octave = (int)(c.key >> 22); k.class_id = -1;
      kp_out[(size_t)frame * kp_cap + pos] = k;
      kscale_out[(size_t)frame * kp_cap + pos] = (int)flag[i] - 1;
    }
    pos++;
  }
  if (threadIdx.x == 0) {
    if (total > kp_cap) atomicOr(&status[frame], 8);
    count_out[frame] = min(total, kp_cap);
  }
  OKB_STAMP(12);
  if (dbg && threadIdx.x == 0) { dbg[blockIdx.x * 16 + 13] = V; dbg[blockIdx.x * 16 + 14] = n; }
#undef OKB_STAMP
}

// ---------------------------------------------------------------------------------------------------------------
// integral image: I[y+1][x+1] = sum of pixels in rows <= y, cols <= x. Row pass (warp per row) then column pass.
__global__ void __launch_bounds__(256) k_integral_rows(const uint8_t* in0, int in_pitch, size_t in_frame_stride, int W, int H,
                                                       int32_t* integral, int ipitch)
{
  const int frame = blockIdx.y;
  const int y = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  int32_t* I = integral + (size_t)frame * ipitch * (H + 1);
  if (blockIdx.x == 0) for (int x = threadIdx.x; x <= W; x += 256) I[x] = 0;
  if (y >= H) return;
  const uint8_t* row = in0 + (size_t)frame * in_frame_stride + (size_t)y * in_pitch;
  int32_t* out = I + (size_t)(y + 1) * ipitch;
  if (lane == 0) out[0] = 0;
  int carry = 0;
  for (int x0 = 0; x0 < W; x0 += 128) {
    const int x = x0 + lane * 4;
    int v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = (x + i < W) ? row[x + i] : 0;
    v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
    int incl = v[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int base = carry + incl - v[3];
#pragma unroll
    for (int i = 0; i < 4; i++) if (x + i < W) out[x + i + 1] = base + v[i];
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}
// Column pass: one CTA per strip of 32 columns, 32 warps each owning a band of rows: band sums -> prefix over the
// bands in shared memory -> every warp rewrites its band with the carried-in prefix (two coalesced sweeps, 32x the
// parallelism of one thread per column).
__global__ void __launch_bounds__(1024) k_integral_cols(int W, int H, int32_t* integral, int ipitch)
{
  __shared__ int band[32][33];
  const int frame = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + lane + 1;
  const int rows = (H + 31) / 32;
  const int y0 = 1 + warp * rows, y1 = min(y0 + rows, H + 1);
  int32_t* I = integral + (size_t)frame * ipitch * (H + 1);
  constexpr int kKeep = 32;   // rows of a band kept in registers (H <= 1024): the band is read once, not twice
  int vals[kKeep];
  int sum = 0;
  const bool in_regs = rows <= kKeep;
  if (x <= W) {
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < kKeep; i++) { vals[i] = (y0 + i < y1) ? I[(size_t)(y0 + i) * ipitch + x] : 0; sum += vals[i]; }
    } else {
      for (int y = y0; y < y1; y++) sum += I[(size_t)y * ipitch + x];
    }
  }
  band[warp][lane] = sum;
  __syncthreads();
  if (warp == 0) {
    int acc = 0;
    for (int w = 0; w < 32; w++) { const int v = band[w][lane]; band[w][lane] = acc; acc += v; }
  }
  __syncthreads();
  if (x > W) return;
  int acc = band[warp][lane];
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < kKeep; i++) if (y0 + i < y1) { acc += vals[i]; I[(size_t)(y0 + i) * ipitch + x] = acc; }
  } else {
    for (int y = y0; y < y1; y++) { acc += I[(size_t)y * ipitch + x]; I[(size_t)y * ipitch + x] = acc; }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// descriptor: one warp per keypoint
__global__ void __launch_bounds__(128) k_describe(const uint8_t* in0, int in_pitch, size_t in_frame_stride, int H,
                                                  const int32_t* integral, int ipitch, const PatternPoint* pattern,
                                                  const uint32_t* short_pairs, const int4* long_pairs,
                                                  okb_keypoint_t* kp, const int32_t* kscale, const int32_t* count,
                                                  int kp_cap, uint8_t* desc)
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
  const int sc = kscale[(size_t)frame * kp_cap + k];
  int* val = values[warp];
  const PatternPoint* pat0 = pattern + ((size_t)sc * kRot) * kPoints;
  for (int i = lane; i < kPoints; i += 32) val[i] = smoothed_intensity(img, in_pitch, I, ipitch, kx, ky, pat0[i]);
  __syncwarp();
  int d0 = 0, d1 = 0;
  for (int q = lane; q < kLongPairs; q += 32) {
    const int4 lp = __ldg(&long_pairs[q]);
    const int dt = val[lp.x] - val[lp.y];
    d0 += dt * lp.z / 1024;
    d1 += dt * lp.w / 1024;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o); }
  float angle = (float)(atan2((double)(float)d1, (double)(float)d0) / 3.14159265358979323846 * 180.0);
  int theta = (int)((double)kRot * ((double)angle / 360.0) + 0.5);
  if (theta < 0) theta += kRot;
  if (theta >= kRot) theta -= kRot;
  if (angle < 0) angle += 360.f;
  __syncwarp();
  const PatternPoint* pat = pattern + ((size_t)sc * kRot + theta) * kPoints;
  for (int i = lane; i < kPoints; i += 32) val[i] = smoothed_intensity(img, in_pitch, I, ipitch, kx, ky, pat[i]);
  __syncwarp();
  uint32_t mine = 0;
#pragma unroll
  for (int w = 0; w < 16; w++) {
    const uint32_t pr = __ldg(&short_pairs[w * 32 + lane]);
    const uint32_t word = __ballot_sync(0xffffffffu, val[pr & 255] > val[pr >> 8]);
    if (lane == w) mine = word;
  }
  if (lane < 16) reinterpret_cast<uint32_t*>(desc + ((size_t)frame * kp_cap + k) * 64)[lane] = mine;
  if (lane == 0) kpp->angle = angle;
}

// ---------------------------------------------------------------------------------------------------------------
// host side
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int detect_init_camera(okb_context* ctx, int cam)
{
  CamWorkspace& ws = ctx->cams[cam];
  const okb_camera_config_t& c = ws.cfg;
  const int W = c.width, H = c.height;
  ws.n_layers = c.octaves == 0 ? 1 : 2 * c.octaves;
  if (ws.n_layers > kMaxLayers || W >= 2048 || H >= 2048 || W < 16 || H < 16) {
    set_error("unsupported geometry: %dx%d, octaves %d (max 2047x2047, 4 octaves)", W, H, c.octaves);
    return OKB_ERR_ARGUMENT;
  }
  auto set = [&](int i, int w, int h, float scale, int parent) {
    LayerGeom& g = ws.geom[i];
    g.w = w; g.h = h; g.pitch = (int)align_up((size_t)w, 64); g.scale = scale;
    g.offset_px = i == 0 ? 0.f : 0.5f * scale - 0.5f; g.parent = parent;
  };
  set(0, W, H, 1.0f, -1);
  if (ws.n_layers > 1) set(1, 2 * (W / 3), 2 * (H / 3), 1.5f, 0);
  for (int i = 2; i < ws.n_layers; i += 2) {
    set(i, ws.geom[i - 2].w / 2, ws.geom[i - 2].h / 2, ws.geom[i - 2].scale * 2, i - 2);
    set(i + 1, ws.geom[i - 1].w / 2, ws.geom[i - 1].h / 2, ws.geom[i - 1].scale * 2, i - 1);
  }
  size_t off = 0;
  ws.ps_bytes = (int64_t)W * H;  // read of the base image
  for (int i = 0; i < ws.n_layers; i++) {
    LayerGeom& g = ws.geom[i];
    if (g.w < 8 || g.h < 8) { set_error("layer %d too small (%dx%d)", i, g.w, g.h); return OKB_ERR_ARGUMENT; }
    g.offset = off; off += align_up((size_t)g.pitch * g.h, 256);
    ws.ps_bytes += (int64_t)g.w * g.h * (i == 0 ? 1 : 2);  // score map write (+ layer image write for i > 0)
    if (i > 0) {
      const LayerGeom& p = ws.geom[g.parent];
      const double sx = 1. / ((double)g.w / p.w), sy = 1. / ((double)g.h / p.h);
      g.fast2 = (sx == 2.0 && sy == 2.0) ? 1 : 0;
      if (!g.fast2) {
        AreaAxis ax, ay; build_area_axis(p.w, g.w, ax); build_area_axis(p.h, g.h, ay);
        OKB_CUDA(cudaMalloc(&g.d_xs, g.w * 4)); OKB_CUDA(cudaMalloc(&g.d_xn, g.w * 4)); OKB_CUDA(cudaMalloc(&g.d_xa, g.w * 16));
        OKB_CUDA(cudaMalloc(&g.d_ys, g.h * 4)); OKB_CUDA(cudaMalloc(&g.d_yn, g.h * 4)); OKB_CUDA(cudaMalloc(&g.d_ya, g.h * 16));
        OKB_CUDA(cudaMemcpy(g.d_xs, ax.start.data(), g.w * 4, cudaMemcpyHostToDevice));
        OKB_CUDA(cudaMemcpy(g.d_xn, ax.count.data(), g.w * 4, cudaMemcpyHostToDevice));
        OKB_CUDA(cudaMemcpy(g.d_xa, ax.alpha.data(), g.w * 16, cudaMemcpyHostToDevice));
        OKB_CUDA(cudaMemcpy(g.d_ys, ay.start.data(), g.h * 4, cudaMemcpyHostToDevice));
        OKB_CUDA(cudaMemcpy(g.d_yn, ay.count.data(), g.h * 4, cudaMemcpyHostToDevice));
        OKB_CUDA(cudaMemcpy(g.d_ya, ay.alpha.data(), g.h * 16, cudaMemcpyHostToDevice));
      }
    }
  }
  ws.dl.n = ws.n_layers; ws.dl.frame_stride = (uint32_t)off;
  for (int i = 0; i < ws.n_layers; i++) {
    const LayerGeom& g = ws.geom[i];
    ws.dl.l[i] = DeviceLayer{g.w, g.h, g.pitch, (uint32_t)g.offset, g.scale, g.offset_px};
  }
  const int B = c.max_batch;
  {
    // candidate list: one region per layer, sized by the layer's share of the pixels with 6x slack
    const int base = std::max(16384, (W * H / 32 + 1023) / 1024 * 1024);
    long long area = 0;
    for (int i = 0; i < ws.n_layers; i++) area += (long long)ws.geom[i].w * ws.geom[i].h;
    int acc = 0;
    for (int i = 0; i < kMaxLayers; i++) {
      ws.cand_off[i] = acc;
      if (i < ws.n_layers)   // coarse layers carry more corners per pixel: generous floors, capped at the single-layer size
        acc += (int)align_up((size_t)std::min((long long)base, std::max(4096LL, 6LL * base * ws.geom[i].w * ws.geom[i].h / area)), 128);
    }
    ws.cand_off[kMaxLayers] = acc;
    ws.cand_cap = acc;
  }
  ws.kp_cap = c.max_keypoints > 0 ? (int)align_up((size_t)c.max_keypoints, 64) : kSortCap;
  OKB_CUDA(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
  OKB_CUDA(cudaStreamCreateWithFlags(&ws.stream2, cudaStreamNonBlocking));
  OKB_CUDA(cudaEventCreateWithFlags(&ws.ev_fork, cudaEventDisableTiming));
  OKB_CUDA(cudaEventCreateWithFlags(&ws.ev_join, cudaEventDisableTiming));
  OKB_CUDA(cudaMalloc(&ws.d_in, (size_t)W * H * B));
  OKB_CUDA(cudaMalloc(&ws.d_img, off * B));
  OKB_CUDA(cudaMalloc(&ws.d_score, off * B));
  OKB_CUDA(cudaMalloc(&ws.d_touch, off * B * 4));
  OKB_CUDA(cudaMemset(ws.d_touch, 0, off * B * 4));
  OKB_CUDA(cudaMalloc(&ws.d_tie_cells, (size_t)kCellWordsPerFrame * B * 4));
  OKB_CUDA(cudaMemset(ws.d_score, 0, off * B));
  OKB_CUDA(cudaMemset(ws.d_img, 0, off * B));
  OKB_CUDA(cudaMalloc(&ws.d_integral, (size_t)(W + 1) * (H + 1) * 4 * B));
  OKB_CUDA(cudaMalloc(&ws.d_cand, (size_t)ws.cand_cap * 4 * B));
  OKB_CUDA(cudaMalloc(&ws.d_cand_count, 4 * kMaxLayers * B));
  OKB_CUDA(cudaMalloc(&ws.d_rec, (size_t)ws.cand_cap * sizeof(CandRecord) * B));
  OKB_CUDA(cudaMalloc(&ws.d_kp, (size_t)ws.kp_cap * sizeof(okb_keypoint_t) * B));
  OKB_CUDA(cudaMalloc(&ws.d_kscale, (size_t)ws.kp_cap * 4 * B));
  OKB_CUDA(cudaMalloc(&ws.d_desc, (size_t)ws.kp_cap * 64 * B));
  OKB_CUDA(cudaMalloc(&ws.d_count, 4 * B));
  OKB_CUDA(cudaMalloc(&ws.d_status, 4 * B));
  OKB_CUDA(cudaMalloc(&ws.d_rays, (size_t)ws.kp_cap * 24 * B));
  OKB_CUDA(cudaMalloc(&ws.d_rays_valid, (size_t)ws.kp_cap * B));
  OKB_CUDA(cudaEventCreateWithFlags(&ws.ev_done, cudaEventDisableTiming));
  OKB_CUDA(cudaMalloc(&ws.d_dbg, (size_t)16 * 8 * B));
  OKB_CUDA(cudaMemset(ws.d_dbg, 0, (size_t)16 * 8 * B));
  OKB_CUDA(cudaMemset(ws.d_count, 0, 4 * B));
  OKB_CUDA(cudaMemset(ws.d_status, 0, 4 * B));
  OKB_CUDA(cudaMallocHost(&ws.h_img, (size_t)W * H * B));
  OKB_CUDA(cudaMallocHost(&ws.h_kp, (size_t)ws.kp_cap * sizeof(okb_keypoint_t) * B));
  OKB_CUDA(cudaMallocHost(&ws.h_desc, (size_t)ws.kp_cap * 64 * B));
  OKB_CUDA(cudaMallocHost(&ws.h_count, 4 * B));
  OKB_CUDA(cudaMallocHost(&ws.h_status, 4 * B));
  OKB_CUDA(cudaMallocHost(&ws.h_rays, (size_t)ws.kp_cap * 24 * B));
  OKB_CUDA(cudaMallocHost(&ws.h_rays_valid, (size_t)ws.kp_cap * B));
  for (int i = 0; i < 4; i++) OKB_CUDA(cudaEventCreate(&ws.ev[i]));
  OKB_CUDA(cudaEventCreate(&ws.ev_mid));
  OKB_CUDA(cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinalizeSmem));
  OKB_CUDA(cudaFuncSetAttribute(k_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, kResolveSmem));
  return OKB_OK;
}

void detect_free_camera(okb_context* ctx, int cam)
{
  CamWorkspace& ws = ctx->cams[cam];
  for (int i = 0; i < kMaxLayers; i++) {
    LayerGeom& g = ws.geom[i];
    cudaFree(g.d_xs); cudaFree(g.d_xn); cudaFree(g.d_xa); cudaFree(g.d_ys); cudaFree(g.d_yn); cudaFree(g.d_ya);
  }
  cudaFree(ws.d_in); cudaFree(ws.d_img); cudaFree(ws.d_score); cudaFree(ws.d_touch); cudaFree(ws.d_tie_cells); cudaFree(ws.d_integral); cudaFree(ws.d_cand);
  cudaFree(ws.d_cand_count); cudaFree(ws.d_rec); cudaFree(ws.d_kp); cudaFree(ws.d_kscale); cudaFree(ws.d_desc);
  cudaFree(ws.d_count); cudaFree(ws.d_status); cudaFree(ws.d_m1_rows); cudaFree(ws.m_d); if (ws.m_h) cudaFreeHost(ws.m_h); cudaFree(ws.d_dbg); cudaFree(ws.d_rays); cudaFree(ws.d_rays_valid);
  if (ws.ev_done) cudaEventDestroy(ws.ev_done);
  cudaFreeHost(ws.h_img); cudaFreeHost(ws.h_kp); cudaFreeHost(ws.h_desc); cudaFreeHost(ws.h_count); cudaFreeHost(ws.h_status); cudaFreeHost(ws.h_rays); cudaFreeHost(ws.h_rays_valid);
  for (int i = 0; i < 4; i++) if (ws.ev[i]) cudaEventDestroy(ws.ev[i]);
  if (ws.ev_mid) cudaEventDestroy(ws.ev_mid);
  if (ws.ev_fork) cudaEventDestroy(ws.ev_fork);
  if (ws.ev_join) cudaEventDestroy(ws.ev_join);
  if (ws.stream2) cudaStreamDestroy(ws.stream2);
  if (ws.stream) cudaStreamDestroy(ws.stream);
}

static void collect_timing(okb_context* ctx, CamWorkspace& ws)
{
  if (!ws.pending_timing) return;
  cudaEventSynchronize(ws.ev[3]);
  float a = 0, b = 0;
  float cth = 0;
  cudaEventElapsedTime(&a, ws.ev[0], ws.ev[1]);
  cudaEventElapsedTime(&b, ws.ev[0], ws.ev[3]);
  cudaEventElapsedTime(&cth, ws.ev_mid, ws.ev[1]);
  ws.ps_ms += a; ws.total_ms += b; ws.score_ms += cth;
  ws.pending_timing = 0;
  (void)ctx;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_encode_tried = 0;

static bool encode_map(CUtensorMap* m, const void* base, int w, int h, size_t pitch, size_t frame_stride, int frames)
{
  if (!g_encode) return false;
  if ((((uintptr_t)base) & 15) || (pitch & 15) || (frame_stride & 15) || pitch == 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
  const cuuint32_t box[3] = {(cuuint32_t)kImgW, (cuuint32_t)kImgH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tensor maps of all layers for this call (layer 0 lives in the caller's buffer, so its map is re-encoded per call;
// the others are cached in the workspace). Layers whose geometry TMA cannot address fall back to plain loads.
static int build_tma_maps(CamWorkspace& ws, const uint8_t* d_images, int src_pitch, size_t in_stride, int frames, TmaMaps& maps)
{
  if (!g_encode_tried) {
    g_encode_tried = 1;
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
  }
  memset(&maps, 0, sizeof(maps));
  if (!ws.tma_ready) {
    for (int i = 1; i < ws.n_layers; i++) {
      const LayerGeom& g = ws.geom[i];
      ws.tma_use[i] = encode_map(&ws.tma[i], ws.d_img + g.offset, g.w, g.h, (size_t)g.pitch, (size_t)ws.dl.frame_stride, ws.cfg.max_batch) ? 1 : 0;
    }
    ws.tma_ready = 1;
  }
  for (int i = 1; i < ws.n_layers; i++) { maps.m[i] = ws.tma[i]; maps.use[i] = ws.tma_use[i]; }
  maps.use[0] = encode_map(&maps.m[0], d_images, ws.geom[0].w, ws.geom[0].h, (size_t)src_pitch, in_stride, frames) ? 1 : 0;
  return OKB_OK;
}

// all frames are device resident: d_images = n_frames x H x src_pitch
int detect_run_device(okb_context* ctx, int cam, int n_frames, const uint8_t* d_images, int src_pitch)
{
  CamWorkspace& ws = ctx->cams[cam];
  const okb_camera_config_t& c = ws.cfg;
  const int W = c.width, H = c.height, B = n_frames;
  cudaStream_t st = ws.stream;
  if (ctx->timers_on) { collect_timing(ctx, ws); cudaEventRecord(ws.ev[0], st); }
  ws.epoch++;
  if (ws.epoch >= 127) {  // epoch field is 7 bits: recycle
    OKB_CUDA(cudaMemsetAsync(ws.d_touch, 0, (size_t)ws.dl.frame_stride * c.max_batch * 4, st));
    ws.epoch = 1;
  }
  const size_t in_stride = (size_t)src_pitch * H;
  // ---- fork: the integral image only needs the input frames; it runs on a side stream underneath the detection
  //      kernels (several of which are latency-bound single-CTA-per-frame kernels that leave most SMs idle)
  const int ipitch = W + 1;
  OKB_CUDA(cudaEventRecord(ws.ev_fork, st));
  OKB_CUDA(cudaStreamWaitEvent(ws.stream2, ws.ev_fork, 0));
  k_integral_rows<<<dim3((H + 7) / 8, B), 256, 0, ws.stream2>>>(d_images, src_pitch, in_stride, W, H, ws.d_integral, ipitch);
  k_integral_cols<<<dim3((W + 31) / 32, B), 1024, 0, ws.stream2>>>(W, H, ws.d_integral, ipitch);
  OKB_CUDA(cudaEventRecord(ws.ev_join, ws.stream2));
  // ---- pyramid: layers whose parents are complete can share a launch
  for (int i = 1; i < ws.n_layers;) {
    ResizeJobs jobs; jobs.n = 0;
    int total = 0;
    auto add = [&](int li) {
      const LayerGeom& g = ws.geom[li]; const LayerGeom& p = ws.geom[g.parent];
      ResizeJob& J = jobs.j[jobs.n];
      if (g.parent == 0) { J.src = d_images; J.src_pitch = src_pitch; J.src_frame_stride = in_stride; }
      else { J.src = ws.d_img + p.offset; J.src_pitch = p.pitch; J.src_frame_stride = ws.dl.frame_stride; }
      J.dst = ws.d_img + g.offset; J.dst_pitch = g.pitch; J.dst_frame_stride = ws.dl.frame_stride;
      J.dw = g.w; J.dh = g.h; J.fast2 = g.fast2;
      J.xs = g.d_xs; J.xn = g.d_xn; J.ys = g.d_ys; J.yn = g.d_yn; J.xa = g.d_xa; J.ya = g.d_ya;
      if (g.fast2) { jobs.tiles_x[jobs.n] = (g.w + 127) / 128; jobs.tiles_y[jobs.n] = (g.h + 7) / 8; }
      else { jobs.tiles_x[jobs.n] = (g.w + 31) / 32; jobs.tiles_y[jobs.n] = (g.h + 31) / 32; }
      total += jobs.tiles_x[jobs.n] * jobs.tiles_y[jobs.n];
      jobs.n++;
    };
    // layer i (odd, from i-2 or 0) and layer i+1 (even, from i-1): both parents have index < i
    add(i);
    if (i + 1 < ws.n_layers) add(i + 1);
    if (jobs.n == 1) { jobs.tiles_x[1] = jobs.tiles_y[1] = 0; }
    k_resize<<<dim3(total, B), 256, 0, st>>>(jobs);
    ctx->launches++; if (ctx->timers_on) ws.ps_launches++;
    i += 2;
  }
  if (ctx->timers_on) cudaEventRecord(ws.ev_mid, st);
  // ---- scores
  TileMap tm; tm.n_layers = ws.n_layers; tm.tile_prefix[0] = 0;
  for (int i = 0; i < ws.n_layers; i++) {
    tm.tiles_x[i] = (ws.geom[i].w + kTileW - 1) / kTileW;
    tm.tile_prefix[i + 1] = tm.tile_prefix[i] + tm.tiles_x[i] * ((ws.geom[i].h + kTileH - 1) / kTileH);
  }
  for (int i = ws.n_layers + 1; i <= kMaxLayers; i++) tm.tile_prefix[i] = tm.tile_prefix[ws.n_layers];
  const int n_tiles = tm.tile_prefix[ws.n_layers];
  OKB_CUDA(cudaMemsetAsync(ws.d_cand_count, 0, 4 * kMaxLayers * B, st));
  CandRegions cr; for (int i = 0; i <= kMaxLayers; i++) cr.off[i] = ws.cand_off[i];
  OKB_CUDA(cudaMemsetAsync(ws.d_status, 0, 4 * B, st));
  OKB_CUDA(cudaMemsetAsync(ws.d_tie_cells, 0, (size_t)kCellWordsPerFrame * B * 4, st));
  TmaMaps maps;
  { int rc = build_tma_maps(ws, d_images, src_pitch, in_stride, c.max_batch, maps); if (rc) return rc; }
  k_score_nms<<<dim3(n_tiles, B), kScoreThreads, 0, st>>>(maps, ws.dl, tm, d_images, src_pitch, in_stride, ws.d_img, ws.d_score,
                                                          ws.d_cand, ws.d_cand_count, ws.cand_cap, cr, c.threshold, ws.d_status, ws.d_tie_cells);
  ctx->launches++; if (ctx->timers_on) ws.ps_launches++;
  if (ctx->timers_on) cudaEventRecord(ws.ev[1], st);
  // ---- candidates, refinement, tie resolution, selection
  k_refine<<<dim3((ws.cand_cap + 127) / 128, B), 128, 0, st>>>(ws.dl, d_images, src_pitch, in_stride, ws.d_img, ws.d_score,
                                                               ws.d_touch, ws.d_cand, ws.d_cand_count, ws.cand_cap, cr,
                                                               ws.d_rec, c.threshold, ws.epoch, ws.d_tie_cells);
  k_resolve<<<B, 512, kResolveSmem, st>>>(ws.dl, ws.d_score, ws.d_touch, ws.d_cand_count, ws.cand_cap, cr, ws.d_rec, ws.epoch, c.threshold,
                                          ws.d_status, ws.d_dbg);
  k_finalize<<<B, 1024, kFinalizeSmem, st>>>(ws.dl, ws.d_cand_count, ws.cand_cap, cr, ws.d_rec, ctx->d_scale_bounds, ctx->d_size_list,
                                            W, H, c.max_keypoints, ws.kp_cap, ws.d_kp, ws.d_kscale, ws.d_count, ws.d_status, ws.d_dbg);
  ctx->launches += 3;
  if (ctx->timers_on) cudaEventRecord(ws.ev[2], st);
  // ---- descriptors (the integral image was produced on the side stream, see the fork above)
  OKB_CUDA(cudaStreamWaitEvent(st, ws.ev_join, 0));
  k_describe<<<dim3((ws.kp_cap + 3) / 4, B), 128, 0, st>>>(d_images, src_pitch, in_stride, H, ws.d_integral, ipitch,
                                                           ctx->d_pattern, ctx->d_short_pairs, ctx->d_long_pairs, ws.d_kp,
                                                           ws.d_kscale, ws.d_count, ws.kp_cap, ws.d_desc);
  ctx->launches += 3;
  { int rc = camera_backproject_batch(ctx, cam, B); if (rc) return rc; }
  if (ctx->timers_on) { cudaEventRecord(ws.ev[3], st); ws.pending_timing = 1; }
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

void detect_collect_timing(okb_context* ctx, int cam) { collect_timing(ctx, ctx->cams[cam]); }

}  // namespace okb
