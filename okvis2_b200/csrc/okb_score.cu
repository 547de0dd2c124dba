// okb_score.cu -- the pyramid + score pass of the detector (sm_100a): two launches per batch of frames.
//
//   k_pyramid     every reduced layer of the scale space in ONE launch. A work item owns a 128 x 32 tile of the first
//                 layer of a chain and carries it down the chain in shared memory:
//                   chain A: layer 2 = 2x2 mean of layer 0 (integer, (a+b+c+d+2)>>2), then layers 4, 6 by halving;
//                   chain B: layer 1 = cv::resize(INTER_AREA) 2/3-sample of layer 0 (table driven float taps, the source
//                            region staged in shared memory by 16-byte loads), then layers 3, 5, 7 by halving.
//                 Every layer is written once, with 16-byte stores. Layers whose size is not an exact half of their
//                 parent (odd sizes) fall back to the table-driven k_resize.
//   k_score_nms   dense AGAST 9-16 score b0 of every layer + 3x3 non-max candidates, persistent CTAs that walk the
//                 (tile, frame) items; the image tile of the NEXT item is in flight (TMA bulk tensor copy, mbarrier) while
//                 the current one is scored from shared memory.
// Behind okvis::Frame::detect (reference okvis_cv/include/okvis/implementation/Frame.hpp:140-154); arithmetic =
// OpenCV-BRISK's, checked bit for bit against cv2 goldens by tests/test_gpu_detect.py.
#include <cuda.h>   // CUtensorMap type and enums only; the encoder is resolved through cudaGetDriverEntryPoint
#include <string.h>

#include <algorithm>
#include <vector>

#include "okb_detect.h"

namespace okb {

// ---------------------------------------------------------------------------------------------------------------
// fallback pyramid kernel (layers that are not an exact 2x2 reduction of an even-sized parent)
struct ResizeJob {
  const uint8_t* src; int src_pitch; size_t src_frame_stride;
  uint8_t* dst; int dst_pitch; size_t dst_frame_stride;
  int dw, dh, fast2;
  const int *xs, *xn, *ys, *yn; const float *xa, *ya;
};

__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ ResizeJob J, int tiles_x)
{
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int frame = blockIdx.y;
  const uint8_t* src = J.src + (size_t)frame * J.src_frame_stride;
  uint8_t* dst = J.dst + (size_t)frame * J.dst_frame_stride;
  // tile = 32 x 32 destination pixels, 4 rows per thread
  const int x = tx * 32 + (threadIdx.x & 31);
  const int y0 = ty * 32 + (threadIdx.x >> 5) * 4;
  if (x >= J.dw) return;
  if (J.fast2) {
#pragma unroll
    for (int r = 0; r < 4; r++) if (y0 + r < J.dh) dst[(size_t)(y0 + r) * J.dst_pitch + x] = half_pixel(src, J.src_pitch, x, y0 + r);
    return;
  }
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

// ---------------------------------------------------------------------------------------------------------------
// chained pyramid
constexpr int kPyrW = 128, kPyrH = 32;     // tile of the first layer of a chain
constexpr int kPyrThreads = 256;
constexpr int kSrcRows = 64, kSrcPitch = 256;   // staged source region of a general (INTER_AREA) tile
constexpr int kSrcPad = 4;                       // zero-weight taps may read up to 3 rows / bytes past the region

struct PyrChain {
  int n;                  // layers of the chain
  int layer[4];           // their indices
  int general;            // first layer: 1 = general INTER_AREA tables from layer 0, 0 = 2x2 mean of layer 0
  int taps;               // general: largest tap count of the tables (3 for the 2/3-sample layer)
  int tiles_x, tiles;     // tiles of the first layer
  const int *xs, *xn, *ys, *yn; const float *xa, *ya;   // general: tap tables of the first layer
};
struct PyrArgs { PyrChain c[2]; int n_chains; };

__device__ __forceinline__ float u8_to_float(uint32_t b)
{ // exact, on the ALU/FMA pipes (I2F would go to the quarter-rate XU pipe): 2^23 + b is representable
  return __uint_as_float(0x4B000000u | b) - 8388608.0f;
}

// one destination pixel of cv::resize(INTER_AREA) from the staged source bytes with a FIXED tap count: the taps beyond the
// table's count carry weight 0 (build_area_axis leaves them 0), and x + w * 0 = x exactly, so the value is the one
// area_pixel computes -- without data-dependent loop bounds. p = first tap of the first row.
template <int TAPS>
__device__ __forceinline__ uint8_t area_pixel_fixed(const uint8_t* p, int pitch, const float (&xa)[4], const float* wy)
{
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < TAPS; j++) {
    float buf = 0.f;
#pragma unroll
    for (int i = 0; i < TAPS; i++) buf += u8_to_float(p[j * pitch + i]) * xa[i];
    if (j == 0) sum = wy[0] * buf; else sum += wy[j] * buf;
  }
  const int q = __float2int_rn(sum);
  return (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
}

__global__ void __launch_bounds__(kPyrThreads) k_pyramid(const __grid_constant__ PyrArgs args, const __grid_constant__ DeviceLayers dl,
                                                         const uint8_t* in0, int in_pitch, size_t in_frame_stride, uint8_t* img_block)
{
  __shared__ __align__(16) uint8_t lv[kPyrH * kPyrW + (kPyrH / 2) * (kPyrW / 2) + (kPyrH / 4) * (kPyrW / 4) + (kPyrH / 8) * (kPyrW / 8)];
  __shared__ __align__(16) uint8_t stage[kSrcRows + kSrcPad][kSrcPitch + 16];
  __shared__ int s_sy[kPyrH];
  __shared__ float s_wy[kPyrH][4];
  const int frame = blockIdx.y;
  int t = blockIdx.x, ci = 0;
  if (t >= args.c[0].tiles) { t -= args.c[0].tiles; ci = 1; }
  const PyrChain& ch = args.c[ci];
  const int tx = t % ch.tiles_x, ty = t / ch.tiles_x;
  const uint8_t* src = in0 + (size_t)frame * in_frame_stride;
  uint8_t* out_frame = img_block + (size_t)frame * dl.frame_stride;
  const DeviceLayer d0 = dl.l[ch.layer[0]];
  const int x0 = tx * kPyrW, y0 = ty * kPyrH;
  const int tw = min(kPyrW, d0.w - x0), th = min(kPyrH, d0.h - y0);   // valid part of the tile
  uint8_t* t0 = lv;
  const int tid = threadIdx.x;
  const bool vec = ((in_pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (!ch.general) {
    // ---- layer 2 tile = 2x2 means of the 256 x 64 region of layer 0: 8 pixels per task from two 16-byte loads
#pragma unroll
    for (int k = 0; k < (kPyrW / 8) * kPyrH / kPyrThreads; k++) {
      const int task = tid + k * kPyrThreads;
      const int r = task / (kPyrW / 8), c8 = (task % (kPyrW / 8)) * 8;
      uint2 o = make_uint2(0, 0);
      if (r < th && c8 < tw) {
        const uint8_t* r0 = src + (size_t)(2 * (y0 + r)) * in_pitch + 2 * (x0 + c8);
        if (vec && c8 + 8 <= tw) {
          const uint4 a = *reinterpret_cast<const uint4*>(r0), b = *reinterpret_cast<const uint4*>(r0 + in_pitch);
          // per 32-bit word: the four 2x2 sums of (a, b) as two 16x2 lanes, + 2, >> 2, packed back to two bytes
          auto px2 = [](uint32_t u, uint32_t v) {
            const uint32_t lo = (u & 0x00ff00ffu) + ((u >> 8) & 0x00ff00ffu) + (v & 0x00ff00ffu) + ((v >> 8) & 0x00ff00ffu) + 0x00020002u;
            return __byte_perm(lo >> 2, 0, 0x4420);   // bytes 0 and 2 (each sum < 1024: no carry between the lanes)
          };
          o.x = px2(a.x, b.x) | (px2(a.y, b.y) << 16);
          o.y = px2(a.z, b.z) | (px2(a.w, b.w) << 16);
        } else {
          uint32_t w[2] = {0, 0};
          for (int i = 0; i < 8 && c8 + i < tw; i++) w[i >> 2] |= (uint32_t)half_pixel(src, in_pitch, x0 + c8 + i, y0 + r) << (8 * (i & 3));
          o.x = w[0]; o.y = w[1];
        }
      }
      *reinterpret_cast<uint2*>(&t0[r * kPyrW + c8]) = o;
    }
  } else {
    // ---- layer 1 tile by the INTER_AREA tap tables; the source region goes through shared memory
    const int ys0 = ch.ys[y0], yl = y0 + th - 1, ys1 = ch.ys[yl] + ch.yn[yl];       // source rows [ys0, ys1)
    const int xl = x0 + tw - 1;
    const int xs0 = ch.xs[x0] & ~15, xs1 = ch.xs[xl] + ch.xn[xl];                   // source columns [xs0, xs1)
    const int rows = ys1 - ys0, cols = xs1 - xs0;
    const bool staged = rows <= kSrcRows && cols <= kSrcPitch && ch.taps <= 4;
    if (staged) {
      const int chunks = (cols + 15) >> 4;
      for (int i = tid; i < rows * chunks; i += kPyrThreads) {
        const int r = i / chunks, c = (i - r * chunks) * 16;
        const uint8_t* g = src + (size_t)(ys0 + r) * in_pitch + xs0 + c;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (vec && xs0 + c + 16 <= in_pitch) v = *reinterpret_cast<const uint4*>(g);
        else {
          uint32_t w[4] = {0, 0, 0, 0};
          for (int b = 0; b < 16; b++) if (xs0 + c + b < dl.l[0].w) w[b >> 2] |= (uint32_t)g[b] << (8 * (b & 3));
          v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        *reinterpret_cast<uint4*>(&stage[r][c]) = v;
      }
      if (tid < kPyrH) {
        const int y = min(y0 + tid, d0.h - 1);
        s_sy[tid] = ch.ys[y] - ys0;
#pragma unroll
        for (int j = 0; j < 4; j++) s_wy[tid][j] = ch.ya[y * 4 + j];
      }
    }
    __syncthreads();
    // thread -> column x0 + (tid & 127), rows (tid >> 7) + 2 i: the x taps are loaded once per thread
    const int cx = tid & (kPyrW - 1);
    const bool col_ok = cx < tw;
    int sx = 0, nx = 0; float xa[4] = {0.f, 0.f, 0.f, 0.f};
    if (col_ok) {
      sx = ch.xs[x0 + cx]; nx = ch.xn[x0 + cx];
#pragma unroll
      for (int i = 0; i < 4; i++) xa[i] = ch.xa[(x0 + cx) * 4 + i];
    }
    if (staged) {
      const uint8_t* pcol = &stage[0][col_ok ? sx - xs0 : 0];
#pragma unroll 4
      for (int r = tid >> 7; r < kPyrH; r += kPyrThreads / kPyrW) {
        const uint8_t* p = pcol + s_sy[r] * (kSrcPitch + 16);
        uint8_t val = ch.taps <= 3 ? area_pixel_fixed<3>(p, kSrcPitch + 16, xa, s_wy[r]) : area_pixel_fixed<4>(p, kSrcPitch + 16, xa, s_wy[r]);
        if (!(col_ok && r < th)) val = 0;
        t0[r * kPyrW + cx] = val;
      }
    } else {
      for (int r = tid >> 7; r < kPyrH; r += kPyrThreads / kPyrW) {
        uint8_t val = 0;
        if (col_ok && r < th) {
          const int y = y0 + r;
          float ya[4];
#pragma unroll
          for (int i = 0; i < 4; i++) ya[i] = ch.ya[y * 4 + i];
          val = area_pixel(src, in_pitch, sx, nx, xa, ch.ys[y], ch.yn[y], ya);
        }
        t0[r * kPyrW + cx] = val;
      }
    }
  }
  __syncthreads();
  // ---- write the first layer, then halve down the chain inside shared memory
  uint8_t* cur = t0; int cw = kPyrW, chh = kPyrH, cx0 = x0, cy0 = y0;
  for (int li = 0; li < ch.n; li++) {
    const DeviceLayer d = dl.l[ch.layer[li]];
    uint8_t* dst = out_frame + d.offset;
    // 16-byte stores of the tile rows (pitch and tile origin are multiples of 16; columns >= d.w hold zeros)
    const int lc = 3 - li;   // log2 of the 16-byte chunks per row: 8, 4, 2, 1
    for (int i = tid; i < (chh << lc); i += kPyrThreads) {
      const int r = i >> lc, c = (i & ((1 << lc) - 1)) * 16;
      if (cy0 + r < d.h && cx0 + c < d.pitch)
        *reinterpret_cast<uint4*>(dst + (size_t)(cy0 + r) * d.pitch + cx0 + c) = *reinterpret_cast<const uint4*>(&cur[r * cw + c]);
    }
    if (li + 1 == ch.n) break;
    const DeviceLayer dn = dl.l[ch.layer[li + 1]];
    uint8_t* nxt = cur + cw * chh;
    const int nw = cw >> 1, nh = chh >> 1, nx0 = cx0 >> 1, ny0 = cy0 >> 1;
    const int lw = 6 - li;   // log2(nw): 64, 32, 16
    // two destination pixels per task from two 32-bit loads
    for (int i = tid; i < (nw * nh) >> 1; i += kPyrThreads) {
      const int r = i >> (lw - 1), c = (i & ((nw >> 1) - 1)) * 2;
      const uint32_t u = *reinterpret_cast<const uint32_t*>(&cur[(2 * r) * cw + 2 * c]);
      const uint32_t v = *reinterpret_cast<const uint32_t*>(&cur[(2 * r + 1) * cw + 2 * c]);
      const uint32_t lo = (u & 0x00ff00ffu) + ((u >> 8) & 0x00ff00ffu) + (v & 0x00ff00ffu) + ((v >> 8) & 0x00ff00ffu) + 0x00020002u;
      uint32_t o = __byte_perm(lo >> 2, 0, 0x4420);
      if (ny0 + r >= dn.h) o = 0;
      if (nx0 + c >= dn.w) o = 0; else if (nx0 + c + 1 >= dn.w) o &= 0xffu;
      *reinterpret_cast<uint16_t*>(&nxt[r * nw + c]) = (uint16_t)o;
    }
    __syncthreads();
    cur = nxt; cw = nw; chh = nh; cx0 = nx0; cy0 = ny0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// score map tiling: tiles of kTileW x TH pixels, TH = 8 rows per warp (TH = 64: 256 threads, TH = 32: 128 threads)
constexpr int kTileW = 64;
constexpr int kRowsPerWarp = 8;   // the strong-pixel bits of a lane's 8 rows live in two bytes of one register

// (layer, x0, y0) of every tile of a frame, built once per workspace
struct TileEntry { int layer, x0, y0, pad; };

// Dense AGAST 9-16 score b0 = clamp(B*, 0, 254) of two horizontally adjacent pixels in 16x2 SIMD:
// B* = max(max_arcs min_arc(ring) - p, p - min_arcs max_arc(ring)) - 1 over the 16 arcs of 9 ring pixels.
// 80 VIMNMX3.U16x2 for the window minima / maxima (min3 of min3, then max3 over the arcs) and 4 instructions for the rest:
// with ~p = -p - 1 per half,  b0 = max(bb + ~p, p + ~bd, 0)  (VIADD.16x2, VIADDMNMX.S16x2.RELU); b0 <= 254 by itself.
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
  const uint32_t bright = __vadd2(bb, ~p);                  // bb - p - 1 per half (signed)
  return __viaddmax_s16x2_relu(p, ~bd, bright);             // max(p - bd - 1, bb - p - 1, 0)
}

// ---- TMA (cp.async.bulk.tensor) staging of the image tiles -----------------------------------------------------
// One 3-D tensor map per layer: (x: w bytes, y: h rows of `pitch` bytes, frame). A single elected thread issues one
// bulk tensor copy of the kImgW x (TH + 6) box per tile; out-of-image elements are zero-filled by the hardware, the
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

// Fused score + non-max suppression, persistent: CTA c walks the items c, c + gridDim.x, ... (item = tile * frames + frame;
// tile = kTileW x TH pixels of one layer). Per item:
//   1. the image tile with its ring halo (3 rows above/below, 16 bytes left/right: TMA wants 16-byte granular boxes AND
//      box origins) is in shared memory, delivered by ONE TMA bulk tensor copy (zero outside the image) that was issued
//      while the previous item was being scored;
//   2. the bytes are expanded once into two 16x2 planes, E[r][k] = (p[2k], p[2k+1]) and O[r][k] = (p[2k+1], p[2k+2]):
//      every ring sample of a horizontal pixel PAIR is then a single conflict-free LDS.32 (E for even dx, O for odd dx).
//      The raw tile is free after that: the TMA copy of the next item starts here;
//   3. a warp owns a tile row per iteration (lane = pixel pair, 8 rows per warp, fully unrolled: every LDS has an immediate
//      offset): 84 packed min/max/add instructions per pair (b0_pair; the ALU pipe is this kernel's limiter), scores go
//      to a shared score tile, the "score >= threshold" bits of the lane's pixels are shifted into a register (no ballots,
//      no atomics inside the loop). Tiles that touch the 3-pixel margin of the layer run a row-checked copy of the loop;
//   4. the strong pixels (a few per cent) are listed (one shared atomic per warp), the score tile is copied to the global
//      score map with 16-byte stores, and every strong pixel gets the 3x3 non-max test from the shared score tile, one per
//      thread. A strong pixel ON the tile border whose in-tile neighbours do not already beat it is emitted with the
//      PENDING flag (bit 30): its out-of-tile neighbours are scores of another tile, so k_refine finishes the test from the
//      global map (complete by then). There is no score halo, i.e. no pixel is scored twice.
//   Candidate word: time key | tie << 31 | pending << 30; the tile's candidates are appended with one atomicAdd.
constexpr int kHaloX = 16;           // measured on B200: the innermost TMA coordinate must be a multiple of 16 bytes
                                     // (bench/tma_probe.cu: x = -8, 8, 376 raise "illegal instruction", -16, 0, 384 work)
constexpr int kImgW = kTileW + 2 * kHaloX;   // bytes per staged image row: x0-16 .. x0+79
constexpr int kExp0 = 3, kExp1 = 21; // staged words (4 bytes) that are expanded: bytes 12 .. 83 cover x0-4 .. x0+67
constexpr int kPlaneW = 2 * (kExp1 - kExp0);   // 16x2 words per plane row; plane word j holds staged bytes 2j+12 (E) / 2j+13 (O)

constexpr int kScPitch = kTileW + 16;   // score tile row: 64 scores + 16 zero bytes (x = -1 of a row is byte 79 of the row above)

template <int TH, bool CHECKED>
__device__ __forceinline__ uint32_t score_rows(const uint32_t* p, uint16_t* sout, uint32_t colmask, uint32_t thr2, int warp, int row_lo, int row_hi)
{
  constexpr int PW = kPlaneW, OO = (TH + 6) * kPlaneW;   // row pitch of a plane, offset of the O plane (words)
  constexpr int kWarps = TH / kRowsPerWarp;
  uint32_t bits = 0;   // bit 8 + i / 24 + i: even / odd pixel of the lane's i-th row (row = warp + kWarps i) is strong
#pragma unroll
  for (int i = 0; i < kRowsPerWarp; i++) {
    const int o = i * kWarps * PW;
    uint32_t s = 0;
    if (!CHECKED || (warp + i * kWarps >= row_lo && warp + i * kWarps < row_hi)) {   // warp-uniform
      uint32_t v[16];
      v[0] = p[o + OO + 3 * PW - 2];  v[1] = p[o + OO + 2 * PW - 2];  v[2] = p[o + 1 * PW - 1];   v[3] = p[o + OO - 1];
      v[4] = p[o];                    v[5] = p[o + OO];               v[6] = p[o + 1 * PW + 1];   v[7] = p[o + OO + 2 * PW + 1];
      v[8] = p[o + OO + 3 * PW + 1];  v[9] = p[o + OO + 4 * PW + 1];  v[10] = p[o + 5 * PW + 1];  v[11] = p[o + OO + 6 * PW];
      v[12] = p[o + 6 * PW];          v[13] = p[o + OO + 6 * PW - 1]; v[14] = p[o + 5 * PW - 1];  v[15] = p[o + OO + 4 * PW - 2];
      s = b0_pair(v, p[o + 3 * PW]);
      if (CHECKED) s &= colmask;
    }
    sout[i * kWarps * (kScPitch / 2)] = (uint16_t)__byte_perm(s, 0, 0x4420);
    // strong pixels (score >= threshold; scores <= 254 so the 16-bit adds cannot carry)
    bits = (bits >> 1) | (__vadd2(s, thr2) & 0x80008000u);
  }
  return bits;
}

template <int TH>
__global__ void __launch_bounds__(TH * 4, 256 / TH) k_score_nms(const __grid_constant__ TmaMaps maps, const __grid_constant__ DeviceLayers dl,
                                                      const TileEntry* __restrict__ tiles, int n_tiles, int n_frames,
                                                      const uint8_t* in0, int in_pitch, size_t in_frame_stride,
                                                      uint8_t* img_block, uint8_t* score_block, uint32_t* cand,
                                                      int32_t* cand_count, int cand_cap, const __grid_constant__ CandRegions cr,
                                                      int threshold, int32_t* status, uint32_t* tie_cells)
{
  constexpr int kThreads = TH * 4, kWarps = TH / kRowsPerWarp, kImgH = TH + 6;
  __shared__ __align__(128) uint8_t tile[kImgH][kImgW];
  __shared__ __align__(16) uint32_t planes[2][kImgH][kPlaneW];   // E and O
  // score tile with a zero frame: row 0 and row TH + 1 stay zero, bytes 64..79 of every row stay zero, so the 8 neighbours
  // of any tile pixel can be read without bounds checks (outside the tile = 0 = never beats or ties a strong pixel)
  // (16 zero bytes in front: the upper-left neighbour of tile pixel (0, 0) is the byte before row 0)
  __shared__ __align__(16) uint8_t sc_raw[16 + (TH + 2) * kScPitch];
  uint8_t (*sc)[kScPitch] = reinterpret_cast<uint8_t (*)[kScPitch]>(sc_raw + 16);
  uint32_t (*pe)[kPlaneW] = planes[0];
  uint32_t (*po)[kPlaneW] = planes[1];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint16_t strong[TH * kTileW];   // (thread << 5) | bit index of a strong pixel
  __shared__ int n_strong2[2], tma_failed;   // the counter alternates between items (reset two barriers before its next use)
  const int n_items = n_tiles * n_frames;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // item = tile * n_frames + frame, advanced by gridDim.x per iteration without divisions
  const int step_t = gridDim.x / n_frames, step_f = gridDim.x - step_t * n_frames;
  int item = blockIdx.x;
  if (item >= n_items) return;
  int tl = item / n_frames, frame = item - tl * n_frames;
  auto issue = [&](const TileEntry& e, int f) {   // one thread
    mbar_expect_tx(&bar, kImgH * kImgW);
    tma_load_3d(&tile[0][0], &maps.m[e.layer], &bar, e.x0 - kHaloX, e.y0 - 3, f);
  };
  if (threadIdx.x == 0) { mbar_init(&bar, 1); n_strong2[0] = n_strong2[1] = 0; tma_failed = 0; }
  for (int i = threadIdx.x; i < (16 + (TH + 2) * kScPitch) / 4; i += kThreads) reinterpret_cast<uint32_t*>(sc_raw)[i] = 0u;
  __syncthreads();
  TileEntry te;
  { const int4 q = __ldg(reinterpret_cast<const int4*>(tiles) + tl); te.layer = q.x; te.x0 = q.y; te.y0 = q.z; te.pad = 0; }
  if (threadIdx.x == 0 && maps.use[te.layer]) issue(te, frame);
  uint32_t phase = 0;
  int par = 0;
  const uint32_t thr2 = (uint32_t)(0x8000 - min(max(threshold, 1), 0x7fff)) * 0x00010001u;
  for (; item < n_items; item += gridDim.x) {
    const int layer = te.layer, x0 = te.x0, y0 = te.y0;
    const DeviceLayer d = dl.l[layer];
    // next item of this CTA
    int ntl = tl + step_t, nframe = frame + step_f;
    if (nframe >= n_frames) { nframe -= n_frames; ntl++; }
    const bool has_next = item + (int)gridDim.x < n_items;
    TileEntry nte = te;
    if (has_next) { const int4 q = __ldg(reinterpret_cast<const int4*>(tiles) + ntl); nte.layer = q.x; nte.x0 = q.y; nte.y0 = q.z; }
    if (maps.use[layer]) {
      // every warp waits for the bulk copy on the mbarrier itself (try_wait suspends in hardware): no CTA barrier
      const long long t_start = clock64();
      while (!mbar_try_wait(&bar, phase)) {
        if (clock64() - t_start > 400000000ll) {   // ~0.2 s: never spin forever on a broken descriptor
          if (lane == 0) { atomicOr(&status[frame], 16); tma_failed = 1; }
          break;
        }
      }
      phase ^= 1u;
    } else {
      const uint8_t* img; int pitch;
      if (layer == 0) { img = in0 + (size_t)frame * in_frame_stride; pitch = in_pitch; }
      else { img = img_block + (size_t)frame * dl.frame_stride + d.offset; pitch = d.pitch; }
      const bool word_ok = ((pitch & 3) == 0) && ((((uintptr_t)img) & 3) == 0);
      __syncthreads();   // (the expansion of the previous item has long finished; this orders the plain stores below after it)
      for (int i = threadIdx.x; i < kImgH * (kImgW / 4); i += kThreads) {
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
    // ---- expansion into the two 16x2 planes: 8 staged bytes per task (kExp1 - kExp0 is even)
    static_assert(((kExp1 - kExp0) & 1) == 0, "expansion works on 8-byte groups");
    constexpr int kGroups = (kExp1 - kExp0) / 2;   // 9 groups of 8 bytes per row
    for (int i = threadIdx.x; i < kImgH * kGroups; i += kThreads) {
      const int r = i / kGroups, g = i - r * kGroups;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&tile[r][4 * (2 * g + kExp0)]);   // 4-byte aligned (kExp0 odd: not 8)
      const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
      uint4 e, o;
      e.x = __byte_perm(w0, 0, 0x4140); e.y = __byte_perm(w0, 0, 0x4342); e.z = __byte_perm(w1, 0, 0x4140); e.w = __byte_perm(w1, 0, 0x4342);
      o.x = __byte_perm(w0, 0, 0x4241); o.y = __byte_perm(w0, w1, 0x7473) & 0x00ff00ffu;   // (w0.b3, w1.b0)
      o.z = __byte_perm(w1, 0, 0x4241); o.w = __byte_perm(w1, w2, 0x7473) & 0x00ff00ffu;
      *reinterpret_cast<uint4*>(&pe[r][4 * g]) = e;
      *reinterpret_cast<uint4*>(&po[r][4 * g]) = o;
    }
    __syncthreads();   // planes complete; every warp is past the non-max phase of the previous item (sc, strong are free)
    if (tma_failed) return;
    // the raw tile is free: the copy of the next item's tile overlaps the scoring of this one
    if (threadIdx.x == 0 && has_next && maps.use[nte.layer]) issue(nte, nframe);
    // ---- dense scores: warp = tile row, lane = pixel pair (x0 + 2*lane, +1). Every ring sample is one LDS.32 at a
    //      compile-time offset from a single per-lane pointer (the two planes are one array).
    const int xg = x0 + 2 * lane;
    const int rows_here = min(TH, d.h - y0);
    const int row_lo = max(3 - y0, 0), row_hi = min(rows_here, d.h - 3 - y0);   // rows [row_lo, row_hi) carry scores
    uint32_t bits;
    {
      const uint32_t* p = &planes[0][warp][2 + lane];          // plane word of this pair (staged byte 16 + 2*lane), tile row `warp`
      uint16_t* sout = reinterpret_cast<uint16_t*>(&sc[warp + 1][2 * lane]);
      if (row_lo == 0 && row_hi == TH && x0 >= 3 && x0 + kTileW <= d.w - 3) bits = score_rows<TH, false>(p, sout, 0xffffffffu, thr2, warp, row_lo, row_hi);
      else {
        // scores are zero in the 3-pixel margin of the layer: per-lane column mask; margin rows are not scored at all
        const uint32_t colmask = ((xg >= 3 && xg < d.w - 3) ? 0x0000ffffu : 0u) | ((xg + 1 >= 3 && xg + 1 < d.w - 3) ? 0xffff0000u : 0u);
        bits = score_rows<TH, true>(p, sout, colmask, thr2, warp, row_lo, row_hi);
      }
    }
    // list the strong pixels (a few per cent of the pixels): warp prefix sum of the counts, one shared atomic per warp
    {
      const int cnt = __popc(bits);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total) {   // warp-uniform
        int base = 0;
        if (lane == 31) base = atomicAdd(&n_strong2[par], total);
        int pos = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
        const uint32_t me = threadIdx.x << 5;
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          strong[pos++] = (uint16_t)(me | (uint32_t)b);
        }
      }
    }
    __syncthreads();   // score tile and strong list complete
    // ---- the score tile goes to the global map with 16-byte stores (pitch and x0 are multiples of 64)
    {
      uint8_t* score = score_block + (size_t)frame * dl.frame_stride + d.offset;
      const int r = threadIdx.x >> 2, c = (threadIdx.x & 3) * 16;
      if (r < rows_here) *reinterpret_cast<uint4*>(score + (size_t)(y0 + r) * d.pitch + x0 + c) = *reinterpret_cast<const uint4*>(&sc[r + 1][c]);
    }
    // ---- 3x3 non-max test of the strong pixels, one per thread; the survivors of a warp are appended to the frame's
    //      candidate list with one global atomic
    const int ns = n_strong2[par];
    if (threadIdx.x == 0) n_strong2[par ^ 1] = 0;   // its readers (previous item) are all past the barrier before this one
    const int room = cr.off[layer + 1] - cr.off[layer];
    for (int i0 = warp * 32; i0 < ns; i0 += kThreads) {   // warp-uniform trip count
      const int i = i0 + lane;
      bool is_c = false, tie = false, pending = false;
      int r = 0, x = 0;
      if (i < ns) {
        const uint32_t e = strong[i];
        const int b = e & 31, t = e >> 5;
        r = (t >> 5) + ((b & 15) - 8) * kWarps; x = 2 * (t & 31) + (b >> 4);
        const uint8_t* q = &sc[r + 1][x];
        const uint32_t c = q[0];
        const uint32_t n0 = __vimax3_u32(q[-kScPitch - 1], q[-kScPitch], q[-kScPitch + 1]);
        const uint32_t n1 = __vimax3_u32(q[-1], q[1], q[kScPitch - 1]);
        const uint32_t nmax = __vimax3_u32(n0, n1, max((uint32_t)q[kScPitch], (uint32_t)q[kScPitch + 1]));
        is_c = nmax <= c; tie = nmax == c;
        pending = r == 0 || r == TH - 1 || x == 0 || x == kTileW - 1;   // some neighbours are scores of another tile
      }
      const unsigned m = __ballot_sync(0xffffffffu, is_c);
      if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&cand_count[frame * kMaxLayers + layer], __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is_c) {
          const int pos = base + __popc(m & ((1u << lane) - 1u));
          if (pos < room) cand[(size_t)frame * cand_cap + cr.off[layer] + pos] = time_key(layer, x0 + x, y0 + r) | (pending ? kCandPending : (tie ? kCandTie : 0u));
          if (pending || tie)
            cells_flag(tie_cells + (size_t)frame * kCellWordsPerFrame + layer * kCellWordsPerLayer, d.w, x0 + x - 2, x0 + x + 2,
                       y0 + r - 2, y0 + r + 2);
        }
      }
    }
    te = nte; tl = ntl; frame = nframe; par ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_encode_tried = 0;

static bool encode_map(CUtensorMap* m, const void* base, int w, int h, size_t pitch, size_t frame_stride, int frames, int tile_h)
{
  if (!g_encode) return false;
  if ((((uintptr_t)base) & 15) || (pitch & 15) || (frame_stride & 15) || pitch == 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
  const cuuint32_t box[3] = {(cuuint32_t)kImgW, (cuuint32_t)(tile_h + 6), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tensor maps of all layers for this call (layer 0 lives in the caller's buffer, so its map is re-encoded per call;
// the others are cached in the workspace). Layers whose geometry TMA cannot address fall back to plain loads.
static void build_tma_maps(CamWorkspace& ws, const uint8_t* d_images, int src_pitch, size_t in_stride, int frames, int tile_h, TmaMaps& maps)
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
      ws.tma_use[i] = encode_map(&ws.tma[i], ws.d_img + g.offset, g.w, g.h, (size_t)g.pitch, (size_t)ws.dl.frame_stride, ws.cfg.max_batch, tile_h) ? 1 : 0;
    }
    ws.tma_ready = 1;
  }
  for (int i = 1; i < ws.n_layers; i++) { maps.m[i] = ws.tma[i]; maps.use[i] = ws.tma_use[i]; }
  maps.use[0] = encode_map(&maps.m[0], d_images, ws.geom[0].w, ws.geom[0].h, (size_t)src_pitch, in_stride, frames, tile_h) ? 1 : 0;
}

static int g_ctas_per_sm[2] = {0, 0}, g_sm_count = 0;

int pyramid_score_run(okb_context* ctx, CamWorkspace& ws, const uint8_t* d_images, int src_pitch, size_t in_stride, int n_frames,
                      const CandRegions& cr)
{
  const okb_camera_config_t& c = ws.cfg;
  const int B = n_frames;
  cudaStream_t st = ws.stream;
  // ---- pyramid: the chains that can be carried in shared memory go into ONE launch, the rest layer by layer. It runs on the side
  //      stream underneath the scoring of layer 0 (more than half of all pixels), which needs nothing from it: the pyramid kernel is
  //      latency-bound on loads, the score kernel on the ALU pipe, so they share the SMs well. The scoring of layers >= 1 waits for it.
  static const int overlap_env = getenv("OKB_PYR_OVERLAP") ? atoi(getenv("OKB_PYR_OVERLAP")) : 1;   // tuning hook
  const bool overlap = ws.n_layers > 1 && overlap_env != 0;
  cudaStream_t sp = overlap ? ws.stream2 : st;
  if (overlap) { OKB_CUDA(cudaEventRecord(ws.ev_fork, st)); OKB_CUDA(cudaStreamWaitEvent(sp, ws.ev_fork, 0)); }
  bool chained[kMaxLayers] = {false};
  PyrArgs pa; memset(&pa, 0, sizeof(pa));
  if (ws.n_layers > 1) {
    auto add_chain = [&](int first, bool general) {
      PyrChain ch; memset(&ch, 0, sizeof(ch));
      ch.general = general ? 1 : 0;
      for (int l = first; l < ws.n_layers && ch.n < 4; l += 2) {
        if (l != first && !ws.geom[l].fast2) break;
        ch.layer[ch.n++] = l; chained[l] = true;
      }
      const LayerGeom& g = ws.geom[first];
      ch.tiles_x = (g.w + kPyrW - 1) / kPyrW; ch.tiles = ch.tiles_x * ((g.h + kPyrH - 1) / kPyrH);
      ch.taps = g.max_taps;
      ch.xs = g.d_xs; ch.xn = g.d_xn; ch.ys = g.d_ys; ch.yn = g.d_yn; ch.xa = g.d_xa; ch.ya = g.d_ya;
      pa.c[pa.n_chains++] = ch;
    };
    if (ws.n_layers > 2 && ws.geom[2].fast2) add_chain(2, false);
    if (!ws.geom[1].fast2) add_chain(1, true);
    if (pa.n_chains == 1) pa.c[1].tiles = 0;
    if (pa.n_chains > 0) {
      k_pyramid<<<dim3(pa.c[0].tiles + pa.c[1].tiles, B), kPyrThreads, 0, sp>>>(pa, ws.dl, d_images, src_pitch, in_stride, ws.d_img);
      ctx->launches++; if (ctx->timers_on) ws.ps_launches++;
    }
    for (int i = 1; i < ws.n_layers; i++) {
      if (chained[i]) continue;
      const LayerGeom& g = ws.geom[i]; const LayerGeom& p = ws.geom[g.parent];
      ResizeJob J;
      if (g.parent == 0) { J.src = d_images; J.src_pitch = src_pitch; J.src_frame_stride = in_stride; }
      else { J.src = ws.d_img + p.offset; J.src_pitch = p.pitch; J.src_frame_stride = ws.dl.frame_stride; }
      J.dst = ws.d_img + g.offset; J.dst_pitch = g.pitch; J.dst_frame_stride = ws.dl.frame_stride;
      J.dw = g.w; J.dh = g.h; J.fast2 = g.fast2;
      J.xs = g.d_xs; J.xn = g.d_xn; J.ys = g.d_ys; J.yn = g.d_yn; J.xa = g.d_xa; J.ya = g.d_ya;
      const int tiles_x = (g.w + 31) / 32;
      k_resize<<<dim3(tiles_x * ((g.h + 31) / 32), B), 256, 0, sp>>>(J, tiles_x);
      ctx->launches++; if (ctx->timers_on) ws.ps_launches++;
    }
  }
  if (overlap) OKB_CUDA(cudaEventRecord(ws.ev_pyr, sp));
  if (ctx->timers_on) cudaEventRecord(ws.ev_mid, st);
  // ---- scores
  const int TH = ws.score_tile_h;
  if (!ws.d_tiles) {
    std::vector<TileEntry> tab;
    for (int i = 0; i < ws.n_layers; i++)
      for (int y0 = 0; y0 < ws.geom[i].h; y0 += TH)
        for (int x0 = 0; x0 < ws.geom[i].w; x0 += kTileW) tab.push_back(TileEntry{i, x0, y0, 0});
    ws.n_tiles = (int)tab.size();
    ws.n_tiles0 = 0;
    for (const TileEntry& e : tab) ws.n_tiles0 += e.layer == 0;
    OKB_CUDA(cudaMalloc(&ws.d_tiles, tab.size() * sizeof(TileEntry)));
    OKB_CUDA(cudaMemcpy(ws.d_tiles, tab.data(), tab.size() * sizeof(TileEntry), cudaMemcpyHostToDevice));
  }
  TmaMaps maps;
  build_tma_maps(ws, d_images, src_pitch, in_stride, c.max_batch, TH, maps);
  if (!g_sm_count) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_ctas_per_sm[0], k_score_nms<64>, 256, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_ctas_per_sm[1], k_score_nms<32>, 128, 0);
    for (int i = 0; i < 2; i++) if (g_ctas_per_sm[i] < 1) g_ctas_per_sm[i] = 1;
  }
  const TileEntry* tiles = (const TileEntry*)ws.d_tiles;
  auto launch = [&](const TileEntry* t0, int n_t) {
    const int items = n_t * B;
    if (TH == 64) {
      const int grid = std::min(items, g_sm_count * g_ctas_per_sm[0]);
      k_score_nms<64><<<grid, 256, 0, st>>>(maps, ws.dl, t0, n_t, B, d_images, src_pitch, in_stride, ws.d_img, ws.d_score, ws.d_cand,
                                            ws.d_cand_count, ws.cand_cap, cr, c.threshold, ws.d_status, ws.d_tie_cells);
    } else {
      const int grid = std::min(items, g_sm_count * g_ctas_per_sm[1]);
      k_score_nms<32><<<grid, 128, 0, st>>>(maps, ws.dl, t0, n_t, B, d_images, src_pitch, in_stride, ws.d_img, ws.d_score, ws.d_cand,
                                            ws.d_cand_count, ws.cand_cap, cr, c.threshold, ws.d_status, ws.d_tie_cells);
    }
    ctx->launches++; if (ctx->timers_on) ws.ps_launches++;
  };
  if (overlap) {
    launch(tiles, ws.n_tiles0);                                   // layer 0 (the table is layer-major), next to the pyramid
    OKB_CUDA(cudaStreamWaitEvent(st, ws.ev_pyr, 0));
    launch(tiles + ws.n_tiles0, ws.n_tiles - ws.n_tiles0);        // the reduced layers
  } else {
    launch(tiles, ws.n_tiles);
  }
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

}  // namespace okb
