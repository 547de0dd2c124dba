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
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "okb_detect.h"

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

constexpr int kMaxTies = 4096;   // per frame (k_resolve keeps them in shared memory)

__global__ void __launch_bounds__(128) k_refine(const __grid_constant__ DeviceLayers dl, const uint8_t* in0, int in_pitch, size_t in_frame_stride,
                                                uint8_t* img_block, uint8_t* score_block, uint32_t* touch_block,
                                                const uint32_t* cand, const int32_t* cand_count, int cand_cap,
                                                const __grid_constant__ CandRegions cr, uint32_t* fkey, float4* fval, int threshold, const uint32_t* d_epoch,
                                                const uint32_t* tie_cells, TieEntry* tie_list, int32_t* tie_count, int cpw)
{
  // cpw candidates per warp (lanes cpw.. only help with the touch emission): 32 when the batch fills the GPU; fewer for a single
  // frame, whose ~150 warps of 32 divergent candidates each would leave three quarters of the schedulers idle
  const int frame = blockIdx.y;
  int prefix[kMaxLayers + 1];
  const int n = cand_total(cr, cand_count + frame * kMaxLayers, dl.n, prefix);
  if ((int)blockIdx.x * 4 * cpw >= n) return;
  const uint32_t epoch = *d_epoch;
  __shared__ FrameViews v;  // dynamically indexed by layer: keep it out of local memory
  if (threadIdx.x == 0) make_views(dl, in0, in_pitch, in_frame_stride, img_block, score_block, touch_block, frame, v);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int i = ((int)blockIdx.x * 4 + (int)(threadIdx.x >> 5)) * cpw + lane;
  const bool have = lane < cpw && i < n;
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
    // what the selection kernel reads, field by field: key | keep << 30 | decided-maximum << 31 (ties: set by k_resolve), values
    fkey[(size_t)frame * cand_cap + i] = key | ((uint32_t)(r.keep != 0) << 30) | ((uint32_t)(tie == 0) << 31);
    fval[(size_t)frame * cand_cap + i] = make_float4(r.x, r.y, r.size, r.response);
    if (tie) {
      // the tie resolution works on a compact list of its own (order irrelevant: it sorts by time key)
      const int p = atomicAdd(&tie_count[frame], 1);
      if (p < kMaxTies) {
        TieEntry e;
        e.key = key; e.rec = (uint32_t)i;
        e.info.own_touch = r.own_touch; e.info.has_above = r.has_above; e.info.exited = (int8_t)r.above.exited;
        e.info.n_queries = (int8_t)r.above.n_queries; e.info.max_x = r.above.max_x; e.info.max_y = r.above.max_y;
        tie_list[(size_t)frame * kMaxTies + p] = e;
      }
    }
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
constexpr int kMaxBlockers = 12;
constexpr int kResolveThreads = 1024;
constexpr int kResolveSmem = 232448 - 2048;   // all of an SM's shared memory: per-tie state (50 B) + score windows (28 B) while they fit
constexpr int kResolvePerTie = 4 + 2 * kMaxBlockers + 8 + 4 + 4 + 4 + 2;
static_assert(kMaxTies * kResolvePerTie <= kResolveSmem, "k_resolve shared memory");

// 5x5 dense scores around a tie, one byte each, packed into 7 words
struct TieWindow { uint32_t w[7]; };
// the reference's isMax2D on the map the sequential algorithm would see: a pixel below the threshold reads 0 unless an
// earlier maximum's cache touch (bit j of `mask`) had filled it in
__device__ __forceinline__ bool tie_is_max(const TieWindow& t, uint32_t mask, int threshold)
{
  int m[5][5];
#pragma unroll
  for (int j = 0; j < 25; j++) {
    int val = (int)((t.w[j >> 2] >> (8 * (j & 3))) & 255u);
    if (val < threshold && !((mask >> j) & 1u)) val = 0;
    m[j / 5][j % 5] = val;
  }
  return is_max_2d_5x5(m);
}

// 5x5 pixel masks: bit 5 r + c is pixel (origin.x + c, origin.y + r). A mask with origin o re-based to origin n
// (d = o - n per axis); bits that leave the 5x5 frame are dropped.
__device__ __forceinline__ uint32_t rebase5x5(uint32_t m, int dx, int dy)
{
  if (dx <= -5 || dx >= 5 || dy <= -5 || dy >= 5) return 0u;
  uint32_t out = 0;
#pragma unroll
  for (int r = 0; r < 5; r++) {
    const int sr = r - dy;
    uint32_t row = (sr >= 0 && sr < 5) ? ((m >> (5 * sr)) & 31u) : 0u;
    row = dx >= 0 ? (row << dx) : (row >> (-dx));
    out |= (row & 31u) << (5 * r);
  }
  return out;
}
constexpr uint32_t kPatch3x3 = 0x1CE7u, kPatch4x4 = 0x7BDEFu;   // rows of 3 (4) set bits, origin (x - 1, y - 1)

// The ties of one frame in time order, one array per field (what k_tie_gather writes and k_resolve reads, coalesced)
struct TieSorted {
  uint32_t* key;     // time key (layer, y, x)
  uint32_t* rec;     // candidate record index
  uint32_t* tmask;   // bits 0..24: 5x5 window pixels touched by non-tied maxima before this tie; bits 25..26: the outcome (1 maximum,
                     // 2 rejected) if no earlier tie touches a sensitive pixel
  uint32_t* sens;    // window pixels whose visibility matters: 0 < score < threshold and not in tmask
  uint32_t* fp;      // cache touches of this tie if it is a maximum: bits 0..24 = pixels of layer + 1 relative to the box of its
                     // above-layer scan, bits 25..26 = own-layer patch (0 none, 1 = 3x3, 2 = 4x4 around x-1, y-1)
  uint32_t* win;     // [7][kMaxTies] packed 5x5 score bytes
};
constexpr size_t kTieSortedWords = (size_t)kMaxTies * (5 + 7);
__device__ __host__ __forceinline__ TieSorted tie_sorted(uint32_t* base, int frame)
{
  uint32_t* p = base + (size_t)frame * kTieSortedWords;
  return TieSorted{p, p + kMaxTies, p + 2 * kMaxTies, p + 3 * kMaxTies, p + 4 * kMaxTies, p + 5 * kMaxTies};
}

// Per tie, spread over the GPU (grid: tiles of 64 ties x frames): its rank in time order (count of smaller keys: the list
// k_refine appended is unordered), everything the dependency rounds need from global memory -- the touch-time words and the
// dense scores of its 5x5 window (50 scattered loads, the reason this is not done by the one CTA that resolves the frame) --
// and the footprint of its own touches as bit masks. Results are written to position `rank` of the per-field arrays.
__global__ void __launch_bounds__(64) k_tie_gather(const __grid_constant__ DeviceLayers dl, const uint8_t* score_block, const uint32_t* touch_block,
                                                    const TieEntry* tie_list, const int32_t* tie_count, const uint32_t* d_epoch, int threshold,
                                                    uint32_t* sorted_base)
{
  __shared__ uint32_t keys[kMaxTies];
  const int frame = blockIdx.y;
  const int T = min(tie_count[frame], kMaxTies);
  if ((int)(blockIdx.x * blockDim.x) >= T) return;
  const TieEntry* E = tie_list + (size_t)frame * kMaxTies;
  for (int j = threadIdx.x; j < T; j += blockDim.x) keys[j] = E[j].key;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const TieEntry e = E[i];
  const uint32_t key = e.key, epoch = *d_epoch;
  const int layer = (int)(key >> 22), y = (int)((key >> 11) & 2047), x = (int)(key & 2047);
  const DeviceLayer d = dl.l[layer];
  const uint32_t* tm = touch_block + (size_t)frame * dl.frame_stride + d.offset;
  const uint8_t* sc = score_block + (size_t)frame * dl.frame_stride + d.offset;
  uint32_t te[25], sv[25];
#pragma unroll
  for (int j = 0; j < 25; j++) {
    const size_t o = (size_t)(y + j / 5 - 2) * d.pitch + (x + j % 5 - 2);
    te[j] = __ldcg(&tm[o]); sv[j] = __ldcg(&sc[o]);
  }
  int rank = 0;
  for (int j = 0; j < T; j++) rank += keys[j] < key;   // keys are distinct pixels
  uint32_t mask = 0, sens = 0;
  TieWindow w;
#pragma unroll
  for (int j = 0; j < 7; j++) w.w[j] = 0;
#pragma unroll
  for (int j = 0; j < 25; j++) {
    if (touched_before(te[j], epoch, key)) mask |= 1u << j;
    else if (sv[j] > 0u && (int)sv[j] < threshold) sens |= 1u << j;
    w.w[j >> 2] |= sv[j] << (8 * (j & 3));
  }
  // the outcome as long as no earlier tie makes a sensitive pixel visible: the common case, decided here by all SMs
  const uint32_t r0 = tie_is_max(w, mask, threshold) ? 1u : 2u;
  uint32_t fp = (uint32_t)e.info.own_touch << 25;
  if (e.info.has_above) {
    ScanIter it; above_window(layer, x, y, it);
    const int ox = (int)it.x_1 - 1, oy = (int)it.y_1 - 1;
    auto mark = [&](int px, int py) {
      const int c = px - ox, r = py - oy;
      if ((unsigned)c < 5u && (unsigned)r < 5u) fp |= 1u << (5 * r + c);
    };
    for (int q = 0; q < e.info.n_queries; q++) {
      int X, Y; bool blk; above_query_pos(it, q, X, Y, blk);
      mark(X, Y);
      if (blk) { mark(X + 1, Y); mark(X, Y + 1); mark(X + 1, Y + 1); }
    }
    if (!e.info.exited) for (int j = 0; j < 9; j++) mark(e.info.max_x + j % 3 - 1, e.info.max_y + j / 3 - 1);
  }
  const TieSorted S = tie_sorted(sorted_base, frame);
  S.key[rank] = key; S.rec[rank] = e.rec; S.tmask[rank] = mask | (r0 << 25); S.sens[rank] = sens; S.fp[rank] = fp;
#pragma unroll
  for (int j = 0; j < 7; j++) S.win[(size_t)j * kMaxTies + rank] = w.w[j];
}

// shared-memory state of k_resolve
struct ResolveCtx {
  uint32_t* ties;                       // time keys, ascending
  uint16_t (*blockers)[kMaxBlockers];
  short4* box;                          // above-scan window (layer+1 coords)
  uint32_t* tmask; uint32_t* fpm; uint32_t* sens;
  int8_t* state;                        // 0 unresolved, 1 maximum, 2 rejected
  int8_t* n_block;                      // -1: list overflowed, enumerate instead
  int* layer_start;
  const uint32_t* win; int win_stride;  // packed 5x5 score windows, word j of tie i at win[j * win_stride + i] (shared memory while they fit)
};

// window pixels of the tie at (layer, x, y) that winner u would have touched
__device__ __noinline__ uint32_t tie_footprint(const ResolveCtx c, int u, int layer, int x, int y)
{
  const uint32_t uk = c.ties[u], f = c.fpm[u];
  const int ul = (int)(uk >> 22), uy = (int)((uk >> 11) & 2047), ux = (int)(uk & 2047);
  if (ul == layer) {
    const uint32_t own = f >> 25;
    return own ? rebase5x5(own == 2 ? kPatch4x4 : kPatch3x3, (ux - 1) - (x - 2), (uy - 1) - (y - 2)) : 0u;
  }
  const short4 bx = c.box[u];
  return rebase5x5(f & 0x1ffffffu, (int)bx.x - (x - 2), (int)bx.z - (y - 2));
}

// Enumerates the possible blockers of tie ti: earlier ties whose touches can reach its window (same layer within 4 px, or the
// layer below through the box of its above-scan). mode 0: append them to blockers[ti] (returns the count, -1 on overflow);
// mode 1: returns 1 + the first undecided one (0 if none); mode 2: ORs the footprints of the winners among them into `add`.
__device__ __noinline__ int tie_enumerate(const ResolveCtx c, int ti, int mode, uint32_t& add)
{
  const uint32_t key = c.ties[ti];
  const int layer = (int)(key >> 22), y = (int)((key >> 11) & 2047), x = (int)(key & 2047);
  int nb = 0;
  // part 0: own layer, rows y-4 .. y before ti; part 1: the layer below, rows whose above-scan box (rows 2/3 uy - 2.5 .. 2/3 uy + 2.2
  // for a 3:2 step, 3/4 uy - 2.5 .. 3/4 uy + 2.3 for a 4:3 step) can reach [y-2, y+2]: +-8 rows around y times the ratio (the box
  // test is the exact filter)
#pragma unroll 1
  for (int part = 0; part < 2; part++) {
    if (part == 1 && layer == 0) break;
    const int uc = (layer & 1) ? (3 * y) / 2 : (4 * y) / 3;
    const int row_lo = part == 0 ? max(y - 4, 0) : max(uc - 8, 0), row_hi = part == 0 ? 2047 : uc + 8;
    const uint32_t lo_key = time_key(layer - part, 0, min(row_lo, 2047));
    int lo = c.layer_start[layer - part], hi = part == 0 ? ti : c.layer_start[layer];
    const int end = hi;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (c.ties[mid] < lo_key) lo = mid + 1; else hi = mid; }
#pragma unroll 1
    for (int u = lo; u < end; u++) {
      const uint32_t uk = c.ties[u];
      bool hit;
      if (part == 0) hit = abs((int)(uk & 2047) - x) <= 4;   // |dy| <= 4 by the key range
      else {
        if ((int)((uk >> 11) & 2047) > row_hi) break;
        const short4 bx = c.box[u];
        hit = !(x + 2 < bx.x || x - 2 > bx.y || y + 2 < bx.z || y - 2 > bx.w);
      }
      if (!hit) continue;
      if (mode == 0) { if (nb >= kMaxBlockers) return -1; c.blockers[ti][nb++] = (uint16_t)u; }
      else if (mode == 1) { if (c.state[u] == 0) return 1 + u; }
      else if (c.state[u] == 1) add |= tie_footprint(c, u, layer, x, y);
    }
  }
  return mode == 0 ? nb : 0;
}

// One CTA per frame. Resolves, in dependency rounds, the candidates whose 2-D maximum test ties with a neighbour:
// their outcome depends on which sub-threshold scores the sequential algorithm had already cached when it reached them.
// Touches by NON-tied maxima are already in the touch-time map (k_refine -> tmask, gathered by k_tie_gather); touches by
// EARLIER TIES are applied here from shared memory: a tie waits only for the earlier ties whose touches can reach its 5x5
// window, listed once; when they are all decided it ORs the footprint masks of the winners among them (re-based to its
// window) into its "touched" mask. Only if that makes a SENSITIVE pixel visible (0 < score < threshold, TieSorted::sens) does
// the reference's isMax2D have to be evaluated again; otherwise the outcome k_tie_gather computed stands. The kernel sees only
// the ties, in time order, reads global memory coalesced and once; the rounds run out of shared memory. It is latency-bound on
// one SM, instruction fetch included: the code is kept small (helpers not inlined, loops rolled).
__global__ void __launch_bounds__(kResolveThreads) k_resolve(uint32_t* sorted_base, const int32_t* tie_count, uint32_t* fkey, int cand_cap,
                                                             int threshold, int32_t* status, long long* dbg)
{
#define OKB_STAMP(i) if (dbg && threadIdx.x == 0) dbg[blockIdx.x * 16 + (i)] = clock64()
  OKB_STAMP(0);
  if (dbg && threadIdx.x == 0) dbg[blockIdx.x * 16 + 15] = 0;
  extern __shared__ unsigned long long resolve_smem[];
  __shared__ int layer_start[kMaxLayers + 1];
  const int frame = blockIdx.x;
  int T = tie_count[frame];
  if (T > kMaxTies) { if (threadIdx.x == 0) atomicOr(&status[frame], 2); T = kMaxTies; }
  if (T == 0) return;
  const int Tp = (T + 31) & ~31;   // array stride: the arrays are laid out for this frame's tie count
  ResolveCtx c;
  c.ties = reinterpret_cast<uint32_t*>(resolve_smem);
  c.box = reinterpret_cast<short4*>(c.ties + Tp);
  c.tmask = reinterpret_cast<uint32_t*>(c.box + Tp);
  c.fpm = c.tmask + Tp;
  c.sens = c.fpm + Tp;
  c.blockers = reinterpret_cast<uint16_t (*)[kMaxBlockers]>(c.sens + Tp);
  c.state = reinterpret_cast<int8_t*>(c.blockers + Tp);
  c.n_block = c.state + Tp;
  c.layer_start = layer_start;
  uint32_t* swin = reinterpret_cast<uint32_t*>(c.n_block + Tp);
  const bool win_in_smem = (size_t)Tp * (kResolvePerTie + 28) <= (size_t)kResolveSmem;
  const TieSorted S = tie_sorted(sorted_base, frame);
  c.win = win_in_smem ? swin : S.win; c.win_stride = win_in_smem ? Tp : kMaxTies;
  uint32_t* FK = fkey + (size_t)frame * cand_cap;
  if (win_in_smem) {
#pragma unroll 1
    for (int i = threadIdx.x; i < 7 * Tp; i += blockDim.x) { const int j = i / Tp, t = i - j * Tp; swin[i] = t < T ? S.win[(size_t)j * kMaxTies + t] : 0u; }
  }
#pragma unroll 1
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const uint32_t key = S.key[i];
    const int layer = (int)(key >> 22), y = (int)((key >> 11) & 2047), x = (int)(key & 2047);
    c.ties[i] = key; c.tmask[i] = S.tmask[i]; c.fpm[i] = S.fp[i]; c.sens[i] = S.sens[i]; c.state[i] = 0;
    ScanIter it; above_window(layer, x, y, it);
    c.box[i] = make_short4((short)((int)it.x_1 - 1), (short)((int)it.x1 + 2), (short)((int)it.y_1 - 1), (short)((int)it.y1 + 2));
  }
  __syncthreads();
  if (threadIdx.x <= kMaxLayers) {
    // first tie index of every layer (layer is the top field of the key)
    const int l = threadIdx.x;
    int lo = 0, hi = T;
    const uint32_t target = time_key(l, 0, 0);
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (c.ties[mid] < target) lo = mid + 1; else hi = mid; }
    layer_start[l] = (l >= kMaxLayers) ? T : lo;
  }
  __syncthreads();
  OKB_STAMP(1);
  uint32_t unused = 0;
#pragma unroll 1
  for (int ti = threadIdx.x; ti < T; ti += blockDim.x) {
    const int nb = tie_enumerate(c, ti, 0, unused);
    c.n_block[ti] = (int8_t)nb;
    if (nb == 0) c.state[ti] = (int8_t)(c.tmask[ti] >> 25);   // nothing earlier can reach its window: k_tie_gather's outcome stands
    if (nb < 0 && dbg) atomicAdd((unsigned long long*)&dbg[blockIdx.x * 16 + 15], 1ull);
  }
  __syncthreads();
  OKB_STAMP(2);
  int rounds = 0;
  const int K = (T + kResolveThreads - 1) / kResolveThreads;   // ties per thread (<= 4)
  uint32_t done = 0;   // bit k: tie threadIdx.x + k * kResolveThreads is decided
  int wait_on[kMaxTies / kResolveThreads];   // the undecided blocker tie k was last seen waiting for (-1: none)
#pragma unroll 1
  for (int k = 0; k < K; k++) { const int ti = threadIdx.x + k * kResolveThreads; wait_on[k] = -1; if (ti >= T || c.state[ti] != 0) done |= 1u << k; }
  while (true) {
    uint32_t decided = 0;   // 2 bits per k: 0 = still blocked, 1 = maximum, 2 = rejected
    bool pending = false;
#pragma unroll 1
    for (int k = 0; k < K; k++) {
      if ((done >> k) & 1u) continue;
      if (wait_on[k] >= 0 && c.state[wait_on[k]] == 0) { pending = true; continue; }   // still waiting for the same tie: O(1)
      const int ti = threadIdx.x + k * kResolveThreads;
      const uint32_t key = c.ties[ti];
      const int layer = (int)(key >> 22), y = (int)((key >> 11) & 2047), x = (int)(key & 2047);
      int blocker = -1;
      uint32_t add = 0;   // window pixels the winners among the blockers have touched
      const int nb = c.n_block[ti];
      if (nb >= 0) {
#pragma unroll 1
        for (int i = 0; i < nb; i++) { const int u = c.blockers[ti][i]; if (c.state[u] == 0) { blocker = u; break; } }
        if (blocker < 0) {
#pragma unroll 1
          for (int i = 0; i < nb; i++) { const int u = c.blockers[ti][i]; if (c.state[u] == 1) add |= tie_footprint(c, u, layer, x, y); }
        }
      } else {
        blocker = tie_enumerate(c, ti, 1, add) - 1;
        if (blocker < 0) tie_enumerate(c, ti, 2, add);
      }
      if (blocker >= 0) { wait_on[k] = blocker; pending = true; continue; }
      const uint32_t tm = c.tmask[ti];
      uint32_t r = tm >> 25;
      // an earlier tie made a sub-threshold score of the window visible: evaluate the test on the new map (rare)
      if (add & c.sens[ti]) {
        TieWindow t;
#pragma unroll
        for (int j = 0; j < 7; j++) t.w[j] = c.win[(size_t)j * c.win_stride + ti];
        r = tie_is_max(t, (tm | add) & 0x1ffffffu, threshold) ? 1u : 2u;
      }
      decided |= r << (2 * k);
    }
    __syncthreads();   // every thread has read the states of this round
#pragma unroll 1
    for (int k = 0; k < K; k++) {
      const uint32_t r = (decided >> (2 * k)) & 3u;
      if (r) { c.state[threadIdx.x + k * kResolveThreads] = (int8_t)r; done |= 1u << k; }
    }
    rounds++;
    const int more = __syncthreads_or(pending);
    if (rounds == 1) { OKB_STAMP(3); } else if (rounds == 2) { OKB_STAMP(7); } else if (rounds == 3) { OKB_STAMP(9); }
    if (!more) break;
  }
  OKB_STAMP(4);
#pragma unroll 1
  for (int ti = threadIdx.x; ti < T; ti += blockDim.x) if (c.state[ti] == 1) FK[S.rec[ti]] |= 0x80000000u;   // one owner per record
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
__device__ __noinline__ int block_exclusive_scan_1024(int val, int* sh /*33 ints*/, int& total)
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
// Ordering is a counting sort by (layer, row) followed by a tiny in-place insertion sort inside every row group. The kernel is
// a chain of short block-wide phases: every global read is coalesced (field arrays written by k_refine) or an independent
// gather issued in batches, and the candidates are read once (the slot a candidate takes inside its row comes back from the
// counting atomic and waits in a register for the row offsets).
__global__ void __launch_bounds__(1024) k_finalize(const __grid_constant__ DeviceLayers dl, const int32_t* cand_count, int cand_cap,
                                                   const __grid_constant__ CandRegions cr, const uint32_t* fkey, const float4* fval, uint32_t* fslot,
                                                   const float* scale_bounds, const uint32_t* size_list, int W, int H,
                                                   int max_kp, int kp_cap, okb_keypoint_t* kp_out, int32_t* kscale_out,
                                                   int32_t* count_out, int32_t* status, long long* dbg)
{
#define OKB_STAMP(i) if (dbg && threadIdx.x == 0) dbg[blockIdx.x * 16 + (i)] = clock64()
  OKB_STAMP(8);
  extern __shared__ unsigned long long keys[];  // kSortCap (key << 32 | candidate index), ordered
  uint32_t* resp = reinterpret_cast<uint32_t*>(keys + kSortCap);          // kSortCap response bits
  uint16_t* rowoff = reinterpret_cast<uint16_t*>(resp + kSortCap);        // kMaxRows + 1 group offsets
  uint8_t* flag = reinterpret_cast<uint8_t*>(rowoff + kMaxRows + 2);      // kSortCap keep flags
  __shared__ int n_valid, hist[256], sh_scan[33], rowbase[kMaxLayers + 1];
  __shared__ uint32_t sel_prefix;
  __shared__ int sel_remaining;
  __shared__ float s_bounds[kScales];
  __shared__ uint32_t s_sizes[kScales];
  const int frame = blockIdx.x;
  int prefix_[kMaxLayers + 1];
  const int n = cand_total(cr, cand_count + frame * kMaxLayers, dl.n, prefix_);
  if (threadIdx.x < dl.n && cand_count[frame * kMaxLayers + threadIdx.x] > cr.off[threadIdx.x + 1] - cr.off[threadIdx.x])
    atomicOr(&status[frame], 1);   // a layer's region of the candidate list overflowed
  const uint32_t* FK = fkey + (size_t)frame * cand_cap;
  const float4* FV = fval + (size_t)frame * cand_cap;
  uint32_t* FS = fslot + (size_t)frame * cand_cap;
  if (threadIdx.x == 0) {
    n_valid = 0;
    int acc = 0;
    for (int l = 0; l < kMaxLayers; l++) { rowbase[l] = acc; if (l < dl.n) acc += dl.l[l].h; }
    rowbase[kMaxLayers] = acc;
  }
  if (threadIdx.x < kScales) { s_bounds[threadIdx.x] = threadIdx.x < kScales - 1 ? scale_bounds[threadIdx.x] : 0.f; s_sizes[threadIdx.x] = size_list[threadIdx.x]; }
  __syncthreads();
  const int n_rows = rowbase[kMaxLayers];
  if (n_rows > kMaxRows) { if (threadIdx.x == 0) { atomicOr(&status[frame], 4); count_out[frame] = 0; } return; }
  uint32_t* rowcnt = reinterpret_cast<uint32_t*>(keys);   // counts live in the (still unused) key array during the counting pass
  for (int i = threadIdx.x; i <= n_rows; i += blockDim.x) rowcnt[i] = 0;
  __syncthreads();
  auto row_of = [&](uint32_t key) { return rowbase[(key >> 22) & 7u] + (int)((key >> 11) & 2047); };
  // pass 1: count the valid keypoints per (layer, row); the atomic returns the candidate's slot inside its row group, which
  // waits in the slot array for the row offsets (four independent loads in flight per thread; rolled loops: this CTA pays for
  // every instruction it fetches)
  int my_valid = 0;
#pragma unroll 1
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * 1024) {
    uint32_t k[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int i = i0 + u * 1024; k[u] = i < n ? __ldcg(&FK[i]) : 0u; }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * 1024;
      if (i >= n) continue;
      uint32_t w = 0;
      if ((k[u] >> 30) == 3u) { w = 0x80000000u | atomicAdd(&rowcnt[row_of(k[u])], 1u); my_valid++; }
      FS[i] = w;
    }
  }
  if (my_valid) atomicAdd(&n_valid, my_valid);
  __syncthreads();
  const int V = n_valid;
  if (V > kSortCap) { if (threadIdx.x == 0) { atomicOr(&status[frame], 4); count_out[frame] = 0; } return; }
  // exclusive scan of the row counts -> group offsets (16 bit: V <= 16384)
  {
    const int chunk = (n_rows + (int)blockDim.x - 1) / (int)blockDim.x;
    const int beg = min((int)threadIdx.x * chunk, n_rows), end = min(beg + chunk, n_rows);
    int cnts[8];   // chunk <= 8 for kMaxRows = 8192
    int sum = 0;
    for (int i = beg, j = 0; i < end; i++, j++) { cnts[j] = (int)rowcnt[i]; sum += cnts[j]; }
    int total = 0;
    int run = block_exclusive_scan_1024(sum, sh_scan, total);   // (its barriers separate the reads above from the writes below)
    for (int i = beg, j = 0; i < end; i++, j++) { rowoff[i] = (uint16_t)run; run += cnts[j]; }
    if (threadIdx.x == 0) rowoff[n_rows] = (uint16_t)V;
  }
  __syncthreads();
  // pass 2: scatter into the row groups
#pragma unroll 1
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * 1024) {
    uint32_t w[4], k[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int i = i0 + u * 1024; w[u] = i < n ? FS[i] : 0u; k[u] = i < n ? __ldcg(&FK[i]) : 0u; }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (!(w[u] >> 31)) continue;
      const uint32_t key = k[u] & 0x1ffffffu;
      keys[(int)rowoff[row_of(key)] + (int)(w[u] & 0x7fffffffu)] = ((unsigned long long)key << 32) | (unsigned)(i0 + u * 1024);
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
  // responses of the ordered keypoints (independent gathers, four in flight per thread)
  for (int i0 = threadIdx.x; i0 < V; i0 += 4 * 1024) {
    float rv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int i = i0 + u * 1024; rv[u] = i < V ? __ldcg(&FV[(int)(keys[i] & 0xffffffffu)].w) : 0.f; }
#pragma unroll
    for (int u = 0; u < 4; u++) { const int i = i0 + u * 1024; if (i < V) resp[i] = float_order_bits(rv[u]); }
  }
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
  // values of the kept keypoints of this thread's chunk (chunk <= 16: V <= 16384), gathered four at a time before they are used
  int my_keep = 0;
  for (int j0 = beg; j0 < end; j0 += 4) {
    float4 val[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int i = j0 + u; if (i < end && flag[i]) val[u] = __ldcg(&FV[(int)(keys[i] & 0xffffffffu)]); }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = j0 + u;
      if (i >= end || !flag[i]) continue;
      const int sc = kscale_from_bounds(s_bounds, val[u].z);
      const int bd = (int)s_sizes[sc];
      const bool out = val[u].x < (float)bd || val[u].x >= (float)(W - bd) || val[u].y < (float)bd || val[u].y >= (float)(H - bd);
      flag[i] = out ? 0 : (uint8_t)(sc + 1);
      my_keep += !out;
    }
  }
  int pos = block_exclusive_scan_1024(my_keep, sh_scan, total);
  for (int j0 = beg; j0 < end; j0 += 4) {
    float4 val[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int i = j0 + u; if (i < end && flag[i]) val[u] = __ldcg(&FV[(int)(keys[i] & 0xffffffffu)]); }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = j0 + u;
      if (i >= end || !flag[i]) continue;
      if (pos < kp_cap) {
        okb_keypoint_t k;
        k.x = val[u].x; k.y = val[u].y; k.size = val[u].z; k.angle = -1.f; k.response = val[u].w;
        k.octave = (int)((uint32_t)(keys[i] >> 32) >> 22); k.class_id = -1;
        kp_out[(size_t)frame * kp_cap + pos] = k;
        kscale_out[(size_t)frame * kp_cap + pos] = (int)flag[i] - 1;
      }
      pos++;
    }
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
        g.max_taps = std::max(*std::max_element(ax.count.begin(), ax.count.end()), *std::max_element(ay.count.begin(), ay.count.end()));
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
  if (const char* e = getenv("OKB_SCORE_TILE_H")) ws.score_tile_h = atoi(e) == 64 ? 64 : 32;   // tuning hook
  ws.kp_cap = c.max_keypoints > 0 ? (int)align_up((size_t)c.max_keypoints, 64) : kSortCap;
  OKB_CUDA(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
  OKB_CUDA(cudaStreamCreateWithFlags(&ws.stream2, cudaStreamNonBlocking));
  OKB_CUDA(cudaEventCreateWithFlags(&ws.ev_fork, cudaEventDisableTiming));
  OKB_CUDA(cudaEventCreateWithFlags(&ws.ev_join, cudaEventDisableTiming));
  OKB_CUDA(cudaEventCreateWithFlags(&ws.ev_pyr, cudaEventDisableTiming));
  OKB_CUDA(cudaMalloc(&ws.d_in, (size_t)W * H * B));
  OKB_CUDA(cudaMalloc(&ws.d_img, off * B));
  OKB_CUDA(cudaMalloc(&ws.d_score, off * B));
  OKB_CUDA(cudaMalloc(&ws.d_touch, off * B * 4));
  OKB_CUDA(cudaMemset(ws.d_touch, 0, off * B * 4));
  OKB_CUDA(cudaMemset(ws.d_score, 0, off * B));
  OKB_CUDA(cudaMemset(ws.d_img, 0, off * B));
  OKB_CUDA(cudaMalloc(&ws.d_integral, (size_t)(W + 1) * (H + 1) * 4 * B));
  OKB_CUDA(cudaMalloc(&ws.d_cand, (size_t)ws.cand_cap * 4 * B));
  {
    // per-call counters in one block (one memset per call): candidate counts, status words, tie-cell bitmaps
    const size_t cc = align_up((size_t)4 * kMaxLayers * B, 256), stb = align_up((size_t)4 * B, 256);
    ws.zero_bytes = cc + 2 * stb + (size_t)kCellWordsPerFrame * B * 4;
    uint8_t* blk = nullptr;
    OKB_CUDA(cudaMalloc(&blk, ws.zero_bytes));
    OKB_CUDA(cudaMemset(blk, 0, ws.zero_bytes));
    ws.d_cand_count = (int32_t*)blk; ws.d_status = (int32_t*)(blk + cc); ws.d_tie_count = (int32_t*)(blk + cc + stb);
    ws.d_tie_cells = (uint32_t*)(blk + cc + 2 * stb);
    OKB_CUDA(cudaMalloc(&ws.d_ties, (size_t)kMaxTies * sizeof(TieEntry) * B));
    OKB_CUDA(cudaMalloc(&ws.d_tie_sorted, kTieSortedWords * 4 * B));
  }
  OKB_CUDA(cudaMalloc(&ws.d_fkey, (size_t)ws.cand_cap * 4 * B));
  OKB_CUDA(cudaMalloc(&ws.d_fval, (size_t)ws.cand_cap * 16 * B));
  OKB_CUDA(cudaMalloc(&ws.d_fslot, (size_t)ws.cand_cap * 4 * B));
  OKB_CUDA(cudaMalloc(&ws.d_kp, (size_t)ws.kp_cap * sizeof(okb_keypoint_t) * B));
  OKB_CUDA(cudaMalloc(&ws.d_kscale, (size_t)ws.kp_cap * 4 * B));
  OKB_CUDA(cudaMalloc(&ws.d_desc, (size_t)ws.kp_cap * 64 * B));
  OKB_CUDA(cudaMalloc(&ws.d_count, 4 * B));
  OKB_CUDA(cudaMalloc(&ws.d_rays, (size_t)ws.kp_cap * 24 * B));
  OKB_CUDA(cudaMalloc(&ws.d_rays_valid, (size_t)ws.kp_cap * B));
  OKB_CUDA(cudaEventCreateWithFlags(&ws.ev_done, cudaEventDisableTiming));
  OKB_CUDA(cudaMalloc(&ws.d_epoch, 8));
  OKB_CUDA(cudaMemset(ws.d_epoch, 0, 8));
  OKB_CUDA(cudaMalloc(&ws.d_dbg, (size_t)16 * 8 * B));
  OKB_CUDA(cudaMemset(ws.d_dbg, 0, (size_t)16 * 8 * B));
  OKB_CUDA(cudaMemset(ws.d_count, 0, 4 * B));
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
  if (c.descriptor_bytes == 48) return harris_init_camera(ctx, cam);
  return OKB_OK;
}

void detect_free_camera(okb_context* ctx, int cam)
{
  CamWorkspace& ws = ctx->cams[cam];
  harris_free_camera(ctx, cam);
  for (int i = 0; i < kMaxLayers; i++) {
    LayerGeom& g = ws.geom[i];
    cudaFree(g.d_xs); cudaFree(g.d_xn); cudaFree(g.d_xa); cudaFree(g.d_ys); cudaFree(g.d_yn); cudaFree(g.d_ya);
  }
  cudaFree(ws.m2_d); cudaFree(ws.d_epoch); cudaFree(ws.d_ties); cudaFree(ws.d_tie_sorted); cudaFree(ws.d_tiles); cudaFree(ws.d_ray_map); cudaFree(ws.d_jac_map);
  cudaFree(ws.d_in); cudaFree(ws.d_img); cudaFree(ws.d_score); cudaFree(ws.d_touch); cudaFree(ws.d_integral); cudaFree(ws.d_cand);
  cudaFree(ws.d_cand_count); cudaFree(ws.d_fkey); cudaFree(ws.d_fval); cudaFree(ws.d_fslot); cudaFree(ws.d_kp); cudaFree(ws.d_kscale); cudaFree(ws.d_desc);
  cudaFree(ws.d_count); cudaFree(ws.d_m1_rows); cudaFree(ws.m_d); if (ws.m_h) cudaFreeHost(ws.m_h); cudaFree(ws.m3_d); if (ws.m3_h) cudaFreeHost(ws.m3_h); cudaFree(ws.motion.d); if (ws.motion.h) cudaFreeHost(ws.motion.h); cudaFree(ws.d_dbg); cudaFree(ws.d_rays); cudaFree(ws.d_rays_valid);
  if (ws.ev_done) cudaEventDestroy(ws.ev_done);
  cudaFreeHost(ws.h_img); cudaFreeHost(ws.h_kp); cudaFreeHost(ws.h_desc); cudaFreeHost(ws.h_count); cudaFreeHost(ws.h_status); cudaFreeHost(ws.h_rays); cudaFreeHost(ws.h_rays_valid);
  for (int i = 0; i < 4; i++) if (ws.ev[i]) cudaEventDestroy(ws.ev[i]);
  if (ws.ev_mid) cudaEventDestroy(ws.ev_mid);
  if (ws.ev_fork) cudaEventDestroy(ws.ev_fork);
  if (ws.ev_join) cudaEventDestroy(ws.ev_join);
  if (ws.ev_pyr) cudaEventDestroy(ws.ev_pyr);
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

// epoch[0] = current epoch (1..126), epoch[1] = 1 when this call wrapped it (the touch maps must be cleared)
__global__ void k_epoch_tick(uint32_t* epoch)
{
  if (threadIdx.x == 0) {
    const uint32_t e = epoch[0] + 1;
    const bool wrap = e >= 127;
    epoch[0] = wrap ? 1u : e; epoch[1] = wrap ? 1u : 0u;
  }
}
__global__ void __launch_bounds__(256) k_touch_clear(const uint32_t* epoch, uint32_t* touch, size_t n)
{
  if (!epoch[1]) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) touch[i] = 0u;
}

// the integral images of a batch on `st` (also used by the D = 48 mode, okb_harris.cu)
void integral_run(CamWorkspace& ws, const uint8_t* d_images, int src_pitch, size_t in_stride, int W, int H, int B, cudaStream_t st)
{
  k_integral_rows<<<dim3((H + 7) / 8, B), 256, 0, st>>>(d_images, src_pitch, in_stride, W, H, ws.d_integral, W + 1);
  k_integral_cols<<<dim3((W + 31) / 32, B), 1024, 0, st>>>(W, H, ws.d_integral, W + 1);
}

// all frames are device resident: d_images = n_frames x H x src_pitch
int detect_run_device(okb_context* ctx, int cam, int n_frames, const uint8_t* d_images, int src_pitch, cudaEvent_t input_ready)
{
  CamWorkspace& ws = ctx->cams[cam];
  const okb_camera_config_t& c = ws.cfg;
  if (c.descriptor_bytes == 48) return harris_run_device(ctx, cam, n_frames, d_images, src_pitch, input_ready);
  const int W = c.width, H = c.height, B = n_frames;
  cudaStream_t st = ws.stream;
  if (ctx->timers_on) collect_timing(ctx, ws);
  // the touch-map epoch lives on the device (a captured CUDA graph replays the same launches: nothing per call may be a kernel
  // argument): one thread advances it, the map is cleared when the 7-bit field wraps
  k_epoch_tick<<<1, 32, 0, st>>>(ws.d_epoch);
  k_touch_clear<<<296, 256, 0, st>>>(ws.d_epoch, ws.d_touch, (size_t)ws.dl.frame_stride * c.max_batch);
  ctx->launches += 2;
  // one contiguous block holds the per-call counters: candidate counts, status words, tie counts, tie-cell bitmaps
  OKB_CUDA(cudaMemsetAsync(ws.d_cand_count, 0, ws.zero_bytes, st));
  // a caller that uploads the frames on another stream (okb_process_multiframe) lets the three nodes above run underneath the copy
  if (input_ready) OKB_CUDA(cudaStreamWaitEvent(st, input_ready, 0));
  if (ctx->timers_on) cudaEventRecord(ws.ev[0], st);
  const size_t in_stride = (size_t)src_pitch * H;
  const int ipitch = W + 1;
  // ---- pyramid + dense scores + non-max candidates (okb_score.cu)
  CandRegions cr; for (int i = 0; i <= kMaxLayers; i++) cr.off[i] = ws.cand_off[i];
  { int rc = pyramid_score_run(ctx, ws, d_images, src_pitch, in_stride, B, cr); if (rc) return rc; }
  // ---- fork: the integral image only needs the input frames. It runs on a side stream underneath the refinement / tie
  //      resolution / selection kernels (latency-bound, two of them one CTA per frame: they leave most SMs idle), not
  //      underneath the pyramid + score pass, which fills the GPU by itself
  OKB_CUDA(cudaEventRecord(ws.ev_fork, st));
  OKB_CUDA(cudaStreamWaitEvent(ws.stream2, ws.ev_fork, 0));
  k_integral_rows<<<dim3((H + 7) / 8, B), 256, 0, ws.stream2>>>(d_images, src_pitch, in_stride, W, H, ws.d_integral, ipitch);
  k_integral_cols<<<dim3((W + 31) / 32, B), 1024, 0, ws.stream2>>>(W, H, ws.d_integral, ipitch);
  OKB_CUDA(cudaEventRecord(ws.ev_join, ws.stream2));
  if (ctx->timers_on) cudaEventRecord(ws.ev[1], st);
  // ---- candidates, refinement, tie resolution, selection
  static const int cpw_env = getenv("OKB_REFINE_CPW") ? atoi(getenv("OKB_REFINE_CPW")) : 0;   // tuning hook
  const int cpw = cpw_env > 0 ? cpw_env : (B >= 8 ? 32 : (B >= 4 ? 16 : 8));
  k_refine<<<dim3((ws.cand_cap + 4 * cpw - 1) / (4 * cpw), B), 128, 0, st>>>(ws.dl, d_images, src_pitch, in_stride, ws.d_img, ws.d_score,
                                                               ws.d_touch, ws.d_cand, ws.d_cand_count, ws.cand_cap, cr,
                                                               ws.d_fkey, (float4*)ws.d_fval, c.threshold, ws.d_epoch, ws.d_tie_cells, ws.d_ties, ws.d_tie_count, cpw);
  k_tie_gather<<<dim3(kMaxTies / 64, B), 64, 0, st>>>(ws.dl, ws.d_score, ws.d_touch, ws.d_ties, ws.d_tie_count, ws.d_epoch, c.threshold, ws.d_tie_sorted);
  k_resolve<<<B, kResolveThreads, kResolveSmem, st>>>(ws.d_tie_sorted, ws.d_tie_count, ws.d_fkey, ws.cand_cap, c.threshold, ws.d_status, ws.d_dbg);
  ctx->launches++;
  k_finalize<<<B, 1024, kFinalizeSmem, st>>>(ws.dl, ws.d_cand_count, ws.cand_cap, cr, ws.d_fkey, (const float4*)ws.d_fval, ws.d_fslot,
                                            ctx->d_scale_bounds, ctx->d_size_list,
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
