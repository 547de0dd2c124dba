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

__global__ void __launch_bounds__(128) k_refine(const __grid_constant__ DeviceLayers dl, const uint8_t* in0, int in_pitch, size_t in_frame_stride,
                                                uint8_t* img_block, uint8_t* score_block, uint32_t* touch_block,
                                                const uint32_t* cand, const int32_t* cand_count, int cand_cap,
                                                const __grid_constant__ CandRegions cr, CandRecord* rec, int threshold, const uint32_t* d_epoch,
                                                const uint32_t* tie_cells)
{
  const int frame = blockIdx.y;
  int prefix[kMaxLayers + 1];
  const int n = cand_total(cr, cand_count + frame * kMaxLayers, dl.n, prefix);
  if (blockIdx.x * blockDim.x >= n) return;
  const uint32_t epoch = *d_epoch;
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
                                                 CandRecord* rec, const uint32_t* d_epoch, int threshold, int32_t* status, long long* dbg)
{
  const uint32_t epoch = *d_epoch;
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
    ws.zero_bytes = cc + stb + (size_t)kCellWordsPerFrame * B * 4;
    uint8_t* blk = nullptr;
    OKB_CUDA(cudaMalloc(&blk, ws.zero_bytes));
    OKB_CUDA(cudaMemset(blk, 0, ws.zero_bytes));
    ws.d_cand_count = (int32_t*)blk; ws.d_status = (int32_t*)(blk + cc); ws.d_tie_cells = (uint32_t*)(blk + cc + stb);
  }
  OKB_CUDA(cudaMalloc(&ws.d_rec, (size_t)ws.cand_cap * sizeof(CandRecord) * B));
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
  return OKB_OK;
}

void detect_free_camera(okb_context* ctx, int cam)
{
  CamWorkspace& ws = ctx->cams[cam];
  for (int i = 0; i < kMaxLayers; i++) {
    LayerGeom& g = ws.geom[i];
    cudaFree(g.d_xs); cudaFree(g.d_xn); cudaFree(g.d_xa); cudaFree(g.d_ys); cudaFree(g.d_yn); cudaFree(g.d_ya);
  }
  cudaFree(ws.d_epoch); cudaFree(ws.d_tiles); cudaFree(ws.d_ray_map); cudaFree(ws.d_jac_map);
  cudaFree(ws.d_in); cudaFree(ws.d_img); cudaFree(ws.d_score); cudaFree(ws.d_touch); cudaFree(ws.d_integral); cudaFree(ws.d_cand);
  cudaFree(ws.d_cand_count); cudaFree(ws.d_rec); cudaFree(ws.d_kp); cudaFree(ws.d_kscale); cudaFree(ws.d_desc);
  cudaFree(ws.d_count); cudaFree(ws.d_m1_rows); cudaFree(ws.m_d); if (ws.m_h) cudaFreeHost(ws.m_h); cudaFree(ws.m3_d); if (ws.m3_h) cudaFreeHost(ws.m3_h); cudaFree(ws.motion.d); if (ws.motion.h) cudaFreeHost(ws.motion.h); cudaFree(ws.d_dbg); cudaFree(ws.d_rays); cudaFree(ws.d_rays_valid);
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

// all frames are device resident: d_images = n_frames x H x src_pitch
int detect_run_device(okb_context* ctx, int cam, int n_frames, const uint8_t* d_images, int src_pitch)
{
  CamWorkspace& ws = ctx->cams[cam];
  const okb_camera_config_t& c = ws.cfg;
  const int W = c.width, H = c.height, B = n_frames;
  cudaStream_t st = ws.stream;
  if (ctx->timers_on) collect_timing(ctx, ws);
  // the touch-map epoch lives on the device (a captured CUDA graph replays the same launches: nothing per call may be a kernel
  // argument): one thread advances it, the map is cleared when the 7-bit field wraps
  k_epoch_tick<<<1, 32, 0, st>>>(ws.d_epoch);
  k_touch_clear<<<296, 256, 0, st>>>(ws.d_epoch, ws.d_touch, (size_t)ws.dl.frame_stride * c.max_batch);
  ctx->launches += 2;
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
  k_refine<<<dim3((ws.cand_cap + 127) / 128, B), 128, 0, st>>>(ws.dl, d_images, src_pitch, in_stride, ws.d_img, ws.d_score,
                                                               ws.d_touch, ws.d_cand, ws.d_cand_count, ws.cand_cap, cr,
                                                               ws.d_rec, c.threshold, ws.d_epoch, ws.d_tie_cells);
  k_resolve<<<B, 512, kResolveSmem, st>>>(ws.dl, ws.d_score, ws.d_touch, ws.d_cand_count, ws.cand_cap, cr, ws.d_rec, ws.d_epoch, c.threshold,
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
