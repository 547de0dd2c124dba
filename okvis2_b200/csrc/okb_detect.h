// okb_detect.h -- definitions shared by the two translation units of the detector: okb_score.cu (pyramid + dense score
// + non-max candidates) and okb_detect.cu (refinement, tie resolution, selection, integral image, descriptors).
#pragma once
#include "okb_internal.h"

namespace okb {

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

// candidate word written by the score kernel: time key | tie << 31 | pending << 30
constexpr uint32_t kCandTie = 0x80000000u, kCandPending = 0x40000000u, kCandKeyMask = 0x3fffffffu;

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

// okb_score.cu: the pyramid + score pass of one batch (all launches on ws.stream; records ws.ev_mid before the score
// kernel when the timers are on)
int pyramid_score_run(okb_context* ctx, CamWorkspace& ws, const uint8_t* d_images, int src_pitch, size_t in_stride, int n_frames,
                      const CandRegions& cr);

}  // namespace okb
