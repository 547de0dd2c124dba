// okb_aux.cu -- K1: keyframe-overlap masks on the device (the step immediately after matching).
//
// Replaces the mask painting + counting of Frontend::doWeNeedANewKeyframe (reference okvis_frontend/src/Frontend.cpp:
// 1058-1167) and ViSlamBackend::overlapFraction (okvis_ceres/src/ViSlamBackend.cpp:2341-2426): per camera image two
// (rows/10) x (cols/10) masks -- every keypoint ("detections") / every keypoint with a matched landmark ("matches") drawn
// as cv::circle(mask, pt*0.1, int(min(rows,cols)*kptrad), 255, FILLED) -- and the pixel counts of their AND and OR.
// The reference repeats this for the current frame and for every keyframe of the window, on every frame.
//
// The filled circle of OpenCV's Circle() (midpoint iteration, drawing.cpp) is a fixed shape per radius: a table of row
// half-widths, built on the host by running the same iteration (build_half_widths), stamped clipped to the mask.
// One CTA per view; the two masks are bit planes in shared memory (atomicOr of word-wide span masks), counted by __popc.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "okb_internal.h"

namespace okb {

constexpr int kMaxRadius = 63;
struct OverlapArgs {
  const okb_overlap_view_t* views; const float* xy; const uint8_t* matched; int32_t* out;   // out: n_views x 2
  int8_t hw[2][kMaxRadius + 1];   // half widths for the two possible radii (views may differ in size)
  int radius_key[2];              // radius of table 0 / 1
  double kptrad;
};

// the spans OpenCV's Circle() fills, as half widths per |row offset| (0..radius); -1 = row not touched
static void build_half_widths(int radius, int8_t* hw)
{
  for (int i = 0; i <= kMaxRadius; i++) hw[i] = -1;
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  while (dx >= dy) {
    hw[dy] = (int8_t)std::max<int>(hw[dy], dx);
    hw[dx] = (int8_t)std::max<int>(hw[dx], dy);
    dy++;
    err += plus;
    plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask;
    dx += mask;
    minus -= mask & 2;
  }
}

__global__ void __launch_bounds__(256) k_overlap(const __grid_constant__ OverlapArgs a)
{
  extern __shared__ uint32_t planes[];   // det[rows][wpr] then mat[rows][wpr]
  const okb_overlap_view_t v = a.views[blockIdx.x];
  const int rows = v.image_rows / 10, cols = v.image_cols / 10;
  const int wpr = (cols + 31) >> 5;
  uint32_t* det = planes; uint32_t* mat = planes + rows * wpr;
  for (int i = threadIdx.x; i < 2 * rows * wpr; i += blockDim.x) planes[i] = 0u;
  __syncthreads();
  const int radius = (int)((double)min(rows, cols) * a.kptrad);
  const int8_t* hw = a.hw[radius == a.radius_key[0] ? 0 : 1];
  for (int k = threadIdx.x; k < v.n_keypoints; k += blockDim.x) {
    const size_t g = (size_t)v.first_keypoint + k;
    const float px = (float)((double)a.xy[2 * g] * 0.1), py = (float)((double)a.xy[2 * g + 1] * 0.1);   // Point2f * double
    const int cx = __float2int_rn(px), cy = __float2int_rn(py);                                          // Point2f -> Point
    const bool m = a.matched[g] != 0;
    for (int dy = -radius; dy <= radius; dy++) {
      const int y = cy + dy, h = hw[dy < 0 ? -dy : dy];
      if (h < 0 || y < 0 || y >= rows) continue;
      const int x0 = max(cx - h, 0), x1 = min(cx + h, cols - 1);
      if (x0 > x1) continue;
      for (int w = x0 >> 5; w <= (x1 >> 5); w++) {
        const int lo = max(x0 - 32 * w, 0), hi = min(x1 - 32 * w, 31);
        const uint32_t bits = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
        atomicOr(&det[y * wpr + w], bits);
        if (m) atomicOr(&mat[y * wpr + w], bits);
      }
    }
  }
  __syncthreads();
  int ic = 0, uc = 0;
  for (int i = threadIdx.x; i < rows * wpr; i += blockDim.x) { ic += __popc(det[i] & mat[i]); uc += __popc(det[i] | mat[i]); }
  __shared__ int s_i, s_u;
  if (threadIdx.x == 0) { s_i = 0; s_u = 0; }
  __syncthreads();
  ic = __reduce_add_sync(0xffffffffu, ic); uc = __reduce_add_sync(0xffffffffu, uc);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_i, ic); atomicAdd(&s_u, uc); }
  __syncthreads();
  if (threadIdx.x == 0) { a.out[2 * blockIdx.x] = s_i; a.out[2 * blockIdx.x + 1] = s_u; }
}

// ---- B1: DBoW2 vocabulary-tree descent (TemplatedVocabulary<FBrisk>::transform; DBoW2 is an empty submodule in
// /root/reference, the published algorithm is restated in oracle/bow_oracle.py): from the root, at every level take the
// child with the smallest FBrisk::distance (= H0, okvis_frontend/src/FBrisk.cpp:64-67; first minimum in child order,
// strict <), until a leaf; report its word id / weight and the node passed `levelsup` levels above the leaves.
// One thread per feature; the node descriptors (k^L nodes x D bytes, 39 KB for small_voc) are read through L1/L2.
struct BowArgs {
  int n, D16, k, L, levelsup, n_nodes;
  const uint4* node_desc;       // [n_nodes + 1][D16], index = node id (0 = root, unused)
  const int32_t* children;      // [n_nodes + 1][k], -1 padded, in file order
  const int32_t* word_of;       // [n_nodes + 1], -1 for inner nodes
  const double* weight;         // [n_nodes + 1]
  const uint4* feat; int32_t* out_word; double* out_weight; int32_t* out_node;
};

__global__ void __launch_bounds__(128) k_bow_transform(const __grid_constant__ BowArgs a)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  uint4 f[4];
  for (int w = 0; w < a.D16; w++) f[w] = a.feat[(size_t)i * a.D16 + w];
  const int nid_level = a.L - a.levelsup;
  int final_id = 0, level = 0, nid = 0;
  while (true) {
    const int32_t* ch = a.children + (size_t)final_id * a.k;
    if (ch[0] < 0) break;   // leaf
    ++level;
    int best = -1; uint32_t best_d = 0xffffffffu;
    for (int c = 0; c < a.k; c++) {
      const int id = ch[c];
      if (id < 0) break;
      uint32_t d = 0;
      for (int w = 0; w < a.D16; w++) {
        const uint4 q = a.node_desc[(size_t)id * a.D16 + w];
        d += __popcll(((unsigned long long)(q.x ^ f[w].x) << 32) | (q.y ^ f[w].y));
        d += __popcll(((unsigned long long)(q.z ^ f[w].z) << 32) | (q.w ^ f[w].w));
      }
      if (d < best_d) { best_d = d; best = id; }
    }
    final_id = best;
    if (level == nid_level) nid = final_id;
  }
  a.out_word[i] = a.word_of[final_id];
  a.out_weight[i] = a.weight[final_id];
  a.out_node[i] = nid;
}

struct AuxState {
  uint8_t* d_buf = nullptr; size_t cap = 0; cudaStream_t stream = nullptr;
  // vocabulary
  uint8_t* d_voc = nullptr; int voc_nodes = 0, voc_k = 0, voc_L = 0, voc_D = 0;
  size_t o_children = 0, o_word = 0, o_weight = 0;
  uint8_t* d_bow = nullptr; size_t bow_cap = 0;
};

void aux_free(okb_context* ctx)
{
  if (!ctx->aux) return;
  AuxState* s = static_cast<AuxState*>(ctx->aux);
  cudaFree(s->d_buf); cudaFree(s->d_voc); cudaFree(s->d_bow);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  ctx->aux = nullptr;
}

static int aux_state(okb_context* ctx)
{
  if (ctx->aux) return OKB_OK;
  AuxState* s = new AuxState();
  ctx->aux = s;
  OKB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  return OKB_OK;
}

}  // namespace okb

using namespace okb;

extern "C" int okb_bow_load(okb_context_t* ctx, int D, int k, int L, int n_nodes, const int32_t* node_id, const int32_t* parent_id,
                            const double* weight, const uint8_t* desc, int n_words, const int32_t* word_id, const int32_t* word_node)
{
  if (!ctx || (D != 48 && D != 64) || k < 1 || L < 1 || n_nodes < 1 || !node_id || !parent_id || !weight || !desc ||
      n_words < 0 || (n_words > 0 && (!word_id || !word_node))) { set_error("okb_bow_load: bad arguments"); return OKB_ERR_ARGUMENT; }
  OKB_CUDA(cudaSetDevice(ctx->device));
  { const int rc = aux_state(ctx); if (rc) return rc; }
  AuxState* s = static_cast<AuxState*>(ctx->aux);
  const size_t N = (size_t)n_nodes + 1;
  std::vector<uint8_t> h_desc(N * D, 0); std::vector<int32_t> h_children(N * k, -1), h_word(N, -1); std::vector<double> h_weight(N, 0.0);
  std::vector<int> fill(N, 0);
  for (int i = 0; i < n_nodes; i++) {
    const int id = node_id[i], p = parent_id[i];
    if (id < 1 || id > n_nodes || p < 0 || p > n_nodes || fill[p] >= k) { set_error("okb_bow_load: node %d (id %d, parent %d) is malformed", i, id, p); return OKB_ERR_ARGUMENT; }
    memcpy(&h_desc[(size_t)id * D], desc + (size_t)i * D, D);
    h_weight[id] = weight[i];
    h_children[(size_t)p * k + fill[p]++] = id;   // children in file order, as DBoW2 builds them while loading
  }
  for (int i = 0; i < n_words; i++) {
    if (word_node[i] < 1 || word_node[i] > n_nodes) { set_error("okb_bow_load: word %d names node %d", i, word_node[i]); return OKB_ERR_ARGUMENT; }
    h_word[word_node[i]] = word_id[i];
  }
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  s->o_children = al(N * D); s->o_word = s->o_children + al(N * k * 4); s->o_weight = s->o_word + al(N * 4);
  const size_t total = s->o_weight + al(N * 8);
  OKB_CUDA(cudaStreamSynchronize(s->stream));
  cudaFree(s->d_voc); s->d_voc = nullptr;
  OKB_CUDA(cudaMalloc(&s->d_voc, total));
  OKB_CUDA(cudaMemcpy(s->d_voc, h_desc.data(), N * D, cudaMemcpyHostToDevice));
  OKB_CUDA(cudaMemcpy(s->d_voc + s->o_children, h_children.data(), N * k * 4, cudaMemcpyHostToDevice));
  OKB_CUDA(cudaMemcpy(s->d_voc + s->o_word, h_word.data(), N * 4, cudaMemcpyHostToDevice));
  OKB_CUDA(cudaMemcpy(s->d_voc + s->o_weight, h_weight.data(), N * 8, cudaMemcpyHostToDevice));
  s->voc_nodes = n_nodes; s->voc_k = k; s->voc_L = L; s->voc_D = D;
  return OKB_OK;
}

extern "C" int okb_bow_transform(okb_context_t* ctx, int n, const uint8_t* desc, int levelsup, int32_t* out_word, double* out_weight,
                                 int32_t* out_node)
{
  AuxState* s = ctx ? static_cast<AuxState*>(ctx->aux) : nullptr;
  if (!s || !s->d_voc) { set_error("okb_bow_transform: no vocabulary loaded"); return OKB_ERR_ARGUMENT; }
  if (n < 0 || levelsup < 0 || (n > 0 && (!desc || !out_word))) { set_error("okb_bow_transform: bad arguments"); return OKB_ERR_ARGUMENT; }
  if (n == 0) return OKB_OK;
  OKB_CUDA(cudaSetDevice(ctx->device));
  const size_t D = (size_t)s->voc_D;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t o_w = al(n * D), o_wt = o_w + al((size_t)n * 4), o_n = o_wt + al((size_t)n * 8), total = o_n + al((size_t)n * 4);
  cudaStream_t st = s->stream;
  if (total > s->bow_cap) {
    OKB_CUDA(cudaStreamSynchronize(st));
    cudaFree(s->d_bow); s->d_bow = nullptr; s->bow_cap = 0;
    OKB_CUDA(cudaMalloc(&s->d_bow, total * 2));
    s->bow_cap = total * 2;
  }
  uint8_t* d = s->d_bow;
  OKB_CUDA(cudaMemcpyAsync(d, desc, n * D, cudaMemcpyHostToDevice, st));
  BowArgs a;
  a.n = n; a.D16 = s->voc_D / 16; a.k = s->voc_k; a.L = s->voc_L; a.levelsup = levelsup; a.n_nodes = s->voc_nodes;
  a.node_desc = (const uint4*)s->d_voc; a.children = (const int32_t*)(s->d_voc + s->o_children);
  a.word_of = (const int32_t*)(s->d_voc + s->o_word); a.weight = (const double*)(s->d_voc + s->o_weight);
  a.feat = (const uint4*)d; a.out_word = (int32_t*)(d + o_w); a.out_weight = (double*)(d + o_wt); a.out_node = (int32_t*)(d + o_n);
  k_bow_transform<<<(n + 127) / 128, 128, 0, st>>>(a);
  ctx->launches++;
  OKB_CUDA(cudaMemcpyAsync(out_word, d + o_w, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (out_weight) OKB_CUDA(cudaMemcpyAsync(out_weight, d + o_wt, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (out_node) OKB_CUDA(cudaMemcpyAsync(out_node, d + o_n, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(cudaStreamSynchronize(st));
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

extern "C" int okb_overlap_counts(okb_context_t* ctx, int n_views, const okb_overlap_view_t* views, int n_keypoints, const float* xy,
                                  const uint8_t* matched, double kptrad, int32_t* out_intersection, int32_t* out_union)
{
  if (!ctx || n_views < 0 || n_keypoints < 0 || (n_views > 0 && (!views || !out_intersection || !out_union)) ||
      (n_keypoints > 0 && (!xy || !matched))) { set_error("okb_overlap_counts: bad arguments"); return OKB_ERR_ARGUMENT; }
  if (n_views == 0) return OKB_OK;
  OKB_CUDA(cudaSetDevice(ctx->device));
  OverlapArgs a;
  memset(&a, 0, sizeof(a));
  a.kptrad = kptrad; a.radius_key[0] = a.radius_key[1] = -1;
  size_t smem = 0;
  for (int i = 0; i < n_views; i++) {
    const okb_overlap_view_t& v = views[i];
    if (v.image_rows < 10 || v.image_cols < 10 || v.n_keypoints < 0 || v.first_keypoint < 0 || v.first_keypoint + v.n_keypoints > n_keypoints) {
      set_error("okb_overlap_counts: view %d is malformed", i); return OKB_ERR_ARGUMENT;
    }
    const int rows = v.image_rows / 10, cols = v.image_cols / 10;
    const int radius = (int)((double)std::min(rows, cols) * kptrad);
    if (radius < 0 || radius > kMaxRadius) { set_error("okb_overlap_counts: radius %d unsupported", radius); return OKB_ERR_UNSUPPORTED; }
    int slot = radius == a.radius_key[0] ? 0 : (radius == a.radius_key[1] ? 1 : (a.radius_key[0] < 0 ? 0 : (a.radius_key[1] < 0 ? 1 : -1)));
    if (slot < 0) { set_error("okb_overlap_counts: more than two distinct mask radii in one call"); return OKB_ERR_UNSUPPORTED; }
    if (a.radius_key[slot] != radius) { a.radius_key[slot] = radius; build_half_widths(radius, a.hw[slot]); }
    smem = std::max(smem, (size_t)2 * rows * ((cols + 31) / 32) * 4);
  }
  if (smem > 200 * 1024) { set_error("okb_overlap_counts: image too large for the shared-memory masks"); return OKB_ERR_UNSUPPORTED; }
  { const int rc = aux_state(ctx); if (rc) return rc; }
  AuxState* s = static_cast<AuxState*>(ctx->aux);
  const size_t b_views = (size_t)n_views * sizeof(okb_overlap_view_t), b_xy = (size_t)n_keypoints * 8, b_m = (size_t)n_keypoints;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t o_xy = al(b_views), o_m = o_xy + al(b_xy), o_out = o_m + al(b_m), total = o_out + al((size_t)n_views * 8);
  if (total > s->cap) {
    OKB_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(s->d_buf); s->d_buf = nullptr; s->cap = 0;
    OKB_CUDA(cudaMalloc(&s->d_buf, total * 2));
    s->cap = total * 2;
  }
  uint8_t* d = s->d_buf;
  cudaStream_t st = s->stream;
  OKB_CUDA(cudaMemcpyAsync(d, views, b_views, cudaMemcpyHostToDevice, st));
  if (n_keypoints) {
    OKB_CUDA(cudaMemcpyAsync(d + o_xy, xy, b_xy, cudaMemcpyHostToDevice, st));
    OKB_CUDA(cudaMemcpyAsync(d + o_m, matched, b_m, cudaMemcpyHostToDevice, st));
  }
  a.views = (const okb_overlap_view_t*)d; a.xy = (const float*)(d + o_xy); a.matched = d + o_m; a.out = (int32_t*)(d + o_out);
  if (smem > 48 * 1024) OKB_CUDA(cudaFuncSetAttribute(k_overlap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_overlap<<<n_views, 256, smem, st>>>(a);
  ctx->launches++;
  std::vector<int32_t> out((size_t)n_views * 2);
  OKB_CUDA(cudaMemcpyAsync(out.data(), d + o_out, out.size() * 4, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(cudaStreamSynchronize(st));
  OKB_CUDA(cudaGetLastError());
  for (int i = 0; i < n_views; i++) { out_intersection[i] = out[2 * i]; out_union[i] = out[2 * i + 1]; }
  return OKB_OK;
}
