// okb_match.cu -- brute-force Hamming matchers with the reference's geometric gates (sm_100a).
//
// Replaces the five match loops of okvis::Frontend (reference okvis_frontend/src/Frontend.cpp):
//   M1 matchToMapByThread :1515-1590      M2 matchToMapByThreadUnitialised :1594-1720
//   M3 matchMotionStereo worker :1809-1907  M4 matchStereo :2016-2074   M5 verifyRecognisedPlace :329-355
// and triangulation::triangulateFast (okvis_frontend/src/stereo_triangulation.cpp:50-132).
//
// Layout: one warp per query descriptor (held in registers as uint4 words), candidates streamed through shared
// memory in tiles of 256 (word-plane layout -> conflict-free uint4 reads), distance = __popcll over the XOR.
// M1/M5 have no gate after the Hamming test, so "first minimum in iteration order" is a lexicographic
// (distance, position) minimum reduced with warp shuffles. M2-M4 evaluate their gates only when a candidate beats
// the running best, exactly like the sequential loops: lanes are replayed in candidate order (ballot + ffs), the
// gate is computed warp-uniformly in fp64 (no FMA contraction: compiled with -fmad=false).
#include <math.h>
#include <string.h>

#include <atomic>

#include "okb_internal.h"
#include "okb_umma.h"
#include "okb_gatecos.h"
#include "okb_camdev.h"

namespace okb {

// ---- fp64 geometry with a fixed association order (left to right), mirrored by the oracle ----------------------
struct V3 { double x, y, z; };
__device__ __forceinline__ V3 v3(const double* p) { return V3{p[0], p[1], p[2]}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double norm(V3 a) { return sqrt(dot(a, a)); }
__device__ __forceinline__ V3 normalized(V3 a)
{
  const double z = dot(a, a);
  if (z > 0.0) { const double n = sqrt(z); return V3{a.x / n, a.y / n, a.z / n}; }
  return a;
}

struct Tri { V3 p; bool valid, parallel; };

// triangulation::triangulateFast; cos26 = cos(2.6 sigma), cos6 = cos(6 sigma) come from gate_cos (okb_gatecos.h)
__device__ Tri triangulate_fast(V3 p1, V3 e1, V3 p2, V3 e2, double cos26, double cos6)
{
  Tri r; r.parallel = false; r.valid = true;
  const V3 t12 = p2 - p1;
  const double b0 = dot(t12, e1), b1 = dot(t12, e2);
  const double A00 = dot(e1, e1), A10 = dot(e1, e2), A01 = -A10, A11 = -dot(e2, e2);
  const double det = A00 * A11 - A10 * A01;
  const bool invertible = fabs(det) > 1.0e-12;
  double l0 = 0, l1 = 0;
  if (invertible) {
    const double invdet = 1.0 / det;
    const double i00 = A11 * invdet, i10 = -A10 * invdet, i01 = -A01 * invdet, i11 = A00 * invdet;
    l0 = i00 * b0 + i01 * b1;
    l1 = i10 * b0 + i11 * b1;
  }
  if (!invertible || l0 < 0.01 || l1 < 0.01) {
    r.parallel = true;
    const V3 m = p1 + 0.5 * t12;
    const double nt = norm(t12);
    const V3 mid = m + (40.0 * (0.01 < nt ? nt : 0.01)) * (e1 + e2);
    r.p = mid;
    if (dot(e1, normalized(mid - p1)) < cos26) r.valid = false;
    if (dot(e2, normalized(mid - p2)) < cos26) r.valid = false;
    return r;
  }
  const V3 xm = l0 * e1 + p1;
  const V3 xn = l1 * e2 + p2;
  const V3 s = xm + xn;
  const V3 mid = V3{s.x / 2.0, s.y / 2.0, s.z / 2.0};
  r.p = mid;
  if (dot(e1, normalized(mid - p1)) < cos26) r.valid = false;
  if (dot(e2, normalized(mid - p2)) < cos26) r.valid = false;
  if (dot(normalized(mid - p2), normalized(mid - p1)) > cos6) r.parallel = true;
  return r;
}

// z / w of T_CW * hp with T_CW = [R | t] row-major 3x4 and hp = (p, 1)
__device__ __forceinline__ double depth_in(const double* T, V3 p)
{
  const double z = ((T[8] * p.x + T[9] * p.y) + T[10] * p.z) + T[11] * 1.0;
  return z / 1.0;
}

// ---------------------------------------------------------------------------------------------------------------
enum { MODE_M2 = 2, MODE_M3 = 3, MODE_M4 = 4 };

// -1 / 1 (default): the gate / finish (/ check / commit) of M3 and M4 run as ONE launch per view or pair (k_pair_view, one CTA
// per frame); 0: the separate kernels (okb_m3_set_fused). Both forms give identical results and are in the parity tests.
static std::atomic<int> g_m3_fused{getenv("OKB_M3_FUSED") ? atoi(getenv("OKB_M3_FUSED")) : -1};   // env: tuning hook
// the Hamming scans of the device-resident M3 / M4: 2 (default) tcgen05 + TMEM (k_scan_umma), 1 legacy integer MMA (k_scan_mma),
// 0 POPC (k_m4_scan); identical results (okb_scan_set_mma)
static std::atomic<int> g_scan_mma{getenv("OKB_SCAN_MODE") ? atoi(getenv("OKB_SCAN_MODE")) : 2};   // env: tuning hook

struct MatchArgs {
  int nq, nc;
  const uint8_t* q_desc; const uint8_t* c_desc;
  const uint8_t* q_use;
  // M1
  const double* q_xy; const okb_keypoint_t* q_kp; const int32_t* c_lm; const double* lm_proj; const uint8_t* lm_is3d; double thr_sq;
  // M2
  const double* q_e; const int32_t* q_prev_lm; const double* c_e; const double* c_r;
  double r0[3], r1[3]; double cos26, cos6;
  // M3 / M4
  const double* q_sof; const double* q_cos26; const double* q_cos6;
  const uint8_t* c_valid; const double* c_sof; const double* c_cos26; const double* c_cos6;
  double T0[12], T1[12];
  uint32_t thr;
  // batched device form (blockIdx.y = frame): element strides of the per-frame arrays, 0 for the single-frame host form
  size_t q_stride, proj_stride, c_stride;
  const int32_t* q_count; const int32_t* c_count;   // per-frame counts (device) or nullptr
  uint32_t* out_dist; int32_t* out_idx; double* out_hp; uint8_t* out_init; int32_t* out_ctr;
  // M4 scan/gate split: when set, k_match_gated only runs for frames whose hit list overflowed (hit_cnt[frame] > hit_cap)
  const int32_t* hit_cnt; int hit_cap;
  // M3 sequence (okb_match_motion_stereo_device*): the queries of frame b are the keypoints of the older view
  // views[b * view_stride + view_index]; poses per frame
  const struct M3View* views; int view_stride, view_index; const struct M3Frame* frames;
  // k_m4_scan over several views at once (blockIdx.z = view * scan_chunks + query chunk): per-view slices of q_use (nq apart),
  // of the hit lists (gridDim.y * hit_cap apart) and of hit_cnt (gridDim.y apart); scan_qt = queries per CTA (<= 128)
  int scan_views, scan_chunks, scan_qt;
  // k_scan_mma: optional compacted list of the eligible queries per (view, frame) ([views][frames][nq] indices + [views][frames] counts);
  // null = all queries below the frame's count, q_use tested per hit
  const int32_t* q_list; const int32_t* q_list_cnt;
};

// one older keyframe view (device pointers) and the per-frame pose of the current camera, as uploaded by the M3 sequence
struct M3View { const uint8_t* desc; const double* rays; const uint8_t* valid; const float* size; const uint8_t* use; int n, pad; double Twc[12], Tcw[12]; };
struct M3Frame { double Twc[12], Tcw[12]; };
// z of T * (p, 1) for a pose stored as (C row-major 9, r 3): Transformation::operator*(Vector4d) = C * head + r * s
__device__ __forceinline__ double depth_cr(const double* T, V3 p) { return (((T[6] * p.x + T[7] * p.y) + T[8] * p.z) + T[11] * 1.0) / 1.0; }

constexpr int kTile = 256;

template <int D16>
__device__ __forceinline__ void load_query(const uint8_t* desc, int q, uint4 (&qd)[D16])
{
  const uint4* p = reinterpret_cast<const uint4*>(desc) + (size_t)q * D16;
#pragma unroll
  for (int i = 0; i < D16; i++) qd[i] = __ldg(p + i);
}
template <int D16>
__device__ __forceinline__ void stage_tile(const uint8_t* c_desc, int nc, int tile0, uint4 (*s_desc)[kTile])
{
  const uint4* g = reinterpret_cast<const uint4*>(c_desc);
  for (int i = threadIdx.x; i < kTile * D16; i += blockDim.x) {
    const int c = tile0 + i / D16, w = i % D16;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c < nc) v = __ldg(g + (size_t)c * D16 + w);
    s_desc[w][i / D16] = v;
  }
}
// asynchronous variant (cp.async / LDGSTS, 16 bytes per request, zero fill past the end); completes with cp_async_wait
template <int D16>
__device__ __forceinline__ void stage_tile_async(const uint8_t* c_desc, int nc, int tile0, uint4 (*s_desc)[kTile])
{
  const uint4* g = reinterpret_cast<const uint4*>(c_desc);
  for (int i = threadIdx.x; i < kTile * D16; i += blockDim.x) {
    const int c = tile0 + i / D16, w = i % D16;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_desc[w][i / D16]);
    const bool ok = c < nc;
    const uint4* src = g + (ok ? (size_t)c * D16 + w : 0);
    const int src_bytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int D16>
__device__ __forceinline__ uint32_t hamming(const uint4 (&qd)[D16], uint4 (*s_desc)[kTile], int ci)
{
  uint32_t d = 0;
#pragma unroll
  for (int w = 0; w < D16; w++) {
    const uint4 c = s_desc[w][ci];
    d += __popcll(((unsigned long long)(qd[w].x ^ c.x) << 32) | (qd[w].y ^ c.y));
    d += __popcll(((unsigned long long)(qd[w].z ^ c.z) << 32) | (qd[w].w ^ c.w));
  }
  return d;
}

// Hamming distance with a warp-uniform early exit: POPC is the bound of the gated matchers (16 results/clk/SM), and a
// candidate can only matter if its distance is < `best` (strict, as in the reference loops). The popcount of the first
// half of the descriptor is a lower bound of the distance; when no lane of the warp is still below `best` after that
// half, the second half is skipped for the whole warp. A skipped lane returns its partial count, which is >= best, so
// every `d < best` test downstream decides exactly as with the full distance.
template <int D16>
__device__ __forceinline__ uint32_t hamming_bounded(const uint4 (&qd)[D16], uint4 (*s_desc)[kTile], int ci, bool valid, uint32_t best)
{
  // The LAST 32 bytes go first: BRISK's later short pairs discriminate better (measured on the synthetic stereo set:
  // a 32-candidate chunk still holds a partial distance < 60 in 2 % of the cases after bytes 32..63, in 48 % after
  // bytes 0..31), so the early exit almost always saves half of the popcounts.
  constexpr int H = D16 / 2;
  uint32_t d = 0;
#pragma unroll
  for (int w = H; w < D16; w++) {
    const uint4 c = s_desc[w][ci];
    d += __popcll(((unsigned long long)(qd[w].x ^ c.x) << 32) | (qd[w].y ^ c.y));
    d += __popcll(((unsigned long long)(qd[w].z ^ c.z) << 32) | (qd[w].w ^ c.w));
  }
  if (!valid) d = 0xffffu;
  if (!__any_sync(0xffffffffu, d < best)) return d;
#pragma unroll
  for (int w = 0; w < H; w++) {
    const uint4 c = s_desc[w][ci];
    d += __popcll(((unsigned long long)(qd[w].x ^ c.x) << 32) | (qd[w].y ^ c.y));
    d += __popcll(((unsigned long long)(qd[w].z ^ c.z) << 32) | (qd[w].w ^ c.w));
  }
  return valid ? d : 0xffffu;
}

// M1. The reprojection gate makes the problem sparse: a keypoint can only match pooled descriptors whose landmark projects
// within `thr` pixels of it. Per frame the pool ROWS of 3-D landmarks are binned by their projection into a grid of cells
// >= thr wide (k_m1_rowbin: one CTA per frame, counting sort in shared memory; rows that project more than thr + 1 px
// outside the keypoint extent cannot pass the gate and are dropped; NaN projections -- which the reference's gate does NOT
// reject -- go to an extra cell every keypoint visits). Then one WARP per keypoint (k_m1_match) walks the 3x3 cells around
// it, lanes over the rows: coalesced (projection, row) records, the reference's exact fp64 gate
// (reprDist.dot(reprDist) > thr^2 -> skip), Hamming distance for the rows that pass, and a warp reduction of
// (distance << 32 | row). The lexicographic minimum is exactly "first strict minimum in ascending LandmarkId / descriptor
// order" of the sequential loop, and it is order independent, so the result is bit-identical.
constexpr int kMaxCells = 4096;

struct M1Args {
  int nq;                       // keypoint capacity per frame
  const int32_t* q_count;       // per-frame keypoint count (device) or nullptr
  const uint8_t* q_desc; const uint8_t* q_use;
  const double* q_xy; const okb_keypoint_t* q_kp;
  size_t q_stride, proj_stride; // per-frame strides (elements) for the batched device form
  int nc; const uint8_t* c_desc; const int32_t* c_lm; const double* lm_proj; const uint8_t* lm_is3d;
  double thr_sq; uint32_t thr;
  double thr_px, min_x, min_y, max_x, max_y;   // gate radius and extent of the keypoint cloud (row pre-filter)
  int cell, gx, gy;             // grid
  int32_t* row_off;             // [frames][kMaxCells + 2]: cell offsets into row_list; cell gx*gy = rows with NaN projections
  int32_t* row_list;            // [frames][nc]
  double2* row_xy;              // [frames][nc]: projection of the row's landmark, in row_list order
  uint32_t* out_dist; int32_t* out_idx;
};

__device__ __forceinline__ void m1_kp_xy(const M1Args& a, size_t fq, int k, double& x, double& y)
{
  if (a.q_kp) { x = (double)a.q_kp[fq + k].x; y = (double)a.q_kp[fq + k].y; }  // MultiFrame::getKeypoint: float -> double
  else { x = a.q_xy[2 * k]; y = a.q_xy[2 * k + 1]; }
}
__device__ __forceinline__ int m1_cell_of(const M1Args& a, double x, double y, int& cx, int& cy)
{
  const double fx = floor((x - a.min_x) / a.cell), fy = floor((y - a.min_y) / a.cell);
  cx = fx < 0.0 ? 0 : (fx > (double)(a.gx - 1) ? a.gx - 1 : (int)fx);
  cy = fy < 0.0 ? 0 : (fy > (double)(a.gy - 1) ? a.gy - 1 : (int)fy);
  return cy * a.gx + cx;
}
// cell of a pool row for this frame from its landmark's projection: -1 = cannot match (farther than thr outside the keypoint
// extent), n_cells = NaN projection (passes the reference's gate against every keypoint)
__device__ __forceinline__ int m1_cell_from(const M1Args& a, double px, double py)
{
  if (!(px == px) || !(py == py)) return a.gx * a.gy;
  const double m = a.thr_px + 1.0;
  if (!(px >= a.min_x - m && px <= a.max_x + m && py >= a.min_y - m && py <= a.max_y + m)) return -1;   // also +-inf
  int cx, cy;
  return m1_cell_of(a, px, py, cx, cy);
}
// cells and projections of four pool rows (r0, r0 + stride, ...): the three dependent loads of a row (landmark index -> is3d ->
// projection) are issued for all four rows before any is consumed
__device__ __forceinline__ void m1_rows4(const M1Args& a, int frame, int r0, int stride, int (&cell)[4], double (&px)[4], double (&py)[4])
{
  int lm[4]; bool ok[4];
#pragma unroll
  for (int u = 0; u < 4; u++) { const int r = r0 + u * stride; lm[u] = r < a.nc ? __ldg(&a.c_lm[r]) : -1; }
#pragma unroll
  for (int u = 0; u < 4; u++) ok[u] = lm[u] >= 0 && a.lm_is3d[lm[u]] != 0;   // not 3-D: cannot match
  const double* lp = a.lm_proj + (size_t)frame * a.proj_stride;
#pragma unroll
  for (int u = 0; u < 4; u++) { px[u] = ok[u] ? lp[2 * (size_t)lm[u]] : 0.0; py[u] = ok[u] ? lp[2 * (size_t)lm[u] + 1] : 0.0; }
#pragma unroll
  for (int u = 0; u < 4; u++) cell[u] = ok[u] ? m1_cell_from(a, px[u], py[u]) : -1;
}

__global__ void __launch_bounds__(1024) k_m1_rowbin(M1Args a)
{
  __shared__ int cnt[kMaxCells + 2];
  const int frame = blockIdx.x;
  const int n_cells = a.gx * a.gy;
  for (int i = threadIdx.x; i <= n_cells + 1; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
#pragma unroll 1
  for (int r0 = threadIdx.x; r0 < a.nc; r0 += 4 * 1024) {
    int c[4]; double px[4], py[4];
    m1_rows4(a, frame, r0, 1024, c, px, py);
#pragma unroll
    for (int u = 0; u < 4; u++) if (c[u] >= 0) atomicAdd(&cnt[c[u]], 1);
  }
  __syncthreads();
  // exclusive scan of the cell counts (<= 4097 cells): one warp, a contiguous chunk per lane
  if (threadIdx.x < 32) {
    const int n = n_cells + 1;
    const int per = (n + 31) / 32, beg = min((int)threadIdx.x * per, n), end = min(beg + per, n);
    int sum = 0;
    for (int i = beg; i < end; i++) sum += cnt[i];
    int incl = sum;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += t; }
    int run = incl - sum;
    for (int i = beg; i < end; i++) { const int c = cnt[i]; cnt[i] = run; run += c; }
    if (threadIdx.x == 31) cnt[n] = incl;
  }
  __syncthreads();
  int32_t* off = a.row_off + (size_t)frame * (kMaxCells + 2);
  for (int i = threadIdx.x; i <= n_cells + 1; i += blockDim.x) off[i] = cnt[i];
  __syncthreads();
  int32_t* list = a.row_list + (size_t)frame * a.nc;
  double2* xy = a.row_xy + (size_t)frame * a.nc;
#pragma unroll 1
  for (int r0 = threadIdx.x; r0 < a.nc; r0 += 4 * 1024) {
    int c[4]; double px[4], py[4];
    m1_rows4(a, frame, r0, 1024, c, px, py);
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (c[u] >= 0) { const int pos = atomicAdd(&cnt[c[u]], 1); list[pos] = r0 + u * 1024; xy[pos] = make_double2(px[u], py[u]); }
  }
}

template <int D16>
__global__ void __launch_bounds__(256) k_m1_match(M1Args a)
{
  const int frame = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + warp;
  if (k >= a.nq) return;
  const size_t fq = (size_t)frame * a.q_stride;
  const int nq = a.q_count ? min(a.q_count[frame], a.nq) : a.nq;
  unsigned long long best = ((unsigned long long)a.thr << 32) | 0xffffffffull;
  double x = 0, y = 0;
  bool active = k < nq && (a.q_use == nullptr || a.q_use[k]) && a.nc > 0;
  if (active) { m1_kp_xy(a, fq, k, x, y); active = x == x && y == y; }
  if (active) {
    const int32_t* off = a.row_off + (size_t)frame * (kMaxCells + 2);
    const int32_t* list = a.row_list + (size_t)frame * a.nc;
    const double2* xy = a.row_xy + (size_t)frame * a.nc;
    uint4 qd[D16];
    load_query<D16>(a.q_desc, (int)(fq + k), qd);
    int cx, cy;
    m1_cell_of(a, x, y, cx, cy);
    const int cx0 = max(cx - 1, 0), cx1 = min(cx + 1, a.gx - 1), n_cells = a.gx * a.gy;
    for (int pass = 0; pass < 4; pass++) {
      int beg, end;
      if (pass < 3) {
        const int cyy = cy - 1 + pass;
        if (cyy < 0 || cyy >= a.gy) continue;
        beg = off[cyy * a.gx + cx0]; end = off[cyy * a.gx + cx1 + 1];   // cells of one grid row are contiguous
      } else { beg = off[n_cells]; end = off[n_cells + 1]; }             // NaN projections: the gate never rejects them
      for (int i = beg + lane; i < end; i += 32) {
        const double2 p = __ldg(&xy[i]);
        const double dx = p.x - x, dy = p.y - y;
        const double d2 = dx * dx + dy * dy;
        if (d2 > a.thr_sq) continue;
        const int r = __ldg(&list[i]);
        const uint4* cp = reinterpret_cast<const uint4*>(a.c_desc) + (size_t)r * D16;
        uint32_t d = 0;
#pragma unroll
        for (int w = 0; w < D16; w++) {
          const uint4 cv = __ldg(cp + w);
          d += __popcll(((unsigned long long)(qd[w].x ^ cv.x) << 32) | (qd[w].y ^ cv.y));
          d += __popcll(((unsigned long long)(qd[w].z ^ cv.z) << 32) | (qd[w].w ^ cv.w));
        }
        const unsigned long long key = ((unsigned long long)d << 32) | (unsigned)r;
        if (key < best) best = key;
      }
    }
  }
  // warp minimum of the 64-bit keys (inactive warps keep the "no match" key)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    if (other < best) best = other;
  }
  if (lane == 0) {
    const uint32_t d = (uint32_t)(best >> 32);
    if (d < a.thr) { a.out_dist[fq + k] = d; a.out_idx[fq + k] = a.c_lm[(uint32_t)best]; }
    else { a.out_dist[fq + k] = a.thr; a.out_idx[fq + k] = -1; }
  }
}

static void m1_grid(double thr, double ext_x, double ext_y, int& cell, int& gx, int& gy)
{
  cell = (int)ceil(thr); if (cell < 8) cell = 8;
  for (;;) {
    gx = (int)(ext_x / cell) + 1; gy = (int)(ext_y / cell) + 1;
    if ((long long)gx * gy <= kMaxCells) break;
    cell *= 2;
  }
}

// bin_stream: the row binning only needs the pool and the projections, not the keypoints; a caller that has them early runs it
// on a side stream (bin_done orders the match kernel behind it)
static int m1_launch(okb_context* ctx, M1Args& a, int D, int n_frames, cudaStream_t st, cudaStream_t bin_stream = nullptr, cudaEvent_t bin_done = nullptr)
{
  if (a.nc > 0) {
    k_m1_rowbin<<<n_frames, 1024, 0, bin_stream ? bin_stream : st>>>(a);
    if (bin_stream) { OKB_CUDA(cudaEventRecord(bin_done, bin_stream)); OKB_CUDA(cudaStreamWaitEvent(st, bin_done, 0)); }
  }
  if (D == 64) k_m1_match<4><<<dim3((a.nq + 7) / 8, n_frames), 256, 0, st>>>(a);
  else k_m1_match<3><<<dim3((a.nq + 7) / 8, n_frames), 256, 0, st>>>(a);
  ctx->launches += a.nc > 0 ? 2 : 1;
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

// M5: per landmark (group of descriptors) the best keypoint over (descriptor, k) order
template <int D16>
__global__ void __launch_bounds__(256) k_match_place(int n_lm, const int32_t* lm_offsets, const uint8_t* lm_desc, int n_kp,
                                                     const uint8_t* kp_desc, uint32_t thr, int32_t* out_k, uint32_t* out_dist)
{
  __shared__ uint4 s_desc[D16][kTile];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + warp;
  const bool active = q < n_lm;
  const int d_beg = active ? lm_offsets[q] : 0, d_end = active ? lm_offsets[q + 1] : 0;
  int max_nd = d_end - d_beg;
  // all warps of the CTA must run the same number of descriptor rounds (tiles are staged cooperatively)
  __shared__ int s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  if (lane == 0) atomicMax(&s_max, max_nd);
  __syncthreads();
  max_nd = s_max;
  unsigned long long best = ((unsigned long long)thr << 32);
  for (int tile0 = 0; tile0 < n_kp; tile0 += kTile) {
    __syncthreads();
    stage_tile<D16>(kp_desc, n_kp, tile0, s_desc);
    __syncthreads();
    for (int dd = 0; dd < max_nd; dd++) {
      if (d_beg + dd < d_end) {
        uint4 qd[D16];
        load_query<D16>(lm_desc, d_beg + dd, qd);
        for (int j = 0; j < kTile / 32; j++) {
          const int ci = j * 32 + lane, k = tile0 + ci;
          if (k < n_kp) {
            const uint32_t d = hamming<D16>(qd, s_desc, ci);
            // order position: descriptor-major, then k (fits: dd < 2^12, k < 2^20)
            const unsigned long long key = ((unsigned long long)d << 32) | ((unsigned)dd << 20) | (unsigned)k;
            if (key < best) best = key;
          }
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o); if (t < best) best = t; }
  if (lane == 0) {
    const uint32_t d = (uint32_t)(best >> 32);
    if (d < thr) { out_dist[q] = d; out_k[q] = (int)(best & 0xfffffu); } else { out_dist[q] = thr; out_k[q] = -1; }
  }
}

// M2 / M3 / M4: sequential-order replay with gates
template <int D16, int MODE>
__global__ void __launch_bounds__(256, 4) k_match_gated(MatchArgs a)
{
  __shared__ uint4 s_desc2[2][D16][kTile];
  if (a.hit_cnt && a.hit_cnt[blockIdx.y] <= a.hit_cap) return;   // this frame was handled by the scan/gate split
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 8 + warp;
  // batched device form: blockIdx.y = frame, per-frame strides (elements) and counts; all zero/null for the host form
  const size_t fq = (size_t)blockIdx.y * a.q_stride, fc = (size_t)blockIdx.y * a.c_stride;
  const M3View* view = (MODE == MODE_M3 && a.views) ? &a.views[(size_t)blockIdx.y * a.view_stride + a.view_index] : nullptr;
  const int nq = view ? min(view->n, a.nq) : (a.q_count ? min(a.q_count[blockIdx.y], a.nq) : a.nq);
  const int nc = a.c_count ? min(a.c_count[blockIdx.y], a.nc) : a.nc;
  const uint8_t* c_desc = a.c_desc + fc * (D16 * 16);
  const bool active = q < nq && (a.q_use == nullptr || a.q_use[fq + q]);
  uint4 qd[D16];
  V3 eq = V3{0, 0, 0};
  double q_c26 = 0, q_c6 = 0, q_sof = 0;
  int prev_lm = -1;
  if (active) {
    if (view) load_query<D16>(view->desc, q, qd); else load_query<D16>(a.q_desc, (int)(fq + q), qd);
    eq = v3(a.q_e + 3 * (fq + q));
    if (MODE == MODE_M2) { q_c26 = a.cos26; q_c6 = a.cos6; if (a.q_prev_lm) prev_lm = a.q_prev_lm[fq + q]; }
    else { q_c26 = a.q_cos26[fq + q]; q_c6 = a.q_cos6[fq + q]; q_sof = a.q_sof[fq + q]; }
  }
  uint32_t best = a.thr;
  int best_idx = -1;
  V3 best_hp = V3{0, 0, 0}; bool have_hp = false; bool best_init = false;
  int ctr = 0, skip_lm = -1;
  V3 r0 = v3(a.r0), r1 = v3(a.r1);
  const double* Tcw0 = nullptr; const double* Tcw1 = nullptr;   // M3 sequence: (C, r) poses of this frame
  if (view) { r0 = v3(view->Twc + 9); r1 = v3(a.frames[blockIdx.y].Twc + 9); Tcw0 = view->Tcw; Tcw1 = a.frames[blockIdx.y].Tcw; }
  if (MODE == MODE_M2 && a.frames) r1 = v3(a.frames[blockIdx.y].Twc + 9);   // batched device form: the camera position of this frame
  const int n_tiles = (nc + kTile - 1) / kTile;
  if (n_tiles > 0) stage_tile_async<D16>(c_desc, nc, 0, s_desc2[0]);
  for (int tt = 0; tt < n_tiles; tt++) {
    const int tile0 = tt * kTile;
    uint4 (*s_desc)[kTile] = s_desc2[tt & 1];
    if (tt + 1 < n_tiles) { stage_tile_async<D16>(c_desc, nc, tile0 + kTile, s_desc2[(tt & 1) ^ 1]); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    for (int j = 0; active && j < kTile / 32; j++) {
      const int ci = j * 32 + lane, c = tile0 + ci;
      if (tile0 + j * 32 >= nc) break;
      int lm = -1;
      bool ok = c < nc;
      if (MODE == MODE_M2 && ok) { lm = a.c_lm[fc + c]; ok = !a.lm_is3d[lm]; }
      const uint32_t d = hamming_bounded<D16>(qd, s_desc, ci, ok, best);
      if (MODE != MODE_M2) {
        // M3 / M4: the gate is a pure function of the pair, so the sequential loop's result is "the smallest distance
        // among the candidates that pass, first index on ties". Every lane whose distance beats the running best
        // evaluates the gate of ITS candidate (the loads and the fp64 triangulation of up to 32 candidates overlap),
        // then the warp takes the minimum of (distance, lane) over the lanes that passed.
        bool pass = false, parallel = false;
        V3 hp = V3{0, 0, 0};
        if (d < best && a.c_valid[fc + c]) {
          const V3 e1 = v3(a.c_e + 3 * (fc + c));
          double c26 = q_c26, c6 = q_c6;
          if (MODE == MODE_M4) { const double s1 = a.c_sof[fc + c]; if (q_sof < s1) { c26 = a.c_cos26[fc + c]; c6 = a.c_cos6[fc + c]; } }
          if (MODE != MODE_M3 || !(dot(eq, e1) < 0.5)) {
            const Tri t = triangulate_fast(r0, eq, r1, e1, c26, c6);
            pass = t.valid; parallel = t.parallel; hp = t.p;
            if (MODE == MODE_M3) {
              if (pass) {
                if (dot(eq, e1) < 0.8) pass = false;
                if (!parallel) {
                  if ((Tcw0 ? depth_cr(Tcw0, hp) : depth_in(a.T0, hp)) < 0.2) pass = false;
                  if ((Tcw1 ? depth_cr(Tcw1, hp) : depth_in(a.T1, hp)) < 0.2) pass = false;
                }
              }
            } else {
              if (!parallel) {
                if (depth_in(a.T0, hp) < 0.05) pass = false;
                if (depth_in(a.T1, hp) < 0.05) pass = false;
                if (dot(eq, e1) < 0.8) pass = false;
              }
            }
          }
        }
        const unsigned key = pass ? ((d << 5) | (unsigned)lane) : 0xffffffffu;
        const unsigned win = __reduce_min_sync(0xffffffffu, key);
        if (win != 0xffffffffu) {
          const int wl = (int)(win & 31u);
          best = win >> 5; best_idx = tile0 + j * 32 + wl;
          best_hp.x = __shfl_sync(0xffffffffu, hp.x, wl); best_hp.y = __shfl_sync(0xffffffffu, hp.y, wl);
          best_hp.z = __shfl_sync(0xffffffffu, hp.z, wl);
          have_hp = true; best_init = !__shfl_sync(0xffffffffu, (int)parallel, wl);
        }
        continue;
      }
      unsigned m = __ballot_sync(0xffffffffu, d < best);
      while (m) {
        const int jl = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t dj = __shfl_sync(0xffffffffu, d, jl);
        const int cj = tile0 + j * 32 + jl;
        if (!(dj < best)) continue;
        int lmj = -1;
        bool pass = true, parallel = false;
        V3 hp = V3{0, 0, 0};
        if (MODE == MODE_M2) {
          lmj = __shfl_sync(0xffffffffu, lm, jl);
          if (lmj == skip_lm) continue;
          const V3 e0 = v3(a.c_e + 3 * (fc + cj)), rr0 = v3(a.c_r + 3 * (fc + cj));
          if (dot(e0, eq) < q_c6) {
            const V3 et = normalized(r1 - rr0);
            const V3 n0 = normalized(cross(e0, et));
            const V3 n1 = normalized(cross(eq, et));
            if (dot(n0, n1) < q_c6) pass = false;
            else if (dot(cross(e0, eq), normalized(n0 + n0)) > 0.0) pass = false;
          }
          if (pass) {
            const Tri t = triangulate_fast(rr0, e0, r1, eq, q_c26, q_c6);
            pass = t.valid; parallel = t.parallel; hp = t.p;
            if (pass && !parallel) {
              if (norm(hp - rr0) < 0.2) pass = false;
              if (norm(hp - r1) < 0.2) pass = false;
            }
          }
          if (pass && lmj == prev_lm && prev_lm >= 0) { ctr++; skip_lm = lmj; continue; }
        } else {
          if (!a.c_valid[fc + cj]) continue;
          const V3 e1 = v3(a.c_e + 3 * (fc + cj));
          double c26 = q_c26, c6 = q_c6;
          if (MODE == MODE_M4) { const double s1 = a.c_sof[fc + cj]; if (q_sof < s1) { c26 = a.c_cos26[fc + cj]; c6 = a.c_cos6[fc + cj]; } }
          if (MODE == MODE_M3) { if (dot(eq, e1) < 0.5) continue; }
          const Tri t = triangulate_fast(r0, eq, r1, e1, c26, c6);
          pass = t.valid; parallel = t.parallel; hp = t.p;
          if (MODE == MODE_M3) {
            if (!pass) continue;
            if (dot(eq, e1) < 0.8) pass = false;
            if (!parallel) {
              if (depth_in(a.T0, hp) < 0.2) pass = false;
              if (depth_in(a.T1, hp) < 0.2) pass = false;
            }
          } else {
            if (!parallel) {
              if (depth_in(a.T0, hp) < 0.05) pass = false;
              if (depth_in(a.T1, hp) < 0.05) pass = false;
              if (dot(eq, e1) < 0.8) pass = false;
            }
          }
        }
        if (!pass) continue;
        best = dj; best_idx = (MODE == MODE_M2) ? lmj : cj;
        if (MODE == MODE_M2) { if (!parallel) { best_hp = hp; have_hp = true; } }
        else { best_hp = hp; have_hp = true; best_init = !parallel; }
        m &= __ballot_sync(0xffffffffu, d < best);
      }
    }
    __syncthreads();  // everyone is done with this buffer before the next prefetch overwrites it
  }
  if (q >= a.nq) return;
  if (lane == 0) {
    if (MODE == MODE_M3 && view && best_idx < 0) best = a.thr;
    a.out_dist[fq + q] = best; a.out_idx[fq + q] = best_idx;
    double* hp = a.out_hp + 4 * (fq + q);
    if (have_hp) { hp[0] = best_hp.x; hp[1] = best_hp.y; hp[2] = best_hp.z; hp[3] = 1.0; }
    else { hp[0] = hp[1] = hp[2] = hp[3] = 0.0; }
    if (a.out_init) a.out_init[fq + q] = best_init ? 1 : 0;
    if (MODE == MODE_M2 && a.out_ctr && ctr) atomicAdd(a.out_ctr + (a.frames ? blockIdx.y : 0), ctr);   // per frame in the batched form
  }
}

// ---- M4, device-resident batched form, as scan + gate ------------------------------------------------------------------
// The gate of M4 is a pure function of the pair, so the loop's result per query is min (distance, candidate index) over
// the candidates with distance < threshold that pass the gate. Pairs below the threshold are rare (a handful per query),
// which splits the work into
//   k_m4_scan    Hamming only, register blocked: a thread keeps ONE candidate descriptor in registers, the queries of the
//                frame stream through shared memory (broadcast LDS); the better-discriminating half of the descriptor is
//                popcounted first and the other half only when some lane of the warp is still below the threshold. Hits
//                (distance < threshold) are appended to a per-frame list. ~30 instructions per (warp, query), 48 registers.
//   k_m4_gate    one thread per hit: the reference's gate (triangulateFast etc., fp64), 64-bit atomicMin of
//                (distance << 32 | candidate) per query = first minimum in candidate order;
//   k_m4_finish  one thread per query: outputs; the triangulated point of the winner is recomputed (same function, same
//                inputs, same result).
// A frame whose list overflows (hit_cnt > hit_cap; pathological inputs) is redone by the sequential-replay kernel
// k_match_gated, which is launched behind and returns at once for all other frames.
struct M4Gate { bool pass, parallel; V3 hp; };
__device__ __forceinline__ M4Gate m4_gate(const MatchArgs& a, size_t fq, size_t fc, int q, int c)
{
  M4Gate g; g.pass = false; g.parallel = false; g.hp = V3{0, 0, 0};
  if (!a.c_valid[fc + c]) return g;
  const V3 eq = v3(a.q_e + 3 * (fq + q)), e1 = v3(a.c_e + 3 * (fc + c));
  double c26 = a.q_cos26[fq + q], c6 = a.q_cos6[fq + q];
  const double s1 = a.c_sof[fc + c];
  if (a.q_sof[fq + q] < s1) { c26 = a.c_cos26[fc + c]; c6 = a.c_cos6[fc + c]; }
  const Tri t = triangulate_fast(v3(a.r0), eq, v3(a.r1), e1, c26, c6);
  g.pass = t.valid; g.parallel = t.parallel; g.hp = t.p;
  if (!g.parallel) {
    if (depth_in(a.T0, g.hp) < 0.05) g.pass = false;
    if (depth_in(a.T1, g.hp) < 0.05) g.pass = false;
    if (dot(eq, e1) < 0.8) g.pass = false;
  }
  return g;
}

// the gate of the M3 worker loop for one (older keypoint, current keypoint) pair (Frontend.cpp:1849-1884)
__device__ __forceinline__ M4Gate m3_gate(const MatchArgs& a, size_t fq, size_t fc, int q, int c, int vi)
{
  M4Gate g; g.pass = false; g.parallel = false; g.hp = V3{0, 0, 0};
  if (!a.c_valid[fc + c]) return g;
  const M3View& view = a.views[(size_t)(fc / a.c_stride) * a.view_stride + vi];
  const M3Frame& fr = a.frames[fc / a.c_stride];
  const V3 eq = v3(a.q_e + 3 * (fq + q)), e1 = v3(a.c_e + 3 * (fc + c));
  if (dot(eq, e1) < 0.5) return g;
  const Tri t = triangulate_fast(v3(view.Twc + 9), eq, v3(fr.Twc + 9), e1, a.q_cos26[fq + q], a.q_cos6[fq + q]);
  g.pass = t.valid; g.parallel = t.parallel; g.hp = t.p;
  if (g.pass) {
    if (dot(eq, e1) < 0.8) g.pass = false;
    if (!g.parallel) {
      if (depth_cr(view.Tcw, g.hp) < 0.2) g.pass = false;
      if (depth_cr(fr.Tcw, g.hp) < 0.2) g.pass = false;
    }
  }
  return g;
}
// fq: first query slot of the (frame, view); vi: the view (M3), default the one of the arguments
template <int MODE>
__device__ __forceinline__ M4Gate pair_gate(const MatchArgs& a, size_t fq, size_t fc, int q, int c, int vi = -1)
{
  if (MODE == MODE_M3) return m3_gate(a, fq, fc, q, c, vi < 0 ? a.view_index : vi);
  return m4_gate(a, fq, fc, q, c);
}

template <int D16>
__global__ void __launch_bounds__(256) k_m4_scan(MatchArgs a, uint2* hits, int32_t* hit_cnt)
{
  constexpr int kQT = 128, H = D16 / 2;
  __shared__ uint4 sq[kQT][D16];
  __shared__ uint8_t s_act[kQT];
  const int frame = blockIdx.y;
  const int vz = a.scan_views > 0 ? blockIdx.z / a.scan_chunks : 0, qz = a.scan_views > 0 ? blockIdx.z % a.scan_chunks : blockIdx.z;
  const int qt = a.scan_qt > 0 ? a.scan_qt : kQT;
  const size_t fq = (size_t)frame * a.q_stride + (size_t)vz * a.nq, fc = (size_t)frame * a.c_stride;
  const M3View* view = a.views ? &a.views[(size_t)frame * a.view_stride + a.view_index + vz] : nullptr;   // M3: queries = an older view
  hits += (size_t)vz * gridDim.y * a.hit_cap; hit_cnt += (size_t)vz * gridDim.y;
  const int nq = view ? min(view->n, a.nq) : min(a.q_count[frame], a.nq), nc = min(a.c_count[frame], a.nc);
  const uint4* q_desc = view ? reinterpret_cast<const uint4*>(view->desc) : reinterpret_cast<const uint4*>(a.q_desc) + fq * D16;
  if ((int)(blockIdx.x * blockDim.x) >= nc) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid_c = c < nc && a.c_valid[fc + c];
  uint4 cd[D16];
#pragma unroll
  for (int w = 0; w < D16; w++) cd[w] = valid_c ? __ldg(reinterpret_cast<const uint4*>(a.c_desc) + (fc + c) * D16 + w) : make_uint4(0, 0, 0, 0);
  {
    const int q0 = qz * qt;   // one chunk of queries per CTA: (candidate tile, frame, query chunk) fills the GPU
    if (q0 >= nq) return;
    for (int i = threadIdx.x; i < qt * D16; i += blockDim.x) {
      const int j = i / D16, w = i % D16;
      sq[j][w] = q0 + j < nq ? __ldg(q_desc + (size_t)(q0 + j) * D16 + w) : make_uint4(0, 0, 0, 0);
    }
    if ((int)threadIdx.x < qt) s_act[threadIdx.x] = (q0 + (int)threadIdx.x < nq && (a.q_use == nullptr || a.q_use[fq + q0 + threadIdx.x])) ? 1 : 0;
    __syncthreads();
    const int jn = min(qt, nq - q0);
    for (int j = 0; j < jn; j++) {
      if (!s_act[j]) continue;   // CTA-uniform
      // popcount of the 256 bits of the second descriptor half through a carry-save adder tree: 8 words -> ones / twos /
      // fours / eights planes (14 LOP3 on the ALU pipe), then 4 POPC instead of 8 -- POPC (XU pipe) is this kernel's bound
      static_assert(D16 == 4, "the adder tree below is written for 2 x uint4 per half");
      uint32_t d;
      {
        const uint4 q0v = sq[j][H], q1v = sq[j][H + 1];
        const uint32_t x0 = q0v.x ^ cd[H].x, x1 = q0v.y ^ cd[H].y, x2 = q0v.z ^ cd[H].z, x3 = q0v.w ^ cd[H].w;
        const uint32_t x4 = q1v.x ^ cd[H + 1].x, x5 = q1v.y ^ cd[H + 1].y, x6 = q1v.z ^ cd[H + 1].z, x7 = q1v.w ^ cd[H + 1].w;
        const uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 | x1));
        const uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 | x4));
        const uint32_t s3 = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 | s2));
        const uint32_t ones = s3 ^ x7, c4 = s3 & x7;
        const uint32_t t1 = c1 ^ c2 ^ c3, f1 = (c1 & c2) | (c3 & (c1 | c2));
        const uint32_t twos = t1 ^ c4, f2 = t1 & c4;
        const uint32_t fours = f1 ^ f2, eights = f1 & f2;
        d = __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
      }
      if (!valid_c) d = 0xffffu;
      if (!__any_sync(0xffffffffu, d < a.thr)) continue;
#pragma unroll
      for (int w = 0; w < H; w++) {
        const uint4 qv = sq[j][w];
        d += __popcll(((unsigned long long)(qv.x ^ cd[w].x) << 32) | (qv.y ^ cd[w].y));
        d += __popcll(((unsigned long long)(qv.z ^ cd[w].z) << 32) | (qv.w ^ cd[w].w));
      }
      if (valid_c && d < a.thr) {
        const int pos = atomicAdd(&hit_cnt[frame], 1);
        if (pos < a.hit_cap) hits[(size_t)frame * a.hit_cap + pos] = make_uint2((d << 20) | (uint32_t)(q0 + j), (uint32_t)c);   // d <= 512, q < 2^20
      }
    }
  }
}

// ---- the Hamming scan on the tensor cores (legacy integer path: mma.sync.m16n8k32.u8 -> IMMA.16832.U8.U8) -------------------
// Hamming(a, b) = popc(a) + popc(b) - 2 popc(a & b), and popc(a & b) over 512 bits is a dot product of 0/1 vectors: 16 IMMAs of
// k = 32 per 16 x 8 tile of pairs, the operands being the descriptor words with one bit plane selected per byte: the candidate
// side keeps the bit where it is, w & (0x01010101 << i) (one LOP3: byte value 2^i), the query side moves it to bit 7 - i (byte
// value 2^(7-i), expanded once per 16 queries), so every common bit adds 128 whatever its plane. The expansion stays in registers
// (the same byte lane of the same word on both sides, so any consistent word assignment gives the exact count). bench/ubench_imma.cu measures 0.48 IMMA.16832 per clock per SM on the B200
// (1.14 POPS): 3.8 pairs per clock per SM against 0.97 for the POPC scan above (profiles/popc_rate.json).
// Layout: CTA = 4 warps x 64 candidates (8 column blocks whose raw words stay in registers); the queries of the CTA's chunk are
// staged in shared memory (row stride 20 words: conflict-free fragment loads) and taken 16 at a time: their 64 plane words are
// expanded once and reused for the 8 column blocks; the candidate planes are expanded per use (2 ALU ops per IMMA operand word).
// Epilogue per tile: 4 distances per thread from the row / column popcounts, hits below the threshold appended as in k_m4_scan.
__device__ __forceinline__ void imma_u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) k_scan_mma(MatchArgs a, uint2* hits, int32_t* hit_cnt)
{
  constexpr int kQT = 256, kRow = 20;
  __shared__ uint32_t sq[kQT][kRow];
  __shared__ uint32_t sc[4][64][kRow];   // per warp: 64 candidates x (16 words + popcount)
  __shared__ int s_qid[kQT];
  const int frame = blockIdx.y;
  const int vz = a.scan_views > 0 ? blockIdx.z / a.scan_chunks : 0, qz = a.scan_views > 0 ? blockIdx.z % a.scan_chunks : blockIdx.z;
  const int qt = a.scan_qt;
  const size_t fq = (size_t)frame * a.q_stride + (size_t)vz * a.nq, fc = (size_t)frame * a.c_stride;
  const M3View* view = a.views ? &a.views[(size_t)frame * a.view_stride + a.view_index + vz] : nullptr;   // M3: queries = an older view
  hits += (size_t)vz * gridDim.y * a.hit_cap; hit_cnt += (size_t)vz * gridDim.y;
  const int nq_all = view ? min(view->n, a.nq) : min(a.q_count[frame], a.nq), nc = min(a.c_count[frame], a.nc);
  const int32_t* list = a.q_list ? a.q_list + ((size_t)vz * gridDim.y + frame) * a.nq : nullptr;
  const int nq = list ? min(a.q_list_cnt[(size_t)vz * gridDim.y + frame], nq_all) : nq_all;
  if ((int)(blockIdx.x * 256) >= nc) return;
  const int q0 = qz * qt;
  if (q0 >= nq) return;
  const uint4* q_desc = view ? reinterpret_cast<const uint4*>(view->desc) : reinterpret_cast<const uint4*>(a.q_desc) + fq * 4;
  // ---- stage the chunk's queries (through the eligible list, if there is one)
  for (int i = threadIdx.x; i < qt * 4; i += blockDim.x) {
    const int j = i >> 2, w = i & 3;
    int qi = -1;
    if (q0 + j < nq) qi = list ? __ldg(&list[q0 + j]) : q0 + j;
    const uint4 v = qi >= 0 ? __ldg(q_desc + (size_t)qi * 4 + w) : make_uint4(0, 0, 0, 0);
    sq[j][4 * w] = v.x; sq[j][4 * w + 1] = v.y; sq[j][4 * w + 2] = v.z; sq[j][4 * w + 3] = v.w;
    if (w == 0) s_qid[j] = qi;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int c_base = blockIdx.x * 256 + warp * 64;
  if (c_base >= nc) return;
  // ---- this warp's 64 candidates: raw words and popcounts in shared memory (private to the warp)
  {
    const uint4* cd = reinterpret_cast<const uint4*>(a.c_desc) + fc * 4;
    for (int i = lane; i < 64 * 4; i += 32) {
      const int j = i >> 2, w = i & 3;
      const uint4 v = c_base + j < nc ? __ldg(cd + (size_t)(c_base + j) * 4 + w) : make_uint4(0, 0, 0, 0);
      uint32_t* row = sc[warp][j];
      row[4 * w] = v.x; row[4 * w + 1] = v.y; row[4 * w + 2] = v.z; row[4 * w + 3] = v.w;
    }
    __syncwarp();
    for (int j = lane; j < 64; j += 32) {
      int pc = 0;
#pragma unroll
      for (int w = 0; w < 16; w++) pc += __popc(sc[warp][j][w]);
      sc[warp][j][16] = (uint32_t)pc;
    }
    __syncwarp();
  }
  const int jn = min(qt, nq - q0);
  constexpr uint32_t M = 0x01010101u;
  const uint32_t* b_rows = &sc[warp][g][t];          // + n * 8 * kRow: words t + 4 j of column n * 8 + g
  const uint32_t* b_pops = &sc[warp][2 * t][16];     // + n * 8 * kRow: popcounts of columns n * 8 + 2t, + kRow: 2t + 1
  const int thr64 = 64 * (int)a.thr;
#pragma unroll 1
  for (int m0 = 0; m0 < jn; m0 += 16) {
    // rows m0 + g and m0 + g + 8: words t + 4 j, their 8 bit planes, and the row popcounts
    uint32_t A[2][4][8];   // [row half][word j][plane i]
    int ra[2];             // 64 (popc(row) - thr): the hit test d < thr reads  128 popc(a & b) > 64 (popc(a) + popc(b) - thr)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      int pr = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t w = sq[m0 + g + 8 * h][t + 4 * j];
        pr += __popc(w);
#pragma unroll
        for (int i = 0; i < 8; i++)   // bit i of every byte moved to bit 7 - i: byte value 2^(7-i)
          A[h][j][i] = (i <= 3 ? (w << (7 - 2 * i)) : (w >> (2 * i - 7))) & (M << (7 - i));
      }
      pr += __shfl_xor_sync(0xffffffffu, pr, 1); pr += __shfl_xor_sync(0xffffffffu, pr, 2);
      ra[h] = 64 * pr - thr64;
    }
#pragma unroll 1
    for (int n = 0; n < 8; n++) {
      uint32_t Bw[4];
#pragma unroll
      for (int j = 0; j < 4; j++) Bw[j] = b_rows[n * 8 * kRow + 4 * j];
      // two independent accumulator chains (a single chain of 16 dependent IMMAs is bound by the MMA latency)
      int c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
#pragma unroll
      for (int jp = 0; jp < 2; jp++)   // word pairs (j, j + 1) feed (a0 | a2) / (b0 | b1)
#pragma unroll
        for (int i = 0; i < 8; i++)
          imma_u8((i & 1) ? c1 : c0, A[0][2 * jp][i], A[1][2 * jp][i], A[0][2 * jp + 1][i], A[1][2 * jp + 1][i], Bw[2 * jp] & (M << i), Bw[2 * jp + 1] & (M << i));
      // element e: row g + 8 (e >> 1), column 2t + (e & 1); every common bit contributed 2^(7-i) * 2^i = 128
      const int cb0 = 64 * (int)b_pops[n * 8 * kRow], cb1 = 64 * (int)b_pops[n * 8 * kRow + kRow];
      const int lim[4] = {ra[0] + cb0, ra[0] + cb1, ra[1] + cb0, ra[1] + cb1};
      bool any = false;
#pragma unroll
      for (int e = 0; e < 4; e++) any |= (c0[e] + c1[e]) > lim[e];
      if (any) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int raw = c0[e] + c1[e];
          if (raw <= lim[e]) continue;
          const int h = e >> 1, cc = e & 1;
          const int j = m0 + g + 8 * h, col = c_base + n * 8 + 2 * t + cc;
          if (j < jn && col < nc) {
            const int q = s_qid[j];
            if ((a.q_use == nullptr || a.q_use[fq + q]) && a.c_valid[fc + col]) {
              // d = popc(a) + popc(b) - 2 popc(a & b), from the same integers: (lim + 64 thr - raw) / 64
              const uint32_t d = (uint32_t)((lim[e] + thr64 - raw) >> 6);
              const int pos = atomicAdd(&hit_cnt[frame], 1);
              if (pos < a.hit_cap) hits[(size_t)frame * a.hit_cap + pos] = make_uint2((d << 20) | (uint32_t)q, (uint32_t)col);   // d <= 512, q < 2^20
            }
          }
        }
      }
    }
  }
}

// ---- the Hamming scan on the 5th-generation tensor cores: tcgen05.mma kind::i8, accumulators in TMEM -----------------------
// Same arithmetic as k_scan_mma (popc(a & b) as a dot product of bit planes), but as a 128 x 128 x 512 integer MMA per tile:
// operands are u8 planes in shared memory (K-major, no swizzle: 8-row x 16-byte core matrices; K byte 32 j + 4 i + b <-> bit
// 8 b + i of descriptor word j, value 128 on both sides, so the accumulator is 16384 popc(a & b)), 16 tcgen05.mma of K = 32 per
// tile issued by ONE thread, D (128 lanes x 128 columns of s32) in TMEM, read back with tcgen05.ld for the hit test.
// CTA = 13 warps, one CTA per SM (192 KB of operands). The CTA keeps ONE candidate tile (128 descriptors, expanded once) and streams
// query tiles of 128 through a double-buffered A operand and a double-buffered accumulator, three roles running concurrently:
//   producers (warps 8-11, thread = tile row): expand query tile t into A[t & 1] (wait for the MMAs of tile t - 2: a_free), arrive
//             on a_full;
//   MMA (warp 12, one thread): waits a_full (and acc_free of tile t - 2), 16 MMAs, tcgen05.commit -> a_free and acc_full;
//   epilogue (warps 0-7: TMEM lanes 32 (w % 4).., column half w / 4): waits acc_full, tcgen05.ld 32 columns at a time, running
//             maximum of acc - 8192 popc(b), and only if that exceeds the row's limit 8192 (popc(a) - thr) the columns are looked
//             at one by one (hits are a handful per query); arrive on acc_free.
// so the MMAs of tile t run under the epilogue of tile t - 1 and the expansion of tile t + 1; the roles share nothing but the
// operands, the accumulators and the four barrier pairs (the epilogue recomputes popc(a) from the 64 bytes the producers just read). Every barrier wait is bounded
// (a timeout raises the context's scan-status word, which okb_sync reports, and ends the kernel). bench/umma_probe.cu checks the tile arithmetic stand-alone.
constexpr int kUmmaRows = 128, kUmmaChunk = kUmmaRows * 16, kUmmaTile = 32 * kUmmaChunk;   // 64 KB per operand tile
constexpr int kUmmaSmem = 3 * kUmmaTile + 1024;

struct UmmaRow { uint4 v[4]; };
__device__ __forceinline__ UmmaRow umma_load_row(const uint4* src /* 4 x uint4, or nullptr: zeros */)
{
  UmmaRow w;
#pragma unroll
  for (int i = 0; i < 4; i++) w.v[i] = src ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
  return w;
}
__device__ __forceinline__ int umma_popc_row(const UmmaRow& w)
{
  int pc = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) pc += __popc(w.v[i].x) + __popc(w.v[i].y) + __popc(w.v[i].z) + __popc(w.v[i].w);
  return pc;
}
// descriptor row r (16 words) -> u8 bit planes of the operand tile: K byte 32 j + 4 i + b <-> bit 8 b + i of word j, value 128
__device__ __forceinline__ void umma_expand_row(uint8_t* tile, int r, const UmmaRow& row)
{
#pragma unroll
  for (int q4 = 0; q4 < 4; q4++) {
    const uint32_t w4[4] = {row.v[q4].x, row.v[q4].y, row.v[q4].z, row.v[q4].w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t w = w4[k];
      const int j = 4 * q4 + k;
      uint4 lo, hi;
      lo.x = (w << 7) & 0x80808080u; lo.y = (w << 6) & 0x80808080u; lo.z = (w << 5) & 0x80808080u; lo.w = (w << 4) & 0x80808080u;
      hi.x = (w << 3) & 0x80808080u; hi.y = (w << 2) & 0x80808080u; hi.z = (w << 1) & 0x80808080u; hi.w = w & 0x80808080u;
      uint8_t* p = tile + (size_t)(2 * j) * kUmmaChunk + (r >> 3) * 128 + (r & 7) * 16;
      *reinterpret_cast<uint4*>(p) = lo;
      *reinterpret_cast<uint4*>(p + kUmmaChunk) = hi;
    }
  }
}

constexpr int kUmmaThreads = 13 * 32;   // warps 0-7 epilogue, 8-11 producers, 12 MMA issue + TMEM allocation

// the queries of one view (M3) or of the frame (M4) as a CTA sees them
struct UmmaView { const uint4* q_desc; const int32_t* list; int q_begin, q_end; size_t fq; uint2* hits; int32_t* hit_cnt; };

__global__ void __launch_bounds__(kUmmaThreads, 1) k_scan_umma(MatchArgs a, uint2* hits, int32_t* hit_cnt, int32_t* status)
{
  using namespace okb::umma;
  extern __shared__ uint8_t umma_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(umma_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem; uint8_t* sA0 = smem + kUmmaTile; uint8_t* sA1 = smem + 2 * kUmmaTile;
  __shared__ uint64_t a_full[2], a_free[2], acc_full[2], acc_free[2];
  __shared__ int s_pb[kUmmaRows];       // 8192 popc(candidate)
  __shared__ uint32_t tmem_base;
  __shared__ int failed;
  const int frame = blockIdx.y;
  const size_t fc = (size_t)frame * a.c_stride;
  const int nc = min(a.c_count[frame], a.nc);
  const int c_base = blockIdx.x * kUmmaRows;
  if (c_base >= nc) return;   // whole CTA
  // which views and which part of their query lists this CTA streams against its candidate tile: scan_qt > 0: one view, one chunk of
  // scan_qt queries (blockIdx.z = view * chunks + chunk; small batches, many short CTAs); scan_qt == 0: every view, all queries
  // (blockIdx.z = 0; the candidate tile is expanded once for all of them)
  const int n_views = a.scan_views > 0 ? a.scan_views : 1;
  const bool merged = a.scan_qt == 0;
  const int v_begin = merged ? 0 : (a.scan_views > 0 ? (int)blockIdx.z / a.scan_chunks : 0), v_end = merged ? n_views : v_begin + 1;
  const int qz = merged ? 0 : (a.scan_views > 0 ? (int)blockIdx.z % a.scan_chunks : (int)blockIdx.z);
  auto view_of = [&](int vz) {
    UmmaView V;
    const M3View* view = a.views ? &a.views[(size_t)frame * a.view_stride + a.view_index + vz] : nullptr;   // M3: queries = an older view
    V.fq = (size_t)frame * a.q_stride + (size_t)vz * a.nq;
    V.q_desc = view ? reinterpret_cast<const uint4*>(view->desc) : reinterpret_cast<const uint4*>(a.q_desc) + V.fq * 4;
    V.list = a.q_list ? a.q_list + ((size_t)vz * gridDim.y + frame) * a.nq : nullptr;
    const int nq_all = view ? min(view->n, a.nq) : min(a.q_count[frame], a.nq);
    const int nq = V.list ? min(a.q_list_cnt[(size_t)vz * gridDim.y + frame], nq_all) : nq_all;
    V.q_begin = merged ? 0 : min(qz * a.scan_qt, nq); V.q_end = merged ? nq : min(nq, (qz + 1) * a.scan_qt);
    V.hits = hits + ((size_t)vz * gridDim.y + frame) * a.hit_cap; V.hit_cnt = hit_cnt + (size_t)vz * gridDim.y + frame;
    return V;
  };
  if (!merged && view_of(v_begin).q_begin >= view_of(v_begin).q_end) return;   // whole CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; b++) { bar_init(&a_full[b], 4); bar_init(&a_free[b], 1); bar_init(&acc_full[b], 1); bar_init(&acc_free[b], 8); }
    bar_init_fence();
    failed = 0;
  }
  if (warp == 12) tmem_alloc(&tmem_base, 256);
  if (warp >= 8 && warp < 12) {   // the candidate tile, once per CTA
    const int r = threadIdx.x - 256, col = c_base + r;
    const UmmaRow row = umma_load_row(col < nc ? reinterpret_cast<const uint4*>(a.c_desc) + (fc + col) * 4 : nullptr);
    umma_expand_row(sB, r, row);
    s_pb[r] = 8192 * umma_popc_row(row);
    fence_smem_to_async();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t td = tmem_base;
  if (warp == 12) {
    // ---- MMA issue: one thread
    if (lane == 0) {
      const uint32_t idesc = idesc_u8(kUmmaRows, kUmmaRows);
      const uint32_t b_addr = smem_addr(sB), a_addr[2] = {smem_addr(sA0), smem_addr(sA1)};
      int t = 0; bool ok = true;
#pragma unroll 1
      for (int vz = v_begin; vz < v_end && ok; vz++) {
        const UmmaView V = view_of(vz);
#pragma unroll 1
        for (int j0 = V.q_begin; j0 < V.q_end; j0 += kUmmaRows, t++) {
          const int buf = t & 1;
          if (!bar_wait(&a_full[buf], (uint32_t)((t >> 1) & 1), &failed)) { ok = false; break; }
          if (t >= 2 && !bar_wait(&acc_free[buf], (uint32_t)(((t >> 1) - 1) & 1), &failed)) { ok = false; break; }
          fence_after_sync();
#pragma unroll 1
          for (int s = 0; s < 16; s++)
            mma_u8(td + buf * kUmmaRows, smem_desc(a_addr[buf] + s * 2 * kUmmaChunk, kUmmaChunk, 128), smem_desc(b_addr + s * 2 * kUmmaChunk, kUmmaChunk, 128),
                   idesc, s > 0);
          mma_commit(&a_free[buf]);     // the A buffer may be overwritten ...
          mma_commit(&acc_full[buf]);   // ... and the accumulator read, once these MMAs have completed
        }
      }
    }
  } else if (warp >= 8) {
    // ---- producers: thread = row of the query tile; the next tile's row is in flight while this one is expanded
    const int r = threadIdx.x - 256;
    int t = 0; bool ok = true;
#pragma unroll 1
    for (int vz = v_begin; vz < v_end && ok; vz++) {
      const UmmaView V = view_of(vz);
      // software pipeline over the tiles: the list index is fetched two tiles ahead, the 64-byte row one tile ahead, so that neither
      // of the two dependent loads stalls the expansion
      auto index_of = [&](int j0) { const int j = j0 + r; return j < V.q_end ? (V.list ? __ldg(&V.list[j]) : j) : -1; };
      auto row_of = [&](int qi) { return umma_load_row(qi >= 0 ? V.q_desc + (size_t)qi * 4 : nullptr); };
      UmmaRow cur = row_of(index_of(V.q_begin));
      int qi_next = index_of(V.q_begin + kUmmaRows);
#pragma unroll 1
      for (int j0 = V.q_begin; j0 < V.q_end; j0 += kUmmaRows, t++) {
        const int buf = t & 1;
        const UmmaRow nxt = row_of(qi_next);
        qi_next = index_of(j0 + 2 * kUmmaRows);
        if (t >= 2 && !bar_wait(&a_free[buf], (uint32_t)(((t >> 1) - 1) & 1), &failed)) { ok = false; break; }
        umma_expand_row(buf ? sA1 : sA0, r, cur);
        fence_smem_to_async();
        __syncwarp();
        if (lane == 0) bar_arrive(&a_full[buf]);
        cur = nxt;
      }
    }
  } else {
    // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31 (= tile rows) and the column half w / 4
    const int r = (warp & 3) * 32 + lane, chalf = (warp >> 2) * 64;
    const int thr = (int)a.thr;
    int t = 0; bool ok = true;
#pragma unroll 1
    for (int vz = v_begin; vz < v_end && ok; vz++) {
      const UmmaView V = view_of(vz);
      // this row's query and its popcount, one tile ahead (the producers pull the same 64 bytes through L1 / L2)
      auto index_of = [&](int j0) { const int j = j0 + r; return j < V.q_end ? (V.list ? __ldg(&V.list[j]) : j) : -1; };
      int q_next = index_of(V.q_begin);
      UmmaRow row_next = umma_load_row(q_next >= 0 ? V.q_desc + (size_t)q_next * 4 : nullptr);
#pragma unroll 1
      for (int j0 = V.q_begin; j0 < V.q_end; j0 += kUmmaRows, t++) {
        const int buf = t & 1;
        const int q = q_next;
        const int pa = umma_popc_row(row_next);
        q_next = index_of(j0 + kUmmaRows);
        row_next = umma_load_row(q_next >= 0 ? V.q_desc + (size_t)q_next * 4 : nullptr);
        const int rowlim = 8192 * (pa - thr);   // hit: acc - 8192 popc(b) > rowlim  <=>  popc(a) + popc(b) - 2 popc(a & b) < thr
        if (!bar_wait(&acc_full[buf], (uint32_t)((t >> 1) & 1), &failed)) { ok = false; break; }
        fence_after_sync();
#pragma unroll 1
        for (int half = 0; half < 2; half++) {   // this warp's 64 columns, 32 at a time (128 registers per thread at 13 warps)
          const int c0 = chalf + 32 * half;
          uint32_t v[32];
          tmem_ld_32x32(td + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(buf * kUmmaRows + c0), v);
          tmem_ld_wait();
          // maxima of acc - 8192 popc(b) over groups of 4 columns; a hit in the 32 columns shows in their maximum. Some lane of a
          // warp has one in most chunks (a few hits per query), so what follows the test is kept short: only the groups of 4 that
          // contain a hit are opened
          int g[8];
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const int4 pb = *reinterpret_cast<const int4*>(&s_pb[c0 + k]);
            g[k >> 2] = max(max((int)v[k] - pb.x, (int)v[k + 1] - pb.y), max((int)v[k + 2] - pb.z, (int)v[k + 3] - pb.w));
          }
          const int m = max(max(max(g[0], g[1]), max(g[2], g[3])), max(max(g[4], g[5]), max(g[6], g[7])));
          if (m > rowlim && q >= 0) {
#pragma unroll
            for (int gi = 0; gi < 8; gi++) {
              if (g[gi] <= rowlim) continue;
#pragma unroll
              for (int kk = 0; kk < 4; kk++) {   // unrolled: v stays in registers
                const int k = 4 * gi + kk, col = c_base + c0 + k;
                if ((int)v[k] - s_pb[c0 + k] > rowlim && col < nc && (a.q_use == nullptr || a.q_use[V.fq + q]) && a.c_valid[fc + col]) {
                  const uint32_t d = (uint32_t)(pa + (s_pb[c0 + k] >> 13) - 2 * (int)(v[k] >> 14));
                  const int pos = atomicAdd(V.hit_cnt, 1);
                  if (pos < a.hit_cap) V.hits[pos] = make_uint2((d << 20) | (uint32_t)q, (uint32_t)col);   // d <= 512, q < 2^20
                }
              }
            }
          }
        }
        fence_before_sync();
        __syncwarp();
        if (lane == 0) bar_arrive(&acc_free[buf]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 12) tmem_dealloc(td, 256);
  if (threadIdx.x == 0 && atomicAdd(&failed, 0)) atomicOr(status, 32);   // one word per context, read by okb_sync
}

// one opt-in to the kernel's 193 KB of dynamic shared memory per device
static int umma_prepare()
{
  static std::mutex m; static std::vector<int> done;
  int dev = 0; OKB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(m);
  for (int d : done) if (d == dev) return OKB_OK;
  OKB_CUDA(cudaFuncSetAttribute(k_scan_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, kUmmaSmem));
  done.push_back(dev);
  return OKB_OK;
}

template <int MODE>
__global__ void __launch_bounds__(128) k_m4_gate(MatchArgs a, const uint2* hits, unsigned long long* best)
{
  const int frame = blockIdx.y;
  const int n = min(a.hit_cnt[frame], a.hit_cap);
  if (a.hit_cnt[frame] > a.hit_cap) return;   // redone by k_match_gated
  const size_t fq = (size_t)frame * a.q_stride, fc = (size_t)frame * a.c_stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint2 e = hits[(size_t)frame * a.hit_cap + i];
    const int q = (int)(e.x & 0xfffffu), c = (int)e.y;
    const uint32_t d = e.x >> 20;
    if (pair_gate<MODE>(a, fq, fc, q, c).pass) atomicMin(&best[fq + q], ((unsigned long long)d << 32) | (unsigned)c);
  }
}

template <int MODE>
__device__ __forceinline__ void m4_finish_one(const MatchArgs& a, const unsigned long long b, size_t fq, size_t fc, int q, int vi = -1)
{
  const uint32_t d = (uint32_t)(b >> 32);
  double* hp = a.out_hp + 4 * (fq + q);
  if (d < a.thr) {
    const int c = (int)(uint32_t)b;
    const M4Gate g = pair_gate<MODE>(a, fq, fc, q, c, vi);
    a.out_dist[fq + q] = d; a.out_idx[fq + q] = c;
    hp[0] = g.hp.x; hp[1] = g.hp.y; hp[2] = g.hp.z; hp[3] = 1.0;
    a.out_init[fq + q] = g.parallel ? 0 : 1;
  } else {
    a.out_dist[fq + q] = a.thr; a.out_idx[fq + q] = -1;
    hp[0] = hp[1] = hp[2] = hp[3] = 0.0;
    a.out_init[fq + q] = 0;
  }
}

template <int MODE>
__global__ void __launch_bounds__(128) k_m4_finish(MatchArgs a, const unsigned long long* best)
{
  const int frame = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.nq || a.hit_cnt[frame] > a.hit_cap) return;
  m4_finish_one<MODE>(a, best[(size_t)frame * a.q_stride + q], (size_t)frame * a.q_stride, (size_t)frame * a.c_stride, q);
}


// ---- M3 as a device-resident sequence over the older keyframes (Frontend::matchMotionStereo, Frontend.cpp:1775-1958) ----
// Per older keyframe index v, for all frames of the batch at once:
//   k_m3_prep    world rays / sigma tables of the view's keypoints (e0_W = (C_WC0 e_C).normalized(), cos(2.6 s), cos(6 s) by
//                gate_cos, use = eligible && back-projection valid) and the candidate mask of the current frame
//                (valid && not yet matched: the reference's compacted set k1s, :1789-1801, in the same ascending order);
//   k_match_gated<., M3>  the worker loop (:1809-1895);
//   k_m3_check   the 4 px re-projection check (:1897-1904, PinholeCamera::projectHomogeneous of T_CW1 * hp_W) and the claim
//                of the matched current keypoint: atomicMin of k0 = "first k0 in ascending order wins" (:1915-1954);
//   k_m3_commit  flags (matching / initialisable / inserted) and the update of the matched mask for the next older keyframe.
struct M3Prep {
  const M3View* views; int n_views; const M3Frame* frames;
  int cap0, cap1; size_t q_stride;   // scratch / output stride per frame (keypoints): n_views * cap0
  double f0;
  double* e0; double* c26; double* c6; uint8_t* use0;            // [frames][n_views][cap0]
  const double* rays1; const uint8_t* valid1; const int32_t* count1; uint8_t* matched1; double* e1; uint8_t* cvalid;   // [frames][cap1]
  const int32_t* m1_lm;       // optional [frames][cap1]: M1's landmark per keypoint; when set, matched1 is initialised from it here
  unsigned long long* best;   // [frames][n_views][cap0] -> all ones
  int32_t* claim;             // [n_views][frames][cap1] -> INT_MAX
  int32_t* hit_cnt;           // [n_views][frames] -> 0
  int32_t* q_list; int32_t* q_list_cnt;   // [n_views][frames][cap0] eligible keypoints of the view (any order), [n_views][frames] counts (zeroed by the caller)
};

// one launch for the whole sequence: grid (keypoint tiles, frames, views). Tables of every view's keypoints, and (by the z = 0
// slice) the world rays and the candidate mask of the current frame; resets the per-view reduction arrays
__global__ void __launch_bounds__(128) k_m3_prep(const __grid_constant__ M3Prep p)
{
  const int frame = blockIdx.y, v = blockIdx.z;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const M3View& V = p.views[(size_t)frame * p.n_views + v];
  if (k == 0) p.hit_cnt[(size_t)v * gridDim.y + frame] = 0;
  if (k < p.cap0) {
    const size_t i = (size_t)frame * p.q_stride + (size_t)v * p.cap0 + k;
    bool use = false;
    if (k < V.n) {
      use = V.valid[k] && (V.use == nullptr || V.use[k]);
      const double x = V.rays[3 * (size_t)k], y = V.rays[3 * (size_t)k + 1], z = V.rays[3 * (size_t)k + 2];
      const double* C = V.Twc;
      const V3 w = V3{(C[0] * x + C[1] * y) + C[2] * z, (C[3] * x + C[4] * y) + C[5] * z, (C[6] * x + C[7] * y) + C[8] * z};
      const V3 e = normalized(w);
      p.e0[3 * i] = e.x; p.e0[3 * i + 1] = e.y; p.e0[3 * i + 2] = e.z;
      const double sigma = (double)V.size[k] / p.f0 * 0.125;
      p.c26[i] = gate_cos(2.6 * sigma); p.c6[i] = gate_cos(6.0 * sigma);
    }
    p.use0[i] = use ? 1 : 0;
    p.best[i] = ~0ull;
    if (use) {
      const size_t vf = (size_t)v * gridDim.y + frame;
      p.q_list[vf * p.cap0 + atomicAdd(&p.q_list_cnt[vf], 1)] = k;
    }
  }
  if (k < p.cap1) {
    const size_t j = (size_t)frame * p.cap1 + k;
    p.claim[((size_t)v * gridDim.y + frame) * p.cap1 + k] = 0x7fffffff;
    if (v == 0) {
      const bool in = k < min(p.count1[frame], p.cap1);
      if (in) {
        const double x = p.rays1[3 * j], y = p.rays1[3 * j + 1], z = p.rays1[3 * j + 2];
        const double* C = p.frames[frame].Twc;
        const V3 e = normalized(V3{(C[0] * x + C[1] * y) + C[2] * z, (C[3] * x + C[4] * y) + C[5] * z, (C[6] * x + C[7] * y) + C[8] * z});
        p.e1[3 * j] = e.x; p.e1[3 * j + 1] = e.y; p.e1[3 * j + 2] = e.z;
      }
      // the reference's compacted candidate set k1s (:1789-1801): valid and not yet matched; k_m3_commit / k_m3_view clear the
      // entries a view inserts, so every view sees the set as its predecessors left it
      bool matched;
      if (p.m1_lm) { matched = in && p.m1_lm[j] >= 0; p.matched1[j] = matched ? 1 : 0; }   // the `landmarkId != 0` test of :1792-1795 on M1's result
      else matched = p.matched1[j] != 0;
      p.cvalid[j] = (in && p.valid1[j] && !matched) ? 1 : 0;
    }
  }
}

struct M3Check {
  const M3View* views; int view_stride, view_index; const M3Frame* frames;
  int cap0, cap1; size_t q_stride;
  Model cam; int width, height; uint32_t thr;
  const okb_keypoint_t* kp1;                     // [frames][cap1]
  const int32_t* k1; const uint32_t* dist; const double* hp; const uint8_t* init;   // matcher outputs (offset by view_index * cap0)
  uint8_t* flags; int32_t* claim; uint8_t* matched1; uint8_t* cvalid;
};

// re-projection check and claim of one (frame, k0): returns the flags (bit 0 matching, bit 1 initialisable)
__device__ __forceinline__ uint8_t m3_check_one(const M3Check& c, int frame, size_t i /* slot of (frame, view, k0) */, int k0, int32_t* claim)
{
  uint8_t fl = 0;
  const int k1 = c.k1[i];
  if (k1 >= 0 && c.dist[i] < c.thr) {
    const double* T = c.frames[frame].Tcw;
    const double x = c.hp[4 * i], y = c.hp[4 * i + 1], z = c.hp[4 * i + 2];
    const double px = ((T[0] * x + T[1] * y) + T[2] * z) + T[9] * 1.0;
    const double py = ((T[3] * x + T[4] * y) + T[5] * z) + T[10] * 1.0;
    const double pz = ((T[6] * x + T[7] * y) + T[8] * z) + T[11] * 1.0;
    double kx, ky;
    const int st = project(c.cam, c.width, c.height, px, py, pz, kx, ky);
    const okb_keypoint_t kp = c.kp1[(size_t)frame * c.cap1 + k1];
    const double dx = (double)kp.x - kx, dy = (double)kp.y - ky;
    const bool matching = st == kProjSuccessful && sqrt(dx * dx + dy * dy) < 4.0;
    fl = (uint8_t)((matching ? 1 : 0) | (c.init[i] ? 2 : 0));
    if (matching) atomicMin(&claim[(size_t)frame * c.cap1 + k1], k0);
  }
  return fl;
}

__global__ void __launch_bounds__(128) k_m3_check(const __grid_constant__ M3Check c)
{
  const int frame = blockIdx.y;
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (k0 >= c.cap0) return;
  c.flags[(size_t)frame * c.q_stride + k0] = m3_check_one(c, frame, (size_t)frame * c.q_stride + k0, k0, c.claim);
}

__global__ void __launch_bounds__(128) k_m3_commit(const __grid_constant__ M3Check c)
{
  const int frame = blockIdx.y;
  const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (k0 >= c.cap0) return;
  const size_t i = (size_t)frame * c.q_stride + k0;
  const uint8_t fl = c.flags[i];
  if (!(fl & 1)) return;
  const size_t j = (size_t)frame * c.cap1 + c.k1[i];
  // cvalid excluded the keypoints matched before this view, so every claim is on an unmatched keypoint: the lowest k0 inserts
  if (c.claim[j] == k0) { c.flags[i] = fl | 4; c.matched1[j] = 1; c.cvalid[j] = 0; }
}

// The per-view steps of the sequence (or the one step of a stereo pair) in ONE launch, one CTA per frame: for every older keyframe
// in order -- gate of the view's hit list -> per-query minimum -> outputs -> 4 px check and claim -> commit -- with block barriers
// where the separate kernels have launch boundaries (the reductions go through L2 atomics and are read back with ld.cg). Same
// device functions as k_m4_gate / k_m4_finish / k_m3_check / k_m3_commit. A frame whose hit list overflowed is matched by brute
// force here (warp per query, minimum over (distance, k1) = first minimum in candidate order). The arguments carry the bases of
// view 0; view v lives nq query slots further in every per-query array, gridDim.x hit lists / counters further, claim_stride claim
// slots further.
template <int D16, int MODE>
__global__ void __launch_bounds__(512) k_pair_view(MatchArgs a, const __grid_constant__ M3Check c, const uint2* hits, unsigned long long* best,
                                                   int v_begin, int v_end, size_t claim_stride)
{
  const int frame = blockIdx.x;
  const size_t fc = (size_t)frame * a.c_stride;
#pragma unroll 1
  for (int v = v_begin; v < v_end; v++) {
    const size_t fq = (size_t)frame * a.q_stride + (size_t)v * a.nq;
    const uint2* hits_v = hits + ((size_t)v * gridDim.x + frame) * a.hit_cap;
    const int n_hits = a.hit_cnt[(size_t)v * gridDim.x + frame];
    int32_t* claim = MODE == MODE_M3 ? c.claim + (size_t)v * claim_stride : nullptr;
    if (n_hits <= a.hit_cap) {
      for (int i = threadIdx.x; i < n_hits; i += blockDim.x) {
        const uint2 e = hits_v[i];
        const int q = (int)(e.x & 0xfffffu), cc = (int)e.y;
        if (pair_gate<MODE>(a, fq, fc, q, cc, v).pass) atomicMin(&best[fq + q], ((unsigned long long)(e.x >> 20) << 32) | (unsigned)cc);
      }
    } else {
      const M3View* view = MODE == MODE_M3 ? &a.views[(size_t)frame * a.view_stride + v] : nullptr;
      const int nq = view ? min(view->n, a.nq) : min(a.q_count[frame], a.nq), nc = min(a.c_count[frame], a.nc);
      const uint4* q_desc = view ? reinterpret_cast<const uint4*>(view->desc) : reinterpret_cast<const uint4*>(a.q_desc) + fq * D16;
      const int lane = threadIdx.x & 31;
      for (int q = threadIdx.x >> 5; q < nq; q += blockDim.x >> 5) {
        if (a.q_use && !a.q_use[fq + q]) continue;
        uint4 qd[D16];
#pragma unroll
        for (int w = 0; w < D16; w++) qd[w] = __ldg(q_desc + (size_t)q * D16 + w);
        unsigned long long mine = ~0ull;
        for (int cc = lane; cc < nc; cc += 32) {
          if (!a.c_valid[fc + cc]) continue;
          uint32_t d = 0;
#pragma unroll
          for (int w = 0; w < D16; w++) {
            const uint4 cv = __ldg(reinterpret_cast<const uint4*>(a.c_desc) + (fc + cc) * D16 + w);
            d += __popc(qd[w].x ^ cv.x) + __popc(qd[w].y ^ cv.y) + __popc(qd[w].z ^ cv.z) + __popc(qd[w].w ^ cv.w);
          }
          if (d < a.thr && pair_gate<MODE>(a, fq, fc, q, cc, v).pass) mine = min(mine, ((unsigned long long)d << 32) | (unsigned)cc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
        if (lane == 0) best[fq + q] = mine;
      }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < a.nq; q += blockDim.x) {
      m4_finish_one<MODE>(a, __ldcg(&best[fq + q]), fq, fc, q, v);
      if (MODE == MODE_M3) c.flags[fq + q] = m3_check_one(c, frame, fq + q, q, claim);   // reads the outputs this thread has just written
    }
    if (MODE != MODE_M3) return;
    __syncthreads();
    for (int q = threadIdx.x; q < a.nq; q += blockDim.x) {
      const uint8_t fl = c.flags[fq + q];
      if (!(fl & 1)) continue;
      const size_t j = fc + c.k1[fq + q];
      if (__ldcg(&claim[j]) == q) { c.flags[fq + q] = fl | 4; c.matched1[j] = 1; c.cvalid[j] = 0; }
    }
    __syncthreads();   // the next view's gate sees this view's insertions
  }
}

// matched[b][k] = lm[b][k] >= 0 (the `landmarkId != 0` test of Frontend.cpp:1792-1795 on the M1 result)
__global__ void __launch_bounds__(256) k_matched_mask(const int32_t* lm, const int32_t* count, int cap, uint8_t* matched)
{
  const int frame = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cap) return;
  const size_t i = (size_t)frame * cap + k;
  matched[i] = (k < min(count[frame], cap) && lm[i] >= 0) ? 1 : 0;
}

// ordered compaction of the matching entries (flags bit 0) of one (frame, view): k0 ascending, as the serial insertion
// loop of Frontend.cpp:1915 walks them. One CTA of 256 threads per (view, frame).
__global__ void __launch_bounds__(256) k_m3_compact(int cap0, int n_older, int cap_m, const int32_t* k1, const double* hp, const uint8_t* flags,
                                                    int32_t* n_match, int32_t* m_k0, int32_t* m_k1, uint8_t* m_flags, double* m_hp)
{
  __shared__ int warp_tot[8], base;
  const int v = blockIdx.x, frame = blockIdx.y;
  const size_t in0 = ((size_t)frame * n_older + v) * cap0, out0 = ((size_t)frame * n_older + v) * cap_m;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < cap0; c0 += 256) {
    const int k0 = c0 + threadIdx.x;
    const bool m = k0 < cap0 && (flags[in0 + k0] & 1);
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int before = base;
    for (int w = 0; w < warp; w++) before += warp_tot[w];
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (m && pos < cap_m) {
      m_k0[out0 + pos] = k0; m_k1[out0 + pos] = k1[in0 + k0]; m_flags[out0 + pos] = flags[in0 + k0];
      for (int i = 0; i < 4; i++) m_hp[4 * (out0 + pos) + i] = hp[4 * (in0 + k0) + i];
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += warp_tot[w]; base += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) n_match[(size_t)frame * n_older + v] = base;
}

void m3_compact_launch(int n_frames, int cap0, int n_older, int cap_m, const int32_t* k1, const double* hp, const uint8_t* flags, int32_t* n_match,
                       int32_t* m_k0, int32_t* m_k1, uint8_t* m_flags, double* m_hp, cudaStream_t st)
{
  k_m3_compact<<<dim3(n_older, n_frames), 256, 0, st>>>(cap0, n_older, cap_m, k1, hp, flags, n_match, m_k0, m_k1, m_flags, m_hp);
}

__global__ void __launch_bounds__(256) k_hamming_matrix(int D16, int na, const uint8_t* A, int nb, const uint8_t* B, uint16_t* out)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= nb || i >= na) return;
  const uint4* a = reinterpret_cast<const uint4*>(A) + (size_t)i * D16;
  const uint4* b = reinterpret_cast<const uint4*>(B) + (size_t)j * D16;
  uint32_t d = 0;
  for (int w = 0; w < D16; w++) {
    const uint4 x = __ldg(a + w), y = __ldg(b + w);
    d += __popcll(((unsigned long long)(x.x ^ y.x) << 32) | (x.y ^ y.y)) + __popcll(((unsigned long long)(x.z ^ y.z) << 32) | (x.w ^ y.w));
  }
  out[(size_t)i * nb + j] = (uint16_t)d;
}

// ---------------------------------------------------------------------------------------------------------------
// host side: staging arena
struct Arena {
  okb_context* ctx; size_t off = 0; bool dry = true;
  uint8_t* h = nullptr; uint8_t* d = nullptr;
  template <class T> const T* in(const T* src, size_t n)
  {
    off = (off + 255) & ~(size_t)255;
    const size_t o = off; off += n * sizeof(T);
    if (dry || !src) return nullptr;
    memcpy(h + o, src, n * sizeof(T));
    return reinterpret_cast<const T*>(d + o);
  }
  template <class T> T* out(size_t n, size_t* host_off)
  {
    off = (off + 255) & ~(size_t)255;
    const size_t o = off; off += n * sizeof(T);
    if (host_off) *host_off = o;
    return dry ? nullptr : reinterpret_cast<T*>(d + o);
  }
};

static int ensure(okb_context* ctx, size_t bytes)
{
  MatchWorkspace& m = match_ws(ctx);
  if (bytes <= m.d_cap) return OKB_OK;
  OKB_CUDA(cudaStreamSynchronize(m.stream));
  if (m.d_buf) cudaFree(m.d_buf);
  if (m.h_buf) cudaFreeHost(m.h_buf);
  m.d_buf = m.h_buf = nullptr; m.d_cap = m.h_cap = 0;
  const size_t cap = bytes + bytes / 2 + (1 << 20);
  OKB_CUDA(cudaMalloc(&m.d_buf, cap));
  OKB_CUDA(cudaMallocHost(&m.h_buf, cap));
  m.d_cap = m.h_cap = cap;
  return OKB_OK;
}

MatchWorkspace& match_ws(okb_context* ctx)
{
  static thread_local const okb_context* bound_ctx = nullptr;
  static thread_local int slot = 0;
  if (bound_ctx != ctx) { slot = ctx->next_slot.fetch_add(1) % kMatchSlots; bound_ctx = ctx; }
  return ctx->match_slots[slot];
}

int match_init(okb_context* ctx)
{
  for (int i = 0; i < kMatchSlots; i++) {
    OKB_CUDA(cudaStreamCreateWithFlags(&ctx->match_slots[i].stream, cudaStreamNonBlocking));
    ctx->match_slots[i].mtx = new std::mutex();
  }
  return OKB_OK;
}
void match_free(okb_context* ctx)
{
  for (int i = 0; i < kMatchSlots; i++) {
    MatchWorkspace& w = ctx->match_slots[i];
    if (w.d_buf) cudaFree(w.d_buf);
    if (w.h_buf) cudaFreeHost(w.h_buf);
    if (w.stream) cudaStreamDestroy(w.stream);
    delete w.mtx;
  }
}

static bool bad_D(int D) { return D != 48 && D != 64; }

// host-side validation of a landmark pool (LandmarkToMatch order, Frontend.cpp:1221-1223): cand_lm indexes lm_* and is
// non-decreasing; a bad index would otherwise become an out-of-bounds device read
static int check_pool(int n_cand, const int32_t* cand_lm, int n_lm, const char* who)
{
  int prev = 0;
  for (int c = 0; c < n_cand; c++) {
    const int lm = cand_lm[c];
    if (lm < 0 || lm >= n_lm || lm < prev) {
      set_error("%s: cand_lm[%d] = %d is out of range [0, %d) or decreasing", who, c, lm, n_lm);
      return OKB_ERR_ARGUMENT;
    }
    prev = lm;
  }
  return OKB_OK;
}

}  // namespace okb

using namespace okb;

#define MW (match_ws(ctx))
#define OKB_LOCK_SLOT std::lock_guard<std::mutex> slot_lock__(*MW.mtx)

#define OKB_CHECK_ARGS(cond, who)                                  \
  if (!(cond)) { set_error("%s: bad arguments", who); return OKB_ERR_ARGUMENT; }

// Runs `body(arena)` twice: a dry pass that sizes the arena and a real pass that fills the pinned buffer.
#define OKB_TWO_PASS(ctx, A, BODY)                                                        \
  Arena A; A.ctx = ctx;                                                                   \
  { A.dry = true; A.off = 0; BODY; }                                                      \
  const size_t in_bytes__ = A.off; (void)in_bytes__;

extern "C" {

int okb_match_map3d(okb_context_t* ctx, int D, int n_kp, const uint8_t* kp_desc, const double* kp_xy, const uint8_t* kp_use,
                    int n_cand, const uint8_t* cand_desc, const int32_t* cand_lm, int n_lm, const double* lm_proj,
                    const uint8_t* lm_is3d, double reprojection_threshold, uint32_t match_threshold, uint32_t* out_dist,
                    int32_t* out_lm)
{
  OKB_CHECK_ARGS(ctx && !bad_D(D) && n_kp >= 0 && n_cand >= 0 && n_lm >= 0 && out_dist && out_lm, "okb_match_map3d");
  OKB_CHECK_ARGS(n_kp == 0 || (kp_desc && kp_xy), "okb_match_map3d");
  OKB_CHECK_ARGS(n_cand == 0 || (cand_desc && cand_lm && lm_proj && lm_is3d), "okb_match_map3d");
  OKB_CHECK_ARGS(reprojection_threshold >= 0.0 && reprojection_threshold < 1e6, "okb_match_map3d");
  { int rcv = check_pool(n_cand, cand_lm, n_lm, "okb_match_map3d"); if (rcv) return rcv; }
  if (n_kp == 0) return OKB_OK;
  OKB_CUDA(cudaSetDevice(ctx->device));
  OKB_LOCK_SLOT;
  cudaStream_t st = MW.stream;
  // extent of the keypoint cloud (sizes the grid and the row pre-filter; not part of the arithmetic)
  double min_x = 0, min_y = 0, max_x = 0, max_y = 0; bool first = true;
  for (int k = 0; k < n_kp; k++) {
    const double x = kp_xy[2 * k], y = kp_xy[2 * k + 1];
    if (!(x == x) || !(y == y)) continue;
    if (first) { min_x = max_x = x; min_y = max_y = y; first = false; }
    if (x < min_x) min_x = x; if (x > max_x) max_x = x; if (y < min_y) min_y = y; if (y > max_y) max_y = y;
  }
  if (!(max_x - min_x < 1e6)) { min_x = -5e5; max_x = 5e5; }
  if (!(max_y - min_y < 1e6)) { min_y = -5e5; max_y = 5e5; }
  M1Args a; memset(&a, 0, sizeof(a));
  m1_grid(reprojection_threshold, max_x - min_x, max_y - min_y, a.cell, a.gx, a.gy);
  a.thr_px = reprojection_threshold; a.min_x = min_x; a.min_y = min_y; a.max_x = max_x; a.max_y = max_y;
  size_t o_dist = 0, o_idx = 0, in_end = 0;
  for (int pass = 0; pass < 2; pass++) {
    Arena A; A.ctx = ctx; A.dry = pass == 0; A.h = (uint8_t*)MW.h_buf; A.d = (uint8_t*)MW.d_buf;
    a.q_desc = A.in(kp_desc, (size_t)n_kp * D); a.q_xy = A.in(kp_xy, (size_t)n_kp * 2);
    a.q_use = kp_use ? A.in(kp_use, (size_t)n_kp) : nullptr;
    a.c_desc = A.in(cand_desc, (size_t)n_cand * D); a.c_lm = A.in(cand_lm, (size_t)n_cand);
    a.lm_proj = A.in(lm_proj, (size_t)n_lm * 2); a.lm_is3d = A.in(lm_is3d, (size_t)n_lm);
    in_end = A.off;
    a.row_off = A.out<int32_t>(kMaxCells + 2, nullptr); a.row_list = A.out<int32_t>((size_t)n_cand + 1, nullptr);
    a.row_xy = A.out<double2>((size_t)n_cand + 1, nullptr);
    a.out_dist = A.out<uint32_t>(n_kp, &o_dist); a.out_idx = A.out<int32_t>(n_kp, &o_idx);
    if (pass == 0) { int rc = ensure(ctx, A.off); if (rc) return rc; }
  }
  a.nq = n_kp; a.nc = n_cand; a.thr = match_threshold; a.thr_sq = reprojection_threshold * reprojection_threshold;
  OKB_CUDA(cudaMemcpyAsync(MW.d_buf, MW.h_buf, in_end, cudaMemcpyHostToDevice, st));
  int rc = m1_launch(ctx, a, D, 1, st);
  if (rc) return rc;
  uint8_t* h = (uint8_t*)MW.h_buf; uint8_t* d = (uint8_t*)MW.d_buf;
  OKB_CUDA(cudaMemcpyAsync(h + o_dist, d + o_dist, (o_idx - o_dist) + (size_t)n_kp * 4, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(cudaStreamSynchronize(st));
  memcpy(out_dist, h + o_dist, (size_t)n_kp * 4); memcpy(out_lm, h + o_idx, (size_t)n_kp * 4);
  return OKB_OK;
}

static int run_gated(okb_context_t* ctx, int mode, int D, MatchArgs& a, size_t in_end, size_t o_dist, size_t o_end,
                     size_t o_idx, size_t o_hp, size_t o_init, size_t o_ctr, uint32_t* out_dist, int32_t* out_idx,
                     double* out_hp, uint8_t* out_init, int32_t* out_ctr)
{
  cudaStream_t st = MW.stream;
  uint8_t* h = (uint8_t*)MW.h_buf; uint8_t* d = (uint8_t*)MW.d_buf;
  OKB_CUDA(cudaMemcpyAsync(d, h, in_end, cudaMemcpyHostToDevice, st));
  if (a.out_ctr) OKB_CUDA(cudaMemsetAsync(a.out_ctr, 0, 4, st));
  const int grid = (a.nq + 7) / 8;
  if (D == 64) {
    if (mode == MODE_M2) k_match_gated<4, MODE_M2><<<grid, 256, 0, st>>>(a);
    else if (mode == MODE_M3) k_match_gated<4, MODE_M3><<<grid, 256, 0, st>>>(a);
    else k_match_gated<4, MODE_M4><<<grid, 256, 0, st>>>(a);
  } else {
    if (mode == MODE_M2) k_match_gated<3, MODE_M2><<<grid, 256, 0, st>>>(a);
    else if (mode == MODE_M3) k_match_gated<3, MODE_M3><<<grid, 256, 0, st>>>(a);
    else k_match_gated<3, MODE_M4><<<grid, 256, 0, st>>>(a);
  }
  ctx->launches++;
  OKB_CUDA(cudaGetLastError());
  OKB_CUDA(cudaMemcpyAsync(h + o_dist, d + o_dist, o_end - o_dist, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(cudaStreamSynchronize(st));
  const size_t n = (size_t)a.nq;
  memcpy(out_dist, h + o_dist, n * 4); memcpy(out_idx, h + o_idx, n * 4);
  if (out_hp) memcpy(out_hp, h + o_hp, n * 32);
  if (out_init) memcpy(out_init, h + o_init, n);
  if (out_ctr) memcpy(out_ctr, h + o_ctr, 4);
  return OKB_OK;
}

int okb_match_map_uninit(okb_context_t* ctx, int D, int n_kp, const uint8_t* kp_desc, const double* kp_e_W,
                         const uint8_t* kp_use, const int32_t* kp_prev_lm, int n_cand, const uint8_t* cand_desc,
                         const int32_t* cand_lm, const double* cand_e_W, const double* cand_r_W, int n_lm,
                         const uint8_t* lm_is3d, const double r_WC1[3], double sigma, uint32_t match_threshold,
                         uint32_t* out_dist, int32_t* out_lm, double* out_hp_W, int32_t* out_ctr)
{
  OKB_CHECK_ARGS(ctx && !bad_D(D) && n_kp >= 0 && n_cand >= 0 && out_dist && out_lm && out_hp_W && r_WC1, "okb_match_map_uninit");
  OKB_CHECK_ARGS(n_kp == 0 || (kp_desc && kp_e_W), "okb_match_map_uninit");
  OKB_CHECK_ARGS(n_cand == 0 || (cand_desc && cand_lm && cand_e_W && cand_r_W && lm_is3d), "okb_match_map_uninit");
  OKB_CHECK_ARGS(n_lm >= 0, "okb_match_map_uninit");
  { int rcv = check_pool(n_cand, cand_lm, n_lm, "okb_match_map_uninit"); if (rcv) return rcv; }
  if (out_ctr) *out_ctr = 0;
  if (n_kp == 0) return OKB_OK;
  OKB_CUDA(cudaSetDevice(ctx->device));
  OKB_LOCK_SLOT;
  MatchArgs a; memset(&a, 0, sizeof(a));
  size_t o_dist = 0, o_idx = 0, o_hp = 0, o_ctr = 0, in_end = 0, o_end = 0;
  for (int pass = 0; pass < 2; pass++) {
    Arena A; A.ctx = ctx; A.dry = pass == 0; A.h = (uint8_t*)MW.h_buf; A.d = (uint8_t*)MW.d_buf;
    a.q_desc = A.in(kp_desc, (size_t)n_kp * D); a.q_e = A.in(kp_e_W, (size_t)n_kp * 3);
    a.q_use = kp_use ? A.in(kp_use, (size_t)n_kp) : nullptr;
    a.q_prev_lm = kp_prev_lm ? A.in(kp_prev_lm, (size_t)n_kp) : nullptr;
    a.c_desc = A.in(cand_desc, (size_t)n_cand * D); a.c_lm = A.in(cand_lm, (size_t)n_cand);
    a.c_e = A.in(cand_e_W, (size_t)n_cand * 3); a.c_r = A.in(cand_r_W, (size_t)n_cand * 3);
    a.lm_is3d = A.in(lm_is3d, (size_t)n_lm);
    in_end = A.off;
    a.out_dist = A.out<uint32_t>(n_kp, &o_dist); a.out_idx = A.out<int32_t>(n_kp, &o_idx);
    a.out_hp = A.out<double>((size_t)n_kp * 4, &o_hp); a.out_ctr = A.out<int32_t>(1, &o_ctr);
    o_end = A.off;
    if (pass == 0) { int rc = ensure(ctx, A.off); if (rc) return rc; }
  }
  a.nq = n_kp; a.nc = n_cand; a.thr = match_threshold;
  for (int i = 0; i < 3; i++) a.r1[i] = r_WC1[i];
  a.cos26 = gate_cos(2.6 * sigma); a.cos6 = gate_cos(6.0 * sigma);  // the libm algorithm, same function as on the device (okb_gatecos.h)
  return run_gated(ctx, MODE_M2, D, a, in_end, o_dist, o_end, o_idx, o_hp, 0, o_ctr, out_dist, out_lm, out_hp_W, nullptr, out_ctr);
}

// ---- M2, device-resident batched form: the queries are the camera's last detected features (descriptors, back-projections) ----
struct M2Prep { const double* rays; const uint8_t* valid; const int32_t* count; const uint8_t* use_in; int cap; const M3Frame* frames; double* e_W; uint8_t* use; };
// e1_W = T_WC1.C() * e1_C.normalized() (Frontend.cpp:1620-1627) and the use mask: inside the frame's count, back-projection valid,
// and the caller's mask (keypoints that already carry a landmark outside loop-closure mode)
__global__ void __launch_bounds__(128) k_m2_prep(const __grid_constant__ M2Prep p)
{
  const int frame = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.cap) return;
  const size_t i = (size_t)frame * p.cap + k;
  const bool in = k < min(p.count[frame], p.cap) && p.valid[i] && (p.use_in == nullptr || p.use_in[i]);
  if (in) {
    const V3 e = normalized(V3{p.rays[3 * i], p.rays[3 * i + 1], p.rays[3 * i + 2]});
    const double* C = p.frames[frame].Twc;
    p.e_W[3 * i] = (C[0] * e.x + C[1] * e.y) + C[2] * e.z;
    p.e_W[3 * i + 1] = (C[3] * e.x + C[4] * e.y) + C[5] * e.z;
    p.e_W[3 * i + 2] = (C[6] * e.x + C[7] * e.y) + C[8] * e.z;
  }
  p.use[i] = in ? 1 : 0;
}

int okb_match_map_uninit_device(okb_context_t* ctx, int cam, int n_frames, int n_cand, const uint8_t* d_cand_desc, const int32_t* d_cand_lm,
                                const double* d_cand_e_W, const double* d_cand_r_W, int n_lm, const uint8_t* d_lm_is3d, const double* T_WC1,
                                double sigma, uint32_t match_threshold, const uint8_t* d_kp_use, const int32_t* d_kp_prev_lm,
                                uint32_t* d_out_dist, int32_t* d_out_lm, double* d_out_hp_W, int32_t* d_out_ctr)
{
  OKB_CHECK_ARGS(ctx && cam >= 0 && cam < ctx->n_cams && n_cand >= 0 && n_lm >= 0 && T_WC1 && d_out_dist && d_out_lm && d_out_hp_W,
                 "okb_match_map_uninit_device");
  CamWorkspace& ws = ctx->cams[cam];
  OKB_CHECK_ARGS(ws.has_model && n_frames >= 1 && n_frames <= ws.cfg.max_batch, "okb_match_map_uninit_device (camera model set? okb_set_camera_model)");
  OKB_CHECK_ARGS(ws.cfg.descriptor_bytes == 64, "okb_match_map_uninit_device (64-byte descriptors)");
  OKB_CHECK_ARGS(n_cand == 0 || (d_cand_desc && d_cand_lm && d_cand_e_W && d_cand_r_W && d_lm_is3d), "okb_match_map_uninit_device");
  OKB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ws.stream;
  const size_t n = (size_t)n_frames * ws.kp_cap;
  const size_t b_frames = (sizeof(M3Frame) * n_frames + 255) & ~(size_t)255, need = b_frames + ((n * 24 + 255) & ~(size_t)255) + n;
  if (need > ws.m2_cap) {
    OKB_CUDA(cudaStreamSynchronize(st));
    cudaFree(ws.m2_d); ws.m2_d = nullptr; ws.m2_cap = 0;
    OKB_CUDA(cudaMalloc(&ws.m2_d, need + need / 4)); ws.m2_cap = need + need / 4;
  }
  M3Frame* d_frames = (M3Frame*)ws.m2_d; double* e_W = (double*)(ws.m2_d + b_frames); uint8_t* use = ws.m2_d + b_frames + ((n * 24 + 255) & ~(size_t)255);
  std::vector<M3Frame> hf(n_frames);   // pageable staging: copied before cudaMemcpyAsync returns
  for (int b = 0; b < n_frames; b++) { memcpy(hf[b].Twc, T_WC1 + 12 * (size_t)b, 96); memset(hf[b].Tcw, 0, 96); }
  OKB_CUDA(cudaMemcpyAsync(d_frames, hf.data(), sizeof(M3Frame) * n_frames, cudaMemcpyHostToDevice, st));
  M2Prep p; p.rays = ws.d_rays; p.valid = ws.d_rays_valid; p.count = ws.d_count; p.use_in = d_kp_use; p.cap = ws.kp_cap; p.frames = d_frames;
  p.e_W = e_W; p.use = use;
  k_m2_prep<<<dim3((ws.kp_cap + 127) / 128, n_frames), 128, 0, st>>>(p);
  MatchArgs a; memset(&a, 0, sizeof(a));
  a.nq = ws.kp_cap; a.q_count = ws.d_count; a.q_stride = (size_t)ws.kp_cap; a.q_desc = ws.d_desc; a.q_e = e_W; a.q_use = use; a.q_prev_lm = d_kp_prev_lm;
  a.nc = n_cand; a.c_stride = 0; a.c_desc = d_cand_desc; a.c_lm = d_cand_lm; a.c_e = d_cand_e_W; a.c_r = d_cand_r_W; a.lm_is3d = d_lm_is3d;
  a.frames = d_frames; a.thr = match_threshold;
  a.cos26 = gate_cos(2.6 * sigma); a.cos6 = gate_cos(6.0 * sigma);
  a.out_dist = d_out_dist; a.out_idx = d_out_lm; a.out_hp = d_out_hp_W; a.out_ctr = d_out_ctr;
  if (d_out_ctr) OKB_CUDA(cudaMemsetAsync(d_out_ctr, 0, (size_t)n_frames * 4, st));
  k_match_gated<4, MODE_M2><<<dim3((ws.kp_cap + 7) / 8, n_frames), 256, 0, st>>>(a);
  ctx->launches += 2;
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

static int stereo_like(okb_context_t* ctx, int mode, int D, int n0, const uint8_t* desc0, const uint8_t* use0,
                       const double* e0_W, const double* sof0, int n1, const uint8_t* desc1, const uint8_t* valid1,
                       const double* e1_W, const double* sof1, const double r_WC0[3], const double r_WC1[3],
                       const double T_CW0[12], const double T_CW1[12], uint32_t thr, int32_t* out_k1, uint32_t* out_dist,
                       double* out_hp_W, uint8_t* out_init, const char* who)
{
  OKB_CHECK_ARGS(ctx && !bad_D(D) && n0 >= 0 && n1 >= 0 && out_k1 && out_dist && out_hp_W && out_init && r_WC0 && r_WC1 && T_CW0 && T_CW1, who);
  OKB_CHECK_ARGS(n0 == 0 || (desc0 && e0_W && sof0), who);
  OKB_CHECK_ARGS(n1 == 0 || (desc1 && e1_W && valid1 && (mode == MODE_M3 || sof1)), who);
  if (n0 == 0) return OKB_OK;
  OKB_CUDA(cudaSetDevice(ctx->device));
  OKB_LOCK_SLOT;
  // per-keypoint cos(2.6 sigma), cos(6 sigma) tables (sigma = size/f * 0.125) by gate_cos: the libm algorithm, the very
  // function k_stereo_prep runs for the device-resident form, so both forms use identical constants by construction
  std::vector<double> c26_0(n0), c6_0(n0), c26_1, c6_1;
  for (int i = 0; i < n0; i++) { const double s = sof0[i] * 0.125; c26_0[i] = gate_cos(2.6 * s); c6_0[i] = gate_cos(6.0 * s); }
  if (mode == MODE_M4) {
    c26_1.resize(n1); c6_1.resize(n1);
    for (int i = 0; i < n1; i++) { const double s = sof1[i] * 0.125; c26_1[i] = gate_cos(2.6 * s); c6_1[i] = gate_cos(6.0 * s); }
  }
  MatchArgs a; memset(&a, 0, sizeof(a));
  size_t o_dist = 0, o_idx = 0, o_hp = 0, o_init = 0, in_end = 0, o_end = 0;
  for (int pass = 0; pass < 2; pass++) {
    Arena A; A.ctx = ctx; A.dry = pass == 0; A.h = (uint8_t*)MW.h_buf; A.d = (uint8_t*)MW.d_buf;
    a.q_desc = A.in(desc0, (size_t)n0 * D); a.q_e = A.in(e0_W, (size_t)n0 * 3);
    a.q_use = use0 ? A.in(use0, (size_t)n0) : nullptr;
    a.q_sof = A.in(sof0, (size_t)n0); a.q_cos26 = A.in(c26_0.data(), (size_t)n0); a.q_cos6 = A.in(c6_0.data(), (size_t)n0);
    a.c_desc = A.in(desc1, (size_t)n1 * D); a.c_e = A.in(e1_W, (size_t)n1 * 3); a.c_valid = A.in(valid1, (size_t)n1);
    if (mode == MODE_M4) { a.c_sof = A.in(sof1, (size_t)n1); a.c_cos26 = A.in(c26_1.data(), (size_t)n1); a.c_cos6 = A.in(c6_1.data(), (size_t)n1); }
    in_end = A.off;
    a.out_dist = A.out<uint32_t>(n0, &o_dist); a.out_idx = A.out<int32_t>(n0, &o_idx);
    a.out_hp = A.out<double>((size_t)n0 * 4, &o_hp); a.out_init = A.out<uint8_t>(n0, &o_init);
    o_end = A.off;
    if (pass == 0) { int rc = ensure(ctx, A.off); if (rc) return rc; }
  }
  a.nq = n0; a.nc = n1; a.thr = thr;
  for (int i = 0; i < 3; i++) { a.r0[i] = r_WC0[i]; a.r1[i] = r_WC1[i]; }
  for (int i = 0; i < 12; i++) { a.T0[i] = T_CW0[i]; a.T1[i] = T_CW1[i]; }
  return run_gated(ctx, mode, D, a, in_end, o_dist, o_end, o_idx, o_hp, o_init, 0, out_dist, out_k1, out_hp_W, out_init, nullptr);
}

int okb_match_motion_stereo(okb_context_t* ctx, int D, int n0, const uint8_t* desc0, const uint8_t* use0, const double* e0_W,
                            const double* size_over_f0, int n1, const uint8_t* desc1, const uint8_t* valid1,
                            const double* e1_W, const double r_WC0[3], const double r_WC1[3], const double T_CW0[12],
                            const double T_CW1[12], uint32_t match_threshold, int32_t* out_k1, uint32_t* out_dist,
                            double* out_hp_W, uint8_t* out_initialisable)
{
  return stereo_like(ctx, MODE_M3, D, n0, desc0, use0, e0_W, size_over_f0, n1, desc1, valid1, e1_W, nullptr, r_WC0, r_WC1,
                     T_CW0, T_CW1, match_threshold, out_k1, out_dist, out_hp_W, out_initialisable, "okb_match_motion_stereo");
}

int okb_match_stereo(okb_context_t* ctx, int D, int n0, const uint8_t* desc0, const uint8_t* valid0, const double* e0_W,
                     const double* size_over_f0, int n1, const uint8_t* desc1, const uint8_t* valid1, const double* e1_W,
                     const double* size_over_f1, const double r_WC0[3], const double r_WC1[3], const double T_CW0[12],
                     const double T_CW1[12], uint32_t match_threshold, int32_t* out_k1, uint32_t* out_dist, double* out_hp_W,
                     uint8_t* out_initialisable)
{
  return stereo_like(ctx, MODE_M4, D, n0, desc0, valid0, e0_W, size_over_f0, n1, desc1, valid1, e1_W, size_over_f1, r_WC0,
                     r_WC1, T_CW0, T_CW1, match_threshold, out_k1, out_dist, out_hp_W, out_initialisable, "okb_match_stereo");
}

int okb_match_map3d_device(okb_context_t* ctx, int cam, int D, int n_frames, int n_cand, const uint8_t* d_cand_desc,
                           const int32_t* d_cand_lm, int n_lm, const double* d_lm_proj, const uint8_t* d_lm_is3d,
                           double reprojection_threshold, uint32_t match_threshold, uint32_t* d_out_dist, int32_t* d_out_lm)
{
  return okb::match_map3d_enqueue(ctx, cam, D, n_frames, n_cand, d_cand_desc, d_cand_lm, n_lm, d_lm_proj, d_lm_is3d, reprojection_threshold,
                                  match_threshold, d_out_dist, d_out_lm, nullptr, nullptr);
}

}  // extern "C"

int okb::match_map3d_enqueue(okb_context_t* ctx, int cam, int D, int n_frames, int n_cand, const uint8_t* d_cand_desc,
                             const int32_t* d_cand_lm, int n_lm, const double* d_lm_proj, const uint8_t* d_lm_is3d,
                             double reprojection_threshold, uint32_t match_threshold, uint32_t* d_out_dist, int32_t* d_out_lm,
                             cudaStream_t bin_stream, cudaEvent_t bin_done)
{
  OKB_CHECK_ARGS(ctx && cam >= 0 && cam < ctx->n_cams && n_cand >= 0 && n_lm >= 0 && d_out_dist && d_out_lm, "okb_match_map3d_device");
  CamWorkspace& ws = ctx->cams[cam];
  if (D != ws.cfg.descriptor_bytes) {   // the pool rows must have the stride of the camera's own descriptors (the queries)
    set_error("okb_match_map3d_device: pool descriptors of %d bytes against camera %d, which describes with %d bytes", D, cam,
              ws.cfg.descriptor_bytes);
    return OKB_ERR_ARGUMENT;
  }
  OKB_CHECK_ARGS(n_frames >= 1 && n_frames <= ws.cfg.max_batch, "okb_match_map3d_device");
  OKB_CHECK_ARGS(n_cand == 0 || (d_cand_desc && d_cand_lm && d_lm_proj && d_lm_is3d), "okb_match_map3d_device");
  OKB_CHECK_ARGS(reprojection_threshold >= 0.0 && reprojection_threshold < 1e6, "okb_match_map3d_device");
  OKB_CUDA(cudaSetDevice(ctx->device));
  M1Args a; memset(&a, 0, sizeof(a));
  m1_grid(reprojection_threshold, ws.cfg.width, ws.cfg.height, a.cell, a.gx, a.gy);
  a.nq = ws.kp_cap; a.q_count = ws.d_count; a.q_stride = (size_t)ws.kp_cap; a.proj_stride = (size_t)n_lm * 2;
  a.q_desc = ws.d_desc; a.q_kp = ws.d_kp;
  a.nc = n_cand; a.c_desc = d_cand_desc; a.c_lm = d_cand_lm; a.lm_proj = d_lm_proj; a.lm_is3d = d_lm_is3d;
  a.thr = match_threshold; a.thr_sq = reprojection_threshold * reprojection_threshold;
  a.thr_px = reprojection_threshold; a.min_x = 0.0; a.min_y = 0.0; a.max_x = ws.cfg.width; a.max_y = ws.cfg.height;
  {
    // per-frame row bins: [n_frames][kMaxCells + 2] offsets, [n_frames][n_cand] rows and projections; grown on demand
    const size_t off_b = (((size_t)n_frames * (kMaxCells + 2) * 4) + 255) & ~(size_t)255;
    const size_t list_b = (((size_t)n_frames * ((size_t)n_cand + 1) * 4) + 255) & ~(size_t)255;
    const size_t need = off_b + list_b + (size_t)n_frames * ((size_t)n_cand + 1) * 16;
    if (need > ws.m1_rows_cap) {
      OKB_CUDA(cudaStreamSynchronize(ws.stream));
      cudaFree(ws.d_m1_rows); ws.d_m1_rows = nullptr; ws.m1_rows_cap = 0;
      OKB_CUDA(cudaMalloc(&ws.d_m1_rows, need + need / 4));
      ws.m1_rows_cap = need + need / 4;
    }
    a.row_off = (int32_t*)ws.d_m1_rows; a.row_list = (int32_t*)(ws.d_m1_rows + off_b); a.row_xy = (double2*)(ws.d_m1_rows + off_b + list_b);
  }
  a.out_dist = d_out_dist; a.out_idx = d_out_lm;
  return m1_launch(ctx, a, D, n_frames, ws.stream, bin_stream, bin_done);
}

extern "C" {

// T_CW = T_WC.inverse() = [C^T | -(C^T r)] as row-major 3x4 (kinematics/implementation/Transformation.hpp:207-209)
static void invert_pose(const double C[9], const double r[3], double T[12])
{
  for (int i = 0; i < 3; i++) {
    T[4 * i] = C[i]; T[4 * i + 1] = C[3 + i]; T[4 * i + 2] = C[6 + i];
    T[4 * i + 3] = -((C[i] * r[0] + C[3 + i] * r[1]) + C[6 + i] * r[2]);
  }
}

int okb_match_stereo_device_ptr(okb_context_t* ctx, int n_frames, int cap0, const okb_keypoint_t* d_kp0, const uint8_t* d_desc0,
                                const int32_t* d_count0, const okb_camera_model_t* model0, const double C_WC0[9],
                                const double r_WC0[3], int cap1, const okb_keypoint_t* d_kp1, const uint8_t* d_desc1,
                                const int32_t* d_count1, const okb_camera_model_t* model1, const double C_WC1[9],
                                const double r_WC1[3], uint32_t match_threshold, void* stream, int32_t* d_out_k1,
                                uint32_t* d_out_dist, double* d_out_hp_W, uint8_t* d_out_initialisable)
{
  OKB_CHECK_ARGS(ctx && n_frames >= 1 && cap0 > 0 && cap1 > 0 && d_kp0 && d_desc0 && d_count0 && model0 && d_kp1 && d_desc1 &&
                 d_count1 && model1 && C_WC0 && r_WC0 && C_WC1 && r_WC1 && d_out_k1 && d_out_dist && d_out_hp_W && d_out_initialisable,
                 "okb_match_stereo_device_ptr");
  OKB_CHECK_ARGS(cap0 < (1 << 20) && cap1 < (1 << 20), "okb_match_stereo_device_ptr (capacity >= 2^20 keypoints per frame)");
  OKB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : MW.stream;
  // scratch: per side rays(3) eW(3) sof c26 c6 doubles + valid bytes
  const size_t n0 = (size_t)n_frames * cap0, n1 = (size_t)n_frames * cap1;
  const int hit_cap = 16 * cap0;   // hits (distance < threshold) per frame kept for the gate pass; more -> sequential replay
  const size_t need_prep = ((n0 + n1) * (9 * 8 + 8) + 255) & ~(size_t)255;
  const size_t need = need_prep + n0 * 8 + (size_t)n_frames * hit_cap * 8 + (size_t)n_frames * 4 + 256;
  if (need > ctx->stereo_cap) {
    OKB_CUDA(cudaDeviceSynchronize());
    if (ctx->stereo_scratch) cudaFree(ctx->stereo_scratch);
    ctx->stereo_scratch = nullptr; ctx->stereo_cap = 0;
    OKB_CUDA(cudaMalloc(&ctx->stereo_scratch, need + need / 4));
    ctx->stereo_cap = need + need / 4;
  }
  double* base = (double*)ctx->stereo_scratch;
  double *rays0 = base, *eW0 = rays0 + 3 * n0, *sof0 = eW0 + 3 * n0, *c26_0 = sof0 + n0, *c6_0 = c26_0 + n0;
  double *rays1 = c6_0 + n0, *eW1 = rays1 + 3 * n1, *sof1 = eW1 + 3 * n1, *c26_1 = sof1 + n1, *c6_1 = c26_1 + n1;
  uint8_t* valid0 = (uint8_t*)(c6_1 + n1); uint8_t* valid1 = valid0 + ((n0 + 7) & ~(size_t)7);
  unsigned long long* best = (unsigned long long*)((uint8_t*)ctx->stereo_scratch + need_prep);
  uint2* hits = (uint2*)(best + n0);
  int32_t* hit_cnt = (int32_t*)(hits + (size_t)n_frames * hit_cap);
  {
    // D4 + world-frame tables of both sides and the reset of the reduction arrays: one launch
    const okb_camera_model_t* const models[2] = {model0, model1}; const double* const Cs[2] = {C_WC0, C_WC1};
    const okb_keypoint_t* const kps[2] = {d_kp0, d_kp1}; const int32_t* const counts[2] = {d_count0, d_count1}; const int caps[2] = {cap0, cap1};
    double* const rays[2] = {rays0, rays1}; uint8_t* const valids[2] = {valid0, valid1}; double* const eWs[2] = {eW0, eW1};
    double* const sofs[2] = {sof0, sof1}; double* const c26s[2] = {c26_0, c26_1}; double* const c6s[2] = {c6_0, c6_1};
    const int rc = okb::camera_stereo_prep_pair(ctx, models, Cs, kps, counts, caps, n_frames, rays, valids, eWs, sofs, c26s, c6s, best, hit_cnt, st);
    if (rc) return rc;
  }
  MatchArgs a; memset(&a, 0, sizeof(a));
  a.nq = cap0; a.nc = cap1; a.q_stride = (size_t)cap0; a.c_stride = (size_t)cap1; a.q_count = d_count0; a.c_count = d_count1;
  a.q_desc = d_desc0; a.c_desc = d_desc1; a.q_use = valid0; a.q_e = eW0; a.q_sof = sof0; a.q_cos26 = c26_0; a.q_cos6 = c6_0;
  a.c_valid = valid1; a.c_e = eW1; a.c_sof = sof1; a.c_cos26 = c26_1; a.c_cos6 = c6_1;
  for (int i = 0; i < 3; i++) { a.r0[i] = r_WC0[i]; a.r1[i] = r_WC1[i]; }
  invert_pose(C_WC0, r_WC0, a.T0); invert_pose(C_WC1, r_WC1, a.T1);
  a.thr = match_threshold;
  a.out_dist = d_out_dist; a.out_idx = d_out_k1; a.out_hp = d_out_hp_W; a.out_init = d_out_initialisable;
  a.hit_cnt = hit_cnt; a.hit_cap = hit_cap;
  const int fused_mode = g_m3_fused.load();
  const bool fused = fused_mode != 0;   // default (-1): the one-launch form for every batch size (measured faster for 1 and for 32 frames)
  const int scan_mode = g_scan_mma.load();
  if (scan_mode == 2) {
    a.scan_qt = n_frames <= 4 ? 128 : 0;   // 0: one CTA per (candidate tile, frame) streams all queries
    { const int rc = umma_prepare(); if (rc) return rc; }
    k_scan_umma<<<dim3((cap1 + 127) / 128, n_frames, a.scan_qt ? (cap0 + a.scan_qt - 1) / a.scan_qt : 1), kUmmaThreads, kUmmaSmem, st>>>(a, hits, hit_cnt, ctx->d_scan_status);
  } else if (scan_mode == 1) {
    a.scan_qt = n_frames <= 4 ? 32 : 256;   // small batches: short query chunks so that one frame's scan still spreads over the SMs
    k_scan_mma<<<dim3((cap1 + 255) / 256, n_frames, (cap0 + a.scan_qt - 1) / a.scan_qt), 128, 0, st>>>(a, hits, hit_cnt);
  } else {
    a.scan_qt = n_frames <= 4 ? 32 : 128;
    k_m4_scan<4><<<dim3((cap1 + 255) / 256, n_frames, (cap0 + a.scan_qt - 1) / a.scan_qt), 256, 0, st>>>(a, hits, hit_cnt);
  }
  if (fused) {
    M3Check none; memset(&none, 0, sizeof(none));
    k_pair_view<4, MODE_M4><<<n_frames, 512, 0, st>>>(a, none, hits, best, 0, 1, 0);
    ctx->launches += 2;
  } else {
    k_m4_gate<MODE_M4><<<dim3(8, n_frames), 128, 0, st>>>(a, hits, best);
    k_m4_finish<MODE_M4><<<dim3((cap0 + 127) / 128, n_frames), 128, 0, st>>>(a, best);
    k_match_gated<4, MODE_M4><<<dim3((cap0 + 7) / 8, n_frames), 256, 0, st>>>(a);   // only frames whose hit list overflowed
    ctx->launches += 4;
  }
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

int okb_match_stereo_device(okb_context_t* ctx, int cam0, int cam1, int n_frames, const double C_WC0[9], const double r_WC0[3],
                            const double C_WC1[9], const double r_WC1[3], uint32_t match_threshold, int32_t* d_out_k1,
                            uint32_t* d_out_dist, double* d_out_hp_W, uint8_t* d_out_initialisable)
{
  OKB_CHECK_ARGS(ctx && cam0 >= 0 && cam0 < ctx->n_cams && cam1 >= 0 && cam1 < ctx->n_cams && cam0 != cam1, "okb_match_stereo_device");
  CamWorkspace& w0 = ctx->cams[cam0]; CamWorkspace& w1 = ctx->cams[cam1];
  OKB_CHECK_ARGS(w0.has_model && w1.has_model && n_frames >= 1 && n_frames <= w0.cfg.max_batch && n_frames <= w1.cfg.max_batch,
                 "okb_match_stereo_device (camera models set? okb_set_camera_model)");
  OKB_CUDA(cudaSetDevice(ctx->device));
  // runs on camera 0's stream once camera 1's features are complete
  OKB_CUDA(cudaEventRecord(w1.ev_done, w1.stream));
  OKB_CUDA(cudaStreamWaitEvent(w0.stream, w1.ev_done, 0));
  int rc = okb_match_stereo_device_ptr(ctx, n_frames, w0.kp_cap, w0.d_kp, desc_slots(w0), w0.d_count, &w0.model, C_WC0, r_WC0, w1.kp_cap,
                                       w1.d_kp, desc_slots(w1), w1.d_count, &w1.model, C_WC1, r_WC1, match_threshold, (void*)w0.stream,
                                       d_out_k1, d_out_dist, d_out_hp_W, d_out_initialisable);
  if (rc) return rc;
  // camera 1 must not overwrite its features before the matcher has read them
  OKB_CUDA(cudaEventRecord(w0.ev_done, w0.stream));
  OKB_CUDA(cudaStreamWaitEvent(w1.stream, w0.ev_done, 0));
  return OKB_OK;
}


int okb_match_motion_stereo_device_ptr(okb_context_t* ctx, int n_frames, int cap1, const okb_keypoint_t* d_kp1, const uint8_t* d_desc1,
                                       const int32_t* d_count1, const okb_camera_model_t* model, int width, int height,
                                       const double* T_WC1, const double* T_CW1, int n_older, const okb_older_view_t* older, int cap0,
                                       uint32_t match_threshold, void* stream, uint8_t* d_matched1, int32_t* d_out_k1,
                                       uint32_t* d_out_dist, double* d_out_hp_W, uint8_t* d_out_flags)
{
  OKB_CHECK_ARGS(ctx, "okb_match_motion_stereo_device_ptr");
  // one scratch area per context for this form: calls of one context must be issued on streams that serialise them
  return okb::motion_sequence(ctx, ctx->motion, n_frames, cap1, d_kp1, d_desc1, d_count1, model, width, height, T_WC1, T_CW1, n_older, older, cap0,
                              match_threshold, stream ? (cudaStream_t)stream : MW.stream, d_matched1, d_out_k1, d_out_dist, d_out_hp_W, d_out_flags, nullptr, nullptr, nullptr);
}

}  // extern "C"

static void fill_m3_tables(M3View* hv, M3Frame* hf, int n_frames, int n_older, const okb_older_view_t* older, const double* T_WC1, const double* T_CW1)
{
  for (size_t i = 0; i < (size_t)n_frames * n_older; i++) {
    const okb_older_view_t& s = older[i];
    hv[i].desc = s.d_desc; hv[i].rays = s.d_rays; hv[i].valid = s.d_valid; hv[i].size = s.d_size; hv[i].use = s.d_use; hv[i].n = s.n; hv[i].pad = 0;
    memcpy(hv[i].Twc, s.T_WC, sizeof(hv[i].Twc)); memcpy(hv[i].Tcw, s.T_CW, sizeof(hv[i].Tcw));
  }
  for (int b = 0; b < n_frames; b++) { memcpy(hf[b].Twc, T_WC1 + 12 * (size_t)b, 96); memcpy(hf[b].Tcw, T_CW1 + 12 * (size_t)b, 96); }
}

// CUDA-graph callers (okb_process_multiframe): the captured copy node re-reads the scratch's page-locked tables at every replay,
// so they are refreshed here, per multiframe, without enqueuing anything. Same argument checks as motion_sequence.
int okb::motion_restage(MotionScratch& ms, int n_frames, int n_older, const okb_older_view_t* older, int cap0, const double* T_WC1, const double* T_CW1)
{
  if (n_older == 0) return OKB_OK;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t b_views = al(sizeof(M3View) * n_frames * n_older), b_frames = al(sizeof(M3Frame) * n_frames);
  OKB_CHECK_ARGS(ms.h && ms.pinned_staging && b_views + b_frames <= ms.h_cap && older && T_WC1 && T_CW1, "motion_restage");
  for (int i = 0; i < n_frames * n_older; i++)
    OKB_CHECK_ARGS(older[i].n >= 0 && older[i].n <= cap0 && (older[i].n == 0 || (older[i].d_desc && older[i].d_rays && older[i].d_valid && older[i].d_size)),
                   "okb_process_multiframe (older view)");
  fill_m3_tables((M3View*)ms.h, (M3Frame*)((uint8_t*)ms.h + b_views), n_frames, n_older, older, T_WC1, T_CW1);
  return OKB_OK;
}

extern "C" void okb_m3_set_fused(int mode) { g_m3_fused.store(mode < 0 ? -1 : (mode ? 1 : 0)); }
extern "C" void okb_scan_set_mma(int mode) { g_scan_mma.store(mode < 0 ? 0 : (mode > 2 ? 2 : mode)); }


int okb::motion_sequence(okb_context_t* ctx, MotionScratch& ms, int n_frames, int cap1, const okb_keypoint_t* d_kp1, const uint8_t* d_desc1,
                         const int32_t* d_count1, const okb_camera_model_t* model, int width, int height, const double* T_WC1, const double* T_CW1,
                         int n_older, const okb_older_view_t* older, int cap0, uint32_t match_threshold, cudaStream_t st, uint8_t* d_matched1,
                         int32_t* d_out_k1, uint32_t* d_out_dist, double* d_out_hp_W, uint8_t* d_out_flags, const double* d_rays1, const uint8_t* d_valid1,
                         const int32_t* d_m1_lm)
{
  OKB_CHECK_ARGS(ctx && n_frames >= 1 && cap1 > 0 && cap1 < (1 << 20) && d_kp1 && d_desc1 && d_count1 && model && T_WC1 && T_CW1 &&
                 n_older >= 0 && (n_older == 0 || older) && cap0 > 0 && cap0 < (1 << 20) && d_matched1 && d_out_k1 && d_out_dist &&
                 d_out_hp_W && d_out_flags && width > 0 && height > 0, "okb_match_motion_stereo_device_ptr");
  for (int i = 0; i < n_frames * n_older; i++)
    OKB_CHECK_ARGS(older[i].n >= 0 && older[i].n <= cap0 && (older[i].n == 0 || (older[i].d_desc && older[i].d_rays && older[i].d_valid && older[i].d_size)),
                   "okb_match_motion_stereo_device_ptr (older view)");
  if (n_older == 0) return OKB_OK;
  OKB_CUDA(cudaSetDevice(ctx->device));
  const size_t nq = (size_t)n_frames * n_older * cap0, n1 = (size_t)n_frames * cap1;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t b_views = al(sizeof(M3View) * n_frames * n_older), b_frames = al(sizeof(M3Frame) * n_frames);
  const int hit_cap = 16 * cap0;   // hits (distance < threshold) per frame and view kept for the gate pass
  const size_t n_hits = (size_t)n_older * n_frames * hit_cap;
  const size_t need = b_views + b_frames + al(nq * 24) + 2 * al(nq * 8) + al(nq) + al(nq) /*init*/ + al(n1 * 24) + al(n1 * 24) + 2 * al(n1) +
                      al(n1 * 4 * n_older) + al(nq * 8) + al(n_hits * 8) + 2 * al((size_t)n_older * n_frames * 4) + al(nq * 4);
  if (need > ms.cap) {
    OKB_CUDA(cudaDeviceSynchronize());
    if (ms.d) cudaFree(ms.d);
    if (ms.h) cudaFreeHost(ms.h);
    ms.d = nullptr; ms.h = nullptr; ms.cap = 0;
    OKB_CUDA(cudaMalloc(&ms.d, need + need / 4));
    OKB_CUDA(cudaMallocHost(&ms.h, b_views + b_frames));
    ms.cap = need + need / 4; ms.h_cap = b_views + b_frames;
  }
  if (b_views + b_frames > ms.h_cap) {
    OKB_CUDA(cudaDeviceSynchronize());
    if (ms.h) cudaFreeHost(ms.h);
    OKB_CUDA(cudaMallocHost(&ms.h, b_views + b_frames)); ms.h_cap = b_views + b_frames;
  }
  uint8_t* base = (uint8_t*)ms.d; size_t o = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base + o; o += al(bytes); return p; };
  M3View* d_views = (M3View*)take(sizeof(M3View) * n_frames * n_older);   // views, then frames: one contiguous upload
  M3Frame* d_frames = (M3Frame*)take(sizeof(M3Frame) * n_frames);
  double* e0 = (double*)take(nq * 24); double* c26 = (double*)take(nq * 8); double* c6 = (double*)take(nq * 8);
  uint8_t* use0 = take(nq); uint8_t* init = take(nq);
  double* rays1 = (double*)take(n1 * 24); double* e1 = (double*)take(n1 * 24);
  uint8_t* valid1 = take(n1); uint8_t* cvalid = take(n1); int32_t* claim = (int32_t*)take(n1 * 4 * n_older);
  unsigned long long* best = (unsigned long long*)take(nq * 8); uint2* hits = (uint2*)take(n_hits * 8);
  int32_t* hit_cnt = (int32_t*)take((size_t)n_older * n_frames * 4);
  int32_t* q_list_cnt = (int32_t*)take((size_t)n_older * n_frames * 4); int32_t* q_list = (int32_t*)take(nq * 4);
  // descriptors of the views and poses. Asynchronous callers: pageable host staging (cudaMemcpyAsync copies it before it
  // returns, so back-to-back calls cannot overwrite each other's tables). Streaming / CUDA-graph callers (ms.pinned_staging: one
  // call in flight, synchronised per multiframe): the scratch's page-locked mirror, which a captured copy node re-reads at replay
  std::vector<uint8_t> hbuf_(ms.pinned_staging ? 0 : b_views + b_frames);
  uint8_t* hb = ms.pinned_staging ? (uint8_t*)ms.h : hbuf_.data();
  M3View* hv = (M3View*)hb;
  M3Frame* hf = (M3Frame*)(hb + b_views);
  fill_m3_tables(hv, hf, n_frames, n_older, older, T_WC1, T_CW1);
  OKB_CUDA(cudaMemcpyAsync(d_views, hb, b_views + sizeof(M3Frame) * (size_t)n_frames, cudaMemcpyHostToDevice, st));
  const Model cam = to_model(*model);
  // D4 of the current keypoints (Frame::computeBackProjections), unless the caller has them already (the detector's own, same model)
  if (d_rays1 && d_valid1) { rays1 = const_cast<double*>(d_rays1); valid1 = const_cast<uint8_t*>(d_valid1); }
  else { k_backproject_ext(cam, d_kp1, d_count1, cap1, n_frames, rays1, valid1, st); ctx->launches++; }
  // ---- once per sequence: tables of all views + the candidate mask, then the Hamming scan of all views (the scan does not depend
  //      on what earlier views insert: it keeps every pair below the threshold whose candidate is unmatched NOW; the gate of
  //      view v re-tests the mask as views 0..v-1 left it)
  const int gmax = (std::max(cap0, cap1) + 127) / 128;
  M3Prep p; memset(&p, 0, sizeof(p));
  p.views = d_views; p.n_views = n_older; p.frames = d_frames; p.cap0 = cap0; p.cap1 = cap1;
  p.q_stride = (size_t)n_older * cap0; p.f0 = 0.5 * (model->fu + model->fv);
  p.e0 = e0; p.c26 = c26; p.c6 = c6; p.use0 = use0;
  p.rays1 = rays1; p.valid1 = valid1; p.count1 = d_count1; p.matched1 = d_matched1; p.e1 = e1; p.cvalid = cvalid;
  p.best = best; p.claim = claim; p.hit_cnt = hit_cnt; p.m1_lm = d_m1_lm; p.q_list = q_list; p.q_list_cnt = q_list_cnt;
  OKB_CUDA(cudaMemsetAsync(q_list_cnt, 0, (size_t)n_older * n_frames * 4, st));
  k_m3_prep<<<dim3(gmax, n_frames, n_older), 128, 0, st>>>(p);
  MatchArgs a; memset(&a, 0, sizeof(a));
  a.nq = cap0; a.nc = cap1; a.q_stride = p.q_stride; a.c_stride = (size_t)cap1; a.c_count = d_count1;
  a.c_desc = d_desc1; a.c_valid = cvalid; a.c_e = e1; a.thr = match_threshold;
  a.views = d_views; a.view_stride = n_older; a.frames = d_frames; a.hit_cap = hit_cap;
  {
    MatchArgs s = a;
    s.q_use = use0; s.view_index = 0; s.scan_views = n_older; s.hit_cnt = hit_cnt;
    static const int qt_env = getenv("OKB_SCAN_QT") ? atoi(getenv("OKB_SCAN_QT")) : 0;   // tuning hook
    const int scan_mode = g_scan_mma.load();
    if (scan_mode == 2) {
      // tcgen05 scan over the eligible keypoints of every view: 128-candidate tiles, query chunks of 512 (128 for small batches:
      // one frame's scan then still spreads over the SMs)
      s.q_list = q_list; s.q_list_cnt = q_list_cnt;
      // batches: one CTA per (candidate tile, frame) streams the eligible keypoints of ALL views against its tile (expanded once);
      // small batches: one CTA per (candidate tile, view, chunk of 128 queries), so that a single frame still spreads over the SMs
      s.scan_qt = n_frames >= 8 ? 0 : 128; s.scan_chunks = s.scan_qt ? (cap0 + s.scan_qt - 1) / s.scan_qt : 1;
      { const int rc = umma_prepare(); if (rc) return rc; }
      k_scan_umma<<<dim3((cap1 + 127) / 128, n_frames, s.scan_qt ? n_older * s.scan_chunks : 1), kUmmaThreads, kUmmaSmem, st>>>(s, hits, hit_cnt, ctx->d_scan_status);
    } else if (scan_mode == 1) {
      // legacy integer MMA scan; small batches: short query chunks
      s.q_list = q_list; s.q_list_cnt = q_list_cnt;
      s.scan_qt = qt_env > 0 ? std::min(qt_env, 256) : (n_frames >= 8 ? 256 : 32); s.scan_chunks = (cap0 + s.scan_qt - 1) / s.scan_qt;
      k_scan_mma<<<dim3((cap1 + 255) / 256, n_frames, n_older * s.scan_chunks), 128, 0, st>>>(s, hits, hit_cnt);
    } else {
      s.scan_qt = qt_env > 0 ? std::min(qt_env, 128) : (n_frames >= 8 ? 128 : 32); s.scan_chunks = (cap0 + s.scan_qt - 1) / s.scan_qt;
      k_m4_scan<4><<<dim3((cap1 + 255) / 256, n_frames, n_older * s.scan_chunks), 256, 0, st>>>(s, hits, hit_cnt);
    }
  }
  ctx->launches += 2;
  // ---- per older keyframe, in order (what a view inserts is invisible to the next one's candidates)
  const int fused_mode = g_m3_fused.load();
  const bool fused = fused_mode != 0;   // default (-1): the one-launch form for every batch size (measured faster for 1 and for 32 frames)
  for (int v = 0; v < n_older; v++) {
    const size_t vo = (size_t)v * cap0;
    a.view_index = v;
    a.q_use = use0 + vo; a.q_e = e0 + 3 * vo; a.q_sof = c26 + vo /* unused by M3 */; a.q_cos26 = c26 + vo; a.q_cos6 = c6 + vo;
    a.out_dist = d_out_dist + vo; a.out_idx = d_out_k1 + vo; a.out_hp = d_out_hp_W + 4 * vo; a.out_init = init + vo;
    a.hit_cnt = hit_cnt + (size_t)v * n_frames;
    const uint2* hits_v = hits + (size_t)v * n_frames * hit_cap;
    unsigned long long* best_v = best + vo;
    M3Check c; memset(&c, 0, sizeof(c));
    c.views = d_views; c.view_stride = n_older; c.view_index = v; c.frames = d_frames; c.cap0 = cap0; c.cap1 = cap1; c.q_stride = p.q_stride;
    c.cam = cam; c.width = width; c.height = height; c.thr = match_threshold; c.kp1 = d_kp1;
    c.k1 = a.out_idx; c.dist = a.out_dist; c.hp = a.out_hp; c.init = a.out_init; c.flags = d_out_flags + vo;
    c.claim = claim + (size_t)v * n1; c.matched1 = d_matched1; c.cvalid = cvalid;
    if (fused) {
      // one launch for all views: with the bases of view 0 (v == 0 here), the kernel walks the views in order
      k_pair_view<4, MODE_M3><<<n_frames, 512, 0, st>>>(a, c, hits, best, 0, n_older, n1);
      ctx->launches++;
      break;
    } else {
      // gate per hit -> outputs; a frame whose hit list overflows is redone by the sequential-replay kernel (which returns at
      // once for all other frames); then the re-projection check / claim and the commit
      k_m4_gate<MODE_M3><<<dim3(8, n_frames), 128, 0, st>>>(a, hits_v, best_v);
      k_m4_finish<MODE_M3><<<dim3((cap0 + 127) / 128, n_frames), 128, 0, st>>>(a, best_v);
      k_match_gated<4, MODE_M3><<<dim3((cap0 + 7) / 8, n_frames), 256, 0, st>>>(a);
      k_m3_check<<<dim3((cap0 + 127) / 128, n_frames), 128, 0, st>>>(c);
      k_m3_commit<<<dim3((cap0 + 127) / 128, n_frames), 128, 0, st>>>(c);
      ctx->launches += 5;
    }
  }
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

extern "C" {

int okb_match_motion_stereo_device(okb_context_t* ctx, int cam, int n_frames, const double* T_WC1, const double* T_CW1, int n_older,
                                   const okb_older_view_t* older, int cap0, uint32_t match_threshold, uint8_t* d_matched1,
                                   int32_t* d_out_k1, uint32_t* d_out_dist, double* d_out_hp_W, uint8_t* d_out_flags)
{
  OKB_CHECK_ARGS(ctx && cam >= 0 && cam < ctx->n_cams, "okb_match_motion_stereo_device");
  CamWorkspace& ws = ctx->cams[cam];
  OKB_CHECK_ARGS(ws.has_model && n_frames >= 1 && n_frames <= ws.cfg.max_batch, "okb_match_motion_stereo_device (camera model set? okb_set_camera_model)");
  // per-camera scratch: the sequences of different cameras run concurrently on their own streams
  return okb::motion_sequence(ctx, ws.motion, n_frames, ws.kp_cap, ws.d_kp, desc_slots(ws), ws.d_count, &ws.model, ws.cfg.width, ws.cfg.height, T_WC1, T_CW1,
                              n_older, older, cap0, match_threshold, ws.stream, d_matched1, d_out_k1, d_out_dist, d_out_hp_W, d_out_flags, ws.d_rays, ws.d_rays_valid, nullptr);
}

int okb_matched_mask_device(okb_context_t* ctx, int cam, int n_frames, const int32_t* d_lm, uint8_t* d_matched)
{
  OKB_CHECK_ARGS(ctx && cam >= 0 && cam < ctx->n_cams && d_lm && d_matched, "okb_matched_mask_device");
  CamWorkspace& ws = ctx->cams[cam];
  OKB_CHECK_ARGS(n_frames >= 1 && n_frames <= ws.cfg.max_batch, "okb_matched_mask_device");
  OKB_CUDA(cudaSetDevice(ctx->device));
  k_matched_mask<<<dim3((ws.kp_cap + 255) / 256, n_frames), 256, 0, ws.stream>>>(d_lm, ws.d_count, ws.kp_cap, d_matched);
  ctx->launches++;
  OKB_CUDA(cudaGetLastError());
  return OKB_OK;
}

int okb_match_motion_stereo_batch(okb_context_t* ctx, int cam, int n_frames, const double* T_WC1, const double* T_CW1, int n_older,
                                  const okb_older_view_t* older, int cap0, uint32_t match_threshold, int cap, uint8_t* matched1,
                                  int cap_m, int32_t* n_match, int32_t* m_k0, int32_t* m_k1, uint8_t* m_flags, double* m_hp_W)
{
  OKB_CHECK_ARGS(ctx && cam >= 0 && cam < ctx->n_cams && matched1 && cap > 0 && cap_m > 0 && n_match && m_k0 && m_k1 && m_flags && m_hp_W &&
                 n_older >= 1 && cap0 > 0, "okb_match_motion_stereo_batch");
  CamWorkspace& ws = ctx->cams[cam];
  OKB_CHECK_ARGS(n_frames >= 1 && n_frames <= ws.cfg.max_batch, "okb_match_motion_stereo_batch");
  OKB_CUDA(cudaSetDevice(ctx->device));
  // device side: mask [frames][kp_cap], dense results, compact lists (own scratch: the camera's m_d staging holds the M1 / M4 results)
  const size_t nq = (size_t)n_frames * n_older * cap0, nm = (size_t)n_frames * n_older * cap_m, n1 = (size_t)n_frames * ws.kp_cap;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t need = al(n1) + al(nq * 4) * 2 + al(nq * 32) + al(nq) + al((size_t)n_frames * n_older * 4) + al(nm * 4) * 2 + al(nm) + al(nm * 32);
  if (need > ws.m3_cap) {
    OKB_CUDA(cudaStreamSynchronize(ws.stream));
    cudaFree(ws.m3_d); if (ws.m3_h) cudaFreeHost(ws.m3_h);
    ws.m3_d = ws.m3_h = nullptr; ws.m3_cap = 0;
    OKB_CUDA(cudaMalloc(&ws.m3_d, need + need / 4));
    OKB_CUDA(cudaMallocHost(&ws.m3_h, need + need / 4));
    ws.m3_cap = need + need / 4;
  }
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t r = o; o += al(bytes); return r; };
  const size_t o_mask = take(n1), o_k1 = take(nq * 4), o_dist = take(nq * 4), o_hp = take(nq * 32), o_fl = take(nq);
  const size_t o_n = take((size_t)n_frames * n_older * 4), o_mk0 = take(nm * 4), o_mk1 = take(nm * 4), o_mf = take(nm), o_mhp = take(nm * 32);
  uint8_t* d = ws.m3_d; uint8_t* h = ws.m3_h;
  cudaStream_t st = ws.stream;
  const int rows = cap < ws.kp_cap ? cap : ws.kp_cap;
  // mask in: the caller's [n_frames][cap] rows -> [n_frames][kp_cap]
  OKB_CUDA(cudaMemsetAsync(d + o_mask, 0, n1, st));
  const bool pin_mask = host_pinned(matched1);
  if (pin_mask) OKB_CUDA(cudaMemcpy2DAsync(d + o_mask, ws.kp_cap, matched1, cap, rows, n_frames, cudaMemcpyHostToDevice, st));
  else {
    for (int b = 0; b < n_frames; b++) memcpy(h + o_mask + (size_t)b * ws.kp_cap, matched1 + (size_t)b * cap, rows);
    OKB_CUDA(cudaMemcpyAsync(d + o_mask, h + o_mask, n1, cudaMemcpyHostToDevice, st));
  }
  int rc = okb_match_motion_stereo_device(ctx, cam, n_frames, T_WC1, T_CW1, n_older, older, cap0, match_threshold, d + o_mask,
                                          (int32_t*)(d + o_k1), (uint32_t*)(d + o_dist), (double*)(d + o_hp), d + o_fl);
  if (rc) return rc;
  k_m3_compact<<<dim3(n_older, n_frames), 256, 0, st>>>(cap0, n_older, cap_m, (const int32_t*)(d + o_k1), (const double*)(d + o_hp), d + o_fl,
                                                       (int32_t*)(d + o_n), (int32_t*)(d + o_mk0), (int32_t*)(d + o_mk1), d + o_mf, (double*)(d + o_mhp));
  ctx->launches++;
  OKB_CUDA(cudaGetLastError());
  // results: the counts first (one extra synchronisation of a few microseconds), then only the filled part of the compact lists
  // (capacity cap_m per (frame, view); a view typically fills a fraction of it) and the mask, through the pinned mirror or straight
  // into page-locked caller buffers
  const bool pin_n = host_pinned(n_match);
  OKB_CUDA(cudaMemcpyAsync(pin_n ? (void*)n_match : (void*)(h + o_n), d + o_n, (size_t)n_frames * n_older * 4, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(wait_stream(ctx, st));
  if (!pin_n) memcpy(n_match, h + o_n, (size_t)n_frames * n_older * 4);
  int max_cnt = 0;
  for (int i = 0; i < n_frames * n_older; i++) {
    if (n_match[i] > cap_m) { set_error("okb_match_motion_stereo_batch: %d matches of view %d exceed the list capacity %d", n_match[i], i, cap_m); return OKB_ERR_CAPACITY; }
    max_cnt = n_match[i] > max_cnt ? n_match[i] : max_cnt;
  }
  const size_t list_rows = (size_t)n_frames * n_older;
  auto out = [&](size_t off, void* dst, size_t elem) -> int {   // 1: landed in the mirror, 0: in the caller's buffer
    if (max_cnt == 0) return 0;
    const bool pinned = host_pinned(dst);
    OKB_CUDA(cudaMemcpy2DAsync(pinned ? dst : (void*)(h + off), (size_t)cap_m * elem, d + off, (size_t)cap_m * elem, (size_t)max_cnt * elem, list_rows,
                               cudaMemcpyDeviceToHost, st));
    return pinned ? 0 : 1;
  };
  int s_k0, s_k1, s_f, s_hp;
  if ((s_k0 = out(o_mk0, m_k0, 4)) < 0 || (s_k1 = out(o_mk1, m_k1, 4)) < 0 || (s_f = out(o_mf, m_flags, 1)) < 0 || (s_hp = out(o_mhp, m_hp_W, 32)) < 0)
    return OKB_ERR_CUDA;
  if (pin_mask) OKB_CUDA(cudaMemcpy2DAsync(matched1, cap, d + o_mask, ws.kp_cap, rows, n_frames, cudaMemcpyDeviceToHost, st));
  else OKB_CUDA(cudaMemcpyAsync(h + o_mask, d + o_mask, n1, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(wait_stream(ctx, st));
  auto trim = [&](int staged, size_t off, void* dst, size_t elem) {
    if (!staged) return;
    for (size_t r = 0; r < list_rows; r++) memcpy((uint8_t*)dst + r * cap_m * elem, h + off + r * cap_m * elem, (size_t)n_match[r] * elem);
  };
  trim(s_k0, o_mk0, m_k0, 4); trim(s_k1, o_mk1, m_k1, 4); trim(s_f, o_mf, m_flags, 1); trim(s_hp, o_mhp, m_hp_W, 32);
  if (!pin_mask) for (int b = 0; b < n_frames; b++) memcpy(matched1 + (size_t)b * cap, h + o_mask + (size_t)b * ws.kp_cap, rows);
  return OKB_OK;
}

// ---- host-buffer batch forms: the queries are the features the last okb_detect_describe[_batch] of the camera left on the
// device, everything else comes from / goes to caller memory (page-locked buffers are used by the copy engines directly)
namespace {
struct CamStage {
  CamWorkspace& ws; cudaStream_t st; size_t off = 0;
  int reserve(size_t bytes)
  {
    if (bytes <= ws.m_cap) return OKB_OK;
    OKB_CUDA(cudaStreamSynchronize(st));
    cudaFree(ws.m_d); if (ws.m_h) cudaFreeHost(ws.m_h);
    ws.m_d = ws.m_h = nullptr; ws.m_cap = 0;
    const size_t cap = bytes + bytes / 4 + (1 << 16);
    OKB_CUDA(cudaMalloc(&ws.m_d, cap));
    OKB_CUDA(cudaMallocHost(&ws.m_h, cap));
    ws.m_cap = cap;
    return OKB_OK;
  }
  size_t take(size_t bytes) { off = (off + 255) & ~(size_t)255; const size_t o = off; off += bytes; return o; }
  int upload(size_t o, const void* src, size_t bytes)
  {
    if (!bytes) return OKB_OK;
    if (host_pinned(src)) { OKB_CUDA(cudaMemcpyAsync(ws.m_d + o, src, bytes, cudaMemcpyHostToDevice, st)); return OKB_OK; }
    memcpy(ws.m_h + o, src, bytes);
    OKB_CUDA(cudaMemcpyAsync(ws.m_d + o, ws.m_h + o, bytes, cudaMemcpyHostToDevice, st));
    return OKB_OK;
  }
  // rows of `row_bytes` (device stride src_rows * elem) into the caller's [n_frames][cap] array
  int download(size_t o, void* dst, size_t elem, int rows, int src_rows, int cap, int n_frames, bool* staged)
  {
    *staged = !host_pinned(dst);
    if (!*staged) {
      OKB_CUDA(cudaMemcpy2DAsync(dst, (size_t)cap * elem, ws.m_d + o, (size_t)src_rows * elem, (size_t)rows * elem, n_frames, cudaMemcpyDeviceToHost, st));
    } else {
      OKB_CUDA(cudaMemcpyAsync(ws.m_h + o, ws.m_d + o, (size_t)src_rows * elem * n_frames, cudaMemcpyDeviceToHost, st));
    }
    return OKB_OK;
  }
  void unstage(size_t o, void* dst, size_t elem, int rows, int src_rows, int cap, int n_frames)
  {
    for (int b = 0; b < n_frames; b++)
      memcpy((uint8_t*)dst + (size_t)b * cap * elem, ws.m_h + o + (size_t)b * src_rows * elem, (size_t)rows * elem);
  }
};
}  // namespace

int okb_match_map3d_batch(okb_context_t* ctx, int cam, int D, int n_frames, int n_cand, const uint8_t* cand_desc, const int32_t* cand_lm,
                          int n_lm, const double* lm_proj, const uint8_t* lm_is3d, double reprojection_threshold,
                          uint32_t match_threshold, int cap, uint32_t* out_dist, int32_t* out_lm)
{
  OKB_CHECK_ARGS(ctx && cam >= 0 && cam < ctx->n_cams && n_cand >= 0 && n_lm >= 0 && cap > 0 && out_dist && out_lm, "okb_match_map3d_batch");
  CamWorkspace& ws = ctx->cams[cam];
  OKB_CHECK_ARGS(n_frames >= 1 && n_frames <= ws.cfg.max_batch, "okb_match_map3d_batch");
  OKB_CHECK_ARGS(n_cand == 0 || (cand_desc && cand_lm && lm_proj && lm_is3d), "okb_match_map3d_batch");
  OKB_CUDA(cudaSetDevice(ctx->device));
  CamStage S{ws, ws.stream};
  OKB_CHECK_ARGS(!bad_D(D), "okb_match_map3d_batch");
  { int rcv = check_pool(n_cand, cand_lm, n_lm, "okb_match_map3d_batch"); if (rcv) return rcv; }
  const size_t o_desc = S.take((size_t)n_cand * D), o_lm = S.take((size_t)n_cand * 4);
  const size_t o_proj = S.take((size_t)n_frames * n_lm * 16), o_3d = S.take((size_t)n_lm);
  const size_t o_dist = S.take((size_t)n_frames * ws.kp_cap * 4), o_idx = S.take((size_t)n_frames * ws.kp_cap * 4);
  int rc = S.reserve(S.off);
  if (rc) return rc;
  if ((rc = S.upload(o_desc, cand_desc, (size_t)n_cand * D)) || (rc = S.upload(o_lm, cand_lm, (size_t)n_cand * 4)) ||
      (rc = S.upload(o_proj, lm_proj, (size_t)n_frames * n_lm * 16)) || (rc = S.upload(o_3d, lm_is3d, (size_t)n_lm)))
    return rc;
  rc = okb_match_map3d_device(ctx, cam, D, n_frames, n_cand, ws.m_d + o_desc, (const int32_t*)(ws.m_d + o_lm), n_lm,
                              (const double*)(ws.m_d + o_proj), ws.m_d + o_3d, reprojection_threshold, match_threshold,
                              (uint32_t*)(ws.m_d + o_dist), (int32_t*)(ws.m_d + o_idx));
  if (rc) return rc;
  const int rows = cap < ws.kp_cap ? cap : ws.kp_cap;
  bool s0, s1;
  if ((rc = S.download(o_dist, out_dist, 4, rows, ws.kp_cap, cap, n_frames, &s0)) || (rc = S.download(o_idx, out_lm, 4, rows, ws.kp_cap, cap, n_frames, &s1)))
    return rc;
  OKB_CUDA(wait_stream(ctx, ws.stream));
  if (s0) S.unstage(o_dist, out_dist, 4, rows, ws.kp_cap, cap, n_frames);
  if (s1) S.unstage(o_idx, out_lm, 4, rows, ws.kp_cap, cap, n_frames);
  return OKB_OK;
}

int okb_match_stereo_batch(okb_context_t* ctx, int cam0, int cam1, int n_frames, const double C_WC0[9], const double r_WC0[3],
                           const double C_WC1[9], const double r_WC1[3], uint32_t match_threshold, int cap, int32_t* out_k1,
                           uint32_t* out_dist, double* out_hp_W, uint8_t* out_initialisable)
{
  OKB_CHECK_ARGS(ctx && cam0 >= 0 && cam0 < ctx->n_cams && cap > 0 && out_k1 && out_dist && out_hp_W && out_initialisable, "okb_match_stereo_batch");
  CamWorkspace& ws = ctx->cams[cam0];
  OKB_CHECK_ARGS(n_frames >= 1 && n_frames <= ws.cfg.max_batch, "okb_match_stereo_batch");
  OKB_CUDA(cudaSetDevice(ctx->device));
  CamStage S{ws, ws.stream};
  const size_t n = (size_t)n_frames * ws.kp_cap;
  const size_t o_k1 = S.take(n * 4), o_dist = S.take(n * 4), o_hp = S.take(n * 32), o_init = S.take(n);
  int rc = S.reserve(S.off);
  if (rc) return rc;
  rc = okb_match_stereo_device(ctx, cam0, cam1, n_frames, C_WC0, r_WC0, C_WC1, r_WC1, match_threshold, (int32_t*)(ws.m_d + o_k1),
                               (uint32_t*)(ws.m_d + o_dist), (double*)(ws.m_d + o_hp), ws.m_d + o_init);
  if (rc) return rc;
  const int rows = cap < ws.kp_cap ? cap : ws.kp_cap;
  bool s[4];
  if ((rc = S.download(o_k1, out_k1, 4, rows, ws.kp_cap, cap, n_frames, &s[0])) || (rc = S.download(o_dist, out_dist, 4, rows, ws.kp_cap, cap, n_frames, &s[1])) ||
      (rc = S.download(o_hp, out_hp_W, 32, rows, ws.kp_cap, cap, n_frames, &s[2])) ||
      (rc = S.download(o_init, out_initialisable, 1, rows, ws.kp_cap, cap, n_frames, &s[3])))
    return rc;
  OKB_CUDA(wait_stream(ctx, ws.stream));
  if (s[0]) S.unstage(o_k1, out_k1, 4, rows, ws.kp_cap, cap, n_frames);
  if (s[1]) S.unstage(o_dist, out_dist, 4, rows, ws.kp_cap, cap, n_frames);
  if (s[2]) S.unstage(o_hp, out_hp_W, 32, rows, ws.kp_cap, cap, n_frames);
  if (s[3]) S.unstage(o_init, out_initialisable, 1, rows, ws.kp_cap, cap, n_frames);
  return OKB_OK;
}

int okb_match_place(okb_context_t* ctx, int D, int n_lm, const int32_t* lm_offsets, const uint8_t* lm_desc, int n_kp,
                    const uint8_t* kp_desc, uint32_t match_threshold, int32_t* out_k, uint32_t* out_dist)
{
  OKB_CHECK_ARGS(ctx && !bad_D(D) && n_lm >= 0 && n_kp >= 0 && n_kp < (1 << 20) && out_k && out_dist, "okb_match_place");
  OKB_CHECK_ARGS(n_lm == 0 || (lm_offsets && lm_desc), "okb_match_place");
  if (n_lm == 0) return OKB_OK;
  const int n_desc = lm_offsets[n_lm];
  for (int i = 0; i < n_lm; i++) OKB_CHECK_ARGS(lm_offsets[i + 1] >= lm_offsets[i] && lm_offsets[i + 1] - lm_offsets[i] < 4096, "okb_match_place");
  OKB_CUDA(cudaSetDevice(ctx->device));
  OKB_LOCK_SLOT;
  cudaStream_t st = MW.stream;
  const int32_t* d_off = nullptr; const uint8_t *d_ld = nullptr, *d_kd = nullptr; int32_t* d_ok = nullptr; uint32_t* d_od = nullptr;
  size_t o_k = 0, o_d = 0, in_end = 0, o_end = 0;
  for (int pass = 0; pass < 2; pass++) {
    Arena A; A.ctx = ctx; A.dry = pass == 0; A.h = (uint8_t*)MW.h_buf; A.d = (uint8_t*)MW.d_buf;
    d_off = A.in(lm_offsets, (size_t)n_lm + 1); d_ld = A.in(lm_desc, (size_t)n_desc * D); d_kd = A.in(kp_desc, (size_t)n_kp * D);
    in_end = A.off;
    d_ok = A.out<int32_t>(n_lm, &o_k); d_od = A.out<uint32_t>(n_lm, &o_d); o_end = A.off;
    if (pass == 0) { int rc = ensure(ctx, A.off); if (rc) return rc; }
  }
  uint8_t* h = (uint8_t*)MW.h_buf; uint8_t* d = (uint8_t*)MW.d_buf;
  OKB_CUDA(cudaMemcpyAsync(d, h, in_end, cudaMemcpyHostToDevice, st));
  const int grid = (n_lm + 7) / 8;
  if (D == 64) k_match_place<4><<<grid, 256, 0, st>>>(n_lm, d_off, d_ld, n_kp, d_kd, match_threshold, d_ok, d_od);
  else k_match_place<3><<<grid, 256, 0, st>>>(n_lm, d_off, d_ld, n_kp, d_kd, match_threshold, d_ok, d_od);
  ctx->launches++;
  OKB_CUDA(cudaGetLastError());
  OKB_CUDA(cudaMemcpyAsync(h + o_k, d + o_k, o_end - o_k, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(cudaStreamSynchronize(st));
  memcpy(out_k, h + o_k, (size_t)n_lm * 4); memcpy(out_dist, h + o_d, (size_t)n_lm * 4);
  return OKB_OK;
}

int okb_hamming_matrix(okb_context_t* ctx, int D, int n_a, const uint8_t* a, int n_b, const uint8_t* b, uint16_t* out_dist)
{
  OKB_CHECK_ARGS(ctx && !bad_D(D) && n_a >= 0 && n_b >= 0 && out_dist, "okb_hamming_matrix");
  if (n_a == 0 || n_b == 0) return OKB_OK;
  OKB_CHECK_ARGS(a && b, "okb_hamming_matrix");
  OKB_CUDA(cudaSetDevice(ctx->device));
  OKB_LOCK_SLOT;
  cudaStream_t st = MW.stream;
  const uint8_t *da = nullptr, *db = nullptr; uint16_t* dout = nullptr; size_t o_out = 0, in_end = 0, o_end = 0;
  for (int pass = 0; pass < 2; pass++) {
    Arena A; A.ctx = ctx; A.dry = pass == 0; A.h = (uint8_t*)MW.h_buf; A.d = (uint8_t*)MW.d_buf;
    da = A.in(a, (size_t)n_a * D); db = A.in(b, (size_t)n_b * D); in_end = A.off;
    dout = A.out<uint16_t>((size_t)n_a * n_b, &o_out); o_end = A.off;
    if (pass == 0) { int rc = ensure(ctx, A.off); if (rc) return rc; }
  }
  uint8_t* h = (uint8_t*)MW.h_buf; uint8_t* d = (uint8_t*)MW.d_buf;
  OKB_CUDA(cudaMemcpyAsync(d, h, in_end, cudaMemcpyHostToDevice, st));
  k_hamming_matrix<<<dim3((n_b + 255) / 256, n_a), 256, 0, st>>>(D / 16, n_a, da, n_b, db, dout);
  ctx->launches++;
  OKB_CUDA(cudaGetLastError());
  OKB_CUDA(cudaMemcpyAsync(h + o_out, d + o_out, o_end - o_out, cudaMemcpyDeviceToHost, st));
  OKB_CUDA(cudaStreamSynchronize(st));
  memcpy(out_dist, h + o_out, (size_t)n_a * n_b * 2);
  return OKB_OK;
}

}  // extern "C"
