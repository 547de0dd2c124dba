// probe: which box shapes / coordinates does cp.async.bulk.tensor.3d accept for a u8 image (pitch 752/768, 3-D map)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
#include <cstring>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct alignas(64) Maps { CUtensorMap m[8]; int use[8]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BW, int BH>
__global__ void k(const __grid_constant__ Maps maps, int layer, int x, int y, int z, uint8_t* out)
{
  __shared__ __align__(128) uint8_t tile[BH][BW];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(&tile[0][0])), "l"(reinterpret_cast<uint64_t>(&maps.m[layer])), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
  }
  uint32_t ok = 0; long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (clock64() - t0 > 100000000ll) break;
  }
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = ok ? tile[i / BW][i % BW] : 0xEE;
}
template <int BW, int BH>
__global__ void kg(const CUtensorMap* __restrict__ mp, int x, int y, int z, uint8_t* out)
{
  __shared__ __align__(128) uint8_t tile[BH][BW];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(&tile[0][0])), "l"(reinterpret_cast<uint64_t>(mp)), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
  }
  uint32_t ok = 0; long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (clock64() - t0 > 100000000ll) break;
  }
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = ok ? tile[i / BW][i % BW] : 0xEE;
}
template <int BW, int BH>
__global__ void k1(const __grid_constant__ CUtensorMap map, int x, int y, int z, uint8_t* out)
{
  __shared__ __align__(128) uint8_t tile[BH][BW];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(&tile[0][0])), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
  }
  uint32_t ok = 0; long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (clock64() - t0 > 100000000ll) break;
  }
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = ok ? tile[i / BW][i % BW] : 0xEE;
}
static int g_mode = 0;
template <int BW, int BH> void run(EncodeTiledFn enc, uint8_t* d_img, int W, int H, int pitch, int frames, const std::vector<uint8_t>& h, int x, int y, int z)
{
  Maps maps; memset(&maps, 0, sizeof(maps));
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
  const cuuint32_t box[3] = {BW, BH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&maps.m[3], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %dx%d W=%d pitch=%d at (%d,%d,%d): encode=%d ", BW, BH, W, pitch, x, y, z, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return; }
  uint8_t* d_out; cudaMalloc(&d_out, BW * BH);
  if (g_mode == 0) k<BW, BH><<<1, 128>>>(maps, 3, x, y, z, d_out);
  else if (g_mode == 1) { CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap)); cudaMemcpy(dm, &maps.m[3], sizeof(CUtensorMap), cudaMemcpyHostToDevice); kg<BW, BH><<<1, 128>>>(dm, x, y, z, d_out); }
  else k1<BW, BH><<<1, 128>>>(maps.m[3], x, y, z, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run=%s ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<uint8_t> o(BW * BH); cudaMemcpy(o.data(), d_out, BW * BH, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < BH; r2++) for (int c = 0; c < BW; c++) {
      int yy = y + r2, xx = x + c; uint8_t want = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? h[((size_t)z * H + yy) * pitch + xx] : 0;
      bad += o[r2 * BW + c] != want;
    }
    printf("mismatches=%d", bad);
  }
  printf("\n");
  cudaFree(d_out);
}
int main(int argc, char** argv)
{
  g_mode = argc > 1 ? atoi(argv[1]) : 0;
  const int only = argc > 2 ? atoi(argv[2]) : -1;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int W = 752, H = 480, frames = 2;
  for (int pitch : {768}) {
    std::vector<uint8_t> h((size_t)pitch * H * frames);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 24);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    const int xs[7] = {-8, -16, 0, 8, 16, 376, 384};
    const int ys[3] = {-3, 0, 32};
    int id = 0;
    for (int xi = 0; xi < 7; xi++) for (int yi = 0; yi < 3; yi++) {
      if (only == id) run<80, 38>(enc, d, W, H, pitch, frames, h, xs[xi], ys[yi], 0);
      if (only == id + 100) run<80, 32>(enc, d, W, H, pitch, frames, h, xs[xi], ys[yi], 0);
      if (only == id + 200) run<64, 38>(enc, d, W, H, pitch, frames, h, xs[xi], ys[yi], 0);
      id++;
    }
    cudaFree(d);
  }
  return 0;
}
