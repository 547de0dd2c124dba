// umma_probe.cu -- stand-alone check of the tcgen05 Hamming tile (the building block of k_scan_umma): one CTA expands 128 query
// and 128 candidate 512-bit descriptors into u8 bit planes in shared memory, issues 16 tcgen05.mma kind::i8 (M = N = 128, K = 32)
// into TMEM and reads the 128 x 128 accumulators back; the host compares 16384 * popc(a & b) with the CPU. Every wait is bounded.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../okvis2_b200/csrc/okb_umma.h"
using namespace okb::umma;

constexpr int kRows = 128, kChunkBytes = kRows * 16;   // one 16-byte K chunk of all 128 rows: 8-row core matrices, 128 B each

// descriptor row r (16 words) -> bit planes: K byte 32 j + 4 i + b  <->  bit 8 b + i of word j, value 128
__device__ __forceinline__ void expand_row(uint8_t* tile, int r, const uint32_t* w16)
{
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const uint32_t w = w16[j];
    uint4 lo, hi;
    lo.x = (w << 7) & 0x80808080u; lo.y = (w << 6) & 0x80808080u; lo.z = (w << 5) & 0x80808080u; lo.w = (w << 4) & 0x80808080u;
    hi.x = (w << 3) & 0x80808080u; hi.y = (w << 2) & 0x80808080u; hi.z = (w << 1) & 0x80808080u; hi.w = w & 0x80808080u;
    uint8_t* p = tile + (size_t)(2 * j) * kChunkBytes + (r >> 3) * 128 + (r & 7) * 16;
    *reinterpret_cast<uint4*>(p) = lo;
    *reinterpret_cast<uint4*>(p + kChunkBytes) = hi;
  }
}

__global__ void __launch_bounds__(160) k_probe(const uint32_t* A, const uint32_t* B, int32_t* D, int* failed_out, int reps, long long* cycles)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem; uint8_t* sb = smem + 65536;
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base;
  __shared__ int failed;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { bar_init(&bar_done, 1); bar_init_fence(); failed = 0; }
  if (warp == 4) tmem_alloc(&tmem_base, 128);
  if (threadIdx.x < 128) {
    expand_row(sa, threadIdx.x, A + (size_t)threadIdx.x * 16);
    expand_row(sb, threadIdx.x, B + (size_t)threadIdx.x * 16);
    fence_smem_to_async();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t td = tmem_base;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = idesc_u8(128, 128);
#pragma unroll 1
    for (int s = 0; s < 16; s++) {
      const uint64_t da = smem_desc(smem_addr(sa) + s * 2 * kChunkBytes, kChunkBytes, 128);
      const uint64_t db = smem_desc(smem_addr(sb) + s * 2 * kChunkBytes, kChunkBytes, 128);
      mma_u8(td, da, db, idesc, s > 0);
    }
    mma_commit(&bar_done);
    if (reps > 0) {
      // issue-rate measurement: reps more tiles of 16 MMAs back to back (same operands, accumulate), one commit at the end
      __shared__ uint64_t bar_rate;
      bar_init(&bar_rate, 1); bar_init_fence();
      const long long t0 = clock64();
      for (int i = 0; i < reps; i++)
#pragma unroll 1
        for (int s = 0; s < 16; s++)
          mma_u8(td, smem_desc(smem_addr(sa) + s * 2 * kChunkBytes, kChunkBytes, 128), smem_desc(smem_addr(sb) + s * 2 * kChunkBytes, kChunkBytes, 128), idesc, 1);
      mma_commit(&bar_rate);
      bar_wait(&bar_rate, 0, &failed);
      *cycles = clock64() - t0;
    }
  }
  if (warp < 4) {
    bar_wait(&bar_done, 0, &failed);
    fence_after_sync();
    if (!atomicAdd(&failed, 0)) {
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(td + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) D[(size_t)(warp * 32 + lane) * 128 + c0 + j] = (int32_t)v[j];
      }
    }
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(td, 128);
  if (threadIdx.x == 0) *failed_out = atomicAdd(&failed, 0);
}

int main()
{
  std::vector<uint32_t> A(128 * 16), B(128 * 16);
  srand(7);
  for (auto& x : A) x = (uint32_t)rand() ^ ((uint32_t)rand() << 16);
  for (auto& x : B) x = (uint32_t)rand() ^ ((uint32_t)rand() << 16);
  for (int i = 0; i < 16; i++) { B[5 * 16 + i] = A[9 * 16 + i]; }          // an identical pair
  for (int i = 0; i < 16; i++) { A[100 * 16 + i] = 0xffffffffu; B[77 * 16 + i] = 0xffffffffu; }   // all ones
  for (int i = 0; i < 16; i++) { A[3 * 16 + i] = 0; }
  uint32_t *dA, *dB; int32_t* dD; int* dF; long long* dC; cudaMalloc(&dC, 8);
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * 128 * 4); cudaMalloc(&dF, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, 128 * 128 * 4);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  k_probe<<<1, 160, 131072>>>(dA, dB, dD, dF, 0, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<int32_t> D(128 * 128); int failed = 0;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&failed, dF, 4, cudaMemcpyDeviceToHost);
  if (failed) { printf("kernel reported a timed-out wait\n"); return 3; }
  long bad = 0;
  for (int r = 0; r < 128; r++)
    for (int c = 0; c < 128; c++) {
      int pc = 0;
      for (int i = 0; i < 16; i++) pc += __builtin_popcount(A[r * 16 + i] & B[c * 16 + i]);
      if (D[r * 128 + c] != pc * 16384) { if (bad < 8) printf("mismatch r %d c %d: got %d want %d (popc %d)\n", r, c, D[r * 128 + c], pc * 16384, pc); bad++; }
    }
  printf("umma_probe: %ld mismatches of 16384; D[9][5] = %d (512 * 16384 = %d)\n", bad, D[9 * 128 + 5], 512 * 16384);
  // issue rate of tcgen05.mma kind::i8 M = N = 128, K = 32 on one SM
  const int reps = 2000;
  k_probe<<<1, 160, 131072>>>(dA, dB, dD, dF, reps, dC);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error in the rate run\n"); return 2; }
  long long cyc = 0; cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  printf("umma_probe: %d x 16 MMAs (128 x 128 x 32, u8) in %lld cycles: %.1f cycles per MMA, %.0f cycles per 128 x 128 x 512 tile, %.0f MAC/clk/SM\n",
         reps, cyc, (double)cyc / (reps * 16.0), (double)cyc / reps, 128.0 * 128 * 32 * 16 * reps / (double)cyc);
  return bad ? 1 : 0;
}
