// throughput microbenchmark: which pipe / rate for packed min/max candidates on sm_100a
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
constexpr int ITERS = 4096;
template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
  if constexpr (OP == 0) return __vimin3_u16x2(a, b, c);
  if constexpr (OP == 1) return __vminu2(a, b);
  if constexpr (OP == 2) return min(a, b);
  if constexpr (OP == 3) { __half2 r = __hmin2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b)); return *reinterpret_cast<uint32_t*>(&r); }
  if constexpr (OP == 4) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c))); return __float_as_uint(r); }
  if constexpr (OP == 5) return __byte_perm(a, b, 0x4140) ;
  if constexpr (OP == 6) { uint32_t r; asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if constexpr (OP == 7) return __vminu4(a, b);
  if constexpr (OP == 8) return __vimin3_u32(a, b, c);
  if constexpr (OP == 9) { uint32_t r; asm volatile("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  if constexpr (OP == 10) return a * b + c;   // IMAD
  if constexpr (OP == 11) return __vimax3_s16x2_relu(a, b, c);
  if constexpr (OP == 12) return (uint32_t)__popc(a ^ b) + c;          // POPC.b32 (+ 1 IADD/LOP on another pipe)
  if constexpr (OP == 13) return (uint32_t)__popcll(((unsigned long long)(a ^ b) << 32) | (a ^ c));   // popc.b64 = 2 POPC + add
  if constexpr (OP == 14) return __viaddmax_s16x2_relu(a, b, c);
  if constexpr (OP == 15) return __vadd2(a, b);
  return 0;
}
template <int OP, int OP2> __global__ void k(uint32_t* out, uint32_t seed, long long* cyc)
{
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
  uint32_t b = seed ^ 0x64136427u, c = seed + 0x64556401u;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP2 < 0 || (i & 1) == 0) r[i] = op<OP>(r[i], b, c); else r[i] = op<OP2 < 0 ? 0 : OP2>(r[i], b, c);
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
static double g_last_rate = 0;
template <int OP, int OP2> void run(const char* name, uint32_t* d, long long* dc)
{
  const int threads = 1024;   // 32 warps / SM = 8 per SMSP
  k<OP, OP2><<<148, threads>>>(d, 12345u, dc);
  cudaDeviceSynchronize();
  k<OP, OP2><<<148, threads>>>(d, 12345u, dc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  double warp_instr_per_sm = (double)ITERS * 8 * (threads / 32);
  g_last_rate = warp_instr_per_sm / c;
  printf("%-34s cycles %8lld  warp-instr/clk/SM %.3f  (cycles per warp-instr per SMSP %.2f)  %s\n", name, c, warp_instr_per_sm / c,
         c / (warp_instr_per_sm / 4), cudaGetErrorString(cudaGetLastError()));
}
int main()
{
  uint32_t* d; long long* dc; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&dc, 8);
  run<0, -1>("vimin3_u16x2", d, dc);
  run<1, -1>("vminu2", d, dc);
  run<2, -1>("min.u32", d, dc);
  run<8, -1>("vimin3_u32", d, dc);
  run<3, -1>("hmin2", d, dc);
  run<9, -1>("min.bf16x2", d, dc);
  run<4, -1>("min.f32 3-input", d, dc);
  run<5, -1>("prmt", d, dc);
  run<6, -1>("hfma2.relu", d, dc);
  run<7, -1>("vminu4", d, dc);
  run<10, -1>("imad", d, dc);
  run<11, -1>("vimax3_s16x2_relu", d, dc);
  run<0, 3>("vimin3_u16x2 + hmin2 (1:1)", d, dc);
  run<0, 6>("vimin3_u16x2 + hfma2.relu (1:1)", d, dc);
  run<0, 5>("vimin3_u16x2 + prmt (1:1)", d, dc);
  run<0, 10>("vimin3_u16x2 + imad (1:1)", d, dc);
  run<3, 6>("hmin2 + hfma2.relu (1:1)", d, dc);
  run<0, 4>("vimin3_u16x2 + fmnmx3 (1:1)", d, dc);
  run<14, -1>("viaddmax_s16x2_relu", d, dc);
  run<15, -1>("vadd2 (VIADD.16x2)", d, dc);
  run<12, -1>("popc.b32 (+ iadd)", d, dc);
  const double popc32 = g_last_rate;   // warp instructions (each = one POPC) per clk per SM
  run<13, -1>("popcll (2 x popc.b32 + add)", d, dc);
  // the matchers' roofline denominator: POPC.b32 lanes per clock per SM (written where bench.py reads it)
  FILE* f = fopen("gpurun_out/popc_rate.json", "w");
  if (f) {
    fprintf(f, "{\"popc_b32_lanes_per_clk_per_sm\": %.3f, \"popc_b32_warp_instr_per_clk_per_sm\": %.4f, \"popcll_warp_instr_per_clk_per_sm\": %.4f, "
               "\"how\": \"bench/ubench_pipes.cu on the B200: 148 x 1024 threads, 8 independent chains per thread, clock64 around 4096 iterations\"}\n",
            popc32 * 32, popc32, g_last_rate);
    fclose(f);
  }
  return 0;
}
