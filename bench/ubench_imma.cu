// ubench_imma.cu -- issue rate of the legacy tensor path for 8-bit integers (mma.sync.m16n8k32.s8 -> IMMA.16832) on sm_100a:
// how many +-1 dot products of 512 terms (= Hamming distances of 512-bit descriptors) per second it could deliver, next to the
// POPC-bound scan (bench/ubench_pipes.cu). Writes gpurun_out/imma_rate.json.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <int ACC>
__global__ void __launch_bounds__(256) k_imma(int iters, int* out)
{
  uint32_t a0 = threadIdx.x * 0x01010101u, a1 = a0 ^ 0x7f7f7f7fu, a2 = a0 + 0x01000100u, a3 = a1 + 0x00010001u;
  uint32_t b0 = blockIdx.x * 0x01010101u + 1u, b1 = b0 ^ 0x55aa55aau;
  int c[ACC][4];
#pragma unroll
  for (int i = 0; i < ACC; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ACC; i++)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0 + i), "r"(b1));
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < ACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  int* out; cudaMalloc(&out, sizeof(int) * sms * 8 * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  double best = 0;
  for (int ctas = 1; ctas <= 8; ctas *= 2) {
    k_imma<8><<<sms * ctas, 256>>>(100, out);
    cudaEventRecord(e0);
    k_imma<8><<<sms * ctas, 256>>>(iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)sms * ctas * 8 /*warps*/ * 8 /*ACC*/ * iters;
    const double macs = mmas * 16 * 8 * 32;
    const double tops = 2 * macs / (ms * 1e-3) / 1e12;
    printf("ctas/SM %d: %.3f ms, %.1f TOPS (int8 dense), %.1f G 512-term dot products/s, %.2f IMMA.16832/clk/SM at %d MHz\n", ctas, ms, tops,
           macs / 512 / (ms * 1e-3) / 1e9, mmas / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
    if (tops > best) best = tops;
  }
  if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  FILE* f = fopen("gpurun_out/imma_rate.json", "w");
  if (f) { fprintf(f, "{\"imma_s8_tops\": %.1f, \"dot512_per_s\": %.4g, \"sms\": %d}\n", best, best * 1e12 / 2 / 512, sms); fclose(f); }
  return 0;
}
