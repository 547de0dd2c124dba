// okb_tables.cu -- uploads the static look-up tables (okb_tables.h) at okb_create.
#include "okb_internal.h"
#include "okb_tables.h"

namespace okb {

int tables_init(okb_context* ctx, float pattern_scale)
{
  ctx->pattern_scale = pattern_scale;
  HostTables T;
  if (!build_host_tables(pattern_scale, T)) {
    set_error("pattern_scale %.3f does not give the 512 short / 870 long pair BRISK pattern", pattern_scale);
    return OKB_ERR_UNSUPPORTED;
  }
  static_assert(sizeof(LongPair) == sizeof(int4), "long pair layout");
  OKB_CUDA(cudaMalloc(&ctx->d_pattern, T.pattern.size() * sizeof(PatternPoint)));
  OKB_CUDA(cudaMemcpy(ctx->d_pattern, T.pattern.data(), T.pattern.size() * sizeof(PatternPoint), cudaMemcpyHostToDevice));
  OKB_CUDA(cudaMalloc(&ctx->d_short_pairs, T.short_pairs.size() * 4));
  OKB_CUDA(cudaMemcpy(ctx->d_short_pairs, T.short_pairs.data(), T.short_pairs.size() * 4, cudaMemcpyHostToDevice));
  OKB_CUDA(cudaMalloc(&ctx->d_long_pairs, T.long_pairs.size() * sizeof(int4)));
  OKB_CUDA(cudaMemcpy(ctx->d_long_pairs, T.long_pairs.data(), T.long_pairs.size() * sizeof(int4), cudaMemcpyHostToDevice));
  OKB_CUDA(cudaMalloc(&ctx->d_scale_bounds, T.scale_bounds.size() * 4));
  OKB_CUDA(cudaMemcpy(ctx->d_scale_bounds, T.scale_bounds.data(), T.scale_bounds.size() * 4, cudaMemcpyHostToDevice));
  OKB_CUDA(cudaMalloc(&ctx->d_size_list, T.size_list.size() * 4));
  OKB_CUDA(cudaMemcpy(ctx->d_size_list, T.size_list.data(), T.size_list.size() * 4, cudaMemcpyHostToDevice));
  for (int i = 0; i < kPoints; i++) ctx->h_pat0[i] = T.pattern[i];
  for (int i = 0; i < kScales; i++) ctx->h_size_list[i] = T.size_list[i];
  ctx->n_short = (int)T.short_pairs.size(); ctx->n_long = (int)T.long_pairs.size();
  return OKB_OK;
}

void tables_free(okb_context* ctx)
{
  cudaFree(ctx->d_pattern); cudaFree(ctx->d_short_pairs); cudaFree(ctx->d_long_pairs);
  cudaFree(ctx->d_scale_bounds); cudaFree(ctx->d_size_list);
  ctx->d_pattern = nullptr;
}

}  // namespace okb
