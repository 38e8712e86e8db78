// magvit_ops.h -- memory-bound kernels of the MAGViT-v2 pixel decoder, NHWC bf16.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace ugk {

// ids int64 [n] -> z bf16 [n, dim] (= NHWC [B,h,w,dim]): +-1, channel c = bit (dim-1-c) of the index
cudaError_t launch_lfq_unpack(const int64_t* ids, void* z_bf16, int64_t n, int dim, cudaStream_t stream);

// LFQ lookup + post_quant_conv (1x1, dim -> dim, w fp32 [dim, dim], bias fp32 [dim] or null) in one pass; z bf16 [n, cpad], channels
// >= dim zero
cudaError_t launch_lfq_postquant(const int64_t* ids, const float* w, const float* bias, void* z_bf16, int64_t n, int dim, int cpad,
                                 cudaStream_t stream);

// y = softmax(scale * x) per row over the first cols_valid columns (the others: zero), bf16 [rows, cols] in and out (fp32 inside);
// cols % 8 == 0, cols_valid % 8 == 0
cudaError_t launch_softmax_rows(const void* x_bf16, void* y_bf16, int64_t rows, int cols, int cols_valid, float scale, cudaStream_t stream);

// GroupNorm(32 groups, eps) (+ swish) over x bf16 NHWC [B, HW, C]; partial: fp32 [B, chunks, 32, 2] scratch, stats: fp32 [B, 32, 2]
size_t groupnorm_workspace_floats(int B, int64_t HW, int C);
// in_bias (nullable, bf16 [C]) is added to x first: the bias of the convolution that produced x
cudaError_t launch_groupnorm_nhwc(const void* x_bf16, const void* in_bias_bf16, const void* gamma_bf16, const void* beta_bf16, void* y_bf16,
                                  float* workspace, int B, int64_t HW, int C, float eps, int swish, int stats_ready, cudaStream_t stream);

// bias_add that also leaves the GroupNorm partial sums of `out` in `workspace` (same layout as the statistics pass): the next
// launch_groupnorm_nhwc on `out` may then be called with stats_ready = 1.  Same shape limits as launch_groupnorm_nhwc.
cudaError_t launch_bias_add_stats_nhwc(const void* a_bf16, const void* bias_a_bf16, const void* r_bf16, const void* bias_r_bf16, void* out_bf16,
                                       float* workspace, int B, int64_t HW, int C, cudaStream_t stream);

// out = a + bias_a[c] (+ r (+ bias_r[c])), bf16 [rows, C]; bias_a, r, bias_r nullable; out may alias a or r
cudaError_t launch_bias_add_nhwc(const void* a_bf16, const void* bias_a_bf16, const void* r_bf16, const void* bias_r_bf16, void* out_bf16,
                                 int64_t rows, int C, cudaStream_t stream);

// nearest-neighbour 2x upsampling, NHWC bf16 [B,H,W,C] -> [B,2H,2W,C]
cudaError_t launch_upsample2x_nhwc(const void* x_bf16, void* y_bf16, int B, int H, int W, int C, cudaStream_t stream);

}  // namespace ugk
