// gemm_tc.h -- bf16 linear layers of the Qwen2 backbone on tcgen05 (2-CTA MMA), with the elementwise step that
// follows each of them in the reference fused into the epilogue.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace ugk {

enum GemmMode : int {
    kGemmPlain = 0,     // out bf16 [M,N]   = A W^T (+ bias)
    kGemmSwiGLU = 1,    // out bf16 [M,N/2] = silu(g) * u, W rows interleaved in blocks of 128: [g_0..127 | u_0..127 | g_128.. ]
    kGemmQkvRope = 2,   // out bf16 [M,N]   = rope(A W^T + bias) on the q and k heads (head_dim 128), v heads copied; k/v rows also
                        // written to the K/V cache
    kGemmResidual = 3,  // x fp32 [M,N] += bf16(A W^T); xb bf16 [M,N] = bf16(x); ss fp32 [M,N/128] = per-128-column sums of x^2
};

struct GemmArgs {
    const void* a;      // bf16 [M,K], row stride lda elements
    const void* w;      // bf16 [N,K], row stride ldw elements
    const void* bias;   // bf16 [N] or null
    void* out;          // see GemmMode
    int64_t lda, ldw, ldo;
    int M, N, K;
    int mode;
    // kGemmQkvRope only
    const float* rope_cos = nullptr;   // fp32 [rope_L, 64]
    const float* rope_sin = nullptr;
    void* kv_cache = nullptr;          // nullable: bf16 [B, cache_L, 2, Hkv, 128]
    int rope_L = 0, rope_Hq = 0, rope_Hkv = 0, cache_L = 0, cache_row0 = 0;
    // RMSNorm folding (plain / SwiGLU / q/k/v modes): row m of the accumulator is scaled by rsqrt(sum_p row_ss[m,p] / K + rms_eps)
    const float* row_ss = nullptr;     // fp32 [M, ss_parts] or null
    int ss_parts = 0;
    float rms_eps = 0.0f;
    // k x k convolution, k = 3 (zero padding 1) or 1, stride 1, as an implicit GEMM -- kGemmPlain only, conv_W > 0:
    // a = NHWC bf16 activation [B,H,W,Cin] (contiguous), out = NHWC [B,H,W,Cout]; w = [N, K] with N = ceil128(Cout) rows (zero rows
    // beyond Cout), K = taps * ceil64(Cin), k = tap * ceil64(Cin) + c (zero beyond Cin); bias (nullable) has N entries.
    // A rows are the pixels of [4 x 32] boxes; partial boxes, image borders and the channel padding are TMA out-of-bounds
    // zero fill (loads) / clipping (stores), so any H, W and any Cin, Cout that are multiples of 8 work.
    int conv_B = 0, conv_H = 0, conv_W = 0, conv_Cin = 0, conv_Cout = 0, conv_taps = 9;
    // kGemmResidual only (contiguous rows of N elements)
    float* res_x = nullptr;
    void* res_xb = nullptr;
    float* res_ss = nullptr;
};

int launch_gemm_tc(const GemmArgs& g, int num_sms, cudaStream_t stream, char* msg, size_t msg_len);

}  // namespace ugk
