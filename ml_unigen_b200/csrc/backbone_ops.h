// backbone_ops.h -- row-wise (HBM-bound) kernels of the Qwen2 decoder stack around the GEMMs.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ugk {

// x[M,H] fp32 residual stream.  If delta != nullptr: x += float(delta) (bf16 [M,H]) first.
// Then h = bf16( w * (x * rsqrt(mean(x^2) + eps)) ).  H % 8 == 0, H <= 8192.
cudaError_t launch_add_rmsnorm(float* x, const void* delta_bf16, const void* w_bf16, void* h_bf16,
                               int64_t M, int H, float eps, cudaStream_t stream);

// in-place rotary embedding on the q and k heads of a fused qkv activation:
// qkv bf16 [M, (Hq + 2*Hkv) * D]; position of row m is m % L; cos/sin fp32 [L, D/2].
// Optional kv_cache bf16 [B, cache_L, 2, Hkv, D]: rotated k and v of row m go to cache row cache_row0 + m % L of
// batch m / L (fused K/V-cache update of the text-prefix reuse).
cudaError_t launch_rope_qk(void* qkv_bf16, const float* cos, const float* sin, int64_t M, int L, int Hq, int Hkv, int D,
                           void* kv_cache, int cache_L, int cache_row0, cudaStream_t stream);

// act[M,I] = bf16( silu(g) * u ) with gu bf16 [M, 2*I] = [g | u].  I % 8 == 0.
cudaError_t launch_swiglu(const void* gu_bf16, void* act_bf16, int64_t M, int I, cudaStream_t stream);

// x[m,:] = float(table[token[m], :]) ; table bf16 [vocab,H], tokens int64 [M].  With xb / ss (both or neither; H % 256 == 0):
// xb[m,:] = bf16(x[m,:]) and ss[m, j] = sum of x[m, 128 j .. 128 j + 127]^2 (fp32 [M, H/128]) -- the operands of the
// norm-folded GEMMs (gemm_tc.h).
cudaError_t launch_embed_rows(float* x, void* xb_bf16, float* ss, const void* table_bf16, const int64_t* tokens, int64_t M, int H,
                              int64_t vocab, cudaStream_t stream);

}  // namespace ugk
