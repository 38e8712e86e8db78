// backbone_ops.cu -- the HBM-bound glue of the Qwen2.5 decoder layer (transformers
// models/qwen2/modeling_qwen2.py: Qwen2RMSNorm, apply_rotary_pos_emb, Qwen2MLP's act_fn(gate)*up,
// embed_tokens), fused so that every activation is read and written once:
//   residual add + RMSNorm        (fp32 residual stream in, bf16 normalised activations out)
//   RoPE on q and k in place      (fp32 math, one rounding to bf16)
//   SwiGLU                        (silu(g) * u)
//   embedding gather              (bf16 table -> fp32 residual stream)
// All kernels use 16-byte vector accesses and one warp (or a few) per row.
#include "backbone_ops.h"
#include <cuda_bf16.h>

namespace ugk {

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------------
// one warp per row; the row stays in registers between the sum of squares and the scaling
template <int kVecPerLane>      // 8-element vectors per lane: H = 32 * 8 * kVecPerLane
__global__ void __launch_bounds__(256) add_rmsnorm_kernel(float* __restrict__ x, const uint4* __restrict__ delta,
                                                          const uint4* __restrict__ w, uint4* __restrict__ h, int64_t M,
                                                          int H, float eps)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    float v[kVecPerLane][8];
    float ss = 0.0f;
    float4* xr = reinterpret_cast<float4*>(x + row * H);
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
        const int vec = i * 32 + lane;                    // 8-element vector index inside the row
        float4 a = xr[2 * vec], b = xr[2 * vec + 1];
        v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
        v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
        if (delta) {
            uint4 d = delta[row * (H / 8) + vec];
            v[i][0] += bf_lo(d.x); v[i][1] += bf_hi(d.x); v[i][2] += bf_lo(d.y); v[i][3] += bf_hi(d.y);
            v[i][4] += bf_lo(d.z); v[i][5] += bf_hi(d.z); v[i][6] += bf_lo(d.w); v[i][7] += bf_hi(d.w);
            xr[2 * vec] = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
            xr[2 * vec + 1] = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) ss = fmaf(v[i][k], v[i][k], ss);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float r = rsqrtf(ss / (float)H + eps);
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
        const int vec = i * 32 + lane;
        uint4 ww = w[vec];
        uint4 o;
        o.x = pack2(bf_lo(ww.x) * (v[i][0] * r), bf_hi(ww.x) * (v[i][1] * r));
        o.y = pack2(bf_lo(ww.y) * (v[i][2] * r), bf_hi(ww.y) * (v[i][3] * r));
        o.z = pack2(bf_lo(ww.z) * (v[i][4] * r), bf_hi(ww.z) * (v[i][5] * r));
        o.w = pack2(bf_lo(ww.w) * (v[i][6] * r), bf_hi(ww.w) * (v[i][7] * r));
        h[row * (H / 8) + vec] = o;
    }
}

cudaError_t launch_add_rmsnorm(float* x, const void* delta, const void* w, void* h, int64_t M, int H, float eps,
                               cudaStream_t stream)
{
    if (H % 256 != 0 || H > 8192 || M <= 0) return cudaErrorInvalidValue;
    const int vpl = H / 256;
    const int rows_per_block = 8;
    const unsigned grid = (unsigned)((M + rows_per_block - 1) / rows_per_block);
#define UG_RMS_CASE(n)                                                                                                  \
    case n:                                                                                                             \
        add_rmsnorm_kernel<n><<<grid, 256, 0, stream>>>(x, (const uint4*)delta, (const uint4*)w, (uint4*)h, M, H, eps); \
        break;
    switch (vpl) {
        UG_RMS_CASE(1) UG_RMS_CASE(2) UG_RMS_CASE(3) UG_RMS_CASE(4) UG_RMS_CASE(5) UG_RMS_CASE(6) UG_RMS_CASE(7) UG_RMS_CASE(8)
        UG_RMS_CASE(12) UG_RMS_CASE(16) UG_RMS_CASE(20) UG_RMS_CASE(24) UG_RMS_CASE(28) UG_RMS_CASE(32)
        default: return cudaErrorInvalidValue;
    }
#undef UG_RMS_CASE
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// RoPE (rotate_half convention): out[d] = x[d]*cos[d] - x[d+D/2]*sin[d]; out[d+D/2] = x[d+D/2]*cos[d] + x[d]*sin[d]
// one thread handles 8 consecutive d of the first half and the matching 8 of the second half of one head.
// q heads are rotated in place; k heads are rotated in place and, when a K/V cache is attached, also written to
// cache[b, row0 + l, 0, hk, :]; v heads are copied to cache[b, row0 + l, 1, hk, :] (fuses the cache update).
__global__ void __launch_bounds__(256) rope_qk_kernel(uint4* __restrict__ qkv, const float4* __restrict__ cos,
                                                      const float4* __restrict__ sin, int64_t M, int L, int Hq, int Hkv,
                                                      int D, int row_vecs, uint4* __restrict__ cache, int cache_L, int row0)
{
    const int per_head = D / 16;                                   // threads per head
    const int heads = Hq + Hkv + (cache ? Hkv : 0);                // q, k (rotated) and, with a cache, v (copied)
    const int64_t total = M * heads * per_head;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(idx % per_head);
        const int hd = (int)((idx / per_head) % heads);
        const int64_t m = idx / ((int64_t)per_head * heads);
        const int pos = (int)(m % L);
        uint4* base = qkv + m * row_vecs + (int64_t)hd * (D / 8);
        uint4 a = base[t], b = base[t + D / 16];
        uint4* crow = nullptr;
        if (cache && hd >= Hq) {
            const int64_t bb = m / L;
            const int kvh = hd - Hq;                               // 0..Hkv-1: k heads, Hkv..2Hkv-1: v heads
            crow = cache + (((bb * cache_L + row0 + pos) * 2 + (kvh >= Hkv ? 1 : 0)) * Hkv + (kvh % Hkv)) * (int64_t)(D / 8);
        }
        if (hd < Hq + Hkv) {
            const float4* c = cos + ((int64_t)pos * (D / 2) + t * 8) / 4;
            const float4* s = sin + ((int64_t)pos * (D / 2) + t * 8) / 4;
            float4 c0 = c[0], c1 = c[1], s0 = s[0], s1 = s[1];
            float x1[8] = {bf_lo(a.x), bf_hi(a.x), bf_lo(a.y), bf_hi(a.y), bf_lo(a.z), bf_hi(a.z), bf_lo(a.w), bf_hi(a.w)};
            float x2[8] = {bf_lo(b.x), bf_hi(b.x), bf_lo(b.y), bf_hi(b.y), bf_lo(b.z), bf_hi(b.z), bf_lo(b.w), bf_hi(b.w)};
            float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            float o1[8], o2[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                o1[k] = x1[k] * cc[k] - x2[k] * sn[k];
                o2[k] = x2[k] * cc[k] + x1[k] * sn[k];
            }
            a = make_uint4(pack2(o1[0], o1[1]), pack2(o1[2], o1[3]), pack2(o1[4], o1[5]), pack2(o1[6], o1[7]));
            b = make_uint4(pack2(o2[0], o2[1]), pack2(o2[2], o2[3]), pack2(o2[4], o2[5]), pack2(o2[6], o2[7]));
            base[t] = a;
            base[t + D / 16] = b;
        }
        if (crow) { crow[t] = a; crow[t + D / 16] = b; }
    }
}

cudaError_t launch_rope_qk(void* qkv, const float* cos, const float* sin, int64_t M, int L, int Hq, int Hkv, int D,
                           void* kv_cache, int cache_L, int cache_row0, cudaStream_t stream)
{
    if (D % 16 != 0 || M <= 0 || L <= 0) return cudaErrorInvalidValue;
    const int heads = Hq + Hkv + (kv_cache ? Hkv : 0);             // q heads, k heads, v heads are contiguous in the fused row
    const int row_vecs = (Hq + 2 * Hkv) * D / 8;
    const int64_t total = M * heads * (D / 16);
    unsigned grid = (unsigned)((total + 255) / 256);
    if (grid > 148u * 16u) grid = 148u * 16u;
    rope_qk_kernel<<<grid, 256, 0, stream>>>((uint4*)qkv, (const float4*)cos, (const float4*)sin, M, L, Hq, Hkv, D, row_vecs,
                                             (uint4*)kv_cache, cache_L, cache_row0);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// MUFU exponential and reciprocal (2 ulp each, far below the bf16 rounding of the result); the fused SwiGLU epilogue of
// gemm_tc.cu uses the same expression, so the fused and the unfused path agree wherever the GEMM outputs agree
__device__ __forceinline__ float silu_mul(float g, float u) { return __fdividef(g, 1.0f + __expf(-g)) * u; }

__global__ void __launch_bounds__(256) swiglu_kernel(const uint4* __restrict__ gu, uint4* __restrict__ act, int64_t M, int Iv)
{
    const int64_t total = M * Iv;                                  // Iv = I / 8 vectors per output row
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = idx / Iv;
        const int c = (int)(idx % Iv);
        const uint4 g = gu[m * 2 * Iv + c], u = gu[m * 2 * Iv + Iv + c];
        uint4 o;
        o.x = pack2(silu_mul(bf_lo(g.x), bf_lo(u.x)), silu_mul(bf_hi(g.x), bf_hi(u.x)));
        o.y = pack2(silu_mul(bf_lo(g.y), bf_lo(u.y)), silu_mul(bf_hi(g.y), bf_hi(u.y)));
        o.z = pack2(silu_mul(bf_lo(g.z), bf_lo(u.z)), silu_mul(bf_hi(g.z), bf_hi(u.z)));
        o.w = pack2(silu_mul(bf_lo(g.w), bf_lo(u.w)), silu_mul(bf_hi(g.w), bf_hi(u.w)));
        act[idx] = o;
    }
}

cudaError_t launch_swiglu(const void* gu, void* act, int64_t M, int I, cudaStream_t stream)
{
    if (I % 8 != 0 || M <= 0) return cudaErrorInvalidValue;
    const int64_t total = M * (I / 8);
    unsigned grid = (unsigned)((total + 255) / 256);
    if (grid > 148u * 32u) grid = 148u * 32u;
    swiglu_kernel<<<grid, 256, 0, stream>>>((const uint4*)gu, (uint4*)act, M, I / 8);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_rows_kernel(float4* __restrict__ x, const uint4* __restrict__ table,
                                                         const int64_t* __restrict__ tokens, int64_t M, int Hv, int64_t vocab)
{
    const int64_t total = M * Hv;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = idx / Hv;
        const int c = (int)(idx % Hv);
        int64_t tok = tokens[m];
        tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
        const uint4 e = table[tok * Hv + c];
        x[2 * idx] = make_float4(bf_lo(e.x), bf_hi(e.x), bf_lo(e.y), bf_hi(e.y));
        x[2 * idx + 1] = make_float4(bf_lo(e.z), bf_hi(e.z), bf_lo(e.w), bf_hi(e.w));
    }
}

// the same gather, one warp per row, that also leaves what the norm-folded GEMMs consume (gemm_tc.cu): xb = bf16(x) (here simply the
// table row) and the per-128-column partial sums of x^2 (fp32 [M, H/128]); H % 256 == 0
__global__ void __launch_bounds__(256) embed_rows_norm_kernel(float4* __restrict__ x, uint4* __restrict__ xb, float* __restrict__ ss,
                                                              const uint4* __restrict__ table, const int64_t* __restrict__ tokens,
                                                              int64_t M, int Hv, int64_t vocab)
{
    const int lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    int64_t tok = tokens[m];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
    for (int i = 0; i < Hv / 32; ++i) {
        const int c = i * 32 + lane;                       // 8-element vector index inside the row; 16 vectors = one 128-column block
        const uint4 e = table[tok * Hv + c];
        const float v[8] = {bf_lo(e.x), bf_hi(e.x), bf_lo(e.y), bf_hi(e.y), bf_lo(e.z), bf_hi(e.z), bf_lo(e.w), bf_hi(e.w)};
        x[2 * (m * Hv + c)] = make_float4(v[0], v[1], v[2], v[3]);
        x[2 * (m * Hv + c) + 1] = make_float4(v[4], v[5], v[6], v[7]);
        xb[m * Hv + c] = e;
        float sq = v[0] * v[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) sq = fmaf(v[k], v[k], sq);
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        if ((lane & 15) == 0) ss[m * (Hv / 16) + 2 * i + (lane >> 4)] = sq;
    }
}

cudaError_t launch_embed_rows(float* x, void* xb, float* ss, const void* table, const int64_t* tokens, int64_t M, int H, int64_t vocab,
                              cudaStream_t stream)
{
    if (H % 8 != 0 || M <= 0) return cudaErrorInvalidValue;
    if (xb || ss) {
        if (!xb || !ss || H % 256 != 0) return cudaErrorInvalidValue;
        embed_rows_norm_kernel<<<(unsigned)((M + 7) / 8), 256, 0, stream>>>((float4*)x, (uint4*)xb, ss, (const uint4*)table, tokens, M,
                                                                            H / 8, vocab);
        return cudaGetLastError();
    }
    const int64_t total = M * (H / 8);
    unsigned grid = (unsigned)((total + 255) / 256);
    if (grid > 148u * 32u) grid = 148u * 32u;
    embed_rows_kernel<<<grid, 256, 0, stream>>>((float4*)x, (const uint4*)table, tokens, M, H / 8, vocab);
    return cudaGetLastError();
}

}  // namespace ugk
