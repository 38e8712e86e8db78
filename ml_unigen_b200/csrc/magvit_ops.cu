// magvit_ops.cu -- the memory-bound kernels of MAGVITv2.decode_code (apple/ml-unigen
// models/multimodal_encoder/magvitv2.py:217-230 LFQ codebook entry; common_modules.py:19-43 swish, GroupNorm(32, 1e-6),
// nearest-2x Upsample) for NHWC bf16 activations.  The convolutions between them are cuDNN (torch.nn.functional.conv2d on
// channels_last tensors): plain library convolutions, like the plain cuBLAS GEMMs of the backbone.
//   * lfq_unpack:   13 bits of a token id -> 13 values +-1 (no embedding table is read);
//   * GroupNorm+swish: ONE statistics pass (per-block partial sums per group, combined in double by a tiny finalize
//     kernel: deterministic, no atomics on global memory) and ONE apply pass that also applies x*sigmoid(x) -- the
//     reference runs group_norm (2 passes), sigmoid and mul as separate ATen kernels (5 passes over the activation);
//   * the per-channel bias of a convolution is never a pass of its own (torch adds it with a 326 us non-vectorised
//     broadcast kernel per convolution at 16 x 512 x 512 x 128): it is added inside the GroupNorm that consumes the
//     convolution (statistics and apply kernels take an input bias), or together with the residual add of the ResnetBlock
//     (bias_add: out = conv2 + b2 + shortcut (+ b_shortcut), one pass);
//   * upsample2x: every input vector is read once and written four times.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "magvit_ops.h"

namespace ugk {
namespace {

constexpr int kGroups = 32;
constexpr int kGnThreads = 256;
constexpr int kGnRowsPerBlock = 256;          // spatial positions per statistics block

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(256) lfq_unpack_kernel(const int64_t* __restrict__ ids, __nv_bfloat16* __restrict__ z, int64_t n, int dim)
{
    const int64_t total = n * dim;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = idx / dim;
        const int c = (int)(idx - t * dim);
        const int64_t id = ids[t];
        z[idx] = __float2bfloat16(((id >> (dim - 1 - c)) & 1) ? 1.0f : -1.0f);
    }
}

// LFQ lookup + post_quant_conv in one pass (magvitv2.py:217-230 + the 1x1 `post_quant_conv` of :438-442): the codebook entry of a
// token is the +-1 vector of its bits, so the 1x1 convolution is b[c] + sum_k W[c,k] * (+-1) -- no table, no GEMM.  The row is
// written with `cpad` channels (>= dim, zero beyond dim): cpad = 64 feeds the implicit-GEMM conv_in without a padding pass.
__global__ void __launch_bounds__(256) lfq_postquant_kernel(const int64_t* __restrict__ ids, const float* __restrict__ w,
                                                            const float* __restrict__ bias, __nv_bfloat16* __restrict__ z, int64_t n,
                                                            int dim, int cpad)
{
    const int64_t total = n * cpad;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = idx / cpad;
        const int c = (int)(idx - t * cpad);
        float acc = 0.0f;
        if (c < dim) {
            const int64_t id = ids[t];
            acc = bias ? bias[c] : 0.0f;
            for (int k = 0; k < dim; ++k) {
                const float wk = w[c * dim + k];
                acc += ((id >> (dim - 1 - k)) & 1) ? wk : -wk;
            }
        }
        z[idx] = __float2bfloat16(acc);
    }
}

// row softmax of a bf16 matrix (the single attention block of the decoder: 1024 x 1024 scores per image): one warp per row,
// y = softmax(scale * x) in fp32, bf16 out; cols % 8 == 0
__global__ void __launch_bounds__(256) softmax_rows_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t rows, int vecs_all,
                                                           int vecs, float scale_log2)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const uint4* xr = x + row * vecs_all;
    float mx = -INFINITY;
    for (int v = lane; v < vecs; v += 32) {
        const uint4 u = xr[v];
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(bf_lo(u.x), bf_hi(u.x)), fmaxf(bf_lo(u.y), bf_hi(u.y))),
                             fmaxf(fmaxf(bf_lo(u.z), bf_hi(u.z)), fmaxf(bf_lo(u.w), bf_hi(u.w)))));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float m2 = mx * scale_log2;
    float sum = 0.0f;
    for (int v = lane; v < vecs; v += 32) {
        const uint4 u = xr[v];
        sum += exp2f(bf_lo(u.x) * scale_log2 - m2) + exp2f(bf_hi(u.x) * scale_log2 - m2) + exp2f(bf_lo(u.y) * scale_log2 - m2) +
               exp2f(bf_hi(u.y) * scale_log2 - m2) + exp2f(bf_lo(u.z) * scale_log2 - m2) + exp2f(bf_hi(u.z) * scale_log2 - m2) +
               exp2f(bf_lo(u.w) * scale_log2 - m2) + exp2f(bf_hi(u.w) * scale_log2 - m2);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float inv = 1.0f / sum;
    uint4* yr = y + row * vecs_all;
    for (int v = vecs + lane; v < vecs_all; v += 32) yr[v] = make_uint4(0u, 0u, 0u, 0u);      // padded keys
    for (int v = lane; v < vecs; v += 32) {
        const uint4 u = xr[v];
        uint4 o;
        o.x = pack2(exp2f(bf_lo(u.x) * scale_log2 - m2) * inv, exp2f(bf_hi(u.x) * scale_log2 - m2) * inv);
        o.y = pack2(exp2f(bf_lo(u.y) * scale_log2 - m2) * inv, exp2f(bf_hi(u.y) * scale_log2 - m2) * inv);
        o.z = pack2(exp2f(bf_lo(u.z) * scale_log2 - m2) * inv, exp2f(bf_hi(u.z) * scale_log2 - m2) * inv);
        o.w = pack2(exp2f(bf_lo(u.w) * scale_log2 - m2) * inv, exp2f(bf_hi(u.w) * scale_log2 - m2) * inv);
        yr[v] = o;
    }
}

// grid (chunks, B); block kGnThreads.  Thread t owns vector slot t % (C/8) (8 consecutive channels) of the rows
// t / (C/8) + k * (kGnThreads / (C/8)) of its chunk.  Requires C % 8 == 0 and kGnThreads % (C/8) == 0.
// kFused: the pass also PRODUCES the tensor the statistics are taken of: out = x + in_bias (+ r (+ bias_r)) is rounded to bf16,
// stored, and the sums are taken over the rounded values -- the same values, thread mapping and summation order as the plain
// statistics pass over `out` would use, so the partial sums are bit-identical to it and the following GroupNorm skips its own
// statistics pass (the residual add at the end of a ResnetBlock feeds norm1 of the next one).
template <bool kFused>
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const uint4* __restrict__ x, const __nv_bfloat16* __restrict__ in_bias,
                                                              const uint4* __restrict__ r_in, const __nv_bfloat16* __restrict__ bias_r,
                                                              uint4* __restrict__ out, float* __restrict__ partial, int64_t HW, int C)
{
    __shared__ float s_part[kGnThreads * 8], q_part[kGnThreads * 8];
    const int vecs = C / 8;
    const int slot = threadIdx.x % vecs, row0 = threadIdx.x / vecs, rstep = kGnThreads / vecs;
    const int b = blockIdx.y;
    const int64_t r_begin = (int64_t)blockIdx.x * kGnRowsPerBlock;
    const int64_t r_end = min(r_begin + (int64_t)kGnRowsPerBlock, HW);
    float s[8], q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.0f; q[e] = 0.0f; }
    const uint4* xb = x + (int64_t)b * HW * vecs;
    float ib[8], rb[8];                            // per-channel bias of the producing convolution (added here instead of in a pass of its own)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        ib[e] = in_bias ? __bfloat162float(in_bias[slot * 8 + e]) : 0.0f;
        rb[e] = (kFused && bias_r) ? __bfloat162float(bias_r[slot * 8 + e]) : 0.0f;
    }
    for (int64_t r = r_begin + row0; r < r_end; r += rstep) {
        const int64_t idx = (int64_t)b * HW * vecs + r * vecs + slot;
        const uint4 v = xb[r * vecs + slot];
        float f[8] = {bf_lo(v.x), bf_hi(v.x), bf_lo(v.y), bf_hi(v.y), bf_lo(v.z), bf_hi(v.z), bf_lo(v.w), bf_hi(v.w)};
        if (kFused) {
            if (in_bias) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += ib[e];
            }
            if (r_in) {
                const uint4 vr = r_in[idx];
                const float g[8] = {bf_lo(vr.x), bf_hi(vr.x), bf_lo(vr.y), bf_hi(vr.y), bf_lo(vr.z), bf_hi(vr.z), bf_lo(vr.w), bf_hi(vr.w)};
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += g[e] + rb[e];
            }
            const uint4 o = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
            out[idx] = o;
            const float h[8] = {bf_lo(o.x), bf_hi(o.x), bf_lo(o.y), bf_hi(o.y), bf_lo(o.z), bf_hi(o.z), bf_lo(o.w), bf_hi(o.w)};
#pragma unroll
            for (int e = 0; e < 8; ++e) { s[e] += h[e]; q[e] = fmaf(h[e], h[e], q[e]); }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float t = f[e] + ib[e];
                s[e] += t;
                q[e] = fmaf(t, t, q[e]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { s_part[threadIdx.x * 8 + e] = s[e]; q_part[threadIdx.x * 8 + e] = q[e]; }
    __syncthreads();
    // fixed-order block reduction (no atomics: the result does not depend on scheduling): thread g adds the partial sums
    // of the C/32 channels of group g over the kGnThreads/(C/8) row lanes -- 64 terms whatever C is
    if (threadIdx.x < kGroups) {
        const int g = threadIdx.x, cpg = C / kGroups;
        float ts = 0.0f, tq = 0.0f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c)
            for (int k = 0; k < rstep; ++k) {
                const int t = (c >> 3) + vecs * k;
                ts += s_part[t * 8 + (c & 7)];
                tq += q_part[t * 8 + (c & 7)];
            }
        float* p = partial + (((int64_t)b * gridDim.x + blockIdx.x) * kGroups + g) * 2;
        p[0] = ts;
        p[1] = tq;
    }
}

// grid B, block 32 x 32: warp g combines the chunk partials of group g in double -- lane l takes the chunks l, l + 32, ... in order,
// then a fixed shuffle tree: the result does not depend on scheduling.  (One thread per group walking up to 1024 chunks with
// dependent double additions took 20-80 us per call, 42 calls per decoded batch.)
__global__ void __launch_bounds__(1024) gn_finalize_kernel(const float* __restrict__ partial, float* __restrict__ stats, int chunks, double inv_count, float eps)
{
    const int b = blockIdx.x, g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0, q = 0.0;
    for (int c = lane; c < chunks; c += 32) {
        const float2 pv = *reinterpret_cast<const float2*>(partial + (((int64_t)b * chunks + c) * kGroups + g) * 2);
        s += (double)pv.x;
        q += (double)pv.y;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    if (lane == 0) {
        const double mean = s * inv_count;
        double var = q * inv_count - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        stats[(b * kGroups + g) * 2] = (float)mean;
        stats[(b * kGroups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// grid (chunks, B); block kGnThreads; same thread mapping as the statistics pass: thread t owns vector slot t % (C/8) of the rows
// t / (C/8) + k * (kGnThreads / (C/8)) of its chunk, so the per-channel scale and shift (rstd * gamma, beta + (bias - mean) * rstd *
// gamma) are formed ONCE per thread and a pixel costs one 16-byte load, 8 FMAs, the swish and one 16-byte store.  (The first
// version recomputed group index, statistics, gamma, beta and bias per element -- 32 small loads and two integer divisions per
// 16 bytes -- and ran at a third of the HBM rate: 31 % of the pixel decoder.)
constexpr int kGnApplyRows = 512;             // spatial positions per block of the normalisation pass
__global__ void __launch_bounds__(kGnThreads) gn_apply_kernel(const uint4* __restrict__ x, const __nv_bfloat16* __restrict__ in_bias,
                                                              const __nv_bfloat16* __restrict__ gamma,
                                                              const __nv_bfloat16* __restrict__ beta, const float* __restrict__ stats,
                                                              uint4* __restrict__ y, int64_t HW, int C, int swish)
{
    const int vecs = C / 8, cpg = C / kGroups;
    const int slot = threadIdx.x % vecs, row0 = threadIdx.x / vecs, rstep = kGnThreads / vecs;
    const int b = blockIdx.y;
    float a[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = slot * 8 + e, g = c / cpg;
        const float mean = stats[(b * kGroups + g) * 2], rstd = stats[(b * kGroups + g) * 2 + 1];
        const float ib = in_bias ? __bfloat162float(in_bias[c]) : 0.0f;
        a[e] = rstd * __bfloat162float(gamma[c]);
        sh[e] = fmaf(ib - mean, a[e], __bfloat162float(beta[c]));
    }
    const int64_t r_begin = (int64_t)blockIdx.x * kGnApplyRows;
    const int64_t r_end = min(r_begin + (int64_t)kGnApplyRows, HW);
    const uint4* xb = x + (int64_t)b * HW * vecs + slot;
    uint4* yb = y + (int64_t)b * HW * vecs + slot;
#pragma unroll 4
    for (int64_t r = r_begin + row0; r < r_end; r += rstep) {
        const uint4 v = xb[r * vecs];
        const float f[8] = {bf_lo(v.x), bf_hi(v.x), bf_lo(v.y), bf_hi(v.y), bf_lo(v.z), bf_hi(v.z), bf_lo(v.w), bf_hi(v.w)};
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float t = fmaf(f[e], a[e], sh[e]);
            if (swish) t = __fdividef(t, 1.0f + __expf(-t));
            o[e] = t;
        }
        yb[r * vecs] = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
    }
}

__global__ void __launch_bounds__(256) bias_add_kernel(const uint4* __restrict__ a, const __nv_bfloat16* __restrict__ bias_a,
                                                       const uint4* __restrict__ r, const __nv_bfloat16* __restrict__ bias_r,
                                                       uint4* __restrict__ out, int64_t rows, int vecs)
{
    const int64_t total = rows * vecs;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int slot = (int)(idx % vecs);
        const uint4 va = a[idx];
        float f[8] = {bf_lo(va.x), bf_hi(va.x), bf_lo(va.y), bf_hi(va.y), bf_lo(va.z), bf_hi(va.z), bf_lo(va.w), bf_hi(va.w)};
        if (bias_a) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += __bfloat162float(bias_a[slot * 8 + e]);
        }
        if (r) {
            const uint4 vr = r[idx];
            const float g[8] = {bf_lo(vr.x), bf_hi(vr.x), bf_lo(vr.y), bf_hi(vr.y), bf_lo(vr.z), bf_hi(vr.z), bf_lo(vr.w), bf_hi(vr.w)};
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += g[e] + (bias_r ? __bfloat162float(bias_r[slot * 8 + e]) : 0.0f);
        }
        out[idx] = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
    }
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int vecs)
{
    const int64_t total = (int64_t)B * H * W * vecs;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(idx % vecs);
        const int64_t pix = idx / vecs;
        const int xx = (int)(pix % W);
        const int yy = (int)((pix / W) % H);
        const int64_t b = pix / ((int64_t)W * H);
        const uint4 val = x[idx];
        uint4* o = y + (((b * 2 * H + 2 * yy) * 2 * W) + 2 * xx) * (int64_t)vecs + v;
        o[0] = val;
        o[vecs] = val;
        o[(int64_t)2 * W * vecs] = val;
        o[(int64_t)2 * W * vecs + vecs] = val;
    }
}

unsigned grid_for(int64_t total, int threads)
{
    int64_t g = (total + threads - 1) / threads;
    const int64_t cap = 148 * 16;
    return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

cudaError_t launch_lfq_unpack(const int64_t* ids, void* z, int64_t n, int dim, cudaStream_t stream)
{
    if (n < 1 || dim < 1 || dim > 62) return cudaErrorInvalidValue;
    lfq_unpack_kernel<<<grid_for(n * dim, 256), 256, 0, stream>>>(ids, reinterpret_cast<__nv_bfloat16*>(z), n, dim);
    return cudaGetLastError();
}

cudaError_t launch_lfq_postquant(const int64_t* ids, const float* w, const float* bias, void* z, int64_t n, int dim, int cpad,
                                 cudaStream_t stream)
{
    if (n < 1 || dim < 1 || dim > 62 || cpad < dim) return cudaErrorInvalidValue;
    lfq_postquant_kernel<<<grid_for(n * cpad, 256), 256, 0, stream>>>(ids, w, bias, reinterpret_cast<__nv_bfloat16*>(z), n, dim, cpad);
    return cudaGetLastError();
}

cudaError_t launch_softmax_rows(const void* x, void* y, int64_t rows, int cols, int cols_valid, float scale, cudaStream_t stream)
{
    if (rows < 1 || cols < 8 || cols % 8 || cols_valid < 8 || cols_valid % 8 || cols_valid > cols) return cudaErrorInvalidValue;
    softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), rows,
                                                                       cols / 8, cols_valid / 8, scale * 1.4426950408889634f);
    return cudaGetLastError();
}

size_t groupnorm_workspace_floats(int B, int64_t HW, int C)
{
    (void)C;
    const int64_t chunks = (HW + kGnRowsPerBlock - 1) / kGnRowsPerBlock;
    return (size_t)B * (size_t)chunks * kGroups * 2 + (size_t)B * kGroups * 2;
}

cudaError_t launch_bias_add_stats_nhwc(const void* a, const void* bias_a, const void* r, const void* bias_r, void* out, float* workspace,
                                       int B, int64_t HW, int C, cudaStream_t stream)
{
    if (B < 1 || HW < 1 || C % kGroups != 0 || C % 8 != 0 || kGnThreads % (C / 8) != 0) return cudaErrorInvalidValue;
    const int chunks = (int)((HW + kGnRowsPerBlock - 1) / kGnRowsPerBlock);
    gn_stats_kernel<true><<<dim3((unsigned)chunks, (unsigned)B), kGnThreads, 0, stream>>>(
        reinterpret_cast<const uint4*>(a), reinterpret_cast<const __nv_bfloat16*>(bias_a), reinterpret_cast<const uint4*>(r),
        reinterpret_cast<const __nv_bfloat16*>(bias_r), reinterpret_cast<uint4*>(out), workspace, HW, C);
    return cudaGetLastError();
}

cudaError_t launch_groupnorm_nhwc(const void* x, const void* in_bias, const void* gamma, const void* beta, void* y, float* workspace,
                                  int B, int64_t HW, int C, float eps, int swish, int stats_ready, cudaStream_t stream)
{
    if (B < 1 || HW < 1 || C % kGroups != 0 || C % 8 != 0 || kGnThreads % (C / 8) != 0) return cudaErrorInvalidValue;
    if (stats_ready && in_bias) return cudaErrorInvalidValue;        // the producer's sums are those of x itself
    const int chunks = (int)((HW + kGnRowsPerBlock - 1) / kGnRowsPerBlock);
    float* partial = workspace;
    float* stats = workspace + (size_t)B * chunks * kGroups * 2;
    if (!stats_ready)
        gn_stats_kernel<false><<<dim3((unsigned)chunks, (unsigned)B), kGnThreads, 0, stream>>>(
            reinterpret_cast<const uint4*>(x), reinterpret_cast<const __nv_bfloat16*>(in_bias), nullptr, nullptr, nullptr, partial, HW, C);
    gn_finalize_kernel<<<B, kGroups * 32, 0, stream>>>(partial, stats, chunks, 1.0 / ((double)HW * (C / kGroups)), eps);
    gn_apply_kernel<<<dim3((unsigned)((HW + kGnApplyRows - 1) / kGnApplyRows), (unsigned)B), kGnThreads, 0, stream>>>(
        reinterpret_cast<const uint4*>(x), reinterpret_cast<const __nv_bfloat16*>(in_bias), reinterpret_cast<const __nv_bfloat16*>(gamma),
        reinterpret_cast<const __nv_bfloat16*>(beta), stats, reinterpret_cast<uint4*>(y), HW, C, swish);
    return cudaGetLastError();
}

cudaError_t launch_bias_add_nhwc(const void* a, const void* bias_a, const void* r, const void* bias_r, void* out, int64_t rows, int C,
                                 cudaStream_t stream)
{
    if (rows < 1 || C % 8 != 0) return cudaErrorInvalidValue;
    bias_add_kernel<<<grid_for(rows * (C / 8), 256), 256, 0, stream>>>(
        reinterpret_cast<const uint4*>(a), reinterpret_cast<const __nv_bfloat16*>(bias_a), reinterpret_cast<const uint4*>(r),
        reinterpret_cast<const __nv_bfloat16*>(bias_r), reinterpret_cast<uint4*>(out), rows, C / 8);
    return cudaGetLastError();
}

cudaError_t launch_upsample2x_nhwc(const void* x, void* y, int B, int H, int W, int C, cudaStream_t stream)
{
    if (B < 1 || H < 1 || W < 1 || C % 8 != 0) return cudaErrorInvalidValue;
    upsample2x_kernel<<<grid_for((int64_t)B * H * W * (C / 8), 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x),
                                                                                       reinterpret_cast<uint4*>(y), B, H, W, C / 8);
    return cudaGetLastError();
}

}  // namespace ugk
