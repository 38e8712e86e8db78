// ug_math.cuh -- UG-FP32 arithmetic of the decode step.
//
// Everything that decides a token id, a selected probability or a mask bit is built from
// IEEE-754 binary32 +, *, fma and / with round-to-nearest-even -- no MUFU approximations, no
// compiler contraction (every operation is an explicit __f*_rn intrinsic) -- so the results
// are reproducible bit for bit by a CPU (oracle/decode_step_oracle.c follows the same spec).
// MUFU / approximate math is only allowed in the candidate FILTER of the exponential race,
// which is conservative by construction and never changes a result.
//
// The functions are __host__ __device__ so that tests/host_math_check.cu can run the very same
// source on the CPU against the oracle before any GPU time is spent.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>

static inline uint32_t ug_host_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float ug_host_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float ug_host_add(float a, float b) { volatile float r = a + b; return r; }
static inline float ug_host_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float ug_host_div(float a, float b) { volatile float r = a / b; return r; }
#if defined(__CUDA_ARCH__)
#define UG_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define UG_ADD(a, b) __fadd_rn((a), (b))
#define UG_MUL(a, b) __fmul_rn((a), (b))
#define UG_DIV(a, b) __fdiv_rn((a), (b))
#define UG_F2U(x) __float_as_uint(x)
#define UG_U2F(x) __uint_as_float(x)
#else
#define UG_FMA(a, b, c) fmaf((a), (b), (c))
#define UG_ADD(a, b) ug_host_add((a), (b))
#define UG_MUL(a, b) ug_host_mul((a), (b))
#define UG_DIV(a, b) ug_host_div((a), (b))
#define UG_F2U(x) ug_host_f2u(x)
#define UG_U2F(x) ug_host_u2f(x)
#endif

#define UG_HD __host__ __device__ __forceinline__

namespace ug {

constexpr float kExpMagic = 12582912.0f;              // 1.5 * 2^23 (0x4B400000)
constexpr float kL2E = 1.44269502162933349609375f;
constexpr float kLn2Hi = 0.693145751953125f;          // 0x3f317200
constexpr float kLn2Lo = 1.42860677e-6f;              // 0x35bfbe8e
constexpr float kExpC1 = 1.0f;
constexpr float kExpC2 = 0x1.fffffcp-2f;
constexpr float kExpC3 = 0x1.555418p-3f;
constexpr float kExpC4 = 0x1.555824p-5f;
constexpr float kExpC5 = 0x1.1267d8p-7f;
constexpr float kExpC6 = 0x1.6ae6dep-10f;
// rows whose (max - min) logit spread is below this take the single-step 2^j scaling (all results normal)
constexpr float kExpFastSpread = 86.0f;

// exp(d) for d <= 0, any magnitude: exact UG-FP32 definition (two-step scaling, one rounding).
UG_HD float expf_exact(float d)
{
    if (!(d >= -104.0f)) return 0.0f;
    float t = UG_FMA(d, kL2E, kExpMagic);
    float j = UG_ADD(t, -kExpMagic);
    float r = UG_FMA(j, -kLn2Hi, d);
    r = UG_FMA(j, -kLn2Lo, r);
    float p = kExpC6;
    p = UG_FMA(p, r, kExpC5);
    p = UG_FMA(p, r, kExpC4);
    p = UG_FMA(p, r, kExpC3);
    p = UG_FMA(p, r, kExpC2);
    p = UG_FMA(p, r, kExpC1);
    p = UG_FMA(p, r, 1.0f);
    float a = UG_U2F((UG_F2U(t) + 191u) << 23);
    return UG_MUL(UG_MUL(p, a), 0x1p-64f);
}

// Same value as expf_exact for -86 <= d <= 0 (result is a normal number, so adding j to the
// exponent field is exact): 11 FMA-pipe operations + 1 integer operation.
UG_HD float expf_fast(float d)
{
    float t = UG_FMA(d, kL2E, kExpMagic);
    float j = UG_ADD(t, -kExpMagic);
    float r = UG_FMA(j, -kLn2Hi, d);
    r = UG_FMA(j, -kLn2Lo, r);
    float p = kExpC6;
    p = UG_FMA(p, r, kExpC5);
    p = UG_FMA(p, r, kExpC4);
    p = UG_FMA(p, r, kExpC3);
    p = UG_FMA(p, r, kExpC2);
    p = UG_FMA(p, r, kExpC1);
    p = UG_FMA(p, r, 1.0f);
    return UG_U2F(UG_F2U(p) + (UG_F2U(t) << 23));
}

#if defined(__CUDACC__)
// Two elements at a time on the packed fp32x2 pipe (FFMA2/FADD2, sm_100): same per-element
// IEEE operations as expf_fast, half the issue slots.
__device__ __forceinline__ float2 expf_fast2(float2 x, float2 negmax)
{
    const float2 l2e = make_float2(kL2E, kL2E), magic = make_float2(kExpMagic, kExpMagic);
    const float2 nmagic = make_float2(-kExpMagic, -kExpMagic);
    const float2 nhi = make_float2(-kLn2Hi, -kLn2Hi), nlo = make_float2(-kLn2Lo, -kLn2Lo);
    float2 d = __fadd2_rn(x, negmax);
    float2 t = __ffma2_rn(d, l2e, magic);
    float2 j = __fadd2_rn(t, nmagic);
    float2 r = __ffma2_rn(j, nhi, d);
    r = __ffma2_rn(j, nlo, r);
    float2 p = make_float2(kExpC6, kExpC6);
    p = __ffma2_rn(p, r, make_float2(kExpC5, kExpC5));
    p = __ffma2_rn(p, r, make_float2(kExpC4, kExpC4));
    p = __ffma2_rn(p, r, make_float2(kExpC3, kExpC3));
    p = __ffma2_rn(p, r, make_float2(kExpC2, kExpC2));
    p = __ffma2_rn(p, r, make_float2(kExpC1, kExpC1));
    p = __ffma2_rn(p, r, make_float2(1.0f, 1.0f));
    float2 e;
    e.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
    e.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
    return e;
}
#endif

// log(x) for positive normal x (callers clamp at 1e-20).
UG_HD float logf_det(float x)
{
    int32_t ix = (int32_t)UG_F2U(x);
    int32_t e = (ix - 0x3f3504f3) >> 23;
    float m = UG_U2F((uint32_t)(ix - (e << 23)));
    float f = UG_ADD(m, -1.0f);
    float fe = (float)e;
    float z = UG_MUL(f, f);
    float y = 7.0376836292E-2f;
    y = UG_FMA(y, f, -1.1514610310E-1f);
    y = UG_FMA(y, f, 1.1676998740E-1f);
    y = UG_FMA(y, f, -1.2420140846E-1f);
    y = UG_FMA(y, f, 1.4249322787E-1f);
    y = UG_FMA(y, f, -1.6668057665E-1f);
    y = UG_FMA(y, f, 2.0000714765E-1f);
    y = UG_FMA(y, f, -2.4999993993E-1f);
    y = UG_FMA(y, f, 3.3333331174E-1f);
    y = UG_MUL(y, f);
    y = UG_MUL(y, z);
    y = UG_FMA(fe, -2.12194440e-4f, y);
    y = UG_FMA(-0.5f, z, y);
    float r = UG_ADD(f, y);
    return UG_FMA(fe, 0.693359375f, r);
}

// ---------------------------------------------------------------- Philox4x32-10
struct u32x4 { uint32_t x, y, z, w; };

UG_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
    const uint64_t p = (uint64_t)a * b;                    // one IMAD.WIDE.U32
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
}

UG_HD u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c.x, hi0, lo0);
        mulhilo(0xCD9E8D57u, c.z, hi1, lo1);
        u32x4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// counter layout: (block | kind<<28, row, offset_lo, offset_hi ^ row_hi)
enum : uint32_t { kKindCoarse = 0u, kKindFine = 1u, kKindSelect = 2u };

UG_HD u32x4 noise_call(uint32_t block, uint32_t kind, uint64_t row, uint64_t seed, uint64_t offset)
{
    u32x4 c;
    c.x = block | (kind << 28);
    c.y = (uint32_t)row;
    c.z = (uint32_t)offset;
    c.w = (uint32_t)(offset >> 32) ^ (uint32_t)(row >> 32);
    return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

UG_HD uint32_t word_of(const u32x4& v, uint32_t i)
{
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// exact race noise of element i of row `row`, given its coarse byte
UG_HD float race_q_from_coarse(uint32_t coarse, uint32_t i, uint64_t row, uint64_t seed, uint64_t offset)
{
    u32x4 f = noise_call(i >> 3, kKindFine, row, seed, offset);
    uint32_t fine = (word_of(f, (i & 7u) >> 1) >> (16u * (i & 1u))) & 0xffffu;
    uint32_t k24 = (coarse << 16) | fine;
    uint32_t v = 0x1000000u - k24;
    if (v == 0x1000000u) return 0x1p-25f;
    return -logf_det(UG_MUL((float)v, 0x1p-24f));
}

UG_HD uint32_t coarse_byte(const u32x4& c, uint32_t k /*0..15*/)
{
    return (word_of(c, k >> 2) >> (8u * (k & 3u))) & 0xffu;
}

UG_HD float race_q(uint32_t i, uint64_t row, uint64_t seed, uint64_t offset)
{
    u32x4 c = noise_call(i >> 4, kKindCoarse, row, seed, offset);
    return race_q_from_coarse(coarse_byte(c, i & 15u), i, row, seed, offset);
}

UG_HD float select_u(uint32_t n, uint64_t batch_row, uint64_t seed, uint64_t offset)
{
    u32x4 c = noise_call(n >> 2, kKindSelect, batch_row, seed, offset);
    return UG_MUL((float)(word_of(c, n & 3u) >> 8), 0x1p-24f);
}

// confidence of one token (models/sampling.py:42): log(clamp(p)) + T * gumbel(u)
UG_HD float confidence(float psel, float u, float temperature)
{
    float lp = logf_det(psel < 1e-20f ? 1e-20f : psel);
    float l1 = logf_det(u < 1e-20f ? 1e-20f : u);
    float nl = -l1;
    float g = -logf_det(nl < 1e-20f ? 1e-20f : nl);
    return UG_ADD(lp, UG_MUL(temperature, g));
}

// order-preserving float -> uint key (-0 is canonicalised to +0 first)
UG_HD uint32_t ordered_key(float f)
{
    uint32_t u = UG_F2U(UG_ADD(f, 0.0f));
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// bf16 helpers (bit level, RNE)
UG_HD float bf16_bits_to_f32(uint32_t h) { return UG_U2F(h << 16); }
UG_HD uint32_t f32_to_bf16_bits(float f)
{
    uint32_t u = UG_F2U(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (u >> 16) | 0x40u;
    return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
// classifier-free guidance in bf16, op by op (models/unigen.py:410)
UG_HD uint32_t cfg_mix_bits(uint32_t cond, uint32_t uncond, float scale)
{
    float c = bf16_bits_to_f32(cond), u = bf16_bits_to_f32(uncond);
    float t1 = bf16_bits_to_f32(f32_to_bf16_bits(UG_ADD(c, -u)));
    float t2 = bf16_bits_to_f32(f32_to_bf16_bits(UG_MUL(scale, t1)));
    return f32_to_bf16_bits(UG_ADD(t2, u));
}

}  // namespace ug
