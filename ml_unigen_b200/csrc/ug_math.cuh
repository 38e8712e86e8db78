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

// ---------------------------------------------------------------- noise streams
// counter layout: (block | kind<<28, row, offset_lo, offset_hi ^ row_hi), key = seed
//
// Race noise of element i of a row: a 24-bit integer k24, u = (2^24 - k24) * 2^-24 in (0,1], q = -log(u) ~ Exp(1).
// The top byte of k24 is zero with probability 2^-8 ("survivor": q < 2^-8); only survivors and elements with a
// large probability can win the race, so the stream is DEFINED such that the survivors of a group of 256 elements
// come from ONE Philox call (count ~ Binomial(256, 2^-8) by inverse CDF on a 32-bit word, then distinct uniform
// positions with 16 fine bits each), and the top byte of every other element (uniform on 1..255) and its fine bits
// come from a per-element word.  The joint law is that of iid uniform 24-bit integers up to the 2^-32 granularity of
// the inverse CDF.
enum : uint32_t { kKindGroup = 0u, kKindElem = 1u, kKindSelect = 2u };
constexpr int kSurvMax = 12;
// rounded 2^32 * P(Binomial(256, 1/256) >= k), k = 1..12
#define UG_SURV_T1 2718028156u
#define UG_SURV_T2 1134904942u
#define UG_SURV_T3 343343334u
#define UG_SURV_T4 80524186u
#define UG_SURV_T5 15334731u
#define UG_SURV_T6 2450227u
#define UG_SURV_T7 336494u
#define UG_SURV_T8 40453u
#define UG_SURV_T9 4319u
#define UG_SURV_T10 414u
#define UG_SURV_T11 36u
#define UG_SURV_T12 3u
// q of a non-survivor is >= -log(1 - 2^-8) = 0.0039139...; the filters use this floor (with slack)
constexpr float kQFloor = 0.0039f;

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

// element index <-> (group, position): group = the 256 elements of each 8192 that one lane of a warp holds when a
// row is read with 16-byte loads (lane = (i/8)%32, 32 loads of 8 elements)
UG_HD uint32_t group_of(uint32_t i) { return ((i >> 3) & 31u) | ((i >> 13) << 5); }
UG_HD uint32_t pos_of(uint32_t i) { return (((i >> 8) & 31u) << 3) | (i & 7u); }
UG_HD uint32_t index_of(uint32_t g, uint32_t pos)
{
    return ((g >> 5) << 13) | ((pos >> 3) << 8) | ((g & 31u) << 3) | (pos & 7u);
}

UG_HD int surv_count(uint32_t w0)
{
    if (w0 >= UG_SURV_T1) return 0;
    return 1 + (w0 < UG_SURV_T2) + (w0 < UG_SURV_T3) + (w0 < UG_SURV_T4) + (w0 < UG_SURV_T5) + (w0 < UG_SURV_T6) +
           (w0 < UG_SURV_T7) + (w0 < UG_SURV_T8) + (w0 < UG_SURV_T9) + (w0 < UG_SURV_T10) + (w0 < UG_SURV_T11) +
           (w0 < UG_SURV_T12);
}

// is `pos` among the first nvalid (0..4) bytes of `packed`?
UG_HD bool ug_has_byte(uint32_t packed, int nvalid, uint32_t pos)
{
    return (nvalid > 0 && (packed & 255u) == pos) | (nvalid > 1 && ((packed >> 8) & 255u) == pos) |
           (nvalid > 2 && ((packed >> 16) & 255u) == pos) | (nvalid > 3 && (packed >> 24) == pos);
}

// Survivors of group g: calls f(pos, k24) for each (k24 < 2^16).  Words 1.. of the group call (and of follow-up
// calls block = g | sub<<12) give one candidate each: position = bits 0..7, fine = bits 8..23; a position that was
// already chosen is skipped.
template <class F>
UG_HD void for_each_survivor(uint32_t g, uint64_t row, uint64_t seed, uint64_t offset, F&& f)
{
    u32x4 cur = noise_call(g, kKindGroup, row, seed, offset);
    const int K = surv_count(cur.x);
    if (K == 0) return;
    uint32_t pv0 = 0, pv1 = 0, pv2 = 0;                     // chosen positions, one per byte, in order
    uint32_t widx = 1, sub = 0;
    int n = 0;
    while (n < K) {
        if (widx == 4) { ++sub; cur = noise_call(g | (sub << 12), kKindGroup, row, seed, offset); widx = 0; }
        const uint32_t w = word_of(cur, widx++);
        const uint32_t pos = w & 255u;
        if (ug_has_byte(pv0, n, pos) || (n > 4 && (ug_has_byte(pv1, n - 4, pos) || (n > 8 && ug_has_byte(pv2, n - 8, pos))))) continue;
        if (n < 4) pv0 |= pos << (8 * n); else if (n < 8) pv1 |= pos << (8 * (n - 4)); else pv2 |= pos << (8 * (n - 8));
        f(pos, (w >> 8) & 0xffffu);
        ++n;
    }
}

// k24 of an element that is NOT a survivor: top byte uniform on 1..255 (first non-zero byte, most significant
// first, of word i%4 of the element call; follow-up calls block = i/4 | sub<<16 if all four are zero), fine bits =
// low 16 bits of the first word.
UG_HD uint32_t elem_k24(uint32_t i, uint64_t row, uint64_t seed, uint64_t offset)
{
    uint32_t w = word_of(noise_call(i >> 2, kKindElem, row, seed, offset), i & 3u);
    const uint32_t fine = w & 0xffffu;
    uint32_t top = w >> 24;
    uint32_t sub = 0;
    while (top == 0) {
        ++sub;
        w = word_of(noise_call((i >> 2) | (sub << 16), kKindElem, row, seed, offset), i & 3u);
        top = (w >> 24) ? (w >> 24) : (((w >> 16) & 255u) ? ((w >> 16) & 255u) : (((w >> 8) & 255u) ? ((w >> 8) & 255u) : (w & 255u)));
    }
    return (top << 16) | fine;
}

// k24 of any element (generic, slow: one group call + one element call)
UG_HD uint32_t race_k24(uint32_t i, uint64_t row, uint64_t seed, uint64_t offset)
{
    const uint32_t p = pos_of(i);
    uint32_t found = 0xffffffffu;
    for_each_survivor(group_of(i), row, seed, offset, [&](uint32_t pos, uint32_t k24) { if (pos == p) found = k24; });
    if (found != 0xffffffffu) return found;
    return elem_k24(i, row, seed, offset);
}

UG_HD float q_from_k24(uint32_t k24)
{
    uint32_t v = 0x1000000u - k24;
    if (v == 0x1000000u) return 0x1p-25f;
    return -logf_det(UG_MUL((float)v, 0x1p-24f));
}

UG_HD float race_q(uint32_t i, uint64_t row, uint64_t seed, uint64_t offset)
{
    return q_from_k24(race_k24(i, row, seed, offset));
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
