// omni_attention.cu -- omni-attention forward for sm_100a on tcgen05 / TMEM / TMA.
//
// Replaces the attention call of the Qwen2 backbone as the reference drives it
// (models/unigen.py:260-266 -> transformers sdpa_attention_forward -> F.scaled_dot_product_attention
// with the dense additive [B,1,L,L] mask built by training/prompting_utils.py:975-1020).  The mask
// is never materialised: it is the closed form of two integers per sequence (pad_len P_b,
// soi_idx S_b) and fully masked KV tiles are skipped.
//
// Persistent grid: one CTA per SM walks a static list of work items; an item = 256 query rows (two 128-row tiles A and B) of one
// (batch, q-head), a GQA-packed tail tile, or one tile of a split pair.  384 threads:
//   warps 0-3  softmax / correction / epilogue of tile A: thread r owns query row r (TMEM lane r)
//   warps 4-7  the same for tile B
//   warp  8    TMA producer: Q_A, Q_B per item, then K_t / V_t tiles (2-deep rings), SWIZZLE_128B boxes
//   warps 9,10 MMA issuers of tile A / tile B: one elected thread runs the whole chain of its tile --
//              S_x[hf] = Q_x K_t[hf]^T and O_x += P_x[hf] V_t[hf] with tcgen05.mma (kind::f16, bf16 x bf16 -> fp32), accumulators in TMEM
//   warp  11   idle (keeps the third warpgroup whole)
// TMEM columns: S_A [0,128), S_B [128,256), O_A [256,384), O_B [384,512).  Every query tile is pipelined in two 64-key halves, each a
// chain of its own (scores -> softmax -> P V -> next scores); P_x[hf] (bf16 pairs) overwrites the first 32 columns of S_x[hf] and is
// the A operand of the second GEMM straight from tensor memory, so the probabilities never touch shared memory.  The exponentials
// are taken against a running reference without a max pass (optimistic online softmax: a stale reference shows in the row sum and
// only then O, l and the half are redone -- exact, because O, l and P always share one reference); a quarter of them runs on the
// FMA pipe (polynomial), the rest on MUFU.EX2.
// GQA: the 6 query heads of a KV head read the same K/V tiles (L2 hits); nothing is expanded.  DESIGN.md section 4.2 has the
// measurements, the timeline and the variants that were measured slower.
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>

#include "omni_attention.h"

namespace ugk {

constexpr int kBM = 128;             // query rows per CTA
constexpr int kBN = 128;             // kv rows per tile
constexpr int kD = 128;              // head dim
constexpr uint32_t kTileBytes = kBM * kD * 2;          // 32 KB: one [128 x 128] bf16 tile (two 64-wide swizzle panels)
constexpr uint32_t kPanelBytes = kTileBytes / 2;       // 16 KB: [128 rows x 64 cols] panel, 128-byte rows
constexpr int kKVStages = 2;


// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(s_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(s_u32(bar)), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();       // a broken pipeline aborts the launch instead of hanging the GPU
    }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            s_u32(dst)),
        "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// executed by the whole (converged) MMA warp; one elected lane commits
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(s_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same load without the wait: several can be in flight; tmem_ld_wait() + pin32() before the registers are used
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// zero-instruction barrier: makes every later use of v[] depend on this point (i.e. on the wait above)
__device__ __forceinline__ void pin32(uint32_t (&v)[32])
{
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// UMMA shared-memory descriptors (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 (Blackwell) | [61,64) layout type, 2 = SWIZZLE_128B
// K-major panel [rows x 64 bf16]: 128-byte rows, 8-row swizzle atoms 1024 B apart (SBO); LBO unused (1).
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major operand [K rows x 128 bf16 along MN] stored as two 64-wide panels 16 KB apart (LBO),
// 128-byte rows, 8-row groups 1024 B apart (SBO).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kPanelBytes >> 4) << 16) |
           ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
// a_major bit 15, b_major bit 16, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn_major, int b_mn_major)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
struct AttnParams {
    AttnArgs a;
    float scale_log2;       // scale * log2(e)
    int n_qpairs;           // CTAs along the sequence: 256 query rows each
    int n_split;            // the last n_split pairs of the linear grid are cut into two one-tile CTAs, see the kernel
    int tail_rows;          // > 0: the last pair holds only this many query rows (Lq % 256) and is GQA-packed, see the kernel
    float* dbg;             // optional: dumps S and P of (b=0, h=0, q tile 0, kv tile 0)
    int st256;              // the output rows are 32-byte aligned: the epilogue may use 32-byte stores
};

// Descriptor words.  The high word of every shared-memory descriptor used here is the same constant
// (SBO = 1024 B, version 1, SWIZZLE_128B); the low word is (address >> 4) | (LBO >> 4) << 16, so stepping through
// a tile is one 32-bit add on the low word -- the single issuing thread must not spend ~100 dependent cycles of
// 64-bit arithmetic per MMA (measured: that, not the tensor core, was pacing the first versions of this kernel).
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo_kmajor(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_mnmajor(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | ((kPanelBytes >> 4) << 16); }
// These wrappers are executed by ALL 32 lanes of the (converged) MMA warp; elect.sync picks the one lane that
// issues.  Issuing from inside an `if (lane == 0)` region instead makes ptxas wrap every UTCHMMA in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop (operands must be uniform registers), which costs more than the MMA.
__device__ __forceinline__ void tc_mma_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
__device__ __forceinline__ void tc_mma_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 db, {%2, %5};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
// single-thread forms (the caller is the one thread elected by elect_one())
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_commit_1t(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_ss_1t(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
__device__ __forceinline__ void tc_mma_ts_1t(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
// TS form: A operand (P, bf16 pairs) read from tensor memory, B from shared memory
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void pin16(uint32_t (&v)[16])
{
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(v[i]));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}


#ifndef UG_ATTN_EARLYPUB
#define UG_ATTN_EARLYPUB 1   // 1: hand a P half to the MMA warp right after its stores (waits for them); 0: behind the next half's first load
                             //    (round 2, after the counters left local memory: 1 is 1-2 % faster; it was a wash before)
#endif
#ifndef UG_ATTN_DUMP
#define UG_ATTN_DUMP 0       // 1: the score / probability tile dump behind `dbg` (failure diagnosis in tests/test_omni_attention_gpu.py)
#endif
#ifndef UG_ATTN_TIMELINE
#define UG_ATTN_TIMELINE 0   // 1: cycle stamps of CTA 0 into the debug buffer (tools/attn_timeline.py builds this variant)
#endif
#ifndef UG_ATTN_POLY
#define UG_ATTN_POLY 1       // of every 4 pairs of exponentials, how many go to the FMA pipe (polynomial); 0 = all on MUFU
                             //    (round 2: 1 = a quarter of the exponentials: window pass 0.173 -> 0.164 ms; 2 is slower again: FMA issue)
#endif
// 2^x for two values on the FMA pipe (round-to-nearest split x = n + f via the 1.5*2^23 constant, degree-3 minimax
// polynomial for 2^f on [-0.5, 0.5] with relative error 7.5e-5 -- far below the bf16 rounding of P -- and an integer
// add of n into the exponent).  MUFU.EX2 costs 8 cycles per warp instruction and sub-partition, and the two warpgroups share it:
// the exponentials of a kv tile take as long as its tensor-core work.  A quarter of them (UG_ATTN_POLY = 1) is computed here;
// a larger share loses again to the extra FMA-pipe instructions (measured).
__device__ __forceinline__ float2 ex2_poly2(float2 x)
{
    const float2 magic = make_float2(12582912.0f, 12582912.0f);
    x.x = fmaxf(x.x, -125.0f);                                 // keeps the exponent arithmetic in range; 2^-125 ~ 0 for P
    x.y = fmaxf(x.y, -125.0f);
    const float2 t = __fadd2_rn(x, magic);                     // low mantissa bits of t hold n = rint(x)
    const float2 nn = __fadd2_rn(magic, make_float2(-t.x, -t.y));   // -n
    const float2 f = __fadd2_rn(x, nn);
    float2 p = make_float2(0x1.c3f76p-5f, 0x1.c3f76p-5f);
    p = __ffma2_rn(p, f, make_float2(0x1.f0de1ap-3f, 0x1.f0de1ap-3f));
    p = __ffma2_rn(p, f, make_float2(0x1.62f31ap-1f, 0x1.62f31ap-1f));
    p = __ffma2_rn(p, f, make_float2(0x1.fff692p-1f, 0x1.fff692p-1f));
    float2 r;
    r.x = __uint_as_float(__float_as_uint(t.x) * 0x800000u + __float_as_uint(p.x));     // one IMAD: p * 2^n
    r.y = __uint_as_float(__float_as_uint(t.y) * 0x800000u + __float_as_uint(p.y));
    return r;
}

// ---- softmax inner loops, specialised on "tile needs masking" so the hot path has no per-element branches ----
template <bool kMasked>
__device__ __forceinline__ float row_max_chunk(const uint32_t (&v)[32], int cbase, int clo, int chi, float mx)
{
    float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
    for (int q = 0; q < 32; q += 4) {
        float a = __uint_as_float(v[q]), b = __uint_as_float(v[q + 1]), c = __uint_as_float(v[q + 2]), d = __uint_as_float(v[q + 3]);
        if (kMasked) {
            const int col = cbase + q;
            a = (col >= clo && col < chi) ? a : -INFINITY;
            b = (col + 1 >= clo && col + 1 < chi) ? b : -INFINITY;
            c = (col + 2 >= clo && col + 2 < chi) ? c : -INFINITY;
            d = (col + 3 >= clo && col + 3 < chi) ? d : -INFINITY;
        }
        m0 = fmaxf(m0, a); m1 = fmaxf(m1, b); m2 = fmaxf(m2, c); m3 = fmaxf(m3, d);
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// p = 2^(s*c - m) for 32 scores -> 16 packed bf16 pairs; returns the fp32 sum of the 32 probabilities.
// Of every 4 pairs, UG_ATTN_POLY go through the FMA-pipe polynomial and the rest through MUFU.EX2.
template <bool kMasked>
__device__ __forceinline__ float prob_chunk(const uint32_t (&v)[32], uint32_t (&pk)[16], float c, float m_use, int cbase,
                                            int clo, int chi)
{
    const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_use, -m_use);
    float2 acc0 = make_float2(0.0f, 0.0f), acc1 = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int q = 0; q < 32; q += 2) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[q]), __uint_as_float(v[q + 1])), c2, nm2);
        float2 e;
        if (((q >> 1) & 3) < UG_ATTN_POLY) e = ex2_poly2(x);              // compile-time choice per pair
        else e = make_float2(ex2(x.x), ex2(x.y));
        if (kMasked) {
            const int col = cbase + q;
            e.x = (col >= clo && col < chi) ? e.x : 0.0f;
            e.y = (col + 1 >= clo && col + 1 < chi) ? e.y : 0.0f;
        }
        pk[q >> 1] = pack_bf16(e.x, e.y);
        if (q & 2) acc1 = __fadd2_rn(acc1, e);
        else acc0 = __fadd2_rn(acc0, e);
    }
    return (acc0.x + acc0.y) + (acc1.x + acc1.y);
}

// Origin of the kv tiles of one sequence: tile t covers the keys [t * 128 + off, t * 128 + off + 128).  off = 0 aligns the
// tiles to the start of the sequence and leaves the ragged tile at the END, where every image row pays for a nearly empty
// half (L = 139 + N for text_len-128 prompts: 11 valid keys in the last tile).  off = (L mod 128) - 128 aligns them to the
// end: the ragged tile is the FIRST one, and as the sequences are left-padded and no row that matters attends to a pad it
// is fully masked -- and skipped -- once pad_len >= L mod 128.  The end alignment is chosen when it needs strictly fewer
// 64-key halves for the keys [P, L) the image block attends to (measured on one box: -4.5 % per launch at P = 57; a ragged
// FIRST tile that still has work costs +2.7 %, hence the tie goes to off = 0).  TMA zero-fills the rows before key 0.
__device__ __forceinline__ int kv_tile_offset(int L, int P)
{
    const int r = L % kBN;
    if (r == 0) return 0;
    const int off = r - kBN;
    const int halves_start = (L + 63) / 64 - P / 64;
    const int halves_end = (L - off) / 64 - (P - off) / 64;
    return halves_end < halves_start ? off : 0;
}

// kv tile range [t0, t1) needed by the query rows [i0, i1) under the omni mask (P = pad_len, S = soi_idx)
__device__ __forceinline__ void kv_range(int i0, int i1, int P, int S, int L, int& t0, int& t1)
{
    const int off = kv_tile_offset(L, P);
    const int jlo = (i0 < P) ? 0 : P;
    const int jhi = (i1 > S) ? L : i1;
    t0 = (jlo - off) / kBN;
    t1 = (jhi - off + kBN - 1) / kBN;
}

// Shared memory: Q tile A | Q tile B | K ring (2) | V ring (2) | barriers.  P lives in tensor memory.
constexpr uint32_t kOff2QA = 0;
constexpr uint32_t kOff2QB = kOff2QA + kTileBytes;
constexpr uint32_t kOff2K = kOff2QB + kTileBytes;
constexpr uint32_t kOff2V = kOff2K + kKVStages * kTileBytes;
constexpr uint32_t kOff2Bar = kOff2V + kKVStages * kTileBytes;
constexpr uint32_t kOff2Items = kOff2Bar + 512;
constexpr uint32_t kAttn2Smem = kOff2Items + 32 * 80 + 1024;
constexpr int kAttn2Threads = 384;           // 2 softmax warpgroups (one per query tile) + TMA warp + 2 MMA warps + a spare warp
constexpr float kStaleSum = 65536.0f;         // a 64-column row sum above this means the softmax reference lags the scores by > ~10 bits

struct ItemInfo {                 // one work item of the persistent grid, decoded once per CTA (see the kernel)
    unsigned long long hmask[2];
    int h, b, qp, lA, R, P, S;
    int tA0, tA1, tB0, tB1, tU0, tU1, kv_off, flags;       // flags: 1 packed tail tile, 2 has tile B, 4 nothing to do
};
constexpr int kItemTable = 32;

struct Attn2Bars {
    uint64_t q_full;
    uint64_t k_full[kKVStages], k_empty[kKVStages];
    uint64_t v_full[kKVStages], v_empty[kKVStages];
    uint64_t s_full[2][2];       // [tile][half]: 64 score columns landed in TMEM (tcgen05.commit)
    uint64_t p_full[2][2];       // [tile][half]: 64 probability columns written to TMEM by the 128 softmax threads
    uint64_t pv_done[2];         // per query tile: one phase per accumulated P V half (O may be rescaled after it)
    uint64_t o_done[2];          // per query tile: last P V accumulated
    uint64_t q_empty;            // the last Q K^T of the item has completed: the Q tiles may be overwritten (persistent grid)
    uint64_t o_free[2];          // per query tile: the epilogue has read O out (128 arrivals): the next item may accumulate
    uint32_t tmem_base;
};

// TMEM columns: S_A [0,128) S_B [128,256) O_A [256,384) O_B [384,512); P_X aliases the first 64 columns of S_X
// (bf16 pairs): the P chunk written after reading S chunk c only covers columns already consumed.
//
// PERSISTENT GRID (round 2): one CTA per SM walks the work items v = blockIdx.x, blockIdx.x + gridDim.x, ... (an item = what a CTA
// of the round-1 grid did: one pair of query tiles, or one tile of a split pair).  TMEM allocation, barrier initialisation and
// descriptor fetches happen once; every mbarrier keeps a running phase, and two barriers carry the cross-item dependencies:
//   q_empty  -- the last Q K^T of an item has completed: the producer may overwrite the Q tiles with the next item's
//   o_free[x]-- the softmax warpgroup has read O_x out: the MMA warp may start accumulating the next item into it
// so the Q/K/V loads and the first score GEMM of item i+1 run while the warpgroups are still in the epilogue of item i (round 1
// paid ~5 k cycles of prologue per CTA: 15 % of a 10-tile CTA at N = 1024, a third of a 4-tile CTA at N = 256).
__global__ void __launch_bounds__(kAttn2Threads, 1)
omni_attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_qt,
                      const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const AttnParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Attn2Bars* bars = reinterpret_cast<Attn2Bars*>(smem + kOff2Bar);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.a.L;
    if (tid == 0) {
        bar_init(&bars->q_full, 1);
        bar_init(&bars->q_empty, 2);
        for (int s = 0; s < kKVStages; ++s) {
            bar_init(&bars->k_full[s], 1); bar_init(&bars->k_empty[s], 2);
            bar_init(&bars->v_full[s], 1); bar_init(&bars->v_empty[s], 2);
        }
        for (int x = 0; x < 2; ++x) {
            for (int hf = 0; hf < 2; ++hf) { bar_init(&bars->s_full[x][hf], 1); bar_init(&bars->p_full[x][hf], 128); }
            bar_init(&bars->o_done[x], 1);
            bar_init(&bars->pv_done[x], 1);
            bar_init(&bars->o_free[x], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {       // TMEM: all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&bars->tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 9 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_qt) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_v) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    // instrumentation (compiled in with -DUG_ATTN_TIMELINE=1 only, tools/attn_timeline.py): cycle stamps of the first item of CTA 0 at
    // dbg[32768 + role*512 + iter*8 + event].  In the product build the macros are empty: their run-time conditions alone were
    // 7 % of the kernel's executed instructions, most of them in the softmax warps' inner loop.
#if UG_ATTN_TIMELINE
    bool tl = false;
    const uint32_t clk0 = (uint32_t)clock();
#define UG_TL(role, iter, ev)                                                                            \
    do {                                                                                                 \
        if (tl && (iter) < 64) p.dbg[32768 + (role) * 512 + (iter) * 8 + (ev)] = (float)((uint32_t)clock() - clk0); \
    } while (0)
#else
#define UG_TL(role, iter, ev) do { } while (0)
#endif

    // running state of the roles across items (every thread walks the same item sequence)
    int n_items = 0;                                  // items started by this CTA: phase of q_full / q_empty
    int jj = 0;                                       // kv tiles of the items before this one: stage and phase of the K/V rings
    // (scalars, not arrays: dynamically indexed arrays live in local memory, and a dependent load/store round trip per counter update
    //  -- ~250 cycles each -- kept the MMA warp 6 k cycles between two items and cost the softmax threads ~300 cycles per half)
    int cnt_x0 = 0, cnt_x1 = 0;                       // items in which query tile 0 / 1 carried work: phase of o_done / o_free
    int np_0 = 0, np_1 = 0;                           // MMA warp of tile x: P halves 0 / 1 consumed (phase of p_full[x][hf])
    int ns_0 = 0, ns_1 = 0;                           // softmax warpgroup x: score halves 0 / 1 processed (phase of s_full[x][hf])
    int n_pub = 0;                                    // softmax: P halves published so far by this warpgroup (phase of pv_done)

    // Item list, heaviest first, walked round-robin by the CTAs (a static schedule: items of one class cost the same up to the
    // pads of their sequence, so no dynamic ticket is needed):
    //   [0, Wh - n_split)          pairs of query tiles (b, h, qp), qp fastest and descending (image rows first)
    //   [.., + Wt)                 GQA-packed tail tiles (b, kv head): when the last pair holds only R = Lq % 256 rows and
    //                              R * G <= 128, the R rows of all G query heads of one KV head share ONE tile (row r of the tile =
    //                              head h + r / R, sequence row r % R): they read the same K/V anyway.  With N + 2 = 1026 query rows
    //                              per image this replaces 12 nearly empty tiles per image (11 % of all tiles) by 2
    //   [.., + n_split) x 2        the last n_split pairs, which would fall into a final, mostly idle round, cut in two: tile A only,
    //                              then tile B only -- the tail of the launch costs one tile, not two
    const int G = p.a.Hq / p.a.Hkv;
    const int n_heavy = p.tail_rows > 0 ? p.n_qpairs - 1 : p.n_qpairs;      // non-packed pairs per (b, h)
    const int Wh = n_heavy * p.a.Hq * p.a.B;
    const int Wt = p.tail_rows > 0 ? p.a.Hkv * p.a.B : 0;
    const int n_virtual = Wh + Wt + p.n_split;
    const int Lq = p.a.Lq, q0 = p.a.q0;
    const int iEnd = min(q0 + Lq, L);
    // The items of this CTA are decoded ONCE, by one thread each, into a table in shared memory (up to kItemTable at a time) while
    // the tensor memory is being allocated: between two items a role only reads a few words.  (Decoding per item in every warp
    // -- divisions, kv ranges, masks, the pad_len / soi_idx loads -- kept the MMA warp off the tensor core for ~5 k cycles.)
    auto decode_item = [&](int v, ItemInfo& it) {
        int only = -1;                                      // -1: both tiles; 0: tile A only; 1: tile B only
        bool packed = false;
        int qp, h, b;
        if (v >= Wh - p.n_split && v < Wh - p.n_split + Wt) {   // packed tail tile
            const int u = v - (Wh - p.n_split);
            packed = true;
            qp = p.n_qpairs - 1;
            h = (u % p.a.Hkv) * G;
            b = u / p.a.Hkv;
        } else {
            int w = v;
            if (v >= Wh + Wt) { w = v - Wt - p.n_split; only = 1; }
            else if (v >= Wh - p.n_split + Wt) { w = v - Wt; only = 0; }
            qp = n_heavy - 1 - w % n_heavy;
            h = (w / n_heavy) % p.a.Hq;
            b = w / (n_heavy * p.a.Hq);
        }
        const int R = packed ? p.tail_rows : kBM;        // sequence rows covered by tile A
        int P = p.a.pad_len[b], S = p.a.soi_idx[b];
        P = max(0, min(P, L));
        S = max(P, min(S, L));
        // the two query tiles of the item (local rows lA/lB of q, absolute sequence rows iA/iB) and the kv tiles they need
        const bool pair_has_b = !packed && qp * 2 * kBM + kBM < Lq;
        bool skip = false;
        if (!pair_has_b) {                                  // nothing to split
            if (only == 1) skip = true;
            only = -1;
        }
        const int lA = qp * 2 * kBM + (only == 1 ? kBM : 0), lB = lA + kBM;
        const int iA = q0 + lA, iB = q0 + lB;
        const bool hasB = only < 0 && pair_has_b;
        int tA0, tA1, tB0 = 0, tB1 = 0;
        kv_range(iA, min(iA + R, iEnd), P, S, L, tA0, tA1);
        if (hasB) kv_range(iB, min(iB + kBM, iEnd), P, S, L, tB0, tB1);
        const int tU0 = hasB ? min(tA0, tB0) : tA0;
        const int tU1 = hasB ? max(tA1, tB1) : tA1;
        // union [jlo, jhi) of the allowed kv columns over the rows of each tile: a 64-column half outside it is skipped
        const int jloA = (iA < P) ? 0 : P, jhiA = (min(iA + R, iEnd) > S) ? L : min(iA + R, iEnd);
        const int jloB = (iB < P) ? 0 : P, jhiB = (min(iB + kBM, iEnd) > S) ? L : min(iB + kBM, iEnd);
        const int kv_off = kv_tile_offset(L, P);
        // bit 2*(t - tU0) + hf of hmask[x]: half (t, hf) of tile x carries work (up to 32 kv tiles = L <= 4096).  Half g = 2 t + hf
        // covers the keys [64 g + kv_off, 64 g + kv_off + 64): it carries work iff it belongs to a kv tile of the query tile and
        // meets the union [jlo, jhi) of the allowed keys of its rows -- a contiguous range of g: two shifts, not a loop
        auto range_mask = [&](int t0x, int t1x, int jlo, int jhi) -> unsigned long long {
            const int g_lo = max(2 * t0x, (jlo - kv_off) / 64), g_hi = min(2 * t1x, (jhi - kv_off + 63) / 64);
            const int n = g_hi - g_lo;
            if (n <= 0) return 0ull;
            return (n >= 64 ? ~0ull : ((1ull << n) - 1ull)) << (g_lo - 2 * tU0);
        };
        it.hmask[0] = range_mask(tA0, tA1, jloA, jhiA);
        it.hmask[1] = hasB ? range_mask(tB0, tB1, jloB, jhiB) : 0ull;
        it.h = h; it.b = b; it.qp = qp; it.lA = lA; it.R = R; it.P = P; it.S = S;
        it.tA0 = tA0; it.tA1 = tA1; it.tB0 = tB0; it.tB1 = tB1; it.tU0 = tU0; it.tU1 = tU1; it.kv_off = kv_off;
        it.flags = (packed ? 1 : 0) | (hasB ? 2 : 0) | (skip ? 4 : 0);
    };
    ItemInfo* items = reinterpret_cast<ItemInfo*>(smem + kOff2Items);
    const int n_mine = (int)blockIdx.x < n_virtual ? (n_virtual - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    for (int base = 0; base < n_mine; base += kItemTable) {
    __syncthreads();                                    // base > 0: every role is done with the previous table
    if (tid < kItemTable && base + tid < n_mine) decode_item((int)blockIdx.x + (base + tid) * (int)gridDim.x, items[tid]);
    __syncthreads();
    for (int ik = 0; ik < min(kItemTable, n_mine - base); ++ik) {
    const ItemInfo& it = items[ik];
    if (it.flags & 4) continue;
    const bool packed = it.flags & 1, hasB = it.flags & 2;
    const int h = it.h, b = it.b, qp = it.qp, lA = it.lA, lB = it.lA + kBM, R = it.R, P = it.P, S = it.S;
    const int hk = h / G;
    const int iA = q0 + lA, iB = q0 + lB;
    const int tA0 = it.tA0, tA1 = it.tA1, tB0 = it.tB0, tB1 = it.tB1, tU0 = it.tU0, tU1 = it.tU1, kv_off = it.kv_off;
    const unsigned long long hmask0 = it.hmask[0], hmask1 = it.hmask[1];
    auto half_on = [&](int x, int t, int hf) -> bool { return ((x ? hmask1 : hmask0) >> (2 * (t - tU0) + hf)) & 1ull; };
#if UG_ATTN_TIMELINE
    tl = p.dbg && blockIdx.x == 0 && n_items == 0;
    // item-boundary stamps of CTA 0: dbg[32768 + 1536 + role * 128 + item * 8 + event], 16 items, 4 roles
    const bool tli = p.dbg && blockIdx.x == 0 && n_items < 16;
#define UG_TLI(role, ev)                                                                                             \
    do {                                                                                                             \
        if (tli) p.dbg[32768 + 1536 + (role) * 128 + n_items * 8 + (ev)] = (float)((uint32_t)clock() - clk0);        \
    } while (0)
#else
#define UG_TLI(role, ev) do { } while (0)
#endif

    if (warp == 8) {
        // =========================================================== TMA producer
        if (lane == 0) UG_TLI(0, 0);
        if (n_items > 0) bar_wait(&bars->q_empty, (uint32_t)(n_items - 1) & 1u);       // the previous item's last Q K^T is done
        if (lane == 0) UG_TLI(0, 1);
        if (packed) {      // rows R*G .. 127 of the packed tile are not loaded: zero scores instead of stale shared memory
            for (uint32_t o = (uint32_t)lane * 16u; o < kTileBytes; o += 32u * 16u)
                *reinterpret_cast<uint4*>(smem + kOff2QA + o) = make_uint4(0u, 0u, 0u, 0u);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) {
            if (packed) {      // box [64 d, R rows, G heads]: tile row g * R + r
                bar_expect_tx(&bars->q_full, 2u * (uint32_t)(R * G) * 128u);
                tma_load_4d(smem + kOff2QA, &tm_qt, &bars->q_full, 0, lA, h, b);
                tma_load_4d(smem + kOff2QA + kPanelBytes, &tm_qt, &bars->q_full, 64, lA, h, b);
            } else {
                bar_expect_tx(&bars->q_full, hasB ? 2 * kTileBytes : kTileBytes);
                tma_load_4d(smem + kOff2QA, &tm_q, &bars->q_full, 0, lA, h, b);
                tma_load_4d(smem + kOff2QA + kPanelBytes, &tm_q, &bars->q_full, 64, lA, h, b);
            }
            if (hasB) {
                tma_load_4d(smem + kOff2QB, &tm_q, &bars->q_full, 0, lB, h, b);
                tma_load_4d(smem + kOff2QB + kPanelBytes, &tm_q, &bars->q_full, 64, lB, h, b);
            }
            for (int t = tU0; t < tU1; ++t) {
                const int j = jj + t - tU0;
                const int s = j % kKVStages;
                const uint32_t ph = (uint32_t)(j / kKVStages) & 1u;
                const int j0 = t * kBN + kv_off;                       // may be negative for tile 0: those rows are zero-filled
                bar_wait(&bars->k_empty[s], ph ^ 1u);
                bar_expect_tx(&bars->k_full[s], kTileBytes);
                tma_load_4d(smem + kOff2K + s * kTileBytes, &tm_k, &bars->k_full[s], 0, j0, hk, b);
                tma_load_4d(smem + kOff2K + s * kTileBytes + kPanelBytes, &tm_k, &bars->k_full[s], 64, j0, hk, b);
                bar_wait(&bars->v_empty[s], ph ^ 1u);
                bar_expect_tx(&bars->v_full[s], kTileBytes);
                tma_load_4d(smem + kOff2V + s * kTileBytes, &tm_v, &bars->v_full[s], 0, j0, hk, b);
                tma_load_4d(smem + kOff2V + s * kTileBytes + kPanelBytes, &tm_v, &bars->v_full[s], 64, j0, hk, b);
            }
            UG_TLI(0, 2);
        }
        __syncwarp();
    } else if (warp == 9 || warp == 10) {
        // =========================================================== MMA issuers: warp 9 drives query tile A, warp 10 tile B
        // Every query tile is pipelined in two 64-key halves, each a chain of its own:
        //     S_x[hf] = Q_x K_t[hf]^T  ->  softmax warpgroup x  ->  O_x += P_x[hf] V_t[hf]  ->  S_x[hf] = Q_x K_(t+1)[hf]^T  -> ...
        // (the new scores go behind the P V that consumes the probabilities they overwrite).  While warpgroup x works on one
        // half, the tensor core runs the other half's P V and next scores, so the warpgroup never waits for its own round trip.
        // Round 1 served all four chains from ONE warp in a fixed order with per-instruction election: a ready chain waited
        // for another chain's softmax and ~50 cycles of issue overhead per MMA made that warp the pacing role.  Now each query
        // tile has its own issuing warp, and inside it ONE elected thread runs the whole chain (waits, MMAs, commits): in a
        // region guarded by elect.sync ptxas keeps the MMA operands in uniform registers and the UTCHMMAs go out back to back.
        // K/V stages and the Q tiles are handed back when BOTH warps have committed (barrier count 2); a warp whose tile does
        // not need a kv tile just passes it on.
        {
            const int x = warp - 9;
            constexpr uint32_t idesc_qk = make_idesc(kBM, 64, 0, 0);        // S half: M=128, N=64
            constexpr uint32_t idesc_pv = make_idesc(kBM, kD, 0, 1);        // O: M=128, N=128, K=64 per half
            const uint32_t ql = desc_lo_kmajor(s_u32(smem + (x == 0 ? kOff2QA : kOff2QB)));
            const uint32_t k_lo0 = desc_lo_kmajor(s_u32(smem + kOff2K)), v_lo0 = desc_lo_mnmajor(s_u32(smem + kOff2V));
            const uint32_t s_tmem = tmem + (uint32_t)x * 128u, o_tmem = tmem + 256u + (uint32_t)x * 128u;
            const unsigned long long hm = x ? hmask1 : hmask0;
            const int nT = tU1 - tU0;
            // the contiguous range [g0, ge) of 64-key halves (relative to kv tile tU0) that carry work for this query tile
            const int g0 = hm ? __ffsll((long long)hm) - 1 : 0, ge = hm ? 64 - __clzll((long long)hm) : 0;
            const int ta = hm ? g0 >> 1 : nT, tb = hm ? (ge - 1) >> 1 : nT - 1;      // its kv tiles [ta, tb]
            uint32_t npl0 = (uint32_t)np_0, npl1 = (uint32_t)np_1;
            const int cnt_me = x == 0 ? cnt_x0 : cnt_x1;
            if (elect_one()) {
                // scores of half g (kv tile g >> 1, keys 64 * (g & 1) ..): 8 MMAs; hands the K stage back after the tile's last half
                auto issue_qk = [&](int g) {
                    const int tl_ = g >> 1, hf = g & 1;
                    const int j = jj + tl_, s = j % kKVStages;
                    if (g == g0 || hf == 0) {
                        bar_wait(&bars->k_full[s], (uint32_t)(j / kKVStages) & 1u);
                        tc_fence_after();
                    }
                    // K tile: two 64-wide d panels of [128 kv rows x 128 B]; kv rows 64.. start 8 KB into each panel
                    const uint32_t kl = k_lo0 + (uint32_t)s * (kTileBytes >> 4) + (uint32_t)hf * ((64u * 128u) >> 4);
                    const uint32_t d = s_tmem + (uint32_t)hf * 64u;
#pragma unroll
                    for (int k = 0; k < kD / 16; ++k) {
                        // 16 bf16 along d = 32 bytes inside the 128-byte swizzle row; panel switch at k = 4
                        const uint32_t off = (uint32_t)(k >> 2) * (kPanelBytes >> 4) + (uint32_t)(k & 3) * 2u;
                        tc_mma_ss_1t(d, ql + off, kl + off, idesc_qk, k > 0 ? 1u : 0u);
                    }
                    tc_commit_1t(&bars->s_full[x][hf]);
                    if (hf == 1 || g + 1 == ge) tc_commit_1t(&bars->k_empty[s]);
                    if (g + 1 == ge) tc_commit_1t(&bars->q_empty);            // the item's last Q K^T of this query tile
                };
                auto pass_on = [&](int tl_) {         // a kv tile this query tile does not need
                    const int j = jj + tl_, s = j % kKVStages;
                    const uint32_t ph = (uint32_t)(j / kKVStages) & 1u;
                    bar_wait(&bars->k_full[s], ph);
                    tc_commit_1t(&bars->k_empty[s]);
                    bar_wait(&bars->v_full[s], ph);
                    tc_commit_1t(&bars->v_empty[s]);
                };
                if (x == 0) UG_TLI(1, 0);
                bar_wait(&bars->q_full, (uint32_t)n_items & 1u);
                if (x == 0) UG_TLI(1, 1);
                for (int tl_ = 0; tl_ < min(ta, nT); ++tl_) pass_on(tl_);
                if (hm) {
                    issue_qk(g0);
                    if (g0 + 1 < ge) issue_qk(g0 + 1);
                    if (x == 0) UG_TLI(1, 2);
                    if (cnt_me > 0) bar_wait(&bars->o_free[x], (uint32_t)(cnt_me - 1) & 1u);   // O_x: the previous item's epilogue has read it
                    for (int g = g0; g < ge; ++g) {
                        const int tl_ = g >> 1, hf = g & 1;
                        const int j = jj + tl_, s = j % kKVStages;
                        if (x == 0) UG_TL(2, g - g0, 0);
                        bar_wait(&bars->p_full[x][hf], (hf ? npl1 : npl0) & 1u);
                        if (x == 0) UG_TL(2, g - g0, 1);
                        if (g == g0 || hf == 0) bar_wait(&bars->v_full[s], (uint32_t)(j / kKVStages) & 1u);
                        tc_fence_after();
                        // V tile rows 64*hf.. (MN-major): 16 kv rows per MMA; the P half aliases its S half
                        const uint32_t vl = v_lo0 + (uint32_t)s * (kTileBytes >> 4) + (uint32_t)hf * ((64u * 128u) >> 4);
                        const uint32_t p_tmem = s_tmem + (uint32_t)hf * 64u;
                        const uint32_t acc0 = g > g0 ? 1u : 0u;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_ts_1t(o_tmem, p_tmem + (uint32_t)k * 8u, vl + (uint32_t)k * ((16u * 128u) >> 4), idesc_pv, k > 0 ? 1u : acc0);
                        tc_commit_1t(&bars->pv_done[x]);
                        if (hf) ++npl1; else ++npl0;
                        if (g + 1 == ge) tc_commit_1t(&bars->o_done[x]);
                        if (hf == 1 || g + 1 == ge) tc_commit_1t(&bars->v_empty[s]);
                        if (x == 0) UG_TL(2, g - g0, 2);
                        if (g + 2 < ge) issue_qk(g + 2);
                        if (x == 0) UG_TL(2, g - g0, 3);
                    }
                } else {
                    tc_commit_1t(&bars->q_empty);
                }
                for (int tl_ = tb + 1; tl_ < nT; ++tl_) pass_on(tl_);
                if (x == 0) UG_TLI(1, 3);
            }
            __syncwarp();
            np_0 += ((ge + 1) >> 1) - ((g0 + 1) >> 1);          // every lane keeps the books (phases of p_full): even and odd g in [g0, ge)
            np_1 += (ge >> 1) - (g0 >> 1);
        }
    } else if (warp == 11) {
        // spare warp (keeps the third warpgroup whole)
    } else {
        // =========================================================== softmax / correction / epilogue, one warpgroup per query tile
        const int x = warp >> 2;                         // 0: tile A, 1: tile B
        const int r = tid & 127;                         // row inside the tile == TMEM lane
        const int i0 = x == 0 ? iA : iB;
        const int rl = packed ? r % R : r;               // sequence row inside the tile
        const int ho = packed ? h + r / R : h;           // query head of this row
        const bool row_stored = !packed || r < R * G;
        const int i1 = min(i0 + (x == 0 ? R : kBM), iEnd);
        const int i = i0 + rl;
        const int lrow = (x == 0 ? lA : lB) + rl;          // row of q / out
        const int t0 = x == 0 ? tA0 : tB0, t1 = x == 0 ? tA1 : tB1;
        const bool active = x == 0 || hasB;
        if (active) {
            int lo, hi;                                  // allowed kv range [lo, hi) of this row
            if (i >= S) { lo = P; hi = L; }
            else if (i >= P) { lo = P; hi = i + 1; }
            else { lo = 0; hi = i + 1; }
            hi = min(hi, L);
            const int max_lo = (i1 > P) ? P : 0;
            const int min_hi = (i0 < S) ? (i0 + 1) : L;
            const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
            const uint32_t s_addr = tmem + lane_base + (uint32_t)x * 128u;      // S_x; P_x half hf in the first 32 columns of S half hf
            const uint32_t o_addr = tmem + lane_base + 256u + (uint32_t)x * 128u;
            const float c = p.scale_log2;
            float m_ref = -INFINITY;                     // the (possibly stale) max the exponentials are taken against
            float l_run = 0.0f;
            // debug dump of S and P of (b = 0, h = 0, q tile 0, first kv tile): compiled in with -DUG_ATTN_DUMP=1 only
            const bool dump = UG_ATTN_DUMP && p.dbg && x == 0 && qp == 0 && h == 0 && b == 0;
            bool o_live = false;                          // O_x holds (or is about to hold) an accumulation of THIS item

            // Optimistic online softmax: the exponentials of a 64-column half are taken against the current reference m_ref
            // without looking for the row max first (any reference is exact as long as O, l and P share it).  The row sum of
            // the half tells whether the reference went stale (sum > 2^16, inf/NaN, or no reference yet); only then -- the
            // first half of a row and after rare jumps of the scores -- slow_half() reloads the scores, takes the true max,
            // rescales O and l and redoes the half.
            auto slow_half = [&](uint32_t sh_addr, bool full, int clo, int chi, bool stale, uint32_t (&pa)[16], uint32_t (&pb)[16],
                                 float& alpha) -> float {
                uint32_t w0[32], w1[32];
                tmem_ld32_async(sh_addr, w0);
                tmem_ld32_async(sh_addr + 32, w1);
                tmem_ld_wait();
                pin32(w0); pin32(w1);
                float mx = -INFINITY;
                if (full) { mx = row_max_chunk<false>(w0, 0, clo, chi, mx); mx = row_max_chunk<false>(w1, 32, clo, chi, mx); }
                else { mx = row_max_chunk<true>(w0, 0, clo, chi, mx); mx = row_max_chunk<true>(w1, 32, clo, chi, mx); }
                const float m_true = fmaxf(m_ref, mx * c);
                const float m_new = stale ? m_true : m_ref;                  // rows that are fine keep their reference
                const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
                alpha = (m_new != m_ref) ? ex2(m_ref - m_use) : 1.0f;        // 0 when there was no reference yet
                // O_x may only be rescaled when no P V of this tile is in flight.  The scores of this kv tile were issued behind
                // every P V of the previous one, so only the P V of the half published just before (half 0 of this kv tile, when
                // this is half 1) can still be running: wait for its phase of pv_done (the barrier is at most one phase behind,
                // so the parity wait is exact).
                if (o_live) {
                    bar_wait(&bars->pv_done[x], (uint32_t)(n_pub - 1) & 1u);
                    tc_fence_after();
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t v[32];
                        tmem_ld32(o_addr + ch * 32, v);
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) * alpha);
                        tmem_st32(o_addr + ch * 32, v);
                    }
                }
                float rs;
                if (full) rs = prob_chunk<false>(w0, pa, c, m_use, 0, clo, chi) + prob_chunk<false>(w1, pb, c, m_use, 32, clo, chi);
                else rs = prob_chunk<true>(w0, pa, c, m_use, 0, clo, chi) + prob_chunk<true>(w1, pb, c, m_use, 32, clo, chi);
                m_ref = m_new;
                return rs;
            };
            // the P half is in tensor memory: hand it to the MMA warp
            auto publish = [&](int hf) {
                tmem_st_wait();
                tc_fence_before();
                bar_arrive(&bars->p_full[x][hf]);
                o_live = true;
                ++n_pub;
            };
            auto dump_p = [&](int hf, const uint32_t (&pa)[16], const uint32_t (&pb)[16]) {
                for (int q = 0; q < 16; ++q) {
                    p.dbg[16384 + r * 128 + hf * 64 + 2 * q] = __uint_as_float(pa[q] << 16);
                    p.dbg[16384 + r * 128 + hf * 64 + 2 * q + 1] = __uint_as_float(pa[q] & 0xffff0000u);
                    p.dbg[16384 + r * 128 + hf * 64 + 32 + 2 * q] = __uint_as_float(pb[q] << 16);
                    p.dbg[16384 + r * 128 + hf * 64 + 32 + 2 * q + 1] = __uint_as_float(pb[q] & 0xffff0000u);
                }
            };

            if (r == 0) UG_TLI(2 + x, 0);
            int pend = -1;               // half whose P stores are in flight and not yet handed to the MMA warp
            int n = 0;
            for (int t = t0; t < t1; ++t) {
              for (int hf = 0; hf < 2; ++hf) {
                if (!half_on(x, t, hf)) continue;
                const int j0 = t * kBN + hf * 64 + kv_off;
                const bool full = (j0 >= max_lo) && (j0 + 64 <= min_hi);
                const int clo = lo - j0, chi = hi - j0;
                const uint32_t sh_addr = s_addr + (uint32_t)hf * 64u;
                if (r == 0) UG_TL(x, n, 0);
                bar_wait(&bars->s_full[x][hf], (uint32_t)(hf ? ns_1 : ns_0) & 1u);
                if (hf) ++ns_1; else ++ns_0;
                tc_fence_after();
                if (r == 0) UG_TL(x, n, 1);
                // Scores come to registers 32 columns at a time, each load in flight while something else happens: the first one
                // while the PREVIOUS half's probabilities are handed over (wait for their tensor-memory stores, fence, arrive), the
                // second one while the first chunk is in the MUFU section.  (tcgen05.wait::ld waits for ALL loads of the thread, so a
                // load is issued right after the wait for its predecessor; pin16 keeps ptxas from sinking the arithmetic below it.)
                uint32_t va[32], vb[32];
                uint32_t pk0[16], pk1[16];
                tmem_ld32_async(sh_addr, va);
                if (pend >= 0) publish(pend);
                tmem_ld_wait();
                pin32(va);
                tmem_ld32_async(sh_addr + 32, vb);
                if (r == 0) UG_TL(x, n, 2);
                if (dump && t == t0)
                    for (int q = 0; q < 32; ++q) p.dbg[r * 128 + hf * 64 + q] = __uint_as_float(va[q]);
                float rs = 0.0f, alpha = 1.0f;
                bool stale = (m_ref == -INFINITY);
                const bool opt = !__any_sync(0xffffffffu, stale);
                if (opt) {
                    rs = full ? prob_chunk<false>(va, pk0, c, m_ref, 0, clo, chi) : prob_chunk<true>(va, pk0, c, m_ref, 0, clo, chi);
                    pin16(pk0);
                }
                tmem_ld_wait();
                pin32(vb);
                if (r == 0) UG_TL(x, n, 3);
                if (dump && t == t0)
                    for (int q = 0; q < 32; ++q) p.dbg[r * 128 + hf * 64 + 32 + q] = __uint_as_float(vb[q]);
                if (opt) {
                    rs += full ? prob_chunk<false>(vb, pk1, c, m_ref, 32, clo, chi) : prob_chunk<true>(vb, pk1, c, m_ref, 32, clo, chi);
                    stale = !(rs <= kStaleSum);
                }
                if (__any_sync(0xffffffffu, stale)) rs = slow_half(sh_addr, full, clo, chi, stale, pk0, pk1, alpha);
                if (dump && t == t0) dump_p(hf, pk0, pk1);
                // P half -> tensor memory (A operand of the second GEMM), over the score columns already consumed
                tmem_st16(sh_addr, pk0);
                tmem_st16(sh_addr + 16, pk1);
                l_run = l_run * alpha + rs;
                if (UG_ATTN_EARLYPUB) publish(hf);
                else pend = hf;
                if (r == 0) UG_TL(x, n, 4);
                ++n;
              }
            }
            if (pend >= 0) publish(pend);
            // epilogue: O / l -> bf16 -> global
            if (r == 0) UG_TLI(2 + x, 1);
            bar_wait(&bars->o_done[x], (uint32_t)(x == 0 ? cnt_x0 : cnt_x1) & 1u);
            if (r == 0) UG_TLI(2 + x, 2);
            tc_fence_after();
            const float inv_l = 1.0f / l_run;
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.a.out) + (int64_t)b * p.a.o_sb + (int64_t)lrow * p.a.o_sl + (int64_t)ho * p.a.o_sh;
            // O row -> registers 32 columns at a time, two loads in flight (ping-pong); O_x is handed back to the MMA warp as soon as
            // the last chunk is in registers, before its stores
            auto store_chunk = [&](const uint32_t (&v)[32], int ch) {
                if (lrow < Lq && row_stored && p.st256) {
                    // 32-byte stores (STG.256): a lane writes whole 32-byte sectors of its own row.  The rows of a warp are 3 KB apart, so a
                    // store instruction touches 32 sectors whatever its width: half as many instructions = half as many sector accesses
                    // (the 16-byte version spent ~4 k cycles per item here, the epilogue's whole time)
#pragma unroll
                    for (int q = 0; q < 32; q += 16) {
                        uint32_t w[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) w[e] = pack_bf16(__uint_as_float(v[q + 2 * e]) * inv_l, __uint_as_float(v[q + 2 * e + 1]) * inv_l);
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + ch * 32 + q), "r"(w[0]), "r"(w[1]),
                                     "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                                     : "memory");
                    }
                } else if (lrow < Lq && row_stored) {
#pragma unroll
                    for (int q = 0; q < 32; q += 8) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(v[q]) * inv_l, __uint_as_float(v[q + 1]) * inv_l);
                        o.y = pack_bf16(__uint_as_float(v[q + 2]) * inv_l, __uint_as_float(v[q + 3]) * inv_l);
                        o.z = pack_bf16(__uint_as_float(v[q + 4]) * inv_l, __uint_as_float(v[q + 5]) * inv_l);
                        o.w = pack_bf16(__uint_as_float(v[q + 6]) * inv_l, __uint_as_float(v[q + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + ch * 32 + q) = o;
                    }
                }
            };
            {
                uint32_t ve[32], vf[32];
                tmem_ld32_async(o_addr, ve);
                tmem_ld32_async(o_addr + 32, vf);
                tmem_ld_wait();
                pin32(ve); pin32(vf);
                store_chunk(ve, 0);
                tmem_ld32_async(o_addr + 64, ve);
                store_chunk(vf, 1);
                tmem_ld32_async(o_addr + 96, vf);
                tmem_ld_wait();
                pin32(ve); pin32(vf);
                tc_fence_before();
                bar_arrive(&bars->o_free[x]);            // O_x has been read out: the next item may accumulate into it
                store_chunk(ve, 2);
                store_chunk(vf, 3);
            }
            if (r == 0) UG_TLI(2 + x, 3);
        }
    }
    // every thread keeps the same books
    jj += tU1 - tU0;
    cnt_x0 += 1;
    cnt_x1 += hasB ? 1 : 0;
    ++n_items;
    }       // items of the table
    }       // tables
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side: TMA descriptors through the driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

static bool make_map(EncodeTiledFn enc, CUtensorMap* m, const void* base, int L, int H, int B, int64_t sb, int64_t sl, int64_t sh,
                     int box_rows = kBM, int box_heads = 1)
{
    cuuint64_t dims[4] = {(cuuint64_t)kD, (cuuint64_t)L, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)sl * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_heads, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}


int launch_omni_attention(const AttnArgs& a, int num_sms, cudaStream_t stream, char* msg, size_t msg_len)
{
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { snprintf(msg, msg_len, "cuTensorMapEncodeTiled entry point not available"); return -6; }
    CUtensorMap tq, tqt, tk, tv;
    const int G = a.Hq / a.Hkv;
    const int tail = a.Lq % (2 * kBM);
    const int tail_rows = (tail > 0 && tail * G <= kBM) ? tail : 0;       // GQA-packed tail tile
    if (!make_map(enc, &tq, a.q, a.Lq, a.Hq, a.B, a.q_sb, a.q_sl, a.q_sh) ||
        !make_map(enc, &tqt, a.q, a.Lq, a.Hq, a.B, a.q_sb, a.q_sl, a.q_sh, tail_rows ? tail_rows : kBM, tail_rows ? G : 1) ||
        !make_map(enc, &tk, a.k, a.L, a.Hkv, a.B, a.kv_sb, a.kv_sl, a.kv_sh) ||
        !make_map(enc, &tv, a.v, a.L, a.Hkv, a.B, a.kv_sb, a.kv_sl, a.kv_sh)) {
        snprintf(msg, msg_len, "cuTensorMapEncodeTiled rejected the q/k/v layout");
        return -3;
    }
    AttnParams p;
    p.a = a;
    p.scale_log2 = a.scale * 1.4426950408889634f;
    p.n_qpairs = (a.Lq + 2 * kBM - 1) / (2 * kBM);
    p.tail_rows = tail_rows;
    // pairs that would fall into a last, partly filled round (one CTA per SM; a packed tail tile counts as half a pair)
    const int n_heavy = tail_rows ? p.n_qpairs - 1 : p.n_qpairs;
    const int Wh = n_heavy * a.Hq * a.B, Wt = tail_rows ? a.Hkv * a.B : 0;
    {
        int k = num_sms > 0 ? (Wh + Wt / 2) % num_sms : 0;
        if (k > num_sms / 2 || k > Wh) k = 0;
        static const char* ov = getenv("UG_ATTN_NSPLIT");        // kernel experiments only (tools/bench_attention.py)
        if (ov && atoi(ov) >= 0 && atoi(ov) <= Wh) k = atoi(ov);
        p.n_split = k;
    }
    p.dbg = a.dbg;
    p.st256 = (reinterpret_cast<uintptr_t>(a.out) % 32 == 0 && a.o_sb % 16 == 0 && a.o_sl % 16 == 0 && a.o_sh % 16 == 0) ? 1 : 0;
    static bool attr_set[64] = {};                       // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(omni_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttn2Smem);
        if (e != cudaSuccess) { snprintf(msg, msg_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int n_virtual = Wh + Wt + p.n_split;
    dim3 grid((unsigned)(num_sms > 0 && num_sms < n_virtual ? num_sms : n_virtual));      // persistent: one CTA per SM walks the items
    omni_attention_kernel<<<grid, kAttn2Threads, kAttn2Smem, stream>>>(tq, tqt, tk, tv, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(msg, msg_len, "launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // namespace ugk
