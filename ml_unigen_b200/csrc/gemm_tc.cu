// gemm_tc.cu -- the linear layers of the Qwen2 backbone (transformers Qwen2Attention / Qwen2MLP, reached from
// apple/ml-unigen models/unigen.py:260-266) as ONE persistent tcgen05 kernel with the elementwise step that follows
// the GEMM in the reference fused into its epilogue:
//   mode plain   : out = x W^T (+ bias)                           codebook head (and o_proj / down_proj when selected)
//   mode SwiGLU  : act = silu(x Wg^T) * (x Wu^T)                  Qwen2MLP.act_fn(gate_proj(x)) * up_proj(x): the
//                  [M, 2I] gate|up activation (588 MB per layer at config 2) is never written to HBM
//   mode q/k/v   : qkv = rope(x [Wq|Wk|Wv]^T + bias)              q_proj/k_proj/v_proj + apply_rotary_pos_emb, and the rotated
//                  k rows / the v rows written to the K/V cache of the text-prefix reuse in the same epilogue
//   mode residual: x += bf16(a W^T); xb = bf16(x); ss = sum x^2   o_proj / down_proj + the residual add of Qwen2DecoderLayer + the
//                  statistics of the Qwen2RMSNorm that follows: the separate add+RMSNorm pass over HBM no longer exists
// RMSNorm folding (all consumer modes, `row_ss` != null): Qwen2RMSNorm commutes with the linear layer that follows it,
//   (g * x / rms) W^T = (1 / rms) * (x (W diag g)^T), so the A operand is bf16(x) as the residual-mode epilogue left it, g is
//   folded into the weight columns once at load time, and 1 / rms = rsqrt(sum_p ss[m, p] / K + eps) scales row m of the
//   accumulator in the epilogue (the partial sums are added in a fixed order: deterministic).
//
// Shape of the kernel (sm_100a only):
//   * cluster of 2 CTAs = one SM pair; the pair owns a 256 x 256 output tile; `tcgen05.mma.cta_group::2.kind::f16`
//     (M=256, N=256, K=16) issued by one thread of the leader CTA; each CTA stages ITS 128 rows of A and ITS half
//     (128 rows) of the W tile, so every operand byte is read from shared memory by exactly one SM and the per-SM
//     shared-memory traffic is half of what two independent 128 x 256 tiles would need;
//   * 6-stage TMA ring (32 KB per stage per CTA, SWIZZLE_128B K-major boxes of 64 k); both CTAs' loads complete on the
//     LEADER's mbarrier; slots are handed back by a multicast `tcgen05.commit` to both CTAs;
//   * fp32 accumulators in tensor memory, double buffered (2 x 256 columns): the epilogue of tile i (warps 4-7:
//     tcgen05.ld -> registers -> swizzled staging tile in shared memory -> TMA store) overlaps the MMAs of tile i+1;
//   * persistent: cluster c takes tiles c, c+74, ... with the n index fastest, so the 74 tiles in flight share one or
//     two 256-row panels of A and sweep W, which stays L2 resident (55 MB for gate|up);
//   * SwiGLU mode: W is stored with gate and up rows interleaved in blocks of 128, so the leader's half of the W tile
//     is 128 gate rows and the peer's half the matching 128 up rows; accumulator columns [0,128) / [128,256) of one
//     TMEM lane are gate_j / up_j of the same 128 output features.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "gemm_tc.h"

#ifndef UG_RES_PREFETCH
#define UG_RES_PREFETCH 0      // residual epilogue, L2 prefetch of the x tile: 0 none, 1 one bulk request per row, 2 four 2-D tensor requests
#endif

namespace ugk {
namespace {

constexpr int kGM = 128;                 // rows of the output tile per CTA (256 per pair)
constexpr int kGN = 256;                 // columns of the pair's tile = N of one MMA
constexpr int kGK = 64;                  // k per stage: one 128-byte swizzle row of bf16
constexpr int kStages = 6;
constexpr uint32_t kABytes = kGM * kGK * 2;            // 16 KB
constexpr uint32_t kBBytes = (kGN / 2) * kGK * 2;      // 16 KB: this CTA's half of the W tile
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr uint32_t kOutBytes = 128 * 128 * 2;          // staging tile of the epilogue: [128 rows x 128 bf16], two 64-wide panels
constexpr uint32_t kOffOut = kStages * kStageBytes;
constexpr uint32_t kOffBars = kOffOut + kOutBytes;
constexpr int kThreads = 384;              // warps 0-3: TMA, MMA, TMEM alloc, spare; warps 4-11: epilogue (two per TMEM lane quarter)

struct GemmBars {
    uint64_t full[kStages];      // leader's: bytes of BOTH CTAs' loads of a stage have landed
    uint64_t empty[kStages];     // per CTA: the MMAs reading this stage have completed (multicast commit)
    uint64_t acc_full[2];        // per CTA: accumulator buffer complete (multicast commit)
    uint64_t acc_empty[2];       // leader's: both CTAs' epilogue warps have drained the buffer (16 arrivals)
    uint32_t tmem_base;
};
constexpr uint32_t kSmemBytes = kOffBars + (uint32_t)sizeof(GemmBars) + 1024;     // + slack for the 1024-byte alignment

struct GemmParams {
    int M, N, K;
    int n_mblk, n_nblk, n_kblk;
    int mode;
    int n_rounds, tail_tiles, split_tail;     // see the unit scheduler in the kernel
    const __nv_bfloat16* bias;
    // q/k/v + RoPE mode: row m is position m % rope_L of batch m / rope_L; cos/sin fp32 [rope_L, 64]
    const float* rope_cos;
    const float* rope_sin;
    __nv_bfloat16* kv_cache;                  // nullable: bf16 [B, cache_L, 2, Hkv, 128]
    int rope_L, rope_Hq, rope_Hkv, cache_L, cache_row0;
    // RMSNorm folding: fp32 [M, ss_parts] partial sums of squares of the rows of A's fp32 source (nullable)
    const float* row_ss;
    int ss_parts;
    float rms_eps;
    // residual mode: x fp32 [M, N] (+= bf16(acc)), xb bf16 [M, N] = bf16(x), ss_out fp32 [M, N / 128]
    float* res_x;
    __nv_bfloat16* res_xb;
    float* res_ss;
    // every unit is a 256 x 128 half tile (N % 256 == 128, e.g. the 128-channel layers of the pixel decoder): n_nhalf = N / 128
    int all_half, n_nhalf;
    // 3x3 convolution as an implicit GEMM (plain mode, conv_W > 0): A rows are the pixels of [4 rows x 32 columns] boxes of an NHWC
    // activation [conv_B, conv_H, conv_W, Cin]; k block kb = tap * conv_cblk + channel block; the box of tap (ky, kx) is the
    // output box shifted by (ky - 1, kx - 1), rows / columns outside the image zero-filled by TMA
    int conv_B, conv_H, conv_W, conv_cblk, conv_boxes;
    int conv_bpr, conv_bpi, conv_taps;       // boxes per row of boxes / per image (partial boxes are clipped by TMA), 9 or 1 taps
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t n_clusters_x()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(s_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Arrive on a barrier of the leader CTA.  What the arrival hands over is a TMEM accumulator buffer whose reads (tcgen05.ld +
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync) have completed in this thread -- no global or shared-memory data -- so a
// RELAXED arrive is sufficient.  A releasing arrive compiles to a memory barrier (cluster scope: MEMBAR.ALL.GPU, the default CTA
// scope: MEMBAR.ALL.CTA) that first drains every outstanding global store of the thread: harmless next to TMA stores, but with
// the plain x / xb stores of the residual epilogue it was the hottest line of the kernel (28 % of the stall samples,
// profiles/r2_ncu_gemm_residual.md).
__device__ __forceinline__ void bar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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
// TMA tile load of this CTA's box into its own shared memory; the bytes are counted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
#ifndef UG_GEMM_L2HINT
#define UG_GEMM_L2HINT 1     // 1: the weight tiles are loaded with an L2 evict_last policy and the output tiles stored with evict_first
#endif
// L2 cache policies (createpolicy): the output of a launch (294 MB for gate|up) streams through L2 once and is not read back by
// this kernel, while W (55 MB) is swept by every round of tiles -- without a hint the stores evict W and it is fetched from HBM
// several times per launch (ncu: 456 MB read against 105 MB of operands)
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_pair_hint(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
// kStream: the tile is not read back soon (gate|up activation, convolution outputs: larger than L2) -> evict_first; the q/k/v tile
// and the logits are consumed by the launch that follows and keep the default policy (measured: no difference for either)
template <bool kStream = true>
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3)
{
#if UG_GEMM_L2HINT
    if (kStream)
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3), "l"(l2_policy_evict_first())
                 : "memory");
    else
#endif
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
template <bool kStream = true>
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1)
{
#if UG_GEMM_L2HINT
    if (kStream)
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1), "l"(l2_policy_evict_first())
                 : "memory");
    else
#endif
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// high word of every shared-memory descriptor used here: SBO = 1024 B (8 rows of 128 B), version 1, SWIZZLE_128B
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo_kmajor(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
// instruction descriptor: c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

// executed by all 32 lanes of the converged MMA warp of the leader CTA; one elected lane issues
__device__ __forceinline__ void mma_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void commit_pair(uint64_t* bar)
{
    asm volatile("{\n\t.reg .pred e;\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(
                     s_u32(bar))
                 : "memory");
}
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
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void pin16(uint32_t (&v)[16])
{
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(v[i]));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void pin32(uint32_t (&v)[32])
{
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// the arithmetic of ug_swiglu (backbone_ops.cu) on the bf16-rounded GEMM outputs, so that the fused and the unfused
// path differ only by the accumulation order inside the GEMM
__device__ __forceinline__ float silu_mul(float g, float u) { return __fdividef(g, 1.0f + __expf(-g)) * u; }

// 1 / rms of row m (RMSNorm folding); the partial sums are added in index order
__device__ __forceinline__ float row_rstd(const GemmParams& p, int m)
{
    if (!p.row_ss) return 1.0f;
    float s = 0.0f;
    if (m < p.M) {
        const float* sp = p.row_ss + (int64_t)m * p.ss_parts;
        for (int j = 0; j < p.ss_parts; ++j) s += __ldg(sp + j);
    }
    return rsqrtf(s / (float)p.K + p.rms_eps);
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2_tensor(const CUtensorMap* map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_o, const GemmParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    GemmBars* bars = reinterpret_cast<GemmBars*>(smem + kOffBars);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cid = (int)cluster_id_x(), ncl = (int)n_clusters_x();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_wh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_o) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            bar_init(&bars->full[s], 1);
            bar_init(&bars->empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            bar_init(&bars->acc_full[i], 1);
            bar_init(&bars->acc_empty[i], 16);       // 8 epilogue warps x 2 CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {       // all 512 TMEM columns of both SMs of the pair (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&bars->tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    // Work units of this cluster.  Rounds 0 .. n_rounds-1: full 256 x 256 tiles, tile = cid + round * ncl (round robin with
    // the n index fastest: the tiles in flight share a few 256-row panels of A and sweep W, all L2 resident).  The
    // remaining R < ncl tiles would leave most SM pairs idle for a whole tile time; when 2R <= ncl they are cut into
    // 256 x 128 halves (N = 128 MMAs, W boxes of 64 rows per CTA), one half per cluster, and the tail costs half a tile.
    const int n_units = p.n_rounds + ((p.split_tail ? cid < 2 * p.tail_tiles : cid < p.tail_tiles) ? 1 : 0);
    auto unit = [&](int i, int& mb, int& nb, int& half) {
        int tile;
        if (p.all_half) {                       // tiles of 256 x 128: n index (in halves) fastest
            tile = cid + i * ncl;
            mb = tile / p.n_nhalf;
            const int nh = tile - mb * p.n_nhalf;
            nb = nh >> 1;
            half = nh & 1;
            return;
        }
        if (i < p.n_rounds) {
            tile = cid + i * ncl;
            half = -1;
        } else if (p.split_tail) {
            tile = p.n_rounds * ncl + (cid >> 1);
            half = cid & 1;
        } else {
            tile = p.n_rounds * ncl + cid;
            half = -1;
        }
        mb = tile / p.n_nblk;
        nb = tile - mb * p.n_nblk;
    };

    if (warp == 0) {
        // =========================================================== TMA producer (one lane, in both CTAs)
        if (lane == 0) {
            uint32_t it = 0;
#if UG_GEMM_L2HINT
            const uint64_t w_policy = l2_policy_evict_last();
#endif
            for (int ui = 0; ui < n_units; ++ui) {
                int mb, nb, half;
                unit(ui, mb, nb, half);
                const int m0 = mb * 256 + (int)rank * kGM;
                const int n0 = half < 0 ? nb * kGN + (int)rank * (kGN / 2) : nb * kGN + half * 128 + (int)rank * 64;
                const CUtensorMap* wmap = half < 0 ? &tm_w : &tm_wh;
                const uint32_t bytes = half < 0 ? 2 * kStageBytes : 2 * (kABytes + kBBytes / 2);
                // convolution: this CTA's 128 rows are the pixels of box (cb, cy .. cy+3, cx .. cx+31)
                int cb = 0, cy = 0, cx = 0;
                if (p.conv_W) {
                    const int box = mb * 2 + (int)rank, bpr = p.conv_bpr, bpi = p.conv_bpi;
                    cb = box / bpi;
                    const int r2 = box - cb * bpi;
                    cy = (r2 / bpr) << 2;
                    cx = (r2 % bpr) << 5;
                }
                int tap = 0, cblk = 0;
                for (int kb = 0; kb < p.n_kblk; ++kb, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                    bar_wait(&bars->empty[s], ph ^ 1u);
                    if (leader) bar_expect_tx(&bars->full[s], bytes);
                    const uint32_t fb = mapa(s_u32(&bars->full[s]), 0);
                    const uint32_t dst = s_u32(smem + s * kStageBytes);
                    if (p.conv_W) {
                        const int ky = p.conv_taps == 1 ? 1 : tap / 3, kx = p.conv_taps == 1 ? 1 : tap - 3 * (tap / 3);
                        tma_load_4d_pair(dst, &tm_a, fb, cblk * kGK, cx + kx - 1, cy + ky - 1, cb);
                        if (++cblk == p.conv_cblk) { cblk = 0; ++tap; }
                    } else {
                        tma_load_2d_pair(dst, &tm_a, fb, kb * kGK, m0);
                    }
#if UG_GEMM_L2HINT
                    tma_load_2d_pair_hint(dst + kABytes, wmap, fb, kb * kGK, n0, w_policy);
#else
                    tma_load_2d_pair(dst + kABytes, wmap, fb, kb * kGK, n0);
#endif
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== MMA issuer (leader CTA, whole warp converged)
        if (leader) {
            uint32_t it = 0;
            for (int ui = 0; ui < n_units; ++ui) {
                int mb, nb, half;
                unit(ui, mb, nb, half);
                const uint32_t idesc = half < 0 ? make_idesc(256, kGN) : make_idesc(256, kGN / 2);
                const uint32_t as = (uint32_t)ui & 1u, aph = ((uint32_t)ui >> 1) & 1u;
                bar_wait(&bars->acc_empty[as], aph ^ 1u);
                tc_fence_after();
                const uint32_t d = tmem + as * 256u;
                for (int kb = 0; kb < p.n_kblk; ++kb, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
                    bar_wait(&bars->full[s], ph);
                    tc_fence_after();
                    const uint32_t a_lo = desc_lo_kmajor(s_u32(smem + s * kStageBytes));
                    const uint32_t b_lo = desc_lo_kmajor(s_u32(smem + s * kStageBytes + kABytes));
#pragma unroll
                    for (int k = 0; k < kGK / 16; ++k)
                        mma_pair(d, a_lo + 2u * k, b_lo + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    commit_pair(&bars->empty[s]);
                }
                commit_pair(&bars->acc_full[as]);
            }
        }
    } else if (warp >= 4) {
        // =========================================================== epilogue (both CTAs): thread r owns row r of the CTA's 128 rows
        const int ew = warp & 3;                         // the TMEM lane quarter this warp may read
        const int cg = (warp - 4) >> 2;                  // which 64 of 128 columns (= which swizzle panel of the staging tile) this warp converts
        const int r = ew * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(ew * 32) << 16);
        const uint32_t acc_empty_leader0 = mapa(s_u32(&bars->acc_empty[0]), 0);
        const uint32_t acc_empty_leader1 = mapa(s_u32(&bars->acc_empty[1]), 0);
        uint8_t* stg = smem + kOffOut;
        const uint32_t stg_row = (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        const bool issuer = threadIdx.x == 128;
        for (int ui = 0; ui < n_units; ++ui) {
            int mb, nb, half;
            unit(ui, mb, nb, half);
            const int m0 = mb * 256 + (int)rank * kGM;
            const uint32_t as = (uint32_t)ui & 1u, aph = ((uint32_t)ui >> 1) & 1u;
            // q/k/v + RoPE mode: the cos/sin values of this thread's row and columns (32 cg + [0,32)) are the same for both
            // heads of the unit; they are requested BEFORE the wait for the accumulator so that their latency is hidden
            float rope_c[32], rope_s[32];
            int rope_m = 0, rope_pos = 0;
            if (MODE == kGemmQkvRope) {
                rope_m = m0 + r;
                rope_pos = rope_m % p.rope_L;
                if ((nb * kGN >> 7) < p.rope_Hq + p.rope_Hkv) {           // at least the first head of the unit is rotated
                    const float4* cp = reinterpret_cast<const float4*>(p.rope_cos + (int64_t)rope_pos * 64 + cg * 32);
                    const float4* sp = reinterpret_cast<const float4*>(p.rope_sin + (int64_t)rope_pos * 64 + cg * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 cv = __ldg(cp + j), sv = __ldg(sp + j);
                        rope_c[4 * j] = cv.x; rope_c[4 * j + 1] = cv.y; rope_c[4 * j + 2] = cv.z; rope_c[4 * j + 3] = cv.w;
                        rope_s[4 * j] = sv.x; rope_s[4 * j + 1] = sv.y; rope_s[4 * j + 2] = sv.z; rope_s[4 * j + 3] = sv.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) { rope_c[j] = 1.0f; rope_s[j] = 0.0f; }
                }
            }
            const float rstd = row_rstd(p, m0 + r);
            if (MODE == kGemmResidual) {
                // The x tile a unit read-modify-writes is pulled into L2 ONE UNIT AHEAD (the first one at kernel start): when the
                // epilogue is the longer half of the pipeline the accumulator is already waiting, so a prefetch issued here for
                // the tile of THIS unit has no lead time (measured: no effect), and the 16 x 512 B a warp keeps in flight per batch
                // cannot hide DRAM latency (o_proj + residual was latency-bound at 2.4 TB/s of loads).
                auto prefetch_unit = [&](int uj) {
                    int mb2, nb2, half2;
                    unit(uj, mb2, nb2, half2);
                    const int pm0 = mb2 * 256 + (int)rank * kGM;
#if UG_RES_PREFETCH == 1
                    const int et = (int)threadIdx.x - 128, prow = pm0 + (et & 127), ph = et >> 7;
                    if (prow < p.M && (half2 < 0 || ph == half2))
                        prefetch_l2_bulk(p.res_x + (int64_t)prow * p.N + nb2 * kGN + ph * 128, 512u);
#elif UG_RES_PREFETCH == 2
                    if (issuer && pm0 < p.M) {     // tm_o describes x (fp32 [M, N], boxes of 128 rows x 64 columns) in this mode
                        const int c0 = nb2 * kGN + (half2 > 0 ? 128 : 0), nbox = half2 < 0 ? 4 : 2;
                        for (int i = 0; i < nbox; ++i) prefetch_l2_tensor(&tm_o, c0 + 64 * i, pm0);
                    }
#endif
                };
                if (ui == 0) prefetch_unit(0);
                if (ui + 1 < n_units) prefetch_unit(ui + 1);
            }
            // residual mode: the x values this thread read-modify-writes in the second pass (below) do not depend on the
            // accumulator, so the first batch is requested BEFORE the wait for it and every later batch one batch ahead: the
            // loads' latency (the kernel was latency-bound: tensor pipe 48 %, DRAM 43 % on the o_proj shape) hides behind the
            // wait, the tensor-memory loads and the staging pass.
            float4 xv0[8];
            auto load_x = [&](int hh, int bch, float4 (&dst)[8]) {
                const int w8 = warp - 4;
                const int col0 = nb * kGN + hh * 128;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = w8 * 16 + bch * 8 + i;
                    dst[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (m0 + row < p.M) dst[i] = *reinterpret_cast<const float4*>(p.res_x + (int64_t)(m0 + row) * p.N + col0 + lane * 4);
                }
            };
            if (MODE == kGemmResidual) load_x(half < 0 ? 0 : half, 0, xv0);
            bar_wait(&bars->acc_full[as], aph);
            tc_fence_after();
            const uint32_t acc = lane_addr + as * 256u;
            // this warp's share of the accumulator buffer is in registers: hand the buffer back to the MMA warp
            auto release_acc = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive_cluster(as ? acc_empty_leader1 : acc_empty_leader0);
            };
            if (MODE == kGemmSwiGLU) {
                if (issuer) bulk_wait_read0();           // the previous unit's TMA store has finished reading the staging tile
                epi_bar();
#pragma unroll 1
                for (int cc = 2 * cg; cc < 2 * cg + 2; ++cc) {
                    uint32_t g[32], u[32];
                    tmem_ld32_async(acc + cc * 32, g);
                    tmem_ld32_async(acc + 128 + cc * 32, u);
                    tmem_ld_wait();
                    pin32(g);
                    pin32(u);
                    if (cc == 2 * cg + 1) release_acc();
                    uint32_t o[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const float g0 = bf16_round(__uint_as_float(g[2 * q]) * rstd), g1 = bf16_round(__uint_as_float(g[2 * q + 1]) * rstd);
                        const float u0 = bf16_round(__uint_as_float(u[2 * q]) * rstd), u1 = bf16_round(__uint_as_float(u[2 * q + 1]) * rstd);
                        o[q] = pack_bf16(silu_mul(g0, u0), silu_mul(g1, u1));
                    }
                    uint8_t* panel = stg + (cc >> 1) * 16384 + stg_row;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t chunk = (uint32_t)((cc & 1) * 4 + j);
                        *reinterpret_cast<uint4*>(panel + ((chunk ^ sw) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    }
                }
                fence_async_smem();
                epi_bar();
                if (issuer && m0 < p.M) {
                    tma_store_2d(&tm_o, s_u32(stg), nb * 128, m0);
                    tma_store_2d(&tm_o, s_u32(stg) + 16384, nb * 128 + 64, m0);
                    bulk_commit();
                }
            } else if (MODE == kGemmQkvRope) {
                // q/k/v projection: every 128-column half is one head.  Thread r holds columns d and d + 64 of its row
                // (32 + 32 columns per warp group), i.e. both members of each rotate_half pair: bias, bf16 rounding (what
                // the unfused path stores), rotation by the row's position, bf16 -- then the staging tile / TMA store into
                // the fused qkv activation and, for k and v heads, the row's bytes straight into the K/V cache.
                const int h0 = half < 0 ? 0 : half, h1 = half < 0 ? 2 : half + 1;
#pragma unroll 1
                for (int hh = h0; hh < h1; ++hh) {
                    const uint32_t acc_h = acc + (half < 0 ? (uint32_t)hh * 128u : 0u);
                    const int col0 = nb * kGN + hh * 128;
                    const int hd = col0 >> 7;                          // head index in [q heads | k heads | v heads]
                    const bool rot = hd < p.rope_Hq + p.rope_Hkv;
                    __nv_bfloat16* crow = nullptr;
                    if (p.kv_cache && hd >= p.rope_Hq && rope_m < p.M) {
                        const int kvh = hd - p.rope_Hq;                  // [0, Hkv): k heads, [Hkv, 2 Hkv): v heads
                        const int64_t bb = rope_m / p.rope_L;
                        crow = p.kv_cache + (((bb * p.cache_L + p.cache_row0 + rope_pos) * 2 + (kvh >= p.rope_Hkv ? 1 : 0)) * p.rope_Hkv +
                                             kvh % p.rope_Hkv) * 128 + cg * 32;
                    }
                    if (issuer) bulk_wait_read0();
                    epi_bar();
#pragma unroll
                    for (int sc = 0; sc < 2; ++sc) {                   // 16 + 16 columns at a time: d = 32 cg + 16 sc + [0,16) and d + 64
                        uint32_t lo[16], hi[16];
                        tmem_ld16_async(acc_h + cg * 32 + sc * 16, lo);
                        tmem_ld16_async(acc_h + 64 + cg * 32 + sc * 16, hi);
                        tmem_ld_wait();
                        pin16(lo);
                        pin16(hi);
                        if (hh == h1 - 1 && sc == 1) release_acc();
                        uint32_t o1[8], o2[8];
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            uint4 b1 = make_uint4(0u, 0u, 0u, 0u), b2 = b1;
                            if (p.bias) {
                                b1 = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + cg * 32 + sc * 16) + j);
                                b2 = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + 64 + cg * 32 + sc * 16) + j);
                            }
                            const uint32_t bw1[4] = {b1.x, b1.y, b1.z, b1.w}, bw2[4] = {b2.x, b2.y, b2.z, b2.w};
                            float x1[8], x2[8];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                x1[2 * e] = bf16_round(fmaf(__uint_as_float(lo[8 * j + 2 * e]), rstd, __uint_as_float(bw1[e] << 16)));
                                x1[2 * e + 1] = bf16_round(fmaf(__uint_as_float(lo[8 * j + 2 * e + 1]), rstd, __uint_as_float(bw1[e] & 0xffff0000u)));
                                x2[2 * e] = bf16_round(fmaf(__uint_as_float(hi[8 * j + 2 * e]), rstd, __uint_as_float(bw2[e] << 16)));
                                x2[2 * e + 1] = bf16_round(fmaf(__uint_as_float(hi[8 * j + 2 * e + 1]), rstd, __uint_as_float(bw2[e] & 0xffff0000u)));
                            }
                            if (rot) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const float a = x1[e], b = x2[e];
                                    const float cs = rope_c[16 * sc + 8 * j + e], sn = rope_s[16 * sc + 8 * j + e];
                                    x1[e] = a * cs - b * sn;
                                    x2[e] = b * cs + a * sn;
                                }
                            }
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                o1[4 * j + e] = pack_bf16(x1[2 * e], x1[2 * e + 1]);
                                o2[4 * j + e] = pack_bf16(x2[2 * e], x2[2 * e + 1]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t chunk = (uint32_t)(cg * 4 + sc * 2 + j);
                            const uint4 v1 = make_uint4(o1[4 * j], o1[4 * j + 1], o1[4 * j + 2], o1[4 * j + 3]);
                            const uint4 v2 = make_uint4(o2[4 * j], o2[4 * j + 1], o2[4 * j + 2], o2[4 * j + 3]);
                            *reinterpret_cast<uint4*>(stg + stg_row + ((chunk ^ sw) << 4)) = v1;
                            *reinterpret_cast<uint4*>(stg + 16384 + stg_row + ((chunk ^ sw) << 4)) = v2;
                            if (crow) {
                                *reinterpret_cast<uint4*>(crow + sc * 16 + 8 * j) = v1;
                                *reinterpret_cast<uint4*>(crow + 64 + sc * 16 + 8 * j) = v2;
                            }
                        }
                    }
                    fence_async_smem();
                    epi_bar();
                    if (issuer && m0 < p.M) {
                        tma_store_2d<false>(&tm_o, s_u32(stg), col0, m0);
                        tma_store_2d<false>(&tm_o, s_u32(stg) + 16384, col0 + 64, m0);
                        bulk_commit();
                    }
                }
            } else {
                const int h0 = half < 0 ? 0 : half, h1 = half < 0 ? 2 : half + 1;
#pragma unroll 1
                for (int hh = h0; hh < h1; ++hh) {
                    const uint32_t acc_h = acc + (half < 0 ? (uint32_t)hh * 128u : 0u);
                    const int col0 = nb * kGN + hh * 128;             // first output column of this 128-column half
                    if (MODE != kGemmResidual && issuer) bulk_wait_read0();
                    epi_bar();
#pragma unroll 1
                    for (int cc = 2 * cg; cc < 2 * cg + 2; ++cc) {
                        uint32_t v[32];
                        tmem_ld32_async(acc_h + cc * 32, v);
                        tmem_ld_wait();
                        pin32(v);
                        if (hh == h1 - 1 && cc == 2 * cg + 1) release_acc();
                        uint32_t o[16];
                        if (p.bias) {
                            const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0 + cc * 32);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 bv = __ldg(bp + j);
                                const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float b0 = __uint_as_float(bw[e] << 16), b1 = __uint_as_float(bw[e] & 0xffff0000u);
                                    o[4 * j + e] = pack_bf16(fmaf(__uint_as_float(v[8 * j + 2 * e]), rstd, b0), fmaf(__uint_as_float(v[8 * j + 2 * e + 1]), rstd, b1));
                                }
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 16; ++q) o[q] = pack_bf16(__uint_as_float(v[2 * q]) * rstd, __uint_as_float(v[2 * q + 1]) * rstd);
                        }
                        {
                            uint8_t* panel = stg + (cc >> 1) * 16384 + stg_row;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint32_t chunk = (uint32_t)((cc & 1) * 4 + j);
                                *reinterpret_cast<uint4*>(panel + ((chunk ^ sw) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                            }
                        }
                    }
                    if (MODE == kGemmResidual) {
                        // second pass, row-major and coalesced: the staging tile now holds bf16(acc) = the value the unfused
                        // linear layer would have written.  Warp w takes rows 16 w .. 16 w + 15; a lane owns 4 consecutive
                        // columns, so one warp access is one 512-byte row segment of x (fp32) / 256 bytes of xb (bf16).
                        epi_bar();
                        const int w8 = warp - 4;
                        const uint32_t lane_off = (uint32_t)(lane >> 4) * 16384u + (uint32_t)(lane & 1) * 8u;
                        const uint32_t lane_chunk = (uint32_t)(lane & 15) >> 1;
                        float4 xv1[8];
                        load_x(hh, 1, xv1);
                        auto second_pass = [&](int bch, const float4 (&xv)[8]) {
                            float sq[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int row = w8 * 16 + bch * 8 + i;
                                const uint2 d = *reinterpret_cast<const uint2*>(stg + lane_off + (uint32_t)row * 128u +
                                                                                ((lane_chunk ^ (uint32_t)(row & 7)) << 4));
                                float4 y;
                                y.x = xv[i].x + bf16lo(d.x);
                                y.y = xv[i].y + bf16hi(d.x);
                                y.z = xv[i].z + bf16lo(d.y);
                                y.w = xv[i].w + bf16hi(d.y);
                                float q2 = y.x * y.x;
                                q2 = fmaf(y.y, y.y, q2);
                                q2 = fmaf(y.z, y.z, q2);
                                sq[i] = fmaf(y.w, y.w, q2);
                                if (m0 + row < p.M) {
                                    const int64_t gi = (int64_t)(m0 + row) * p.N + col0 + lane * 4;
                                    *reinterpret_cast<float4*>(p.res_x + gi) = y;
                                    *reinterpret_cast<uint2*>(p.res_xb + gi) = make_uint2(pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
                                }
                            }
                            // the 8 row sums of the batch in 9 shuffles instead of 40: a butterfly that halves the number of
                            // rows a lane carries at every step (bits 4, 3, 2 of the lane index pick the row), then the last two
                            // steps on the single remaining value.  Every row is summed by the same tree: deterministic.
                            float s4[4], s2[2];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float keep = (lane & 16) ? sq[j + 4] : sq[j], send = (lane & 16) ? sq[j] : sq[j + 4];
                                s4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                            }
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const float keep = (lane & 8) ? s4[j + 2] : s4[j], send = (lane & 8) ? s4[j] : s4[j + 2];
                                s2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                            }
                            float s1 = ((lane & 4) ? s2[1] : s2[0]) + __shfl_xor_sync(0xffffffffu, (lane & 4) ? s2[0] : s2[1], 4);
                            s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
                            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                            const int ri = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                            const int srow = m0 + w8 * 16 + bch * 8 + ri;
                            if ((lane & 3) == 0 && srow < p.M) p.res_ss[(int64_t)srow * (p.N >> 7) + (col0 >> 7)] = s1;
                        };
                        second_pass(0, xv0);
                        if (hh + 1 < h1) load_x(hh + 1, 0, xv0);      // the next half's first batch: in flight across its staging pass
                        second_pass(1, xv1);
                    } else {
                        fence_async_smem();
                        epi_bar();
                        if (issuer && p.conv_W) {            // the rows of the staging tile are the pixels of one box of the output image
                            const int box = mb * 2 + (int)rank;
                            if (box < p.conv_boxes) {
                                const int bpr = p.conv_bpr, bpi = p.conv_bpi;
                                const int cb = box / bpi, r2 = box - cb * bpi, cy = (r2 / bpr) << 2, cx = (r2 % bpr) << 5;
                                tma_store_4d(&tm_o, s_u32(stg), col0, cx, cy, cb);
                                tma_store_4d(&tm_o, s_u32(stg) + 16384, col0 + 64, cx, cy, cb);
                                bulk_commit();
                            }
                        } else if (issuer && m0 < p.M) {         // codebook head: the decode step reads the logits at once (default policy)
                            tma_store_2d<false>(&tm_o, s_u32(stg), col0, m0);
                            tma_store_2d<false>(&tm_o, s_u32(stg) + 16384, col0 + 64, m0);
                            bulk_commit();
                        }
                    }
                }
            }
        }
        if (issuer) bulk_wait0();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();            // the peer may still receive multicast arrivals / remote arrives until both are here
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
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

// row-major bf16 [rows, cols] with row stride ld elements; box = [box_rows x 64 cols], 128-byte swizzle
// row-major fp32 [rows, cols] (contiguous rows), boxes of 128 rows x 64 columns, no swizzle: only used to PREFETCH tiles of the
// residual stream into L2 (cp.async.bulk.prefetch.tensor), never to move them into shared memory
bool make_map_2d_f32(EncodeTiledFn enc, CUtensorMap* m, const void* base, int64_t rows, int64_t cols)
{
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// NHWC bf16 activation [B, H, W, C] (contiguous): boxes of [4 rows x 32 columns] pixels x 64 channels, 128-byte swizzle -- in
// shared memory a box is a [128 pixels x 64 channels] K-major tile like any other A operand (and like the staging tile of the
// epilogue, for the store)
bool make_map_nhwc(EncodeTiledFn enc, CUtensorMap* m, const void* base, int B, int H, int W, int C)
{
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, 32, 4, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_map_2d(EncodeTiledFn enc, CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows = 128)
{
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int launch_gemm_tc(const GemmArgs& g, int num_sms, cudaStream_t stream, char* msg, size_t msg_len)
{
    if (g.mode != kGemmPlain && g.mode != kGemmSwiGLU && g.mode != kGemmQkvRope && g.mode != kGemmResidual) { snprintf(msg, msg_len, "gemm: unsupported mode %d", g.mode); return -4; }
    const bool all_half = g.N % kGN == 128;               // every unit is a 256 x 128 half tile
    if (g.M < 1 || g.N < 128 || g.N % 128 || g.K < kGK || g.K % kGK || (all_half && g.mode != kGemmPlain)) {
        snprintf(msg, msg_len, "gemm: need N %% 256 == 0 (plain mode: N %% 128 == 0) and K %% 64 == 0 (M=%d N=%d K=%d)", g.M, g.N, g.K);
        return -4;
    }
    const bool conv = g.conv_W > 0;
    const int conv_cblk = conv ? (g.conv_Cin + 63) / 64 : 0;
    if (conv && (g.mode != kGemmPlain || g.conv_Cin % 8 || g.conv_Cout % 8 || (g.conv_taps != 9 && g.conv_taps != 1) ||
                 g.K != g.conv_taps * conv_cblk * 64 || g.N < g.conv_Cout || g.conv_B < 1 || g.conv_H < 1)) {
        snprintf(msg, msg_len, "conv: need Cin %% 8 == 0, Cout %% 8 == 0, taps 9 or 1, K = taps * ceil64(Cin), N = ceil128(Cout) (B=%d H=%d W=%d Cin=%d Cout=%d)",
                 g.conv_B, g.conv_H, g.conv_W, g.conv_Cin, g.conv_Cout);
        return -4;
    }
    if (g.lda % 8 || g.ldw % 8 || (g.mode != kGemmResidual && g.ldo % 8)) { snprintf(msg, msg_len, "gemm: row strides must be multiples of 8 elements"); return -5; }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { snprintf(msg, msg_len, "cuTensorMapEncodeTiled entry point not available"); return -6; }
    const int n_out = g.mode == kGemmSwiGLU ? g.N / 2 : g.N;
    CUtensorMap ta, tw, twh, to;
    const bool residual = g.mode == kGemmResidual;       // writes x / xb / ss with plain stores: no output tensor map
    if (residual && (!g.res_x || !g.res_xb || !g.res_ss)) { snprintf(msg, msg_len, "gemm: the residual mode needs x, xb and ss"); return -1; }
    if (g.row_ss && g.ss_parts < 1) { snprintf(msg, msg_len, "gemm: row_ss needs ss_parts >= 1"); return -4; }
    if (!(conv ? make_map_nhwc(enc, &ta, g.a, g.conv_B, g.conv_H, g.conv_W, g.conv_Cin) : make_map_2d(enc, &ta, g.a, g.M, g.K, g.lda)) ||
        !make_map_2d(enc, &tw, g.w, g.N, g.K, g.ldw) || !make_map_2d(enc, &twh, g.w, g.N, g.K, g.ldw, 64) ||
        !(residual ? make_map_2d_f32(enc, &to, g.res_x, g.M, g.N)
                   : conv ? make_map_nhwc(enc, &to, g.out, g.conv_B, g.conv_H, g.conv_W, g.conv_Cout) : make_map_2d(enc, &to, g.out, g.M, n_out, g.ldo))) {
        snprintf(msg, msg_len, "cuTensorMapEncodeTiled rejected the a/w/out layout");
        return -3;
    }
    GemmParams p;
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.n_mblk = (g.M + 255) / 256;
    p.n_nblk = g.N / kGN;
    p.all_half = all_half ? 1 : 0;
    p.n_nhalf = g.N / 128;
    p.conv_B = g.conv_B; p.conv_H = g.conv_H; p.conv_W = conv ? g.conv_W : 0; p.conv_cblk = conv_cblk;
    p.conv_bpr = conv ? (g.conv_W + 31) / 32 : 0;
    p.conv_bpi = conv ? p.conv_bpr * ((g.conv_H + 3) / 4) : 0;
    p.conv_boxes = conv ? g.conv_B * p.conv_bpi : 0;
    p.conv_taps = g.conv_taps;
    if (conv) { p.n_mblk = (p.conv_boxes + 1) / 2; p.M = p.conv_boxes * 128; }
    p.n_kblk = g.K / kGK;
    p.mode = g.mode;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(g.bias);
    p.rope_cos = g.rope_cos; p.rope_sin = g.rope_sin;
    p.kv_cache = reinterpret_cast<__nv_bfloat16*>(g.kv_cache);
    p.rope_L = g.rope_L; p.rope_Hq = g.rope_Hq; p.rope_Hkv = g.rope_Hkv; p.cache_L = g.cache_L; p.cache_row0 = g.cache_row0;
    p.row_ss = g.row_ss; p.ss_parts = g.ss_parts; p.rms_eps = g.rms_eps;
    p.res_x = g.res_x; p.res_xb = reinterpret_cast<__nv_bfloat16*>(g.res_xb); p.res_ss = g.res_ss;
    if (g.mode == kGemmQkvRope && (!g.rope_cos || !g.rope_sin || g.rope_L < 1 || g.N != (g.rope_Hq + 2 * g.rope_Hkv) * 128)) {
        snprintf(msg, msg_len, "gemm: q/k/v + RoPE mode needs cos/sin tables, rope_L >= 1 and N == (Hq + 2 Hkv) * 128");
        return -4;
    }
    static bool attr_set[64] = {};                       // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<kGemmPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<kGemmSwiGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<kGemmQkvRope>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<kGemmResidual>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) { snprintf(msg, msg_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int max_clusters = num_sms / 2 > 0 ? num_sms / 2 : 1;
    const int n_tiles = all_half ? p.n_mblk * p.n_nhalf : p.n_mblk * p.n_nblk;
    const bool can_split = g.mode != kGemmSwiGLU && !all_half;
    int n_clusters = max_clusters;
    if (n_tiles < max_clusters) n_clusters = can_split && 2 * n_tiles <= max_clusters ? 2 * n_tiles : n_tiles;
    p.n_rounds = n_tiles / n_clusters;
    p.tail_tiles = n_tiles - p.n_rounds * n_clusters;
    p.split_tail = can_split && p.tail_tiles > 0 && 2 * p.tail_tiles <= n_clusters ? 1 : 0;
    const dim3 grid(2u * (unsigned)n_clusters);
    switch (g.mode) {
        case kGemmPlain: gemm_tc_kernel<kGemmPlain><<<grid, kThreads, kSmemBytes, stream>>>(ta, tw, twh, to, p); break;
        case kGemmSwiGLU: gemm_tc_kernel<kGemmSwiGLU><<<grid, kThreads, kSmemBytes, stream>>>(ta, tw, twh, to, p); break;
        case kGemmResidual: gemm_tc_kernel<kGemmResidual><<<grid, kThreads, kSmemBytes, stream>>>(ta, tw, twh, to, p); break;
        default: gemm_tc_kernel<kGemmQkvRope><<<grid, kThreads, kSmemBytes, stream>>>(ta, tw, twh, to, p); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(msg, msg_len, "gemm launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // namespace ugk
