// omni_attention.cu -- omni-attention forward for sm_100a on tcgen05 / TMEM / TMA.
//
// Replaces the attention call of the Qwen2 backbone as the reference drives it
// (models/unigen.py:260-266 -> transformers sdpa_attention_forward -> F.scaled_dot_product_attention
// with the dense additive [B,1,L,L] mask built by training/prompting_utils.py:975-1020).  The mask
// is never materialised: it is the closed form of two integers per sequence (pad_len P_b,
// soi_idx S_b) and fully masked KV tiles are skipped.
//
// One CTA = one 128-row query tile of one (batch, q-head); 192 threads:
//   warps 0-3  softmax / correction / epilogue: thread r owns query row r (TMEM lane r)
//   warp  4    TMA producer: Q once, then K_j / V_j tiles (2-deep rings), SWIZZLE_128B boxes
//   warp  5    MMA issuer (one elected lane): S_j = Q K_j^T and O += P_j V_j with tcgen05.mma
//              (kind::f16, bf16 x bf16 -> fp32, M=128 N=128 K=16), accumulators in TMEM
// TMEM columns: S ping [0,128), S pong [128,256), O [256,384).
// S is double buffered so that QK^T of tile j+1 runs on the tensor core while the softmax warps
// work on tile j; P_j goes through shared memory in the K-major SWIZZLE_128B UMMA layout.
// GQA: the 6 query heads of a KV head read the same K/V tiles (L2 hits); nothing is expanded.
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cmath>

#include "omni_attention.h"

namespace ugk {

constexpr int kBM = 128;             // query rows per CTA
constexpr int kBN = 128;             // kv rows per tile
constexpr int kD = 128;              // head dim
constexpr int kAttnThreads = 192;
constexpr uint32_t kTileBytes = kBM * kD * 2;          // 32 KB: one [128 x 128] bf16 tile (two 64-wide swizzle panels)
constexpr uint32_t kPanelBytes = kTileBytes / 2;       // 16 KB: [128 rows x 64 cols] panel, 128-byte rows
constexpr int kKVStages = 2;

// shared memory map (dynamic, 1024-byte aligned)
constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffK = kOffQ + kTileBytes;
constexpr uint32_t kOffV = kOffK + kKVStages * kTileBytes;
constexpr uint32_t kOffP = kOffV + kKVStages * kTileBytes;
constexpr uint32_t kOffBar = kOffP + kTileBytes;
constexpr uint32_t kAttnSmem = kOffBar + 256 + 1024;   // + barriers + alignment slack

struct AttnBars {
    uint64_t q_full;
    uint64_t k_full[kKVStages], k_empty[kKVStages];
    uint64_t v_full[kKVStages], v_empty[kKVStages];
    uint64_t s_full[2], s_empty[2];
    uint64_t p_full;
    uint64_t pv_done;
    uint32_t tmem_base;
};

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
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
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
    int n_qtiles;
    float* dbg;             // optional: dumps S and P of (b=0, h=0, q tile 0, kv tile 0)
};

__global__ void __launch_bounds__(kAttnThreads, 1)
omni_attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const AttnParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    AttnBars* bars = reinterpret_cast<AttnBars*>(smem + kOffBar);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.a.L;
    // heaviest tiles (image rows) first: reverse the q-tile index
    const int qt = p.n_qtiles - 1 - (int)blockIdx.x;
    const int h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (p.a.Hq / p.a.Hkv);
    const int i0 = qt * kBM;
    const int i1 = min(i0 + kBM, L);
    int P = p.a.pad_len[b], S = p.a.soi_idx[b];
    P = max(0, min(P, L));
    S = max(P, min(S, L));

    // kv range needed by this query tile (closed form of the omni mask)
    const int jlo = (i0 < P) ? 0 : P;
    const int jhi = (i1 > S) ? L : i1;
    const int t0 = jlo / kBN;
    const int t1 = (jhi + kBN - 1) / kBN;
    const int n_tiles = t1 - t0;               // >= 1

    if (tid == 0) {
        bar_init(&bars->q_full, 1);
        for (int s = 0; s < kKVStages; ++s) {
            bar_init(&bars->k_full[s], 1); bar_init(&bars->k_empty[s], 1);
            bar_init(&bars->v_full[s], 1); bar_init(&bars->v_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) { bar_init(&bars->s_full[s], 1); bar_init(&bars->s_empty[s], 128); }
        bar_init(&bars->p_full, 128);
        bar_init(&bars->pv_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {       // TMEM: 512 columns (S ping, S pong, O) -- one CTA per SM, nobody else competes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&bars->tmem_base)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 4) {
        // =========================================================== TMA producer
        if (lane == 0) {
            bar_expect_tx(&bars->q_full, kTileBytes);
            tma_load_4d(smem + kOffQ, &tm_q, &bars->q_full, 0, i0, h, b);
            tma_load_4d(smem + kOffQ + kPanelBytes, &tm_q, &bars->q_full, 64, i0, h, b);
            for (int j = 0; j < n_tiles; ++j) {
                const int s = j % kKVStages;
                const uint32_t ph = (uint32_t)(j / kKVStages) & 1u;
                const int j0 = (t0 + j) * kBN;
                bar_wait(&bars->k_empty[s], ph ^ 1u);
                bar_expect_tx(&bars->k_full[s], kTileBytes);
                tma_load_4d(smem + kOffK + s * kTileBytes, &tm_k, &bars->k_full[s], 0, j0, hk, b);
                tma_load_4d(smem + kOffK + s * kTileBytes + kPanelBytes, &tm_k, &bars->k_full[s], 64, j0, hk, b);
                bar_wait(&bars->v_empty[s], ph ^ 1u);
                bar_expect_tx(&bars->v_full[s], kTileBytes);
                tma_load_4d(smem + kOffV + s * kTileBytes, &tm_v, &bars->v_full[s], 0, j0, hk, b);
                tma_load_4d(smem + kOffV + s * kTileBytes + kPanelBytes, &tm_v, &bars->v_full[s], 64, j0, hk, b);
            }
        }
    } else if (warp == 5) {
        // =========================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc(kBM, kBN, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc(kBM, kD, 0, 1);
            const uint32_t q_addr = s_u32(smem + kOffQ), p_addr = s_u32(smem + kOffP);
            auto issue_qk = [&](int j) {
                const int s = j % kKVStages, sb = j & 1;
                bar_wait(&bars->k_full[s], (uint32_t)(j / kKVStages) & 1u);
                if (j >= 2) bar_wait(&bars->s_empty[sb], (uint32_t)((j >> 1) + 1) & 1u);
                tc_fence_after();
                const uint32_t k_addr = s_u32(smem + kOffK + s * kTileBytes);
#pragma unroll
                for (int k = 0; k < kD / 16; ++k) {
                    // 16 bf16 along K = 32 bytes inside the 128-byte swizzle row; panel switch at k = 4
                    const uint32_t off = (uint32_t)(k >> 2) * kPanelBytes + (uint32_t)(k & 3) * 32u;
                    tc_mma(tmem + (uint32_t)sb * 128u, desc_kmajor(q_addr + off), desc_kmajor(k_addr + off), idesc_qk, k > 0);
                }
                tc_commit(&bars->k_empty[s]);
                tc_commit(&bars->s_full[sb]);
            };
            bar_wait(&bars->q_full, 0);
            issue_qk(0);
            for (int j = 0; j < n_tiles; ++j) {
                if (j + 1 < n_tiles) issue_qk(j + 1);
                const int s = j % kKVStages;
                bar_wait(&bars->p_full, (uint32_t)j & 1u);
                bar_wait(&bars->v_full[s], (uint32_t)(j / kKVStages) & 1u);
                tc_fence_after();
                const uint32_t v_addr = s_u32(smem + kOffV + s * kTileBytes);
#pragma unroll
                for (int k = 0; k < kBN / 16; ++k) {
                    const uint32_t a_off = (uint32_t)(k >> 2) * kPanelBytes + (uint32_t)(k & 3) * 32u;   // P: K-major
                    const uint32_t b_off = (uint32_t)k * 16u * 128u;                                      // V: 16 kv rows
                    tc_mma(tmem + 256u, desc_kmajor(p_addr + a_off), desc_mnmajor(v_addr + b_off), idesc_pv,
                           (j > 0 || k > 0) ? 1u : 0u);
                }
                tc_commit(&bars->v_empty[s]);
                tc_commit(&bars->pv_done);
            }
        }
    } else {
        // =========================================================== softmax / correction / epilogue
        const int r = tid;                          // row inside the tile == TMEM lane
        const int i = i0 + r;
        // allowed kv range [lo, hi) of this row (closed-form omni mask); rows beyond L are harmless
        int lo, hi;
        if (i >= S) { lo = P; hi = L; }
        else if (i >= P) { lo = P; hi = i + 1; }
        else { lo = 0; hi = i + 1; }
        hi = min(hi, L);
        // tile-uniform bounds for the "no masking needed" fast path
        const int max_lo = (i1 > P) ? P : 0;
        const int min_hi = (i0 < S) ? (i0 + 1) : L;
        const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
        const float c = p.scale_log2;
        float m_run = -INFINITY, l_run = 0.0f;
        uint32_t pk[64];

        for (int j = 0; j < n_tiles; ++j) {
            const int sb = j & 1;
            const int j0 = (t0 + j) * kBN;
            const bool full_tile = (j0 >= max_lo) && (j0 + kBN <= min_hi);
            const int clo = lo - j0, chi = hi - j0;           // valid columns [clo, chi)
            bar_wait(&bars->s_full[sb], (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            const uint32_t s_addr = tmem + lane_base + (uint32_t)sb * 128u;
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t v[32];
                tmem_ld32(s_addr + ch * 32, v);
                if (p.dbg && blockIdx.x == gridDim.x - 1 && h == 0 && b == 0 && j == 0)
                    for (int q = 0; q < 32; ++q) p.dbg[r * 128 + ch * 32 + q] = __uint_as_float(v[q]);
                if (full_tile) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) mx = fmaxf(mx, __uint_as_float(v[q]));
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const int col = ch * 32 + q;
                        mx = fmaxf(mx, (col >= clo && col < chi) ? __uint_as_float(v[q]) : -INFINITY);
                    }
                }
            }
            const float m_new = fmaxf(m_run, mx * c);
            const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
            const float alpha = ex2(m_run - m_use);           // 0 when m_run == -inf
            // pass 2: probabilities, row sum, bf16 pack
            float rs = 0.0f;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t v[32];
                tmem_ld32(s_addr + ch * 32, v);
#pragma unroll
                for (int q = 0; q < 32; q += 2) {
                    const int col = ch * 32 + q;
                    float e0 = ex2(fmaf(__uint_as_float(v[q]), c, -m_use));
                    float e1 = ex2(fmaf(__uint_as_float(v[q + 1]), c, -m_use));
                    if (!full_tile) {
                        e0 = (col >= clo && col < chi) ? e0 : 0.0f;
                        e1 = (col + 1 >= clo && col + 1 < chi) ? e1 : 0.0f;
                    }
                    const uint32_t w = pack_bf16(e0, e1);
                    pk[(ch * 32 + q) >> 1] = w;
                    // the row sum uses the bf16-rounded values the tensor core will see
                    rs += __uint_as_float(w << 16) + __uint_as_float(w & 0xffff0000u);
                }
            }
            tc_fence_before();
            bar_arrive(&bars->s_empty[sb]);
            // O must be complete (PV_{j-1}) before it is rescaled and before P is overwritten
            if (j > 0) {
                bar_wait(&bars->pv_done, (uint32_t)(j - 1) & 1u);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t v[32];
                        tmem_ld32(tmem + lane_base + 256u + ch * 32, v);
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) * alpha);
                        tmem_st32(tmem + lane_base + 256u + ch * 32, v);
                    }
                    tmem_st_wait();
                }
            }
            l_run = l_run * alpha + rs;
            m_run = m_new;
            // P -> shared memory, K-major SWIZZLE_128B: 16-byte chunk index XOR (row & 7)
            {
                uint8_t* prow = smem + kOffP + (uint32_t)r * 128u;
#pragma unroll
                for (int cc = 0; cc < 16; ++cc) {             // 16 chunks of 8 columns
                    const uint32_t panel = (uint32_t)(cc >> 3) * kPanelBytes;
                    const uint32_t chunk = (uint32_t)((cc & 7) ^ (r & 7)) << 4;
                    *reinterpret_cast<uint4*>(prow + panel + chunk) = make_uint4(pk[cc * 4], pk[cc * 4 + 1], pk[cc * 4 + 2], pk[cc * 4 + 3]);
                }
            }
            if (p.dbg && blockIdx.x == gridDim.x - 1 && h == 0 && b == 0 && j == 0)
                for (int q = 0; q < 64; ++q) {
                    p.dbg[16384 + r * 128 + 2 * q] = __uint_as_float(pk[q] << 16);
                    p.dbg[16384 + r * 128 + 2 * q + 1] = __uint_as_float(pk[q] & 0xffff0000u);
                }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            bar_arrive(&bars->p_full);
        }
        // epilogue: O / l -> bf16 -> global
        bar_wait(&bars->pv_done, (uint32_t)(n_tiles - 1) & 1u);
        tc_fence_after();
        const float inv_l = 1.0f / l_run;
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.a.out) + (int64_t)b * p.a.o_sb + (int64_t)i * p.a.o_sl + (int64_t)h * p.a.o_sh;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + 256u + ch * 32, v);
            if (i < L) {
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
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
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

static bool make_map(EncodeTiledFn enc, CUtensorMap* m, const void* base, int L, int H, int B, int64_t sb, int64_t sl, int64_t sh)
{
    cuuint64_t dims[4] = {(cuuint64_t)kD, (cuuint64_t)L, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)sl * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)kBM, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static thread_local float* g_attn_dbg = nullptr;
void set_attn_debug(float* p) { g_attn_dbg = p; }

int launch_omni_attention(const AttnArgs& a, int num_sms, cudaStream_t stream, char* msg, size_t msg_len)
{
    (void)num_sms;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { snprintf(msg, msg_len, "cuTensorMapEncodeTiled entry point not available"); return -6; }
    CUtensorMap tq, tk, tv;
    if (!make_map(enc, &tq, a.q, a.L, a.Hq, a.B, a.q_sb, a.q_sl, a.q_sh) ||
        !make_map(enc, &tk, a.k, a.L, a.Hkv, a.B, a.kv_sb, a.kv_sl, a.kv_sh) ||
        !make_map(enc, &tv, a.v, a.L, a.Hkv, a.B, a.kv_sb, a.kv_sl, a.kv_sh)) {
        snprintf(msg, msg_len, "cuTensorMapEncodeTiled rejected the q/k/v layout");
        return -3;
    }
    AttnParams p;
    p.a = a;
    p.scale_log2 = a.scale * 1.4426950408889634f;
    p.n_qtiles = (a.L + kBM - 1) / kBM;
    p.dbg = g_attn_dbg;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(omni_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSmem);
        if (e != cudaSuccess) { snprintf(msg, msg_len, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    dim3 grid((unsigned)p.n_qtiles, (unsigned)a.Hq, (unsigned)a.B);
    omni_attention_kernel<<<grid, kAttnThreads, kAttnSmem, stream>>>(tq, tk, tv, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(msg, msg_len, "launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // namespace ugk
