// decode_step.cu -- fused MaskGIT decode step for sm_100a.
//
// Replaces models/unigen.py:407-453 + models/sampling.py:20-26,41-46 of apple/ml-unigen
// (softmax over the 8192-entry codebook slice, torch.multinomial == exponential race,
// selected-probability gather, mask_len, mask_by_random_topk, re-masking) with ONE launch that
// reads every bf16 logit exactly once from HBM.
//
// Structure:
//   * persistent grid, 2 CTAs per SM, 288 threads = 8 consumer warps + 1 producer warp;
//   * the producer warp claims pairs of rows from a global ticket, skips rows whose token is
//     already known (their logits are never read), and streams each unknown row (V bf16 =
//     16 KB) into a 4-deep shared-memory ring with cp.async.bulk + mbarrier complete_tx;
//   * the 256 consumer threads own one row at a time: 32 logits per thread in registers,
//     packed-bf16 max, fp32x2 (FFMA2) deterministic exp, fixed-order sum, then the race;
//   * the race needs argmax_i p_i/q_i over V Exp(1) draws.  With in-kernel Philox noise only a
//     coarse byte per element is generated (16 elements per Philox call); an element can only
//     win if byte <= 256*16*p_i, so on average ~48 of 8192 elements are evaluated exactly
//     (two IEEE divisions + the fine noise bits).  If the best candidate does not clear the
//     bound 1/16 (probability e^-16 per row) the whole row is evaluated exactly;
//   * every finished row is published as ONE 64-bit word (id+1 | selected probability; 0 = pending), so no
//     fence or atomic is needed; the CTA that processed the last row of image b gathers the N words
//     (spinning on the few still in flight), computes the confidences, finds the mask_len-th smallest
//     with a 4-pass radix select in shared memory and writes mask / ids in place -- no second launch,
//     no host synchronisation.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cfloat>
#include <cstdint>

#include "ug_math.cuh"
#include "ug_ptx.cuh"
#include "decode_step.h"

namespace ugk {

constexpr int kConsumers = 256;
constexpr int kThreads = kConsumers + 32;
constexpr int kVec = 16;
constexpr int kRowsPerTicket = 2;
constexpr int kCandMax = 1024;
constexpr float kRaceBound = 1.0f / 16.0f;      // the candidate lists cover every ratio >= max(best survivor, 1/16)

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ unsigned long long best_key(float r, int i)
{
    return ((unsigned long long)__float_as_uint(r) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);   // r >= 0
}
// ------------------------------------------------------------------------------------------------
struct Best {
    float r;      // ratio p/q
    float p;      // p of that element
    int i;        // element index
};
__device__ __forceinline__ void best_merge(Best& a, float r, float p, int i)
{
    if (r > a.r || (r == a.r && i < a.i)) { a.r = r; a.p = p; a.i = i; }
}
// exact evaluation of one element of the race (cold path: kept out of line on purpose)
__device__ __noinline__ void race_eval_exact(float ek, float S, int i, uint64_t rowg, uint64_t seed, uint64_t offset, Best& best)
{
    const float q = ug::race_q((uint32_t)i, rowg, seed, offset);
    const float p = __fdiv_rn(ek, S);
    const float r = __fdiv_rn(p, q);
    best_merge(best, r, p, i);
}

// shared-memory control block (after the ring buffers)
struct Ctrl {
    uint64_t full[8];
    uint64_t empty[8];
    int task[8];
    float red_max[8];
    float red_min[8];
    float red_sum[8];
    unsigned long long best;     // max over candidates of (ratio bits << 32 | ~index)
    int cand_cnt;
    int sel_unknown;
    int sel_bin;
    int sel_k;
    int hist[256];
    int wtot[8];
    uint32_t cand[kCandMax];     // element index
    uint32_t cand_k[kCandMax];   // its 24-bit noise integer (survivors: from the group call)
    uint32_t survmask[512];      // bit i: element i is a survivor of the current row (V <= 16384)
};

// ------------------------------------------------------------------------------------------------
// mask selection for image b: confidence, k-th smallest (radix select), outputs.  256 threads.
// `probs` are read with ld.cg (they may have been produced by other CTAs of this launch).
__device__ void select_phase(const DecodeArgs& a, int b, Ctrl* c, uint32_t* keys, int tid,
                             const int64_t* mask_len_in /* nullable: standalone top-k */)
{
    const int N = a.N;
    const int lane = tid & 31, warp = tid >> 5;
    const uint64_t gb = (uint64_t)(a.batch_offset + b);
    const size_t base = (size_t)b * N;
    if (tid == 0) c->sel_unknown = 0;
    consumer_bar();
    int unk = 0;
    for (int n = tid; n < N; n += kConsumers) {
        float psel;
        if (mask_len_in) {
            psel = a.psel[base + n];
        } else {
            // rows of this image still in flight on other CTAs are published within a row time
            unsigned long long w;
            uint32_t spins = 0;
            while ((w = ld_relaxed_u64(a.slots + base + n)) == 0ull) {
                if (++spins > (1u << 26)) __trap();
            }
            a.slots[base + n] = 0ull;                                  // re-arm the workspace for the next launch
            psel = __uint_as_float((uint32_t)w);
            a.sampled[base + n] = (long long)(w >> 32) - 1;            // models/unigen.py:419-422
            unk += (a.ids[base + n] == a.mask_id);
        }
        float u = a.u_noise ? a.u_noise[base + n] : ug::select_u((uint32_t)n, gb, a.seed, a.offset);
        float conf = ug::confidence(psel, u, a.temperature);
        keys[n] = ug::ordered_key(conf);
        if (a.conf_out) a.conf_out[base + n] = conf;
        if (a.sel_prob_out) a.sel_prob_out[base + n] = psel;
    }
    if (!mask_len_in) {
        unk = __reduce_add_sync(0xffffffffu, unk);
        if (lane == 0 && unk) atomicAdd(&c->sel_unknown, unk);
    }
    consumer_bar();
    int k;
    if (mask_len_in) {
        long long ml = mask_len_in[b];
        k = ml < 0 ? 0 : (ml > N - 1 ? N - 1 : (int)ml);
    } else {
        k = c->sel_unknown - 1;                       // models/unigen.py:439-441
        if (a.cap < k) k = a.cap;
        if (k < 1) k = 1;
    }
    uint32_t prefix = 0, pmask = 0;
#pragma unroll 1
    for (int shift = 24; shift >= 0; shift -= 8) {
        c->hist[tid] = 0;
        consumer_bar();
        for (int n = tid; n < N; n += kConsumers) {
            uint32_t key = keys[n];
            if ((key & pmask) == prefix) atomicAdd(&c->hist[(key >> shift) & 255u], 1);
        }
        consumer_bar();
        int val = c->hist[tid];
        int incl = val;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) c->wtot[warp] = incl;
        consumer_bar();
        int wofs = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) wofs += (w < warp) ? c->wtot[w] : 0;
        incl += wofs;
        int excl = incl - val;
        if (excl <= k && k < incl) { c->sel_bin = tid; c->sel_k = k - excl; }
        consumer_bar();
        prefix |= (uint32_t)c->sel_bin << shift;
        pmask |= 0xffu << shift;
        k = c->sel_k;
    }
    const uint32_t cut = prefix;                        // key of sorted(confidence)[mask_len]
    for (int n = tid; n < N; n += kConsumers) {
        uint8_t masking = keys[n] < cut;                // models/sampling.py:45 (strict <)
        a.mask[base + n] = masking;
        if (!mask_len_in) a.ids[base + n] = masking ? a.mask_id : a.sampled[base + n];   // models/unigen.py:453 (same thread wrote sampled)
    }
    consumer_bar();
}

// ------------------------------------------------------------------------------------------------
template <int kChunks, bool kInjectQ, bool kCfg>
__global__ void __launch_bounds__(kThreads, 2) decode_step_kernel(const DecodeArgs a, const int stages)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int V = kChunks * kConsumers * kVec;
    constexpr uint32_t kRowBytes = V * 2u;
    constexpr uint32_t kStageBytes = kCfg ? 2u * kRowBytes : kRowBytes;
    uint8_t* ring = smem;
    Ctrl* c = reinterpret_cast<Ctrl*>(smem + (size_t)stages * kStageBytes);
    uint32_t* keys = reinterpret_cast<uint32_t*>(c + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int N = a.N;
    const long long total_rows = (long long)a.B * N;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&c->full[s], 1); mbar_init(&c->empty[s], 1); }
        c->cand_cnt = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int w = tid; w < 512; w += kThreads) c->survmask[w] = 0u;
    __syncthreads();

    if (warp == kConsumers / 32) {
        // =========================================================== producer warp
        int stage = 0;
        uint32_t phase = 0;
        const unsigned total_units = (unsigned)((total_rows + kRowsPerTicket - 1) / kRowsPerTicket);
        while (true) {
            unsigned unit = 0;
            if (lane == 0) unit = atomicAdd(a.ticket, 1u);
            unit = __shfl_sync(0xffffffffu, unit, 0);
            if (unit >= total_units) break;
            const long long r = (long long)unit * kRowsPerTicket + lane;
            const bool valid = lane < kRowsPerTicket && r < total_rows;
            long long cur = 0;
            if (valid) cur = a.ids[r];
            const bool unknown = valid && cur == a.mask_id;
            bool completes = false;
            if (valid && !unknown) {                              // models/unigen.py:422, 433: keep the id, confidence = FLT_MAX
                a.slots[r] = ((unsigned long long)(cur + 1) << 32) | (unsigned long long)__float_as_uint(FLT_MAX);
                completes = (r % N) == N - 1;                     // the image's last row needs no consumer: schedule its selection
            }
            unsigned umask = __ballot_sync(0xffffffffu, unknown);
            unsigned cmask = __ballot_sync(0xffffffffu, completes);
            while (umask) {
                const int l = __ffs(umask) - 1;
                umask &= umask - 1;
                const long long row = (long long)unit * kRowsPerTicket + l;
                mbar_wait(&c->empty[stage], phase ^ 1u);
                if (lane == 0) {
                    c->task[stage] = (int)row;
                    uint8_t* dst = ring + (size_t)stage * kStageBytes;
                    mbar_arrive_expect_tx(&c->full[stage], kStageBytes);
                    bulk_g2s(dst, a.logits + (size_t)row * V, kRowBytes, &c->full[stage]);
                    if (kCfg) bulk_g2s(dst + kRowBytes, a.uncond + (size_t)row * V, kRowBytes, &c->full[stage]);
                }
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            while (cmask) {
                const int l = __ffs(cmask) - 1;
                cmask &= cmask - 1;
                const long long row = (long long)unit * kRowsPerTicket + l;
                mbar_wait(&c->empty[stage], phase ^ 1u);
                if (lane == 0) {
                    c->task[stage] = -2 - (int)(row / N);          // SELECT(b)
                    mbar_arrive(&c->full[stage]);
                }
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
        mbar_wait(&c->empty[stage], phase ^ 1u);
        if (lane == 0) {
            c->task[stage] = -1;                                   // sentinel
            mbar_arrive(&c->full[stage]);
            __threadfence();
            if (atomicAdd(a.cta_done, 1u) == gridDim.x - 1) {      // last producer: re-arm the workspace
                *a.ticket = 0u;
                *a.cta_done = 0u;
            }
        }
        return;
    }

    // =============================================================== consumers (256 threads)
    int stage = 0;
    uint32_t phase = 0;
    while (true) {
        mbar_wait(&c->full[stage], phase);
        const int task = c->task[stage];
        if (task < 0) {                                            // sentinel (-1) or SELECT(b) (<= -2)
            consumer_bar();
            if (tid == 0) mbar_arrive(&c->empty[stage]);
            if (task == -1) break;
            select_phase(a, -2 - task, c, keys, tid, nullptr);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
            continue;
        }
        const long long row = task;
        const int b = (int)(row / N);
        const uint64_t rowg = (uint64_t)(a.batch_offset + b) * (uint64_t)N + (uint64_t)(row - (long long)b * N);
        const uint8_t* srow = ring + (size_t)stage * kStageBytes;   // stays valid until this row is published

        // ---- shared -> registers: 16 consecutive logits per chunk, two 16-byte halves in
        //      swapped order for threads 4..7 of each octet so LDS.128 is bank-conflict free
        uint32_t xb[kChunks * 8];
        {
            const int sw = (tid >> 2) & 1;
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) {
                const uint8_t* p = srow + ((size_t)ch * kConsumers + tid) * (kVec * 2);
                uint4 h0 = *reinterpret_cast<const uint4*>(p + 16 * sw);
                uint4 h1 = *reinterpret_cast<const uint4*>(p + 16 * (sw ^ 1));
                uint4 lo = sw ? h1 : h0, hi = sw ? h0 : h1;
                if (kCfg) {
                    uint4 g0 = *reinterpret_cast<const uint4*>(p + kRowBytes + 16 * sw);
                    uint4 g1 = *reinterpret_cast<const uint4*>(p + kRowBytes + 16 * (sw ^ 1));
                    uint4 ul = sw ? g1 : g0, uh = sw ? g0 : g1;
                    uint32_t cw[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                    uint32_t uw[8] = {ul.x, ul.y, ul.z, ul.w, uh.x, uh.y, uh.z, uh.w};
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        uint32_t m0 = ug::cfg_mix_bits(cw[w] & 0xffffu, uw[w] & 0xffffu, a.guidance);
                        uint32_t m1 = ug::cfg_mix_bits(cw[w] >> 16, uw[w] >> 16, a.guidance);
                        xb[ch * 8 + w] = m0 | (m1 << 16);
                    }
                } else {
                    xb[ch * 8 + 0] = lo.x; xb[ch * 8 + 1] = lo.y; xb[ch * 8 + 2] = lo.z; xb[ch * 8 + 3] = lo.w;
                    xb[ch * 8 + 4] = hi.x; xb[ch * 8 + 5] = hi.y; xb[ch * 8 + 6] = hi.z; xb[ch * 8 + 7] = hi.w;
                }
            }
        }
        // ---- row max / min on packed bf16
        {
            __nv_bfloat162 mx = *reinterpret_cast<__nv_bfloat162*>(&xb[0]);
            __nv_bfloat162 mn = mx;
#pragma unroll
            for (int w = 1; w < kChunks * 8; ++w) {
                __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&xb[w]);
                mx = __hmax2(mx, v);
                mn = __hmin2(mn, v);
            }
            float fmx = fmaxf(__low2float(mx), __high2float(mx));
            float fmn = fminf(__low2float(mn), __high2float(mn));
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                fmx = fmaxf(fmx, __shfl_xor_sync(0xffffffffu, fmx, off));
                fmn = fminf(fmn, __shfl_xor_sync(0xffffffffu, fmn, off));
            }
            if (lane == 0) { c->red_max[warp] = fmx; c->red_min[warp] = fmn; }
        }
        consumer_bar();                                            // (1)
        if (tid == 0) { c->cand_cnt = 0; c->best = 0ull; }
        float m = c->red_max[0], mn = c->red_min[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) { m = fmaxf(m, c->red_max[w]); mn = fminf(mn, c->red_min[w]); }
        const bool fast = (m - mn) <= ug::kExpFastSpread;          // block-uniform

        // ---- softmax numerators e_i = exp(x_i - m) and the fixed-order row sum
        float e[kChunks * kVec];
        float tsum;
        if (fast) {
            const float2 negm = make_float2(-m, -m);
            float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int w = 0; w < kChunks * 8; ++w) {
                float2 x = make_float2(__uint_as_float(xb[w] << 16), __uint_as_float(xb[w] & 0xffff0000u));
                float2 ev = ug::expf_fast2(x, negm);
                e[2 * w] = ev.x;
                e[2 * w + 1] = ev.y;
                acc = __fadd2_rn(acc, ev);
            }
            tsum = __fadd_rn(acc.x, acc.y);
        } else {
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int w = 0; w < kChunks * 8; ++w) {
                float x0 = __uint_as_float(xb[w] << 16), x1 = __uint_as_float(xb[w] & 0xffff0000u);
                e[2 * w] = ug::expf_exact(__fadd_rn(x0, -m));
                e[2 * w + 1] = ug::expf_exact(__fadd_rn(x1, -m));
                a0 = __fadd_rn(a0, e[2 * w]);
                a1 = __fadd_rn(a1, e[2 * w + 1]);
            }
            tsum = __fadd_rn(a0, a1);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) tsum = __fadd_rn(tsum, __shfl_xor_sync(0xffffffffu, tsum, off));
        if (lane == 0) c->red_sum[warp] = tsum;
        consumer_bar();                                            // (2)
        const float S = __fadd_rn(__fadd_rn(__fadd_rn(c->red_sum[0], c->red_sum[1]), __fadd_rn(c->red_sum[2], c->red_sum[3])),
                                  __fadd_rn(__fadd_rn(c->red_sum[4], c->red_sum[5]), __fadd_rn(c->red_sum[6], c->red_sum[7])));

        // ---- the race: argmax_i fl(fl(e_i/S)/q_i), lowest index on ties (torch.multinomial)
        Best best{-1.0f, 0.0f, 0x7fffffff};
        bool need_full = false;
        if (kInjectQ) {
            const float* qrow = a.q_noise + (size_t)row * V;
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) {
                const int i0 = (ch * kConsumers + tid) * kVec;
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    float4 q = ld_stream_f4(qrow + i0 + 4 * v4);
                    float qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float p = __fdiv_rn(e[ch * kVec + 4 * v4 + k], S);
                        float r = __fdiv_rn(p, qq[k]);
                        if (r > best.r) { best.r = r; best.p = p; best.i = i0 + 4 * v4 + k; }
                    }
                }
            }
        } else {
            // Candidates (noise stream v2, ug_math.cuh): (A) the "survivors" of every group of 256 elements (noise
            // integer < 2^16, q < 2^-8; ~V/256 per row, one Philox call per group) and (B) every other element whose
            // probability is large enough to reach the best survivor's ratio although its q >= kQFloor.  Both lists are
            // evaluated with the exact operations, so the filter can never change a result.
            if (tid < 32 * ((V + 8191) / 8192)) {
                ug::for_each_survivor((uint32_t)tid, rowg, a.seed, a.offset, [&](uint32_t pos, uint32_t k24) {
                    const uint32_t i = ug::index_of((uint32_t)tid, pos);
                    if (i >= (uint32_t)V) return;                  // V = 4096: the second half of the group does not exist
                    const int s = atomicAdd(&c->cand_cnt, 1);
                    if (s < kCandMax) { c->cand[s] = i; c->cand_k[s] = k24; }
                    atomicOr(&c->survmask[i >> 5], 1u << (i & 31u));
                });
            }
            consumer_bar();                                        // (3)
            const int cntA = c->cand_cnt < kCandMax ? c->cand_cnt : kCandMax;
            auto eval_cand = [&](int s) {
                const int i = (int)c->cand[s];
                // the logit is still in the ring slot; e_i is recomputed with the same IEEE operations
                uint32_t xbits = *reinterpret_cast<const uint16_t*>(srow + 2 * i);
                if (kCfg) xbits = ug::cfg_mix_bits(xbits, *reinterpret_cast<const uint16_t*>(srow + kRowBytes + 2 * i), a.guidance);
                const float d = __fadd_rn(__uint_as_float(xbits << 16), -m);
                const float ei = fast ? ug::expf_fast(d) : ug::expf_exact(d);
                const float q = ug::q_from_k24(c->cand_k[s]);
                const float p = __fdiv_rn(ei, S);
                const float r = __fdiv_rn(p, q);
                best_merge(best, r, p, i);
            };
            for (int s = tid; s < cntA; s += kConsumers) eval_cand(s);
            if (best.r >= 0.0f) atomicMax(&c->best, best_key(best.r, best.i));
            consumer_bar();                                        // (3b)
            {
                const float bestA = __uint_as_float((uint32_t)(c->best >> 32));
                const float R = fmaxf(bestA * (1.0f - 0x1p-12f), kRaceBound);
                const float thr = R * ug::kQFloor * S * (1.0f - 0x1p-10f);   // e_i < thr  =>  p_i / q_i < R for q_i >= -log(1-2^-8)
#pragma unroll
                for (int ch = 0; ch < kChunks; ++ch) {
                    const int i0 = (ch * kConsumers + tid) * kVec;
#pragma unroll
                    for (int k = 0; k < kVec; ++k) {
                        if (e[ch * kVec + k] >= thr) {
                            const uint32_t i = (uint32_t)(i0 + k);
                            if (!((c->survmask[i >> 5] >> (i & 31u)) & 1u)) {
                                const int s = atomicAdd(&c->cand_cnt, 1);
                                if (s < kCandMax) c->cand[s] = i;
                            }
                        }
                    }
                }
            }
            consumer_bar();                                        // (3c)
            const int cnt = c->cand_cnt;
            need_full = cnt > kCandMax || a.force_full_race != 0;
            const int lim = cnt < kCandMax ? cnt : kCandMax;
            for (int s = cntA + tid; s < lim; s += kConsumers) {
                c->cand_k[s] = ug::elem_k24(c->cand[s], rowg, a.seed, a.offset);
                eval_cand(s);
            }
            for (int s = tid; s < cntA; s += kConsumers) c->survmask[c->cand[s] >> 5] = 0u;   // re-arm for the next row
        }
        // ---- block argmax: one 64-bit shared atomicMax per thread that holds a candidate
        unsigned long long mykey = best.r >= 0.0f ? best_key(best.r, best.i) : 0ull;
        if (mykey) atomicMax(&c->best, mykey);
        consumer_bar();                                            // (4)
        unsigned long long win = c->best;
        if (!kInjectQ && (need_full || __uint_as_float((uint32_t)(win >> 32)) < kRaceBound)) {
            // rare (e^-16 per row): no candidate clears the bound -> evaluate every element of the row exactly
            consumer_bar();
            if (tid == 0) c->best = 0ull;
            best = Best{-1.0f, 0.0f, 0x7fffffff};
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) {
                const int i0 = (ch * kConsumers + tid) * kVec;
#pragma unroll 1
                for (int k = 0; k < kVec; ++k)
                    race_eval_exact(e[ch * kVec + k], S, i0 + k, rowg, a.seed, a.offset, best);
            }
            consumer_bar();
            mykey = best_key(best.r, best.i);
            atomicMax(&c->best, mykey);
            consumer_bar();
            win = c->best;
        }
        // ---- publish: the thread that holds the winner writes one word (id+1 | selected probability); the ring
        //      slot is free now.  The CTA that processed the last row of an image runs that image's mask selection.
        if (mykey == win && mykey) {
            a.slots[row] = ((unsigned long long)(best.i + 1) << 32) | (unsigned long long)__float_as_uint(best.p);
            if (a.rowsum_out) a.rowsum_out[row] = S;
        }
        if (tid == 0) mbar_arrive(&c->empty[stage]);
        if (row - (long long)b * N == N - 1) select_phase(a, b, c, keys, tid, nullptr);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
}

// standalone mask_by_random_topk (models/sampling.py:41-46): one CTA of 256 threads per image
__global__ void __launch_bounds__(kConsumers) topk_kernel(const DecodeArgs a, const int64_t* mask_len)
{
    extern __shared__ __align__(128) uint8_t smem[];
    Ctrl* c = reinterpret_cast<Ctrl*>(smem);
    uint32_t* keys = reinterpret_cast<uint32_t*>(c + 1);
    select_phase(a, blockIdx.x, c, keys, threadIdx.x, mask_len);
}

__global__ void export_noise_kernel(uint64_t seed, uint64_t offset, int64_t batch_offset, int B, int N, int V,
                                    float* q_out, float* u_out)
{
    const size_t total = (size_t)B * N * V;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(idx % V);
        const size_t rn = idx / V;
        const uint64_t gb = (uint64_t)(batch_offset + (int64_t)(rn / N));
        const uint64_t rowg = gb * (uint64_t)N + (uint64_t)(rn % N);
        if (q_out) q_out[idx] = ug::race_q(i, rowg, seed, offset);
        if (u_out && i == 0) u_out[rn] = ug::select_u((uint32_t)(rn % N), gb, seed, offset);
    }
}

// ------------------------------------------------------------------------------------------------
static size_t ctrl_bytes(int N) { return sizeof(Ctrl) + sizeof(uint32_t) * (size_t)N; }

template <int kChunks, bool kInjectQ, bool kCfg>
static cudaError_t launch_one(const DecodeArgs& a, int num_sms, cudaStream_t stream)
{
    constexpr int V = kChunks * kConsumers * kVec;
    const size_t stage_bytes = (kCfg ? 2u : 1u) * (size_t)V * 2u;
    int stages = (int)((100u * 1024u) / stage_bytes);              // ~2 CTAs/SM within 227 KB
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    const size_t smem = (size_t)stages * stage_bytes + ctrl_bytes(a.N);
    auto kern = decode_step_kernel<kChunks, kInjectQ, kCfg>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    int per_sm = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    if (err != cudaSuccess) return err;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    const long long units = ((long long)a.B * a.N + kRowsPerTicket - 1) / kRowsPerTicket;
    long long grid = (long long)num_sms * per_sm;
    if (grid > units) grid = units;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(a, stages);
    return cudaGetLastError();
}

template <int kChunks>
static cudaError_t launch_chunks(const DecodeArgs& a, int num_sms, cudaStream_t stream)
{
    const bool inj = a.q_noise != nullptr, cfg = a.uncond != nullptr;
    if (inj && cfg) return launch_one<kChunks, true, true>(a, num_sms, stream);
    if (inj) return launch_one<kChunks, true, false>(a, num_sms, stream);
    if (cfg) return launch_one<kChunks, false, true>(a, num_sms, stream);
    return launch_one<kChunks, false, false>(a, num_sms, stream);
}

cudaError_t launch_decode_step(const DecodeArgs& a, int num_sms, cudaStream_t stream)
{
    switch (a.V) {
        case 4096: return launch_chunks<1>(a, num_sms, stream);
        case 8192: return launch_chunks<2>(a, num_sms, stream);
        case 16384: return launch_chunks<4>(a, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_topk(const DecodeArgs& a, const int64_t* mask_len, cudaStream_t stream)
{
    const size_t smem = ctrl_bytes(a.N);
    cudaError_t err = cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    topk_kernel<<<a.B, kConsumers, smem, stream>>>(a, mask_len);
    return cudaGetLastError();
}

cudaError_t launch_export_noise(uint64_t seed, uint64_t offset, int64_t batch_offset, int B, int N, int V,
                                float* q_out, float* u_out, cudaStream_t stream)
{
    export_noise_kernel<<<1184, 256, 0, stream>>>(seed, offset, batch_offset, B, N, V, q_out, u_out);
    return cudaGetLastError();
}

// classifier-free guidance as a pass of its own, in place on the conditional logits: x = bf16(bf16(s * bf16(c - u)) + u), the
// op-by-op bf16 arithmetic of models/unigen.py:410 under autocast (ug::cfg_mix_bits, the expression the exact kernel applies
// inline).  The mixed logits are bf16 values again, so the certified fast kernel can stream them: CFG costs one extra
// elementwise pass instead of the exact kernel (the shipped evaluation configs all run with guidance_scale > 1).
__global__ void __launch_bounds__(256) cfg_mix_kernel(uint4* __restrict__ cond, const uint4* __restrict__ uncond, float scale, int64_t n_vec)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 c = cond[i], u = uncond[i];
        const uint32_t cw[4] = {c.x, c.y, c.z, c.w}, uw[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int w = 0; w < 4; ++w)
            o[w] = ug::cfg_mix_bits(cw[w] & 0xffffu, uw[w] & 0xffffu, scale) | (ug::cfg_mix_bits(cw[w] >> 16, uw[w] >> 16, scale) << 16);
        cond[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

cudaError_t launch_cfg_mix(void* cond_inout, const void* uncond, float scale, int64_t n, cudaStream_t stream)
{
    if (n < 8 || n % 8) return cudaErrorInvalidValue;
    int64_t blocks = (n / 8 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cfg_mix_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<uint4*>(cond_inout), static_cast<const uint4*>(uncond), scale, n / 8);
    return cudaGetLastError();
}

}  // namespace ugk
