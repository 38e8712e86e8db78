// decode_step.h -- internal launch interface of the fused decode-step kernel.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ugk {

struct DecodeArgs {
    const uint16_t* logits;      // bf16 [B,N,V]
    const uint16_t* uncond;      // bf16 [B,N,V] or null
    float guidance;
    long long* ids;              // int64 [B,N] in/out
    long long* sampled;          // int64 [B,N]
    uint8_t* mask;               // [B,N]
    float* sel_prob_out;         // nullable
    float* conf_out;             // nullable
    float* rowsum_out;           // nullable (debug: softmax denominators)
    const float* q_noise;        // nullable
    const float* u_noise;        // nullable
    uint64_t seed, offset;
    int64_t batch_offset;
    int B, N, V;
    long long mask_id;
    int cap;
    float temperature;
    int force_full_race;         // test hook, bits: 1 = evaluate every element of every row exactly; 2 = exact evaluation of
                                 // the candidate lists for every row; 4 = exact confidence for every token (fast kernel)
    // workspace
    float* psel;                 // standalone top-k only: probs [B*N]
    unsigned long long* slots;   // [B*N] one word per row: (id + 1) << 32 | bits(selected prob); 0 = not published yet
    unsigned* ticket;            // row-pair ticket
    unsigned* cta_done;          // producers finished
    unsigned* stats;             // nullable test hook: [0] exact candidate evaluations, [1] exhaustive rows,
                                 // [2] selections with an undecided band, [3] rows replayed exactly for the band, [4..7] reasons of [1]: degenerate row, bound not reached, list A / list B overflow
};

cudaError_t launch_decode_step(const DecodeArgs& a, int num_sms, cudaStream_t stream);   // exact arithmetic, one CTA per row
bool decode_fast_supported(const DecodeArgs& a);
cudaError_t launch_decode_fast(const DecodeArgs& a, int num_sms, cudaStream_t stream);   // certified fast path, one warp per row
cudaError_t launch_topk(const DecodeArgs& a, const int64_t* mask_len, cudaStream_t stream);
cudaError_t launch_export_noise(uint64_t seed, uint64_t offset, int64_t batch_offset, int B, int N, int V,
                                float* q_out, float* u_out, cudaStream_t stream);

// cond[i] = bf16(bf16(scale * bf16(cond[i] - uncond[i])) + uncond[i]) in place, n bf16 elements (n % 8 == 0)
cudaError_t launch_cfg_mix(void* cond_inout, const void* uncond, float scale, int64_t n, cudaStream_t stream);

}  // namespace ugk
