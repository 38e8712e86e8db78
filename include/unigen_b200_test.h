/*
 * unigen_b200_test.h -- TEST-ONLY entry points of libunigen_b200.so (symbol prefix ugt_).
 *
 * These are the production entry points of unigen_b200.h with one extra, explicit argument that carries the
 * instrumentation a parity test needs.  They hold no state: nothing a ugt_ call does can change the behaviour of a
 * later ug_ call (round 1 had thread-local ug_debug_set_* switches in the product ABI; they are gone).
 * Product code (ml_unigen_b200/*.py outside the `_debug` keyword of ops.decode_step / ops.omni_attention, bench.py)
 * never calls them.
 */
#ifndef UNIGEN_B200_TEST_H_
#define UNIGEN_B200_TEST_H_

#include "unigen_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ugt_decode_options {
    float* rowsum_out;        /* NULL or float [B,N]: softmax denominators of the step (0 for known rows) */
    int force_exact_bits;     /* fast kernel: 1 = evaluate every element of every row of the race exactly (the e^-16 fallback),
                                 2 = exact evaluation of the candidate lists on every row, 4 = exact confidence for every token
                                 of the mask selection */
    int exact_kernel;         /* != 0: always the exact-arithmetic kernel (decode_step.cu); both kernels return the same ids and masks */
    unsigned* stats;          /* NULL or device uint32[8] counters of the fast kernel's exact paths: [0] rows whose candidates were
                                 evaluated exactly, [1] rows evaluated exhaustively, [2] mask selections with an undecided band,
                                 [3] rows replayed for it, [4..7] reasons of [1] (degenerate row, bound not reached, list A / B overflow) */
} ugt_decode_options;

/* ug_decode_step_bf16 with explicit test options (opts == NULL: identical to ug_decode_step_bf16) */
int ugt_decode_step_bf16(const void* logits, const void* uncond_logits, float guidance_scale,
                         int64_t* ids_inout, int64_t* sampled_out, uint8_t* mask_out,
                         float* sel_prob_out, float* conf_out,
                         const float* q_noise, const float* u_noise,
                         uint64_t philox_seed, uint64_t philox_offset, int64_t batch_offset,
                         int B, int N, int V, int64_t mask_token_id, int mask_len_cap, float temperature,
                         void* workspace, size_t workspace_bytes, void* stream, const ugt_decode_options* opts);

/* ug_omni_attention_rows_bf16 with a dump buffer: float[2*128*128 + 2048] receiving the raw scores S and the bf16
 * probabilities P of the first kv tile of the first query tile of (b=0, h=0), followed by 2048 cycle stamps of CTA (0,0,0)
 * (tools/attn_timeline.py); NULL = identical to ug_omni_attention_rows_bf16 */
int ugt_omni_attention_rows_bf16(const void* q, const void* k, const void* v, void* out,
                                 const int32_t* pad_len, const int32_t* soi_idx,
                                 int B, int L, int Lq, int q_row_offset, int Hq, int Hkv, int D, float scale,
                                 int64_t q_stride_b, int64_t q_stride_l, int64_t q_stride_h,
                                 int64_t kv_stride_b, int64_t kv_stride_l, int64_t kv_stride_h,
                                 int64_t o_stride_b, int64_t o_stride_l, int64_t o_stride_h,
                                 void* stream, float* debug_dump);

#ifdef __cplusplus
}
#endif
#endif /* UNIGEN_B200_TEST_H_ */
