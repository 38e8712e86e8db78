/*
 * unigen_b200.h -- C ABI of libunigen_b200.so: the B200 (sm_100a) implementation of the
 * text-to-image parallel-decoding hot path of apple/ml-unigen.
 *
 * Every entry point takes plain device pointers and sizes; the caller (PyTorch, or anything
 * else that owns CUDA memory) allocates all buffers including the workspace.  The library
 * never allocates, frees, retains a pointer past the call, synchronises with the host,
 * throws, or exits.  All work is enqueued on `stream` (a cudaStream_t passed as void*), so a
 * T-step generation loop can be captured in a CUDA graph.
 *
 * There is no global or thread-local switch that alters the behaviour of a call: everything a call does is determined by
 * its arguments (test instrumentation lives in unigen_b200_test.h behind the ugt_ prefix, passed explicitly per call).
 *
 * Return value: 0 on success; negative UG_ERR_* for rejected arguments; positive = the
 * cudaError_t reported by the launch.  ug_last_error() returns a thread-local message.
 *
 * Reference interfaces replaced (paths relative to the apple/ml-unigen tree):
 *   ug_decode_step_bf16      models/unigen.py:407-453 (the body of the t2i_generate loop after
 *                            the forward pass) + models/sampling.py:20-26,41-46
 *   ug_mask_by_random_topk   models/sampling.py:41-46  mask_by_random_topk(mask_len, probs, temperature, generator)
 *   ug_omni_attention_bf16   the attention call of the Qwen2 backbone as driven by
 *                            models/unigen.py:260-266 with the 4-D mask built by
 *                            training/prompting_utils.py:975-1020 (create_attention_mask_predict_next,
 *                            rm_pad_in_image=True); plug-in point = transformers attention-function
 *                            interface fn(module, q[B,Hq,L,D], k[B,Hkv,L,D], v, attention_mask, scaling=...)
 *   ug_export_noise          (test hook) the exponential / uniform tensors that
 *                            torch.multinomial (unigen.py:419) and gumbel_noise (sampling.py:24-26)
 *                            would have drawn, as generated inside the kernel
 */
#ifndef UNIGEN_B200_H_
#define UNIGEN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UG_ABI_VERSION 2

enum {
    UG_OK = 0,
    UG_ERR_NULL = -1,        /* a required pointer is NULL */
    UG_ERR_SHAPE = -2,       /* unsupported B/N/V/L/heads/head_dim */
    UG_ERR_ALIGN = -3,       /* pointer or stride not aligned as documented */
    UG_ERR_WORKSPACE = -4,   /* workspace too small */
    UG_ERR_ARG = -5,         /* other invalid argument */
    UG_ERR_DEVICE = -6       /* not an sm_100 device / driver entry point missing */
};

int ug_abi_version(void);
const char* ug_last_error(void);

/* ---- decode step -----------------------------------------------------------------------------
 * One MaskGIT step for B images of N tokens over a V-entry codebook.
 *   logits          bf16 [B,N,V], dense, 16-byte aligned          (image rows x codebook columns)
 *   uncond_logits   NULL, or bf16 [B,N,V]: logits = bf16(bf16(scale*bf16(cond-uncond))+uncond)
 *   ids_inout       int64 [B,N]: current ids (mask_token_id = still unknown); rewritten in place
 *                   to  masking ? mask_token_id : sampled
 *   sampled_out     int64 [B,N]: sampled ids with already-known ids kept (the value t2i_generate returns)
 *   mask_out        uint8 [B,N]: 1 where the token is re-masked for the next step
 *   sel_prob_out    NULL or float [B,N]: selected_probs (FLT_MAX for known tokens)
 *   conf_out        NULL or float [B,N]: confidence used by the mask selection
 *   q_noise         NULL -> in-kernel Philox; else float [B,N,V] Exp(1) noise (the tensor
 *                   torch.multinomial would draw)
 *   u_noise         NULL -> in-kernel Philox; else float [B,N] U[0,1) noise (gumbel_noise's)
 *   philox_seed/offset  key and per-step offset of the in-kernel streams; batch_offset is the
 *                   global index of image 0 of this call, so sharding a batch across GPUs
 *                   does not change any image
 *   mask_len_cap    floor(N * mask_ratio) computed by the host in float32 (may be -1)
 *   temperature     annealed temperature of this step (already compounded by the host)
 *   workspace       ug_decode_step_workspace_bytes(B,N,V) bytes, zero-filled ONCE before the
 *                   first call (the kernel leaves it zeroed); one workspace per stream
 * Supported: V in {4096, 8192, 16384}; 2 <= N <= 4096; 1 <= B; B*N < 2^31.
 */
size_t ug_decode_step_workspace_bytes(int B, int N, int V);

int ug_decode_step_bf16(const void* logits, const void* uncond_logits, float guidance_scale,
                        int64_t* ids_inout, int64_t* sampled_out, uint8_t* mask_out,
                        float* sel_prob_out, float* conf_out,
                        const float* q_noise, const float* u_noise,
                        uint64_t philox_seed, uint64_t philox_offset, int64_t batch_offset,
                        int B, int N, int V, int64_t mask_token_id, int mask_len_cap, float temperature,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- mask_by_random_topk ---------------------------------------------------------------------
 *   probs float [B,N]; mask_len int64 [B] (already clamped by the caller, 0 <= mask_len < N);
 *   u_noise NULL -> Philox (seed, offset, batch_offset) else float [B,N]; mask_out uint8 [B,N].
 *   masking = confidence < sort(confidence)[mask_len],
 *   confidence = log(clamp(probs,1e-20)) + temperature * gumbel(u).   2 <= N <= 4096. */
int ug_mask_by_random_topk(const float* probs, const int64_t* mask_len, const float* u_noise,
                           uint64_t philox_seed, uint64_t philox_offset, int64_t batch_offset,
                           float temperature, uint8_t* mask_out, float* conf_out,
                           int B, int N, void* stream);

/* ---- noise export (test hook) ---------------------------------------------------------------- */
int ug_export_noise(uint64_t philox_seed, uint64_t philox_offset, int64_t batch_offset,
                    int B, int N, int V, float* q_out /* [B,N,V] or NULL */, float* u_out /* [B,N] or NULL */,
                    void* stream);

/* Classifier-free guidance as a pass of its own, in place on the conditional logits (models/unigen.py:410 under autocast, op by op
 * in bf16): cond[i] = bf16(bf16(scale * bf16(cond[i] - uncond[i])) + uncond[i]) for n bf16 elements (n % 8 == 0) -- exactly the
 * values ug_decode_step_bf16 forms inline when it is given uncond_logits.  The result is bf16 again, so a following
 * ug_decode_step_bf16(logits = cond, uncond_logits = NULL) runs the certified fast kernel and returns the same ids and masks. */
int ug_cfg_mix_bf16(void* cond_logits_inout, const void* uncond_logits, float guidance_scale, int64_t n, void* stream);

/* ---- omni attention --------------------------------------------------------------------------
 * out = softmax(q k^T * scale + M) v per head, with the omni mask M given in closed form by two
 * integers per sequence:
 *     row i <  pad_len[b]             : columns j <= i                (pad rows; outputs unused)
 *     pad_len[b] <= i < soi_idx[b]    : columns pad_len[b] <= j <= i  (text: causal, never sees pads)
 *     i >= soi_idx[b]                 : columns j >= pad_len[b]       (<soi>, image block, <eoi>: bidirectional)
 *   q   bf16 [B,L,Hq,D]   (element strides q_stride_b, q_stride_l, q_stride_h; D contiguous)
 *   k,v bf16 [B,L,Hkv,D]  (kv_stride_b, kv_stride_l, kv_stride_h)
 *   out bf16 [B,L,Hq,D]   (o_stride_b, o_stride_l, o_stride_h)
 *   pad_len, soi_idx int32 [B] on the device
 * Supported: D = 128, Hq % Hkv == 0, L <= 4096, all strides multiples of 8 elements, 16-byte aligned bases.
 */
int ug_omni_attention_bf16(const void* q, const void* k, const void* v, void* out,
                           const int32_t* pad_len, const int32_t* soi_idx,
                           int B, int L, int Hq, int Hkv, int D, float scale,
                           int64_t q_stride_b, int64_t q_stride_l, int64_t q_stride_h,
                           int64_t kv_stride_b, int64_t kv_stride_l, int64_t kv_stride_h,
                           int64_t o_stride_b, int64_t o_stride_l, int64_t o_stride_h,
                           void* stream);

/* The same for a window of query rows: q and out are [B, Lq, Hq, D] and hold the sequence rows
 * [q_row_offset, q_row_offset + Lq); k, v still span all L rows.  With q_row_offset = soi_idx this is the
 * per-step pass of the text-prefix K/V reuse: rows < soi_idx never see the image block,
 * so their K/V are computed once per generation and only the N+2 image-block rows are recomputed each step. */
int ug_omni_attention_rows_bf16(const void* q, const void* k, const void* v, void* out,
                                const int32_t* pad_len, const int32_t* soi_idx,
                                int B, int L, int Lq, int q_row_offset, int Hq, int Hkv, int D, float scale,
                                int64_t q_stride_b, int64_t q_stride_l, int64_t q_stride_h,
                                int64_t kv_stride_b, int64_t kv_stride_l, int64_t kv_stride_h,
                                int64_t o_stride_b, int64_t o_stride_l, int64_t o_stride_h,
                                void* stream);

/* ---- row-wise kernels of the Qwen2 decoder layer (transformers models/qwen2/modeling_qwen2.py) ------------
 * These four replace the ATen elementwise chains of Qwen2RMSNorm (+ residual add), apply_rotary_pos_emb, Qwen2MLP's
 * act_fn(gate)*up and embed_tokens (models/unigen.py:370, 452).
 *   ug_add_rmsnorm : x fp32 [M,H] (+= delta bf16 [M,H] if non-NULL, written back); out bf16 [M,H] =
 *                    weight * x * rsqrt(mean(x^2) + eps).  H = 256 k, k in 1..8, 12, 16, ..., 32.  (Since round 2 the layer loop
 *                    no longer launches it: the residual add and the norm are GEMM epilogues; it remains as the unfused
 *                    yardstick of those epilogues and for callers that need the normalised activations themselves.)
 *   ug_rope_qk     : in-place rotate_half RoPE on the q and k heads of qkv bf16 [M,(Hq+2Hkv)*D];
 *                    position = m % L; cos/sin fp32 [L, D/2].  With kv_cache (bf16 [B, cache_L, 2, Hkv, D]) the rotated
 *                    k and the v of row m are also written to cache row cache_row_offset + m % L of batch m / L.
 *   ug_swiglu      : act bf16 [M,I] = silu(g) * u, gate_up bf16 [M,2I] = [g | u].
 *   ug_embed_rows  : x fp32 [M,H] = table bf16 [vocab,H][tokens[m]] (ids outside [0, vocab) are clamped for memory safety; the
 *                    host validates them).  With xb_bf16 / ss (both or neither, H % 256 == 0) the same launch also writes
 *                    xb = bf16(x) and ss fp32 [M, H/128] = per-128-column sums of x^2: the operands of the norm-folded GEMMs below.
 */
int ug_add_rmsnorm(float* x, const void* delta_bf16, const void* weight_bf16, void* out_bf16,
                   int64_t M, int H, float eps, void* stream);
int ug_rope_qk(void* qkv_bf16, const float* cos, const float* sin, int64_t M, int L, int Hq, int Hkv, int D,
               void* kv_cache_bf16 /* nullable */, int cache_L, int cache_row_offset, void* stream);
int ug_swiglu(const void* gate_up_bf16, void* act_bf16, int64_t M, int I, void* stream);
int ug_embed_rows(float* x, void* xb_bf16 /* nullable */, float* ss /* nullable */, const void* table_bf16, const int64_t* tokens,
                  int64_t M, int H, int64_t vocab, void* stream);

/* ---- linear layers on tcgen05 with fused epilogues -------------------------------
 * Replaces torch.nn.functional.linear + the elementwise op that follows it in transformers' Qwen2MLP /
 * Qwen2Attention (models/qwen2/modeling_qwen2.py: `down_proj(act_fn(gate_proj(x)) * up_proj(x))`, q/k/v_proj with
 * bias), reached from apple/ml-unigen models/unigen.py:260-266.
 *   a   bf16 [M,K] (row stride lda elements), w bf16 [N,K] (row stride ldw; nn.Linear weight layout), bias bf16 [N] or NULL
 *   mode UG_GEMM_PLAIN  : out bf16 [M,N]   = a w^T + bias
 *   mode UG_GEMM_SWIGLU : out bf16 [M,N/2] = silu(g) * u, with g = a Wg^T and u = a Wu^T each rounded to bf16 first (what
 *                         the unfused autocast path holds) and w = Wg / Wu interleaved in blocks of 128 rows:
 *                         w[256 j + i] = Wg[128 j + i], w[256 j + 128 + i] = Wu[128 j + i], i < 128 (`pack_gate_up`
 *                         in ml_unigen_b200.ops does this once at load time); bias must be NULL
 * Supported: N % 256 == 0 (plain mode: N % 128 == 0), K % 64 == 0, strides multiples of 8 elements, 16-byte aligned bases; fp32
 * accumulation.
 *
 * RMSNorm folding (row_ss != NULL; also ug_qkv_rope_bf16).  Qwen2RMSNorm (input_layernorm / post_attention_layernorm / norm)
 * commutes with the linear layer that follows it: (g * x / rms(x)) W^T = (1 / rms(x)) * (x (W diag g)^T).  The caller folds the
 * gain g into the columns of w once at load time, passes a = bf16(x) and row_ss = fp32 [M, ss_parts] partial sums of squares of
 * the rows of the fp32 residual stream x (as ug_gemm_residual_bf16 / ug_embed_rows leave them); the epilogue multiplies row m of
 * the fp32 accumulator by rsqrt(sum_p row_ss[m,p] / K + rms_eps) before bias / rounding (partial sums added in index order).
 * The separate add + RMSNorm pass over HBM (ug_add_rmsnorm) disappears from the layer.
 */
#define UG_GEMM_PLAIN 0
#define UG_GEMM_SWIGLU 1
int ug_gemm_bf16(const void* a, const void* w, const void* bias, void* out, int M, int N, int K,
                 int64_t lda, int64_t ldw, int64_t ldo, int mode,
                 const float* row_ss /* nullable */, int ss_parts, float rms_eps, void* stream);

/* o_proj / down_proj of a Qwen2DecoderLayer together with the residual add that follows them and the statistics of the next
 * RMSNorm, in the epilogue of the same launch (reached from apple/ml-unigen models/unigen.py:260-266):
 *     x_inout fp32 [M,N] += bf16(a w^T)     (the linear output rounded to bf16 first: what the unfused autocast path adds)
 *     xb_out  bf16 [M,N]  = bf16(x_inout)   (the A operand of the next norm-folded GEMM)
 *     ss_out  fp32 [M,N/128]: ss[m,j] = sum of x_inout[m, 128 j .. 128 j + 127]^2 after the update
 * x_inout, xb_out, ss_out have contiguous rows.  Shape limits of ug_gemm_bf16.  Replaces F.linear + ug_add_rmsnorm. */
int ug_gemm_residual_bf16(const void* a, const void* w, float* x_inout, void* xb_out, float* ss_out, int M, int N, int K,
                          int64_t lda, int64_t ldw, void* stream);

/* Fused q/k/v projection of Qwen2Attention.forward: qkv bf16 [M,(Hq+2Hkv)*128] = a [w_q|w_k|w_v]^T + bias, then
 * apply_rotary_pos_emb (rotate_half convention) on the q and k heads in the epilogue of the same launch, and -- with a
 * K/V cache -- the rotated k row and the v row of sequence row m written to kv_cache[m / L, cache_row_offset + m % L]
 * (bf16 [B, cache_L, 2, Hkv, 128]).  Replaces ug_gemm_bf16 (plain) + ug_rope_qk; same rounding points: the projection is
 * rounded to bf16 before the rotation (what the unfused path stores), the rotation is fp32, the result bf16.
 * Row m is position m % L; cos/sin fp32 [L, 64].  head_dim must be 128. */
int ug_qkv_rope_bf16(const void* a, const void* w_qkv, const void* bias_qkv, void* qkv_out, int M, int K,
                     int64_t lda, int64_t ldw, int64_t ldo, const float* cos, const float* sin, int L, int Hq, int Hkv,
                     void* kv_cache_bf16 /* nullable */, int cache_L, int cache_row_offset,
                     const float* row_ss /* nullable: RMSNorm folding, see ug_gemm_bf16 */, int ss_parts, float rms_eps, void* stream);

/* ---- MAGViT-v2 pixel decoder, memory-bound kernels -----------------------------
 * MAGVITv2.decode_code (models/multimodal_encoder/magvitv2.py:438-442) = LFQuantizer.get_codebook_entry (:217-230) +
 * VQGANDecoder.forward (:371-408).  Activations are NHWC bf16; the convolutions between these kernels are cuDNN.
 *   ug_lfq_unpack_bf16      : z bf16 [n, dim] = +-1, channel c = bit (dim-1-c) of ids[n] (the `embedding` buffer of the
 *                             reference, never materialised)
 *   ug_groupnorm_nhwc_bf16  : y = GroupNorm(32 groups, eps)(x + in_bias) * gamma + beta, then y * sigmoid(y) when swish != 0
 *                             (common_modules.py:19-27); in_bias (nullable, bf16 [C]) = the bias of the convolution that produced
 *                             x, added here instead of in a pass of its own; x, y bf16 [B, HW, C] (may alias), C % 32 == 0,
 *                             C % 8 == 0, 256 % (C/8) == 0; workspace fp32 of ug_groupnorm_workspace_floats(B, HW, C) elements;
 *                             stats_ready != 0: the partial sums of x are already in the workspace (ug_bias_add_stats_nhwc_bf16)
 *   ug_bias_add_nhwc_bf16   : out = a + bias_a (+ r (+ bias_r)), bf16 [rows, C]: conv2 bias + shortcut (+ nin_shortcut bias) of a
 *                             ResnetBlock (common_modules.py:343-360) in one pass; bias_a, r, bias_r nullable; out may alias a or r
 *   ug_upsample2x_nhwc_bf16 : F.interpolate(scale_factor=2, mode="nearest") (common_modules.py:40), [B,H,W,C] -> [B,2H,2W,C]
 */
int ug_lfq_unpack_bf16(const int64_t* ids, void* z_bf16, int64_t n, int dim, void* stream);

/* Round 2: the convolutions and the attention block of the decoder on the library's own kernels (no cuDNN / SDPA call remains).
 *   ug_lfq_postquant_bf16 : LFQ lookup + `post_quant_conv` (1x1, dim -> dim; magvitv2.py:217-230, 438-442) in one pass: the codebook
 *                           entry is the +-1 bit vector of the id, so z[n, c] = bias[c] + sum_k w[c,k] * (+-1) (w fp32 [dim,dim],
 *                           bias fp32 [dim] or NULL); rows of `cpad` >= dim channels, zero beyond dim (cpad = 64 feeds conv_in)
 *   ug_conv_nhwc_bf16     : k x k convolution, stride 1, k = 3 with zero padding 1 (every `Conv2d(.., 3, 1, 1)` of VQGANDecoder,
 *                           common_modules.py ResnetBlock / Upsample) or k = 1 (nin_shortcut, the q / k / v / proj_out of AttnBlock),
 *                           as an IMPLICIT GEMM on the tcgen05 kernel of ug_gemm_bf16: x NHWC bf16 [B,H,W,Cin], y NHWC bf16
 *                           [B,H,W,Cout]; w_packed bf16 [ceil128(Cout), k*k*ceil64(Cin)]: row o, column (ky*k + kx)*ceil64(Cin) + c
 *                           = Conv2d.weight[o, c, ky, kx], zero in the padding (`pack_conv` in ml_unigen_b200.magvit); bias bf16
 *                           [ceil128(Cout)] or NULL.  A rows = the pixels of [4 x 32] boxes fetched per (tap, 64-channel block) by
 *                           4-D TMA with the tap's shift: image borders, partial boxes and channel padding are TMA out-of-bounds
 *                           fill / clipping, so any H, W work; Cin % 8 == 0, Cout % 8 == 0.
 *   ug_softmax_rows_bf16  : y = softmax(scale * x) per row of a bf16 matrix (AttnBlock, common_modules.py:171-215)
 */
int ug_lfq_postquant_bf16(const int64_t* ids, const float* w, const float* bias, void* z_bf16, int64_t n, int dim, int cpad, void* stream);
int ug_conv_nhwc_bf16(const void* x_bf16, const void* w_packed_bf16, const void* bias_bf16, void* y_bf16,
                      int B, int H, int W, int Cin, int Cout, int ksize, void* stream);
/* cols_valid <= cols: the row softmax is taken over the first cols_valid columns, the rest of the output row is zero (padded keys) */
int ug_softmax_rows_bf16(const void* x_bf16, void* y_bf16, int64_t rows, int cols, int cols_valid, float scale, void* stream);
size_t ug_groupnorm_workspace_floats(int B, int64_t HW, int C);
int ug_groupnorm_nhwc_bf16(const void* x_bf16, const void* in_bias_bf16, const void* gamma_bf16, const void* beta_bf16, void* y_bf16,
                           float* workspace, int B, int64_t HW, int C, float eps, int swish, int stats_ready, void* stream);
/* ug_bias_add_nhwc_bf16 on [B, HW, C] that also leaves the GroupNorm partial sums of `out` in `workspace`; the next
 * ug_groupnorm_nhwc_bf16 on `out` with the same workspace may pass stats_ready = 1 and skip its statistics pass (the partial sums
 * are bit-identical to the ones that pass would compute).  Shape limits of ug_groupnorm_nhwc_bf16. */
int ug_bias_add_stats_nhwc_bf16(const void* a_bf16, const void* bias_a_bf16, const void* r_bf16, const void* bias_r_bf16, void* out_bf16,
                                float* workspace, int B, int64_t HW, int C, void* stream);
int ug_bias_add_nhwc_bf16(const void* a_bf16, const void* bias_a_bf16, const void* r_bf16, const void* bias_r_bf16, void* out_bf16,
                          int64_t rows, int C, void* stream);
int ug_upsample2x_nhwc_bf16(const void* x_bf16, void* y_bf16, int B, int H, int W, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIGEN_B200_H_ */
