// ug_capi.cu -- extern "C" boundary of libunigen_b200.so (see include/unigen_b200.h).
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "unigen_b200.h"
#include "unigen_b200_test.h"
#include "decode_step.h"
#include "omni_attention.h"
#include "backbone_ops.h"
#include "gemm_tc.h"
#include "magvit_ops.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what)
{
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// per-device SM count, cached (cudaGetDevice/cudaDeviceGetAttribute do not synchronise)
int sm_count(int* cc_major)
{
    static thread_local int cached_dev = -1, cached_sms = 0, cached_major = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&cached_major, cudaDevAttrComputeCapabilityMajor, dev);
        cached_dev = dev;
    }
    if (cc_major) *cc_major = cached_major;
    return cached_sms;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ug_decode_step_bf16 and its test-only twin ugt_decode_step_bf16 (explicit, stateless options; NULL = production behaviour)
int decode_step_impl(const void* logits, const void* uncond_logits, float guidance_scale,
                     int64_t* ids_inout, int64_t* sampled_out, uint8_t* mask_out,
                     float* sel_prob_out, float* conf_out,
                     const float* q_noise, const float* u_noise,
                     uint64_t philox_seed, uint64_t philox_offset, int64_t batch_offset,
                     int B, int N, int V, int64_t mask_token_id, int mask_len_cap, float temperature,
                     void* workspace, size_t workspace_bytes, void* stream, const ugt_decode_options* opts);

int attention_impl(const void* q, const void* k, const void* v, void* out,
                   const int32_t* pad_len, const int32_t* soi_idx,
                   int B, int L, int Lq, int q_row_offset, int Hq, int Hkv, int D, float scale,
                   int64_t q_stride_b, int64_t q_stride_l, int64_t q_stride_h,
                   int64_t kv_stride_b, int64_t kv_stride_l, int64_t kv_stride_h,
                   int64_t o_stride_b, int64_t o_stride_l, int64_t o_stride_h,
                   void* stream, float* debug_dump);

}  // namespace

extern "C" {

int ug_abi_version(void) { return UG_ABI_VERSION; }

const char* ug_last_error(void) { return g_err; }

size_t ug_decode_step_workspace_bytes(int B, int N, int V)
{
    (void)V;
    if (B <= 0 || N <= 0) return 0;
    // [ticket][done] padded to 256 B | one 64-bit result word per row
    return 256 + sizeof(unsigned long long) * (size_t)B * N;
}

int ug_decode_step_bf16(const void* logits, const void* uncond_logits, float guidance_scale,
                        int64_t* ids_inout, int64_t* sampled_out, uint8_t* mask_out,
                        float* sel_prob_out, float* conf_out,
                        const float* q_noise, const float* u_noise,
                        uint64_t philox_seed, uint64_t philox_offset, int64_t batch_offset,
                        int B, int N, int V, int64_t mask_token_id, int mask_len_cap, float temperature,
                        void* workspace, size_t workspace_bytes, void* stream)
{
    return decode_step_impl(logits, uncond_logits, guidance_scale, ids_inout, sampled_out, mask_out, sel_prob_out, conf_out, q_noise,
                            u_noise, philox_seed, philox_offset, batch_offset, B, N, V, mask_token_id, mask_len_cap, temperature,
                            workspace, workspace_bytes, stream, nullptr);
}

int ugt_decode_step_bf16(const void* logits, const void* uncond_logits, float guidance_scale,
                         int64_t* ids_inout, int64_t* sampled_out, uint8_t* mask_out,
                         float* sel_prob_out, float* conf_out,
                         const float* q_noise, const float* u_noise,
                         uint64_t philox_seed, uint64_t philox_offset, int64_t batch_offset,
                         int B, int N, int V, int64_t mask_token_id, int mask_len_cap, float temperature,
                         void* workspace, size_t workspace_bytes, void* stream, const ugt_decode_options* opts)
{
    return decode_step_impl(logits, uncond_logits, guidance_scale, ids_inout, sampled_out, mask_out, sel_prob_out, conf_out, q_noise,
                            u_noise, philox_seed, philox_offset, batch_offset, B, N, V, mask_token_id, mask_len_cap, temperature,
                            workspace, workspace_bytes, stream, opts);
}

}  // extern "C"

namespace {
int decode_step_impl(const void* logits, const void* uncond_logits, float guidance_scale,
                     int64_t* ids_inout, int64_t* sampled_out, uint8_t* mask_out,
                     float* sel_prob_out, float* conf_out,
                     const float* q_noise, const float* u_noise,
                     uint64_t philox_seed, uint64_t philox_offset, int64_t batch_offset,
                     int B, int N, int V, int64_t mask_token_id, int mask_len_cap, float temperature,
                     void* workspace, size_t workspace_bytes, void* stream, const ugt_decode_options* opts)
{
    if (!logits || !ids_inout || !sampled_out || !mask_out || !workspace)
        return fail(UG_ERR_NULL, "ug_decode_step_bf16: logits, ids_inout, sampled_out, mask_out and workspace are required");
    if (V != 4096 && V != 8192 && V != 16384)
        return fail(UG_ERR_SHAPE, "ug_decode_step_bf16: V=%d unsupported (4096, 8192, 16384)", V);
    if (B < 1 || N < 2 || N > 4096 || (long long)B * N >= (1ll << 31))
        return fail(UG_ERR_SHAPE, "ug_decode_step_bf16: B=%d N=%d unsupported (B>=1, 2<=N<=4096, B*N<2^31)", B, N);
    if (!aligned16(logits) || !aligned16(uncond_logits) || !aligned16(q_noise) || !aligned16(workspace))
        return fail(UG_ERR_ALIGN, "ug_decode_step_bf16: logits, uncond_logits, q_noise and workspace must be 16-byte aligned");
    if (workspace_bytes < ug_decode_step_workspace_bytes(B, N, V))
        return fail(UG_ERR_WORKSPACE, "ug_decode_step_bf16: workspace %zu < %zu bytes", workspace_bytes,
                    ug_decode_step_workspace_bytes(B, N, V));
    int major = 0;
    const int sms = sm_count(&major);
    if (sms <= 0 || major < 10) return fail(UG_ERR_DEVICE, "ug_decode_step_bf16: needs an sm_100 device (cc major %d)", major);

    ugk::DecodeArgs a;
    memset(&a, 0, sizeof a);
    a.logits = static_cast<const uint16_t*>(logits);
    a.uncond = static_cast<const uint16_t*>(uncond_logits);
    a.guidance = guidance_scale;
    a.ids = reinterpret_cast<long long*>(ids_inout);
    a.sampled = reinterpret_cast<long long*>(sampled_out);
    a.mask = mask_out;
    a.sel_prob_out = sel_prob_out;
    a.conf_out = conf_out;
    a.rowsum_out = opts ? opts->rowsum_out : nullptr;
    a.force_full_race = opts ? opts->force_exact_bits : 0;
    a.q_noise = q_noise;
    a.u_noise = u_noise;
    a.seed = philox_seed;
    a.offset = philox_offset;
    a.batch_offset = batch_offset;
    a.B = B; a.N = N; a.V = V;
    a.mask_id = mask_token_id;
    a.cap = mask_len_cap;
    a.temperature = temperature;
    unsigned* w = static_cast<unsigned*>(workspace);
    a.ticket = w;
    a.cta_done = w + 1;
    a.slots = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + 256);
    a.stats = opts ? opts->stats : nullptr;
    // The certified fast path (decode_fast.cu: row kernel + selection kernel, two launches on `stream`) serves the production configuration (in-kernel noise, no CFG,
    // V <= 8192); injected race noise, CFG and V = 16384 run on the exact-arithmetic kernel (decode_step.cu).
    // Both produce the ids and masks of the same specification.
    const bool fast = !(opts && opts->exact_kernel) && ugk::decode_fast_supported(a);
    cudaError_t e = fast ? ugk::launch_decode_fast(a, sms, static_cast<cudaStream_t>(stream))
                         : ugk::launch_decode_step(a, sms, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_decode_step_bf16 launch");
    return UG_OK;
}
}  // namespace

extern "C" {

int ug_mask_by_random_topk(const float* probs, const int64_t* mask_len, const float* u_noise,
                           uint64_t philox_seed, uint64_t philox_offset, int64_t batch_offset,
                           float temperature, uint8_t* mask_out, float* conf_out,
                           int B, int N, void* stream)
{
    if (!probs || !mask_len || !mask_out) return fail(UG_ERR_NULL, "ug_mask_by_random_topk: probs, mask_len, mask_out are required");
    if (B < 1 || N < 2 || N > 4096) return fail(UG_ERR_SHAPE, "ug_mask_by_random_topk: B=%d N=%d unsupported (2<=N<=4096)", B, N);
    ugk::DecodeArgs a;
    memset(&a, 0, sizeof a);
    a.psel = const_cast<float*>(probs);
    a.u_noise = u_noise;
    a.seed = philox_seed;
    a.offset = philox_offset;
    a.batch_offset = batch_offset;
    a.B = B; a.N = N;
    a.temperature = temperature;
    a.mask = mask_out;
    a.conf_out = conf_out;
    cudaError_t e = ugk::launch_topk(a, mask_len, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_mask_by_random_topk launch");
    return UG_OK;
}

int ug_export_noise(uint64_t philox_seed, uint64_t philox_offset, int64_t batch_offset,
                    int B, int N, int V, float* q_out, float* u_out, void* stream)
{
    if (!q_out && !u_out) return fail(UG_ERR_NULL, "ug_export_noise: nothing to write");
    if (B < 1 || N < 1 || V < 1) return fail(UG_ERR_SHAPE, "ug_export_noise: bad shape");
    cudaError_t e = ugk::launch_export_noise(philox_seed, philox_offset, batch_offset, B, N, V, q_out, u_out,
                                             static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_export_noise launch");
    return UG_OK;
}

int ug_omni_attention_rows_bf16(const void* q, const void* k, const void* v, void* out,
                                const int32_t* pad_len, const int32_t* soi_idx,
                                int B, int L, int Lq, int q_row_offset, int Hq, int Hkv, int D, float scale,
                                int64_t q_stride_b, int64_t q_stride_l, int64_t q_stride_h,
                                int64_t kv_stride_b, int64_t kv_stride_l, int64_t kv_stride_h,
                                int64_t o_stride_b, int64_t o_stride_l, int64_t o_stride_h,
                                void* stream)
{
    return attention_impl(q, k, v, out, pad_len, soi_idx, B, L, Lq, q_row_offset, Hq, Hkv, D, scale, q_stride_b, q_stride_l, q_stride_h,
                          kv_stride_b, kv_stride_l, kv_stride_h, o_stride_b, o_stride_l, o_stride_h, stream, nullptr);
}

int ugt_omni_attention_rows_bf16(const void* q, const void* k, const void* v, void* out,
                                 const int32_t* pad_len, const int32_t* soi_idx,
                                 int B, int L, int Lq, int q_row_offset, int Hq, int Hkv, int D, float scale,
                                 int64_t q_stride_b, int64_t q_stride_l, int64_t q_stride_h,
                                 int64_t kv_stride_b, int64_t kv_stride_l, int64_t kv_stride_h,
                                 int64_t o_stride_b, int64_t o_stride_l, int64_t o_stride_h,
                                 void* stream, float* debug_dump)
{
    return attention_impl(q, k, v, out, pad_len, soi_idx, B, L, Lq, q_row_offset, Hq, Hkv, D, scale, q_stride_b, q_stride_l, q_stride_h,
                          kv_stride_b, kv_stride_l, kv_stride_h, o_stride_b, o_stride_l, o_stride_h, stream, debug_dump);
}

}  // extern "C"

namespace {
int attention_impl(const void* q, const void* k, const void* v, void* out,
                   const int32_t* pad_len, const int32_t* soi_idx,
                   int B, int L, int Lq, int q_row_offset, int Hq, int Hkv, int D, float scale,
                   int64_t q_stride_b, int64_t q_stride_l, int64_t q_stride_h,
                   int64_t kv_stride_b, int64_t kv_stride_l, int64_t kv_stride_h,
                   int64_t o_stride_b, int64_t o_stride_l, int64_t o_stride_h,
                   void* stream, float* debug_dump)
{
    if (!q || !k || !v || !out || !pad_len || !soi_idx) return fail(UG_ERR_NULL, "ug_omni_attention_bf16: null pointer");
    if (D != 128 || Hkv < 1 || Hq % Hkv != 0 || B < 1 || L < 1 || L > 4096 || Lq < 1 || q_row_offset < 0 || q_row_offset + Lq > L)
        return fail(UG_ERR_SHAPE, "ug_omni_attention_bf16: B=%d L=%d Lq=%d q0=%d Hq=%d Hkv=%d D=%d unsupported (D=128, Hq%%Hkv==0, "
                    "L<=4096: 32 kv tiles per row of work masks, q0+Lq<=L)", B, L, Lq, q_row_offset, Hq, Hkv, D);
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out) ||
        ((q_stride_b | q_stride_l | q_stride_h | kv_stride_b | kv_stride_l | kv_stride_h | o_stride_b | o_stride_l | o_stride_h) & 7))
        return fail(UG_ERR_ALIGN, "ug_omni_attention_bf16: bases must be 16-byte aligned and strides multiples of 8 elements");
    int major = 0;
    const int sms = sm_count(&major);
    if (sms <= 0 || major < 10) return fail(UG_ERR_DEVICE, "ug_omni_attention_bf16: needs an sm_100 device (cc major %d)", major);
    ugk::AttnArgs a;
    a.q = q; a.k = k; a.v = v; a.out = out;
    a.pad_len = pad_len; a.soi_idx = soi_idx;
    a.B = B; a.L = L; a.Hq = Hq; a.Hkv = Hkv; a.scale = scale;
    a.Lq = Lq; a.q0 = q_row_offset;
    a.q_sb = q_stride_b; a.q_sl = q_stride_l; a.q_sh = q_stride_h;
    a.kv_sb = kv_stride_b; a.kv_sl = kv_stride_l; a.kv_sh = kv_stride_h;
    a.o_sb = o_stride_b; a.o_sl = o_stride_l; a.o_sh = o_stride_h;
    a.dbg = debug_dump;
    char msg[256] = "";
    int rc = ugk::launch_omni_attention(a, sms, static_cast<cudaStream_t>(stream), msg, sizeof msg);
    if (rc != 0) { snprintf(g_err, sizeof g_err, "ug_omni_attention_bf16: %s", msg); return rc; }
    return UG_OK;
}
}  // namespace

extern "C" {

int ug_omni_attention_bf16(const void* q, const void* k, const void* v, void* out,
                           const int32_t* pad_len, const int32_t* soi_idx,
                           int B, int L, int Hq, int Hkv, int D, float scale,
                           int64_t q_stride_b, int64_t q_stride_l, int64_t q_stride_h,
                           int64_t kv_stride_b, int64_t kv_stride_l, int64_t kv_stride_h,
                           int64_t o_stride_b, int64_t o_stride_l, int64_t o_stride_h,
                           void* stream)
{
    return ug_omni_attention_rows_bf16(q, k, v, out, pad_len, soi_idx, B, L, L, 0, Hq, Hkv, D, scale, q_stride_b, q_stride_l,
                                       q_stride_h, kv_stride_b, kv_stride_l, kv_stride_h, o_stride_b, o_stride_l, o_stride_h,
                                       stream);
}

}  // extern "C"

// ---- row-wise backbone kernels ---------------------------------------------------------------------
extern "C" int ug_add_rmsnorm(float* x, const void* delta_bf16, const void* weight_bf16, void* out_bf16,
                              int64_t M, int H, float eps, void* stream)
{
    if (!x || !weight_bf16 || !out_bf16) return fail(UG_ERR_NULL, "ug_add_rmsnorm: x, weight, out are required");
    const int vpl = H / 256;               // the instantiated row widths of add_rmsnorm_kernel
    if (M < 1 || H % 256 != 0 || !((vpl >= 1 && vpl <= 8) || (vpl <= 32 && vpl % 4 == 0)))
        return fail(UG_ERR_SHAPE, "ug_add_rmsnorm: M=%lld H=%d unsupported (H = 256 k with k in 1..8, 12, 16, 20, 24, 28, 32)", (long long)M, H);
    if (!aligned16(x) || !aligned16(delta_bf16) || !aligned16(weight_bf16) || !aligned16(out_bf16))
        return fail(UG_ERR_ALIGN, "ug_add_rmsnorm: pointers must be 16-byte aligned");
    cudaError_t e = ugk::launch_add_rmsnorm(x, delta_bf16, weight_bf16, out_bf16, M, H, eps, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_add_rmsnorm launch");
    return UG_OK;
}

extern "C" int ug_rope_qk(void* qkv_bf16, const float* cos, const float* sin, int64_t M, int L, int Hq, int Hkv, int D,
                          void* kv_cache_bf16, int cache_L, int cache_row_offset, void* stream)
{
    if (!qkv_bf16 || !cos || !sin) return fail(UG_ERR_NULL, "ug_rope_qk: null pointer");
    if (M < 1 || L < 1 || D % 16 != 0 || Hq < 1 || Hkv < 1 || M % L != 0) return fail(UG_ERR_SHAPE, "ug_rope_qk: bad shape");
    if (kv_cache_bf16 && M % L != 0)
        return fail(UG_ERR_SHAPE, "ug_rope_qk: with a K/V cache M (%lld) must be a whole number of sequences of L (%d) rows", (long long)M, L);
    if (kv_cache_bf16 && (cache_row_offset < 0 || cache_row_offset + L > cache_L))
        return fail(UG_ERR_SHAPE, "ug_rope_qk: cache window [%d, %d) outside cache_L=%d", cache_row_offset, cache_row_offset + L, cache_L);
    if (!aligned16(qkv_bf16) || !aligned16(cos) || !aligned16(sin) || !aligned16(kv_cache_bf16))
        return fail(UG_ERR_ALIGN, "ug_rope_qk: pointers must be 16-byte aligned");
    cudaError_t e = ugk::launch_rope_qk(qkv_bf16, cos, sin, M, L, Hq, Hkv, D, kv_cache_bf16, cache_L, cache_row_offset,
                                        static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_rope_qk launch");
    return UG_OK;
}

extern "C" int ug_swiglu(const void* gate_up_bf16, void* act_bf16, int64_t M, int I, void* stream)
{
    if (!gate_up_bf16 || !act_bf16) return fail(UG_ERR_NULL, "ug_swiglu: null pointer");
    if (M < 1 || I % 8 != 0) return fail(UG_ERR_SHAPE, "ug_swiglu: bad shape");
    if (!aligned16(gate_up_bf16) || !aligned16(act_bf16)) return fail(UG_ERR_ALIGN, "ug_swiglu: pointers must be 16-byte aligned");
    cudaError_t e = ugk::launch_swiglu(gate_up_bf16, act_bf16, M, I, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_swiglu launch");
    return UG_OK;
}

extern "C" int ug_embed_rows(float* x, void* xb_bf16, float* ss, const void* table_bf16, const int64_t* tokens, int64_t M, int H,
                             int64_t vocab, void* stream)
{
    if (!x || !table_bf16 || !tokens) return fail(UG_ERR_NULL, "ug_embed_rows: null pointer");
    if ((xb_bf16 == nullptr) != (ss == nullptr)) return fail(UG_ERR_NULL, "ug_embed_rows: xb and ss come together (both or neither)");
    if (M < 1 || H % 8 != 0 || vocab < 1 || (ss && H % 256 != 0))
        return fail(UG_ERR_SHAPE, "ug_embed_rows: bad shape (H %% 8 == 0; H %% 256 == 0 with xb / ss)");
    if (!aligned16(x) || !aligned16(table_bf16) || !aligned16(xb_bf16)) return fail(UG_ERR_ALIGN, "ug_embed_rows: pointers must be 16-byte aligned");
    cudaError_t e = ugk::launch_embed_rows(x, xb_bf16, ss, table_bf16, tokens, M, H, vocab, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_embed_rows launch");
    return UG_OK;
}

extern "C" int ug_gemm_bf16(const void* a, const void* w, const void* bias, void* out, int M, int N, int K,
                            int64_t lda, int64_t ldw, int64_t ldo, int mode, const float* row_ss, int ss_parts, float rms_eps, void* stream)
{
    if (!a || !w || !out) return fail(UG_ERR_NULL, "ug_gemm_bf16: null pointer");
    if (mode != UG_GEMM_PLAIN && mode != UG_GEMM_SWIGLU) return fail(UG_ERR_SHAPE, "ug_gemm_bf16: unknown mode %d", mode);
    if (mode == UG_GEMM_SWIGLU && bias) return fail(UG_ERR_SHAPE, "ug_gemm_bf16: the SwiGLU mode takes no bias");
    if (M < 1 || N < 128 || N % 128 || (N % 256 && mode != UG_GEMM_PLAIN) || K < 64 || K % 64)
        return fail(UG_ERR_SHAPE, "ug_gemm_bf16: need M >= 1, N %% 256 == 0 (plain mode: N %% 128 == 0), K %% 64 == 0 (M=%d N=%d K=%d)", M, N, K);
    const int n_out = mode == UG_GEMM_SWIGLU ? N / 2 : N;
    if (lda < K || ldw < K || ldo < n_out || lda % 8 || ldw % 8 || ldo % 8)
        return fail(UG_ERR_SHAPE, "ug_gemm_bf16: row strides must cover the row and be multiples of 8 elements");
    if (!aligned16(a) || !aligned16(w) || !aligned16(out) || (bias && !aligned16(bias)))
        return fail(UG_ERR_ALIGN, "ug_gemm_bf16: pointers must be 16-byte aligned");
    int major = 0;
    const int sms = sm_count(&major);
    if (major != 10) return fail(UG_ERR_DEVICE, "ug_gemm_bf16 needs an sm_100 device (tcgen05), found compute capability %d.x", major);
    ugk::GemmArgs g;
    g.a = a; g.w = w; g.bias = bias; g.out = out;
    g.lda = lda; g.ldw = ldw; g.ldo = ldo;
    g.M = M; g.N = N; g.K = K; g.mode = mode;
    if (row_ss && ss_parts < 1) return fail(UG_ERR_ARG, "ug_gemm_bf16: row_ss needs ss_parts >= 1");
    g.row_ss = row_ss; g.ss_parts = ss_parts; g.rms_eps = rms_eps;
    char msg[256] = "";
    const int rc = ugk::launch_gemm_tc(g, sms, static_cast<cudaStream_t>(stream), msg, sizeof msg);
    if (rc != 0) return fail(rc, "ug_gemm_bf16: %s", msg);
    return UG_OK;
}

extern "C" int ug_gemm_residual_bf16(const void* a, const void* w, float* x_inout, void* xb_out, float* ss_out, int M, int N, int K,
                                     int64_t lda, int64_t ldw, void* stream)
{
    if (!a || !w || !x_inout || !xb_out || !ss_out) return fail(UG_ERR_NULL, "ug_gemm_residual_bf16: null pointer");
    if (M < 1 || N < 256 || N % 256 || K < 64 || K % 64)
        return fail(UG_ERR_SHAPE, "ug_gemm_residual_bf16: need M >= 1, N %% 256 == 0, K %% 64 == 0 (M=%d N=%d K=%d)", M, N, K);
    if (lda < K || ldw < K || lda % 8 || ldw % 8)
        return fail(UG_ERR_SHAPE, "ug_gemm_residual_bf16: row strides must cover the row and be multiples of 8 elements");
    if (!aligned16(a) || !aligned16(w) || !aligned16(x_inout) || !aligned16(xb_out))
        return fail(UG_ERR_ALIGN, "ug_gemm_residual_bf16: pointers must be 16-byte aligned");
    int major = 0;
    const int sms = sm_count(&major);
    if (major != 10) return fail(UG_ERR_DEVICE, "ug_gemm_residual_bf16 needs an sm_100 device (tcgen05), found compute capability %d.x", major);
    ugk::GemmArgs g;
    g.a = a; g.w = w; g.bias = nullptr; g.out = nullptr;
    g.lda = lda; g.ldw = ldw; g.ldo = N;
    g.M = M; g.N = N; g.K = K; g.mode = ugk::kGemmResidual;
    g.res_x = x_inout; g.res_xb = xb_out; g.res_ss = ss_out;
    char msg[256] = "";
    const int rc = ugk::launch_gemm_tc(g, sms, static_cast<cudaStream_t>(stream), msg, sizeof msg);
    if (rc != 0) return fail(rc, "ug_gemm_residual_bf16: %s", msg);
    return UG_OK;
}

extern "C" int ug_qkv_rope_bf16(const void* a, const void* w_qkv, const void* bias_qkv, void* qkv_out, int M, int K,
                                int64_t lda, int64_t ldw, int64_t ldo, const float* cos, const float* sin, int L, int Hq, int Hkv,
                                void* kv_cache_bf16, int cache_L, int cache_row_offset, const float* row_ss, int ss_parts, float rms_eps,
                                void* stream)
{
    if (!a || !w_qkv || !qkv_out || !cos || !sin) return fail(UG_ERR_NULL, "ug_qkv_rope_bf16: null pointer");
    if (row_ss && ss_parts < 1) return fail(UG_ERR_ARG, "ug_qkv_rope_bf16: row_ss needs ss_parts >= 1");
    if (Hq < 1 || Hkv < 1 || L < 1 || M < 1) return fail(UG_ERR_SHAPE, "ug_qkv_rope_bf16: bad shape");
    const int N = (Hq + 2 * Hkv) * 128;
    if (N % 256 || K < 64 || K % 64) return fail(UG_ERR_SHAPE, "ug_qkv_rope_bf16: need (Hq + 2 Hkv) even and K %% 64 == 0 (N=%d K=%d)", N, K);
    if (lda < K || ldw < K || ldo < N || lda % 8 || ldw % 8 || ldo % 8)
        return fail(UG_ERR_SHAPE, "ug_qkv_rope_bf16: row strides must cover the row and be multiples of 8 elements");
    if (!aligned16(a) || !aligned16(w_qkv) || !aligned16(qkv_out) || (bias_qkv && !aligned16(bias_qkv)) || !aligned16(cos) ||
        !aligned16(sin) || (kv_cache_bf16 && !aligned16(kv_cache_bf16)))
        return fail(UG_ERR_ALIGN, "ug_qkv_rope_bf16: pointers must be 16-byte aligned");
    if (kv_cache_bf16 && (cache_row_offset < 0 || cache_row_offset + L > cache_L))
        return fail(UG_ERR_SHAPE, "ug_qkv_rope_bf16: rows [%d, %d) do not fit a cache of %d rows", cache_row_offset, cache_row_offset + L, cache_L);
    int major = 0;
    const int sms = sm_count(&major);
    if (major != 10) return fail(UG_ERR_DEVICE, "ug_qkv_rope_bf16 needs an sm_100 device (tcgen05), found compute capability %d.x", major);
    ugk::GemmArgs g;
    g.a = a; g.w = w_qkv; g.bias = bias_qkv; g.out = qkv_out;
    g.lda = lda; g.ldw = ldw; g.ldo = ldo;
    g.M = M; g.N = N; g.K = K; g.mode = ugk::kGemmQkvRope;
    g.rope_cos = cos; g.rope_sin = sin; g.kv_cache = kv_cache_bf16;
    g.rope_L = L; g.rope_Hq = Hq; g.rope_Hkv = Hkv; g.cache_L = cache_L; g.cache_row0 = cache_row_offset;
    g.row_ss = row_ss; g.ss_parts = ss_parts; g.rms_eps = rms_eps;
    char msg[256] = "";
    const int rc = ugk::launch_gemm_tc(g, sms, static_cast<cudaStream_t>(stream), msg, sizeof msg);
    if (rc != 0) return fail(rc, "ug_qkv_rope_bf16: %s", msg);
    return UG_OK;
}

extern "C" int ug_lfq_unpack_bf16(const int64_t* ids, void* z_bf16, int64_t n, int dim, void* stream)
{
    if (!ids || !z_bf16) return fail(UG_ERR_NULL, "ug_lfq_unpack_bf16: null pointer");
    if (n < 1 || dim < 1 || dim > 62) return fail(UG_ERR_SHAPE, "ug_lfq_unpack_bf16: bad shape");
    cudaError_t e = ugk::launch_lfq_unpack(ids, z_bf16, n, dim, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_lfq_unpack_bf16 launch");
    return UG_OK;
}

extern "C" int ug_lfq_postquant_bf16(const int64_t* ids, const float* w, const float* bias, void* z_bf16, int64_t n, int dim, int cpad,
                                     void* stream)
{
    if (!ids || !w || !z_bf16) return fail(UG_ERR_NULL, "ug_lfq_postquant_bf16: null pointer");
    if (n < 1 || dim < 1 || dim > 62 || cpad < dim) return fail(UG_ERR_SHAPE, "ug_lfq_postquant_bf16: bad shape");
    cudaError_t e = ugk::launch_lfq_postquant(ids, w, bias, z_bf16, n, dim, cpad, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_lfq_postquant_bf16 launch");
    return UG_OK;
}

extern "C" int ug_softmax_rows_bf16(const void* x_bf16, void* y_bf16, int64_t rows, int cols, int cols_valid, float scale, void* stream)
{
    if (!x_bf16 || !y_bf16) return fail(UG_ERR_NULL, "ug_softmax_rows_bf16: null pointer");
    if (rows < 1 || cols < 8 || cols % 8 || cols_valid < 8 || cols_valid % 8 || cols_valid > cols)
        return fail(UG_ERR_SHAPE, "ug_softmax_rows_bf16: need cols %% 8 == 0 and 8 <= cols_valid <= cols, cols_valid %% 8 == 0");
    if (!aligned16(x_bf16) || !aligned16(y_bf16)) return fail(UG_ERR_ALIGN, "ug_softmax_rows_bf16: pointers must be 16-byte aligned");
    cudaError_t e = ugk::launch_softmax_rows(x_bf16, y_bf16, rows, cols, cols_valid, scale, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_softmax_rows_bf16 launch");
    return UG_OK;
}

extern "C" int ug_conv_nhwc_bf16(const void* x_bf16, const void* w_packed_bf16, const void* bias_bf16, void* y_bf16,
                                 int B, int H, int W, int Cin, int Cout, int ksize, void* stream)
{
    if (!x_bf16 || !w_packed_bf16 || !y_bf16) return fail(UG_ERR_NULL, "ug_conv_nhwc_bf16: null pointer");
    if (B < 1 || H < 1 || W < 1 || Cin < 8 || Cin % 8 || Cout < 8 || Cout % 8 || (ksize != 3 && ksize != 1))
        return fail(UG_ERR_SHAPE, "ug_conv_nhwc_bf16: need Cin %% 8 == 0, Cout %% 8 == 0, ksize 3 or 1 (B=%d H=%d W=%d Cin=%d Cout=%d k=%d)",
                    B, H, W, Cin, Cout, ksize);
    const long long boxes = (long long)B * ((H + 3) / 4) * ((W + 31) / 32);
    if (boxes * 128 >= (1ll << 31)) return fail(UG_ERR_SHAPE, "ug_conv_nhwc_bf16: too many pixels");
    if (!aligned16(x_bf16) || !aligned16(w_packed_bf16) || !aligned16(y_bf16) || (bias_bf16 && !aligned16(bias_bf16)))
        return fail(UG_ERR_ALIGN, "ug_conv_nhwc_bf16: pointers must be 16-byte aligned");
    int major = 0;
    const int sms = sm_count(&major);
    if (major != 10) return fail(UG_ERR_DEVICE, "ug_conv_nhwc_bf16 needs an sm_100 device (tcgen05), found compute capability %d.x", major);
    const int taps = ksize * ksize, cin_pad = (Cin + 63) / 64 * 64, n_pad = (Cout + 127) / 128 * 128;
    ugk::GemmArgs g;
    g.a = x_bf16; g.w = w_packed_bf16; g.bias = bias_bf16; g.out = y_bf16;
    g.M = (int)(boxes * 128); g.N = n_pad; g.K = taps * cin_pad;
    g.lda = Cin; g.ldw = g.K; g.ldo = Cout;
    g.mode = ugk::kGemmPlain;
    g.conv_B = B; g.conv_H = H; g.conv_W = W; g.conv_Cin = Cin; g.conv_Cout = Cout; g.conv_taps = taps;
    char msg[256] = "";
    const int rc = ugk::launch_gemm_tc(g, sms, static_cast<cudaStream_t>(stream), msg, sizeof msg);
    if (rc != 0) return fail(rc, "ug_conv_nhwc_bf16: %s", msg);
    return UG_OK;
}

extern "C" size_t ug_groupnorm_workspace_floats(int B, int64_t HW, int C)
{
    if (B < 1 || HW < 1 || C < 1) return 0;
    return ugk::groupnorm_workspace_floats(B, HW, C);
}

extern "C" int ug_bias_add_nhwc_bf16(const void* a_bf16, const void* bias_a_bf16, const void* r_bf16, const void* bias_r_bf16, void* out_bf16,
                                     int64_t rows, int C, void* stream)
{
    if (!a_bf16 || !out_bf16) return fail(UG_ERR_NULL, "ug_bias_add_nhwc_bf16: null pointer");
    if (bias_r_bf16 && !r_bf16) return fail(UG_ERR_ARG, "ug_bias_add_nhwc_bf16: bias_r without r");
    if (rows < 1 || C < 8 || C % 8) return fail(UG_ERR_SHAPE, "ug_bias_add_nhwc_bf16: need C %% 8 == 0");
    if (!aligned16(a_bf16) || !aligned16(out_bf16) || (r_bf16 && !aligned16(r_bf16)))
        return fail(UG_ERR_ALIGN, "ug_bias_add_nhwc_bf16: pointers must be 16-byte aligned");
    cudaError_t e = ugk::launch_bias_add_nhwc(a_bf16, bias_a_bf16, r_bf16, bias_r_bf16, out_bf16, rows, C, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_bias_add_nhwc_bf16 launch");
    return UG_OK;
}

extern "C" int ug_bias_add_stats_nhwc_bf16(const void* a_bf16, const void* bias_a_bf16, const void* r_bf16, const void* bias_r_bf16, void* out_bf16,
                                           float* workspace, int B, int64_t HW, int C, void* stream)
{
    if (!a_bf16 || !out_bf16 || !workspace) return fail(UG_ERR_NULL, "ug_bias_add_stats_nhwc_bf16: null pointer");
    if (bias_r_bf16 && !r_bf16) return fail(UG_ERR_ARG, "ug_bias_add_stats_nhwc_bf16: bias_r without r");
    if (B < 1 || HW < 1 || C < 32 || C % 32 || C % 8 || 256 % (C / 8))
        return fail(UG_ERR_SHAPE, "ug_bias_add_stats_nhwc_bf16: need C %% 32 == 0 and 256 %% (C/8) == 0 (C=%d)", C);
    if (!aligned16(a_bf16) || !aligned16(out_bf16) || (r_bf16 && !aligned16(r_bf16)))
        return fail(UG_ERR_ALIGN, "ug_bias_add_stats_nhwc_bf16: pointers must be 16-byte aligned");
    cudaError_t e = ugk::launch_bias_add_stats_nhwc(a_bf16, bias_a_bf16, r_bf16, bias_r_bf16, out_bf16, workspace, B, HW, C,
                                                    static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_bias_add_stats_nhwc_bf16 launch");
    return UG_OK;
}

extern "C" int ug_groupnorm_nhwc_bf16(const void* x_bf16, const void* in_bias_bf16, const void* gamma_bf16, const void* beta_bf16, void* y_bf16,
                                      float* workspace, int B, int64_t HW, int C, float eps, int swish, int stats_ready, void* stream)
{
    if (!x_bf16 || !gamma_bf16 || !beta_bf16 || !y_bf16 || !workspace) return fail(UG_ERR_NULL, "ug_groupnorm_nhwc_bf16: null pointer");
    if (B < 1 || HW < 1 || C < 32 || C % 32 || C % 8 || 256 % (C / 8))
        return fail(UG_ERR_SHAPE, "ug_groupnorm_nhwc_bf16: need C %% 32 == 0 and 256 %% (C/8) == 0 (C=%d)", C);
    if (!aligned16(x_bf16) || !aligned16(y_bf16)) return fail(UG_ERR_ALIGN, "ug_groupnorm_nhwc_bf16: x and y must be 16-byte aligned");
    if (stats_ready && in_bias_bf16) return fail(UG_ERR_ARG, "ug_groupnorm_nhwc_bf16: stats_ready excludes in_bias");
    cudaError_t e = ugk::launch_groupnorm_nhwc(x_bf16, in_bias_bf16, gamma_bf16, beta_bf16, y_bf16, workspace, B, HW, C, eps, swish, stats_ready,
                                               static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_groupnorm_nhwc_bf16 launch");
    return UG_OK;
}

extern "C" int ug_upsample2x_nhwc_bf16(const void* x_bf16, void* y_bf16, int B, int H, int W, int C, void* stream)
{
    if (!x_bf16 || !y_bf16) return fail(UG_ERR_NULL, "ug_upsample2x_nhwc_bf16: null pointer");
    if (B < 1 || H < 1 || W < 1 || C < 8 || C % 8) return fail(UG_ERR_SHAPE, "ug_upsample2x_nhwc_bf16: need C %% 8 == 0");
    if (!aligned16(x_bf16) || !aligned16(y_bf16)) return fail(UG_ERR_ALIGN, "ug_upsample2x_nhwc_bf16: pointers must be 16-byte aligned");
    cudaError_t e = ugk::launch_upsample2x_nhwc(x_bf16, y_bf16, B, H, W, C, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_upsample2x_nhwc_bf16 launch");
    return UG_OK;
}

extern "C" int ug_cfg_mix_bf16(void* cond_logits_inout, const void* uncond_logits, float guidance_scale, int64_t n, void* stream)
{
    if (!cond_logits_inout || !uncond_logits) return fail(UG_ERR_NULL, "ug_cfg_mix_bf16: null pointer");
    if (n < 8 || n % 8) return fail(UG_ERR_SHAPE, "ug_cfg_mix_bf16: the element count must be a positive multiple of 8");
    if (!aligned16(cond_logits_inout) || !aligned16(uncond_logits)) return fail(UG_ERR_ALIGN, "ug_cfg_mix_bf16: pointers must be 16-byte aligned");
    cudaError_t e = ugk::launch_cfg_mix(cond_logits_inout, uncond_logits, guidance_scale, n, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "ug_cfg_mix_bf16 launch");
    return UG_OK;
}
