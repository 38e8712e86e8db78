// omni_attention.h -- internal launch interface of the tcgen05 omni-attention kernel.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace ugk {

struct AttnArgs {
    const void* q; const void* k; const void* v; void* out;
    const int32_t* pad_len; const int32_t* soi_idx;
    int B, L, Hq, Hkv;
    int Lq;          // query / output rows per sequence (q and out are [B, Lq, Hq, D])
    int q0;          // absolute sequence index of query row 0 (0 for full attention; soi_idx for suffix-only passes)
    float scale;
    int64_t q_sb, q_sl, q_sh;
    int64_t kv_sb, kv_sl, kv_sh;
    int64_t o_sb, o_sl, o_sh;
    // test hook (ugt_omni_attention_rows_bf16): float[2*128*128 + 2048] receiving S (raw q.k) and P (bf16-rounded probabilities)
    // of kv tile 0 of the FIRST q tile (rows 0..127) of (b=0, h=0), then cycle stamps; nullptr = off
    float* dbg = nullptr;
};

// returns 0, a negative UG_ERR_* code, or a positive cudaError_t; msg receives a description
int launch_omni_attention(const AttnArgs& a, int num_sms, cudaStream_t stream, char* msg, size_t msg_len);

}  // namespace ugk
