"""Tensor-level wrappers over the C ABI (include/unigen_b200.h).

PyTorch is used here only for device memory and streams: every function takes CUDA tensors,
passes `data_ptr()`s and the current stream to libunigen_b200.so and returns without
synchronising.  Nothing here computes on the host or falls back to torch operators.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

_FLT_MAX = torch.finfo(torch.float32).max


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.UnigenB200Error("unigen-b200 kernels need CUDA tensors (there is no CPU path)")


DECODE_STEP_LAUNCHES = 2          # decode_fast_kernel + select_kernel (the production path of ug_decode_step_bf16)


class DecodeWorkspace:
    """Caller-owned workspace of the decode step (zeroed once; the kernel re-arms it)."""

    def __init__(self, B: int, N: int, V: int, device):
        self.B, self.N, self.V = B, N, V
        self.nbytes = _lib.lib().ug_decode_step_workspace_bytes(B, N, V)
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)


def decode_step(logits: torch.Tensor, ids_inout: torch.Tensor, *, mask_token_id: int, mask_len_cap: int,
                temperature: float, workspace: DecodeWorkspace,
                sampled_out: Optional[torch.Tensor] = None, mask_out: Optional[torch.Tensor] = None,
                sel_prob_out: Optional[torch.Tensor] = None, conf_out: Optional[torch.Tensor] = None,
                uncond_logits: Optional[torch.Tensor] = None, guidance_scale: float = 0.0,
                q_noise: Optional[torch.Tensor] = None, u_noise: Optional[torch.Tensor] = None,
                seed: int = 0, offset: int = 0, batch_offset: int = 0, _debug: Optional[dict] = None):
    """One fused MaskGIT step (models/unigen.py:407-453).  `ids_inout` int64 [B,N] is updated in
    place; returns (sampled_ids int64 [B,N], masking uint8 [B,N]).
    `_debug` (tests only): dict with rowsum_out / stats tensors, force_exact_bits, exact_kernel -> the call goes through
    `ugt_decode_step_bf16` (include/unigen_b200_test.h)."""
    _need_cuda(logits, ids_inout)
    if logits.dtype != torch.bfloat16 or logits.dim() != 3 or not logits.is_contiguous():
        raise ValueError("logits must be a contiguous bf16 [B,N,V] tensor")
    B, N, V = logits.shape
    if ids_inout.dtype != torch.int64 or tuple(ids_inout.shape) != (B, N) or not ids_inout.is_contiguous():
        raise ValueError("ids_inout must be a contiguous int64 [B,N] tensor")
    if (workspace.B, workspace.N, workspace.V) != (B, N, V):
        raise ValueError("workspace was sized for a different shape")
    dev = logits.device
    if sampled_out is None:
        sampled_out = torch.empty(B, N, dtype=torch.int64, device=dev)
    if mask_out is None:
        mask_out = torch.empty(B, N, dtype=torch.uint8, device=dev)
    for t, dt, shp in ((uncond_logits, torch.bfloat16, (B, N, V)), (q_noise, torch.float32, (B, N, V)),
                       (u_noise, torch.float32, (B, N)), (sel_prob_out, torch.float32, (B, N)),
                       (conf_out, torch.float32, (B, N)), (sampled_out, torch.int64, (B, N)),
                       (mask_out, torch.uint8, (B, N))):
        if t is not None and (t.dtype != dt or tuple(t.shape) != shp or not t.is_contiguous() or t.device != dev):
            raise ValueError(f"bad optional tensor: expected contiguous {dt} {shp} on {dev}")
    args = (_ptr(logits), _ptr(uncond_logits), float(guidance_scale),
            _ptr(ids_inout), _ptr(sampled_out), _ptr(mask_out), _ptr(sel_prob_out), _ptr(conf_out),
            _ptr(q_noise), _ptr(u_noise), seed & (2**64 - 1), offset & (2**64 - 1), batch_offset,
            B, N, V, mask_token_id, int(mask_len_cap), float(temperature),
            _ptr(workspace.buf), workspace.nbytes, _stream())
    if _debug is None:
        rc = _lib.lib().ug_decode_step_bf16(*args)
    else:
        import ctypes
        o = _lib.DecodeTestOptions(_ptr(_debug.get("rowsum_out")), int(_debug.get("force_exact_bits", 0)),
                                   int(bool(_debug.get("exact_kernel", False))), _ptr(_debug.get("stats")))
        rc = _lib.lib().ugt_decode_step_bf16(*args, ctypes.byref(o))
    _lib.check(rc)
    return sampled_out, mask_out


def mask_by_random_topk(probs: torch.Tensor, mask_len: torch.Tensor, temperature: float, *,
                        u_noise: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0,
                        batch_offset: int = 0, conf_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """models/sampling.py:41-46 on the device.  probs float32 [B,N]; mask_len int64 [B] or [B,1]."""
    _need_cuda(probs, mask_len)
    if probs.dtype != torch.float32 or probs.dim() != 2:
        raise ValueError("probs must be float32 [B,N]")
    probs = probs.contiguous()
    B, N = probs.shape
    ml = mask_len.reshape(-1).to(torch.int64).contiguous()
    if ml.numel() != B:
        raise ValueError("mask_len must have one entry per row")
    out = torch.empty(B, N, dtype=torch.uint8, device=probs.device)
    rc = _lib.lib().ug_mask_by_random_topk(_ptr(probs), _ptr(ml), _ptr(u_noise), seed & (2**64 - 1),
                                           offset & (2**64 - 1), batch_offset, float(temperature),
                                           _ptr(out), _ptr(conf_out), B, N, _stream())
    _lib.check(rc)
    return out.view(torch.bool)


def export_noise(seed: int, offset: int, B: int, N: int, V: int, device, batch_offset: int = 0,
                 want_q: bool = True, want_u: bool = True):
    """(q float32 [B,N,V], u float32 [B,N]) exactly as the Philox mode of decode_step draws them."""
    q = torch.empty(B, N, V, dtype=torch.float32, device=device) if want_q else None
    u = torch.empty(B, N, dtype=torch.float32, device=device) if want_u else None
    rc = _lib.lib().ug_export_noise(seed & (2**64 - 1), offset & (2**64 - 1), batch_offset, B, N, V,
                                    _ptr(q), _ptr(u), _stream())
    _lib.check(rc)
    return q, u


def omni_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, pad_len: torch.Tensor,
                   soi_idx: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None, q_offset: int = 0,
                   _debug_dump: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Omni attention.  q [B,Lq,Hq,D], k/v [B,L,Hkv,D] bf16 (last dim contiguous, arbitrary other strides);
    pad_len / soi_idx int32 [B].  Query row r is sequence row q_offset + r (Lq = L, q_offset = 0 for the full
    pass; q_offset = soi_idx for the image-block-only pass of the prefix K/V reuse).  Returns out [B,Lq,Hq,D] bf16."""
    _need_cuda(q, k, v, pad_len, soi_idx)
    B, Lq, Hq, D = q.shape
    L, Hkv = k.shape[1], k.shape[2]
    if q.dtype != torch.bfloat16 or k.dtype != torch.bfloat16 or v.dtype != torch.bfloat16:
        raise ValueError("q, k, v must be bf16")
    if q.stride(3) != 1 or k.stride(3) != 1 or v.stride(3) != 1:
        raise ValueError("head_dim must be contiguous")
    if k.shape != v.shape or any(sk != sv for sk, sv, n in zip(k.stride(), v.stride(), k.shape) if n > 1):
        raise ValueError("k and v must share shape and strides")           # (the stride of a size-1 dimension is immaterial)
    if pad_len.dtype != torch.int32 or soi_idx.dtype != torch.int32:
        raise ValueError("pad_len / soi_idx must be int32")
    if out is None:
        out = torch.empty(B, Lq, Hq, D, dtype=torch.bfloat16, device=q.device)
    args = (_ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(pad_len), _ptr(soi_idx),
            B, L, Lq, int(q_offset), Hq, Hkv, D, float(scale),
            q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2),
            out.stride(0), out.stride(1), out.stride(2), _stream())
    if _debug_dump is None:
        rc = _lib.lib().ug_omni_attention_rows_bf16(*args)
    else:                                       # tests / tools only (include/unigen_b200_test.h)
        rc = _lib.lib().ugt_omni_attention_rows_bf16(*args, _ptr(_debug_dump))
    _lib.check(rc)
    return out


# ------------------------------------------------------------------------------------------------
# row-wise kernels of the decoder layer
def add_rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float, delta: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x fp32 [M,H] residual stream (updated in place with `delta` bf16 [M,H] when given);
    returns bf16 [M,H] = weight * rmsnorm(x)."""
    _need_cuda(x, weight, delta)
    M, H = x.shape
    if x.dtype != torch.float32 or not x.is_contiguous() or weight.dtype != torch.bfloat16:
        raise ValueError("x must be contiguous fp32 [M,H], weight bf16 [H]")
    if delta is not None and (delta.dtype != torch.bfloat16 or tuple(delta.shape) != (M, H) or not delta.is_contiguous()):
        raise ValueError("delta must be contiguous bf16 [M,H]")
    if out is None:
        out = torch.empty(M, H, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().ug_add_rmsnorm(_ptr(x), _ptr(delta), _ptr(weight), _ptr(out), M, H, float(eps), _stream()))
    return out


def rope_qk_(qkv: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, L: int, Hq: int, Hkv: int, D: int,
             kv_cache: Optional[torch.Tensor] = None, cache_row_offset: int = 0) -> torch.Tensor:
    """In-place RoPE on the q and k heads of the fused qkv activation bf16 [M,(Hq+2Hkv)*D].  With `kv_cache`
    (bf16 [B, cache_L, 2, Hkv, D]) the rotated k and the v rows are also written into the cache window
    [cache_row_offset, cache_row_offset + L) -- the K/V-cache update of the text-prefix reuse, fused."""
    _need_cuda(qkv, cos, sin, kv_cache)
    M = qkv.shape[0]
    if qkv.dtype != torch.bfloat16 or not qkv.is_contiguous() or qkv.shape[1] != (Hq + 2 * Hkv) * D:
        raise ValueError("qkv must be contiguous bf16 [M,(Hq+2Hkv)*D]")
    if cos.dtype != torch.float32 or tuple(cos.shape) != (L, D // 2) or not cos.is_contiguous() or sin.shape != cos.shape:
        raise ValueError("cos/sin must be contiguous fp32 [L, D/2]")
    cache_L = 0
    if kv_cache is not None:
        if kv_cache.dtype != torch.bfloat16 or not kv_cache.is_contiguous() or tuple(kv_cache.shape[2:]) != (2, Hkv, D) \
                or kv_cache.shape[0] * L != M:
            raise ValueError("kv_cache must be contiguous bf16 [M/L, cache_L, 2, Hkv, D]")
        cache_L = kv_cache.shape[1]
    _lib.check(_lib.lib().ug_rope_qk(_ptr(qkv), _ptr(cos), _ptr(sin), M, L, Hq, Hkv, D, _ptr(kv_cache), cache_L,
                                     int(cache_row_offset), _stream()))
    return qkv


def swiglu(gate_up: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 [M,2I] = [gate | up] -> bf16 [M,I] = silu(gate) * up."""
    _need_cuda(gate_up)
    M, I2 = gate_up.shape
    if gate_up.dtype != torch.bfloat16 or not gate_up.is_contiguous() or I2 % 16:
        raise ValueError("gate_up must be contiguous bf16 [M,2I]")
    if out is None:
        out = torch.empty(M, I2 // 2, dtype=torch.bfloat16, device=gate_up.device)
    _lib.check(_lib.lib().ug_swiglu(_ptr(gate_up), _ptr(out), M, I2 // 2, _stream()))
    return out


def embed_rows(table: torch.Tensor, tokens: torch.Tensor, out: Optional[torch.Tensor] = None, *, with_norm_operands: bool = False):
    """table bf16 [vocab,H], tokens int64 [M] -> fp32 [M,H].  With `with_norm_operands` returns (x, xb, ss): also
    xb = bf16(x) and ss fp32 [M, H/128] = per-128-column sums of x^2, the operands of the norm-folded GEMMs."""
    _need_cuda(table, tokens)
    vocab, H = table.shape
    tokens = tokens.reshape(-1)
    if table.dtype != torch.bfloat16 or not table.is_contiguous() or tokens.dtype != torch.int64 or not tokens.is_contiguous():
        raise ValueError("table must be contiguous bf16, tokens contiguous int64")
    M = tokens.numel()
    if out is None:
        out = torch.empty(M, H, dtype=torch.float32, device=table.device)
    xb = ss = None
    if with_norm_operands:
        xb = torch.empty(M, H, dtype=torch.bfloat16, device=table.device)
        ss = torch.empty(M, H // 128, dtype=torch.float32, device=table.device)
    _lib.check(_lib.lib().ug_embed_rows(_ptr(out), _ptr(xb), _ptr(ss), _ptr(table), _ptr(tokens), M, H, vocab, _stream()))
    return (out, xb, ss) if with_norm_operands else out


def norm_operands(x: torch.Tensor):
    """fp32 [M,H] residual stream -> (xb bf16 [M,H], ss fp32 [M,H/128]) with plain torch ops (host-side helper for
    callers that hand in their own embeddings; the hot loop gets these from the GEMM / embedding epilogues)."""
    M, H = x.shape
    return x.bfloat16(), (x.float() ** 2).view(M, H // 128, 128).sum(-1).contiguous()


def cfg_mix_(cond_logits: torch.Tensor, uncond_logits: torch.Tensor, guidance_scale: float) -> torch.Tensor:
    """In place on `cond_logits` (bf16, contiguous): guidance_scale * (cond - uncond) + uncond with a bf16 rounding after
    every op (models/unigen.py:410 under autocast) -- what `decode_step(..., uncond_logits=...)` forms inline.  Afterwards
    `decode_step(cond_logits, ...)` without `uncond_logits` takes the certified fast kernel."""
    _need_cuda(cond_logits, uncond_logits)
    if cond_logits.dtype != torch.bfloat16 or uncond_logits.dtype != torch.bfloat16 or cond_logits.shape != uncond_logits.shape \
            or not cond_logits.is_contiguous() or not uncond_logits.is_contiguous() or cond_logits.numel() % 8:
        raise ValueError("cond/uncond logits must be contiguous bf16 tensors of the same shape, element count % 8 == 0")
    _lib.check(_lib.lib().ug_cfg_mix_bf16(_ptr(cond_logits), _ptr(uncond_logits), float(guidance_scale), cond_logits.numel(), _stream()))
    return cond_logits


GEMM_PLAIN, GEMM_SWIGLU = 0, 1


def pack_gate_up(w_gate: torch.Tensor, w_up: torch.Tensor) -> torch.Tensor:
    """[I,K] gate_proj and up_proj weights -> [2I,K] with rows interleaved in blocks of 128 (the layout
    `gemm(..., mode=GEMM_SWIGLU)` expects: block 2j = gate rows [128j, 128j+128), block 2j+1 = the same up rows)."""
    I, K = w_gate.shape
    if w_up.shape != w_gate.shape or I % 128:
        raise ValueError("gate/up weights must have the same shape [I,K] with I % 128 == 0")
    return torch.stack([w_gate.reshape(I // 128, 128, K), w_up.reshape(I // 128, 128, K)], 1).reshape(2 * I, K).contiguous()


def _row_ss_args(row_ss: Optional[torch.Tensor], M: int, eps: float):
    if row_ss is None:
        return None, 0, 0.0
    if row_ss.dtype != torch.float32 or row_ss.dim() != 2 or row_ss.shape[0] != M or not row_ss.is_contiguous():
        raise ValueError("row_ss must be contiguous fp32 [M, parts]")
    return row_ss.data_ptr(), int(row_ss.shape[1]), float(eps)


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, mode: int = GEMM_PLAIN,
         out: Optional[torch.Tensor] = None, row_ss: Optional[torch.Tensor] = None, rms_eps: float = 1e-6) -> torch.Tensor:
    """bf16 a [M,K] x w [N,K]^T on the tcgen05 kernel.  mode GEMM_PLAIN: bf16 [M,N] (+bias);
    GEMM_SWIGLU: w from `pack_gate_up`, returns bf16 [M,N/2] = silu(gate) * up.
    `row_ss` fp32 [M, parts]: RMSNorm folding -- row m of the accumulator is scaled by rsqrt(sum(row_ss[m]) / K + rms_eps)
    (a = bf16(x), the norm gain folded into the columns of w; include/unigen_b200.h)."""
    _need_cuda(a, w, bias)
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or a.dim() != 2 or w.dim() != 2:
        raise ValueError("a and w must be 2-D bf16")
    if a.stride(1) != 1 or w.stride(1) != 1:
        raise ValueError("a and w must have contiguous rows")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError("inner dimensions differ")
    n_out = N // 2 if mode == GEMM_SWIGLU else N
    if out is None:
        out = torch.empty(M, n_out, dtype=torch.bfloat16, device=a.device)
    elif out.shape != (M, n_out) or out.dtype != torch.bfloat16 or out.stride(1) != 1:
        raise ValueError("out must be bf16 [M, n_out] with contiguous rows")
    ssp, parts, eps = _row_ss_args(row_ss, M, rms_eps)
    _lib.check(_lib.lib().ug_gemm_bf16(_ptr(a), _ptr(w), _ptr(bias), _ptr(out), M, N, K,
                                       a.stride(0), w.stride(0), out.stride(0), mode, ssp, parts, eps, _stream()))
    return out


def gemm_residual_(a: torch.Tensor, w: torch.Tensor, x: torch.Tensor, xb: Optional[torch.Tensor] = None,
                   ss: Optional[torch.Tensor] = None):
    """o_proj / down_proj + residual add + statistics of the next RMSNorm in one launch: x fp32 [M,N] += bf16(a w^T) in place;
    returns (xb bf16 [M,N] = bf16(x), ss fp32 [M, N/128] = per-128-column sums of x^2)."""
    _need_cuda(a, w, x)
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or a.dim() != 2 or w.dim() != 2 or a.stride(1) != 1 or w.stride(1) != 1:
        raise ValueError("a and w must be 2-D bf16 with contiguous rows")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K or x.dtype != torch.float32 or tuple(x.shape) != (M, N) or not x.is_contiguous():
        raise ValueError("x must be contiguous fp32 [M, N]; inner dimensions of a and w must agree")
    if xb is None:
        xb = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    if ss is None:
        ss = torch.empty(M, N // 128, dtype=torch.float32, device=a.device)
    if xb.dtype != torch.bfloat16 or tuple(xb.shape) != (M, N) or not xb.is_contiguous() or ss.dtype != torch.float32 \
            or tuple(ss.shape) != (M, N // 128) or not ss.is_contiguous():
        raise ValueError("xb must be contiguous bf16 [M,N], ss contiguous fp32 [M,N/128]")
    _lib.check(_lib.lib().ug_gemm_residual_bf16(_ptr(a), _ptr(w), _ptr(x), _ptr(xb), _ptr(ss), M, N, K,
                                                a.stride(0), w.stride(0), _stream()))
    return xb, ss


def qkv_rope(a: torch.Tensor, w_qkv: torch.Tensor, bias_qkv: Optional[torch.Tensor], cos: torch.Tensor, sin: torch.Tensor,
             L: int, Hq: int, Hkv: int, kv_cache: Optional[torch.Tensor] = None, cache_row_offset: int = 0,
             out: Optional[torch.Tensor] = None, row_ss: Optional[torch.Tensor] = None, rms_eps: float = 1e-6) -> torch.Tensor:
    """Fused q/k/v projection + RoPE (+ K/V-cache update): bf16 a [M,K] -> bf16 qkv [M,(Hq+2Hkv)*128] with the q and k
    heads rotated (row m = position m % L; cos/sin fp32 [L,64]); with `kv_cache` (bf16 [M/L, cache_L, 2, Hkv, 128]) the
    rotated k and the v rows also land in cache rows [cache_row_offset, cache_row_offset + L).  One launch instead of
    `gemm` + `rope_qk_`."""
    _need_cuda(a, w_qkv, bias_qkv, cos, sin, kv_cache)
    M, K = a.shape
    N = (Hq + 2 * Hkv) * 128
    if a.dtype != torch.bfloat16 or w_qkv.dtype != torch.bfloat16 or a.stride(1) != 1 or w_qkv.stride(1) != 1 \
            or tuple(w_qkv.shape) != (N, K):
        raise ValueError("a must be bf16 [M,K] and w_qkv bf16 [(Hq+2Hkv)*128, K], rows contiguous")
    if cos.dtype != torch.float32 or tuple(cos.shape) != (L, 64) or not cos.is_contiguous() or sin.shape != cos.shape \
            or sin.dtype != torch.float32 or not sin.is_contiguous():
        raise ValueError("cos/sin must be contiguous fp32 [L, 64]")
    cache_L = 0
    if kv_cache is not None:
        if kv_cache.dtype != torch.bfloat16 or not kv_cache.is_contiguous() or tuple(kv_cache.shape[2:]) != (2, Hkv, 128) \
                or kv_cache.shape[0] * L != M:
            raise ValueError("kv_cache must be contiguous bf16 [M/L, cache_L, 2, Hkv, 128]")
        cache_L = kv_cache.shape[1]
    if out is None:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    elif tuple(out.shape) != (M, N) or out.dtype != torch.bfloat16 or out.stride(1) != 1:
        raise ValueError("out must be bf16 [M, (Hq+2Hkv)*128] with contiguous rows")
    ssp, parts, eps = _row_ss_args(row_ss, M, rms_eps)
    _lib.check(_lib.lib().ug_qkv_rope_bf16(_ptr(a), _ptr(w_qkv), _ptr(bias_qkv), _ptr(out), M, K, a.stride(0), w_qkv.stride(0),
                                           out.stride(0), _ptr(cos), _ptr(sin), L, Hq, Hkv, _ptr(kv_cache), cache_L,
                                           int(cache_row_offset), ssp, parts, eps, _stream()))
    return out
