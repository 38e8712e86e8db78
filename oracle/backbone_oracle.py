"""oracle/backbone_oracle.py -- TEST INFRASTRUCTURE ONLY.

Plain PyTorch fp32 restatement of the floating-point half of the hot path:

* the omni attention mask in closed form, pinned against
  `create_attention_mask_predict_next(..., rm_pad_in_image=True)`
  (training/prompting_utils.py:975-1020) by tests/golden/omni_mask.npz;
* eager attention softmax(q k^T * scale + M) v with GQA (transformers
  models/qwen2/modeling_qwen2.py `eager_attention_forward` / `repeat_kv`; the reference selects
  the equivalent sdpa path, models/unigen.py:67);
* the Qwen2 decoder stack as `Qwen2Model.forward(inputs_embeds=, attention_mask=4-D)` computes it
  (third-party: transformers==4.51.0 pinned by scripts/requirements.txt:17, 5.5.0 installed):
  RMSNorm -> q/k/v proj (+bias) -> RoPE(theta, positions arange(L)) -> attention -> o_proj ->
  residual -> RMSNorm -> SwiGLU -> residual; final RMSNorm; tied LM head restricted to the
  codebook rows [text_vocab, text_vocab + V).  Pinned against transformers' own Qwen2Model in
  tests/test_backbone_oracle.py and against the unmodified reference UniGen in
  tests/golden/backbone_tiny.npz.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def omni_allowed(pad_len: torch.Tensor, soi_idx: torch.Tensor, L: int) -> torch.Tensor:
    """bool [B,L,L]: True where query row i may attend to column j."""
    dev = pad_len.device
    i = torch.arange(L, device=dev)[None, :, None]
    j = torch.arange(L, device=dev)[None, None, :]
    P = pad_len.long()[:, None, None]
    S = soi_idx.long()[:, None, None]
    pad_rows = (i < P) & (j <= i)
    text_rows = (i >= P) & (i < S) & (j >= P) & (j <= i)
    img_rows = (i >= S) & (j >= P)
    return pad_rows | text_rows | img_rows


def omni_additive_mask(pad_len, soi_idx, L, dtype=torch.float32) -> torch.Tensor:
    """[B,1,L,L] additive mask with the reference's fill value (-9.223372e18, prompting_utils.py:1013-1018)."""
    allowed = omni_allowed(pad_len, soi_idx, L)
    m = torch.zeros(allowed.shape, dtype=torch.float32, device=allowed.device)
    m.masked_fill_(~allowed, float(torch.iinfo(torch.int64).min))
    return m[:, None].to(dtype)


def attention_ref(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, pad_len, soi_idx, scale: float) -> torch.Tensor:
    """q [B,L,Hq,D], k/v [B,L,Hkv,D] (any float dtype) -> out [B,L,Hq,D] float32, computed in fp32."""
    B, L, Hq, D = q.shape
    Hkv = k.shape[2]
    rep = Hq // Hkv
    qf = q.float().permute(0, 2, 1, 3)
    kf = k.float().permute(0, 2, 1, 3).repeat_interleave(rep, dim=1)
    vf = v.float().permute(0, 2, 1, 3).repeat_interleave(rep, dim=1)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    s = s + omni_additive_mask(pad_len, soi_idx, L).to(s.device)
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, vf).permute(0, 2, 1, 3).contiguous()


# --------------------------------------------------------------------------------------------- Qwen2
def rms_norm(x, w, eps):
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    return w * (xf * torch.rsqrt(var + eps)).to(x.dtype)


def rope_cos_sin(L: int, D: int, theta: float, device, dtype=torch.float32):
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.float32, device=device) / D))
    pos = torch.arange(L, dtype=torch.float32, device=device)
    fr = torch.outer(pos, inv)
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def apply_rope(x, cos, sin):
    """x [B,L,H,D]; rotate_half convention of transformers' apply_rotary_pos_emb."""
    D = x.shape[-1]
    x1, x2 = x[..., : D // 2], x[..., D // 2:]
    rot = torch.cat([-x2, x1], dim=-1)
    return x * cos[None, :, None, :] + rot * sin[None, :, None, :]


def qwen2_forward(weights: Dict[str, torch.Tensor], cfg: dict, embeds: torch.Tensor, pad_len, soi_idx) -> torch.Tensor:
    """hidden states after the final norm, [B,L,H] (dtype of `embeds`, fp32 recommended)."""
    B, L, H = embeds.shape
    Hq, Hkv, D = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"]
    eps = cfg["rms_norm_eps"]
    cos, sin = rope_cos_sin(L, D, cfg["rope_theta"], embeds.device, embeds.dtype)
    x = embeds
    for l in range(cfg["num_hidden_layers"]):
        pre = f"layers.{l}."
        h = rms_norm(x, weights[pre + "input_layernorm.weight"], eps)
        q = F.linear(h, weights[pre + "self_attn.q_proj.weight"], weights[pre + "self_attn.q_proj.bias"]).view(B, L, Hq, D)
        k = F.linear(h, weights[pre + "self_attn.k_proj.weight"], weights[pre + "self_attn.k_proj.bias"]).view(B, L, Hkv, D)
        v = F.linear(h, weights[pre + "self_attn.v_proj.weight"], weights[pre + "self_attn.v_proj.bias"]).view(B, L, Hkv, D)
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        a = attention_ref(q, k, v, pad_len, soi_idx, D ** -0.5).to(x.dtype).reshape(B, L, Hq * D)
        x = x + F.linear(a, weights[pre + "self_attn.o_proj.weight"])
        h = rms_norm(x, weights[pre + "post_attention_layernorm.weight"], eps)
        g = F.linear(h, weights[pre + "mlp.gate_proj.weight"])
        u = F.linear(h, weights[pre + "mlp.up_proj.weight"])
        x = x + F.linear(F.silu(g) * u, weights[pre + "mlp.down_proj.weight"])
    return rms_norm(x, weights["norm.weight"], eps)


def codebook_logits(weights, hidden_img: torch.Tensor, text_vocab: int, V: int) -> torch.Tensor:
    """logits[:, -(N+1):-1, text_vocab:-1] of the tied LM head (models/unigen.py:287, 416) without the other columns."""
    return F.linear(hidden_img, weights["embed_tokens.weight"][text_vocab:text_vocab + V])


def smoke_attention(dev) -> None:
    """One small omni-attention call on the GPU against attention_ref (used by __graft_entry__.smoke)."""
    from ml_unigen_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, L, Hq, Hkv, D = 2, 300, 12, 2, 128
    q = torch.randn(B, L, Hq, D, generator=g).bfloat16().to(dev)
    k = torch.randn(B, L, Hkv, D, generator=g).bfloat16().to(dev)
    v = torch.randn(B, L, Hkv, D, generator=g).bfloat16().to(dev)
    P = torch.tensor([0, 17], dtype=torch.int32, device=dev)
    S = torch.tensor([40, 41], dtype=torch.int32, device=dev)
    out = ops.omni_attention(q, k, v, P, S, D ** -0.5).float()
    ref = attention_ref(q, k, v, P, S, D ** -0.5)
    err = (out - ref).abs().max().item()
    assert err < 2e-2, f"omni attention max-abs error {err}"
    print(f"smoke: omni attention max-abs error {err:.3e} vs fp32 reference")


def smoke_gemm(dev) -> None:
    """One small fused gate|up + SwiGLU launch and one q/k/v + RoPE launch of the tcgen05 linear-layer kernel on the GPU
    against the fp32 formulas (used by __graft_entry__.smoke)."""
    from ml_unigen_b200 import ops
    g = torch.Generator().manual_seed(1)
    M, K, I = 300, 256, 384
    x = torch.randn(M, K, generator=g).bfloat16().to(dev)
    wg = (torch.randn(I, K, generator=g) * K ** -0.5).bfloat16().to(dev)
    wu = (torch.randn(I, K, generator=g) * K ** -0.5).bfloat16().to(dev)
    got = ops.gemm(x, ops.pack_gate_up(wg, wu), mode=ops.GEMM_SWIGLU).float()
    ref = torch.nn.functional.silu(x.float() @ wg.float().T) * (x.float() @ wu.float().T)
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 8e-3, f"fused SwiGLU GEMM relative L2 error {rel}"
    B, L, Hq, Hkv, D = 2, 150, 2, 1, 128
    w = (torch.randn((Hq + 2 * Hkv) * D, K, generator=g) * K ** -0.5).bfloat16().to(dev)
    b = torch.randn((Hq + 2 * Hkv) * D, generator=g).bfloat16().to(dev)
    cos, sin = rope_cos_sin(L, D, 1e6, dev)
    qkv = ops.qkv_rope(x, w, b, cos[:, : D // 2].contiguous(), sin[:, : D // 2].contiguous(), L, Hq, Hkv)
    full = torch.nn.functional.linear(x.float(), w.float(), b.float())
    ref_q = apply_rope(full[:, : Hq * D].view(B, L, Hq, D), cos, sin)
    rel_q = ((qkv[:, : Hq * D].view(B, L, Hq, D).float() - ref_q).norm() / ref_q.norm()).item()
    assert rel_q < 6e-3, f"fused q/k/v + RoPE relative L2 error {rel_q}"
    print(f"smoke: tcgen05 GEMM epilogues: SwiGLU rel-L2 {rel:.2e}, q/k/v + RoPE rel-L2 {rel_q:.2e} vs fp32 formulas")
    # residual epilogue (o_proj / down_proj + residual add + RMSNorm statistics) and the folded RMSNorm of the consumer GEMM
    N = 256
    wd = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().to(dev)
    x0 = torch.randn(M, N, generator=g).to(dev)
    xr = x0.clone()
    xb, ss = ops.gemm_residual_(x, wd, xr)
    want = x0 + (x.float() @ wd.float().T).bfloat16().float()
    e_x = ((xr - want).norm() / want.norm()).item()
    assert e_x < 2e-3 and torch.equal(xb, xr.bfloat16()), f"residual epilogue: rel-L2 {e_x}"
    e_ss = ((ss.sum(-1) - xr.pow(2).sum(-1)).abs() / xr.pow(2).sum(-1)).max().item()
    assert e_ss < 1e-5, f"residual epilogue: row statistics off by {e_ss}"
    gain = (1 + 0.1 * torch.randn(N, generator=g)).to(dev)
    wn = (torch.randn(256, N, generator=g) * N ** -0.5).to(dev)
    got = ops.gemm(xb, (wn * gain[None, :]).bfloat16(), row_ss=ss, rms_eps=1e-6).float()
    ref = rms_norm(xr, gain, 1e-6) @ wn.T
    rel_n = ((got - ref).norm() / ref.norm()).item()
    assert rel_n < 8e-3, f"folded RMSNorm GEMM relative L2 error {rel_n}"
    print(f"smoke: residual epilogue rel-L2 {e_x:.2e} (xb exact, row sums {e_ss:.1e}); folded RMSNorm + GEMM rel-L2 {rel_n:.2e}")
