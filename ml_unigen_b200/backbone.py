"""Qwen2.5 decoder stack as the T2I path runs it (transformers Qwen2Model.forward with
`inputs_embeds` + the 4-D omni mask, reached from apple/ml-unigen models/unigen.py:260-266),
laid out for one B200:

* weights resident in HBM as bf16: fused [q|k|v] projection (+bias), o_proj, fused [gate|up], down,
  RMSNorm gains, tied embedding / LM head (3.1 GB at 1.5 B parameters -- replicated per GPU);
* fp32 residual stream (what the reference's bf16-autocast mode keeps), bf16 activations between
  GEMMs, every elementwise chain fused into one hand-written row kernel (ops.add_rmsnorm, rope_qk_,
  swiglu, embed_rows);
* attention through the tcgen05 omni-attention kernel on strided views of the fused qkv activation --
  no repeat_kv, no dense mask;
* the LM head is applied only to the N image rows and only to the V codebook rows of the tied
  embedding (models/unigen.py:416 slices exactly that block out of the full-vocabulary logits),
  producing the dense bf16 [B,N,V] tensor the decode-step kernel streams;
* linear layers: gate_proj/up_proj + SwiGLU are ONE launch of the tcgen05 2-CTA kernel
  (`ops.gemm(..., mode=GEMM_SWIGLU)` on block-interleaved gate|up weights: the [M, 2I] activation never reaches HBM),
  the q/k/v projection + bias + RoPE + K/V-cache update are ONE launch of the same kernel (`ops.qkv_rope`), the codebook
  head runs on it too; o_proj and down_proj stay on cuBLAS (`F.linear`), which measured 3-4 % faster than the own
  kernel on those two shapes and has nothing to fuse (tools/bench_gemm.py, profiles/r1_gemm_bench.json).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ops


@dataclass
class QwenConfig:
    hidden_size: int = 1536
    intermediate_size: int = 8960
    num_hidden_layers: int = 28
    num_attention_heads: int = 12
    num_key_value_heads: int = 2
    head_dim: int = 128
    vocab_size: int = 159867          # len(tokenizer) + 8192 + 1 (evaluation/inference_geneval.py:104)
    rope_theta: float = 1000000.0
    rms_norm_eps: float = 1e-6

    def as_dict(self):
        return dict(self.__dict__)


class QwenBackboneB200:
    """Inference-only decoder stack.  `weights` uses transformers' Qwen2Model parameter names
    (`embed_tokens.weight`, `layers.{i}.self_attn.q_proj.weight`, ..., `norm.weight`)."""

    def __init__(self, cfg: QwenConfig, weights: Dict[str, torch.Tensor], device):
        if not torch.cuda.is_available():
            raise RuntimeError("QwenBackboneB200 needs a CUDA device (no CPU path)")
        self.cfg = cfg
        self.device = torch.device(device)
        bf = lambda t: t.detach().to(self.device, torch.bfloat16).contiguous()
        self.embed = bf(weights["embed_tokens.weight"])
        self.final_norm = bf(weights["norm.weight"])
        self.layers = []
        for l in range(cfg.num_hidden_layers):
            p = f"layers.{l}."
            self.layers.append(dict(
                ln1=bf(weights[p + "input_layernorm.weight"]),
                wqkv=bf(torch.cat([weights[p + "self_attn.q_proj.weight"], weights[p + "self_attn.k_proj.weight"],
                                   weights[p + "self_attn.v_proj.weight"]], 0)),
                bqkv=bf(torch.cat([weights[p + "self_attn.q_proj.bias"], weights[p + "self_attn.k_proj.bias"],
                                   weights[p + "self_attn.v_proj.bias"]], 0)),
                wo=bf(weights[p + "self_attn.o_proj.weight"]),
                ln2=bf(weights[p + "post_attention_layernorm.weight"]),
                wgu=ops.pack_gate_up(bf(weights[p + "mlp.gate_proj.weight"]), bf(weights[p + "mlp.up_proj.weight"])),
                wd=bf(weights[p + "mlp.down_proj.weight"]),
            ))
        self._rope_cache = {}
        self.own_plain_gemms = os.environ.get("UG_PLAIN_GEMM", "cublas") == "own"      # experiment switch, see DESIGN.md 4.5
        self.attn_events = None        # bench hook: list receiving (start, end) CUDA events of every attention launch
        self.mlp_events = None         # bench hook: the same for every fused gate|up + SwiGLU GEMM launch
        self.plain_events = None       # bench hook: the same for every o_proj / down_proj launch

    # -------------------------------------------------------------------------------------------
    @classmethod
    def random_init(cls, cfg: QwenConfig, device, seed: int = 0, std: float = 0.02) -> "QwenBackboneB200":
        """Random-init weights of the named architecture, generated directly on the GPU (bf16)."""
        g = torch.Generator(device=device).manual_seed(seed)
        H, I, D = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
        Hq, Hkv = cfg.num_attention_heads, cfg.num_key_value_heads
        rn = lambda *s: (torch.randn(*s, device=device, generator=g, dtype=torch.float32) * std).bfloat16()
        w = {"embed_tokens.weight": rn(cfg.vocab_size, H), "norm.weight": torch.ones(H, device=device)}
        for l in range(cfg.num_hidden_layers):
            p = f"layers.{l}."
            w[p + "input_layernorm.weight"] = torch.ones(H, device=device)
            w[p + "post_attention_layernorm.weight"] = torch.ones(H, device=device)
            w[p + "self_attn.q_proj.weight"] = rn(Hq * D, H)
            w[p + "self_attn.k_proj.weight"] = rn(Hkv * D, H)
            w[p + "self_attn.v_proj.weight"] = rn(Hkv * D, H)
            w[p + "self_attn.q_proj.bias"] = rn(Hq * D)
            w[p + "self_attn.k_proj.bias"] = rn(Hkv * D)
            w[p + "self_attn.v_proj.bias"] = rn(Hkv * D)
            w[p + "self_attn.o_proj.weight"] = rn(H, Hq * D)
            w[p + "mlp.gate_proj.weight"] = rn(I, H)
            w[p + "mlp.up_proj.weight"] = rn(I, H)
            w[p + "mlp.down_proj.weight"] = rn(H, I)
        return cls(cfg, w, device)

    @classmethod
    def from_qwen2_model(cls, qwen2_model, device) -> "QwenBackboneB200":
        """From a transformers `Qwen2Model` (e.g. the reference's `UniGen.llm.model`)."""
        c = qwen2_model.config
        cfg = QwenConfig(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                         num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                         num_key_value_heads=c.num_key_value_heads,
                         head_dim=getattr(c, "head_dim", None) or c.hidden_size // c.num_attention_heads,
                         vocab_size=c.vocab_size, rope_theta=_rope_theta(c), rms_norm_eps=c.rms_norm_eps)
        return cls(cfg, dict(qwen2_model.state_dict()), device)

    # -------------------------------------------------------------------------------------------
    def _rope(self, L: int):
        if L not in self._rope_cache:
            D = self.cfg.head_dim
            inv = 1.0 / (self.cfg.rope_theta ** (torch.arange(0, D, 2, dtype=torch.float32, device=self.device) / D))
            fr = torch.outer(torch.arange(L, dtype=torch.float32, device=self.device), inv)      # positions arange(L)
            self._rope_cache[L] = (fr.cos().contiguous(), fr.sin().contiguous())
        return self._rope_cache[L]

    def embed_tokens(self, token_ids: torch.Tensor) -> torch.Tensor:
        """int64 [B,L] -> fp32 residual stream [B,L,H]."""
        B, L = token_ids.shape
        return ops.embed_rows(self.embed, token_ids.contiguous()).view(B, L, self.cfg.hidden_size)

    def forward(self, x: torch.Tensor, pad_len: torch.Tensor, soi_idx: torch.Tensor, rows: Optional[slice] = None,
                kv_cache: Optional[list] = None) -> torch.Tensor:
        """x: fp32 [B,L,H] embeddings (consumed as the residual stream, modified in place).
        Returns the final-norm hidden states, bf16 [B,L,H] (or [B,rows,H] when `rows` is given).
        With `kv_cache` (from `new_kv_cache`) the post-RoPE K/V of every layer are kept for `forward_suffix`."""
        cfg = self.cfg
        B, L, H = x.shape
        Hq, Hkv, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        M = B * L
        x2 = x.view(M, H)
        cos, sin = self._rope(L)
        scale = D ** -0.5
        delta = None
        for li, ly in enumerate(self.layers):
            h = ops.add_rmsnorm(x2, ly["ln1"], cfg.rms_norm_eps, delta)
            qkv = ops.qkv_rope(h, ly["wqkv"], ly["bqkv"], cos, sin, L, Hq, Hkv,
                               kv_cache=None if kv_cache is None else kv_cache[li])
            q = qkv[:, : Hq * D].view(B, L, Hq, D)
            if kv_cache is not None:
                k, v = kv_cache[li][:, :, 0], kv_cache[li][:, :, 1]
            else:
                k = qkv[:, Hq * D: (Hq + Hkv) * D].view(B, L, Hkv, D)
                v = qkv[:, (Hq + Hkv) * D:].view(B, L, Hkv, D)
            a = self._attention(q, k, v, pad_len, soi_idx, scale, 0)
            o = self._plain_linear(a.view(M, Hq * D), ly["wo"])
            h = ops.add_rmsnorm(x2, ly["ln2"], cfg.rms_norm_eps, o)
            act = self._gate_up(h, ly["wgu"])
            delta = self._plain_linear(act, ly["wd"])
        hn = ops.add_rmsnorm(x2, self.final_norm, cfg.rms_norm_eps, delta).view(B, L, H)
        return hn if rows is None else hn[:, rows]

    def new_kv_cache(self, B: int, L: int) -> list:
        """Per-layer post-RoPE K/V, bf16 [B, L, 2, Hkv, D] (19 MB per layer at B=16, L=1163)."""
        cfg = self.cfg
        return [torch.empty(B, L, 2, cfg.num_key_value_heads, cfg.head_dim, dtype=torch.bfloat16, device=self.device)
                for _ in self.layers]

    def forward_suffix(self, x: torch.Tensor, pad_len: torch.Tensor, soi_idx: torch.Tensor, s0: int, kv_cache: list,
                       rows: Optional[slice] = None) -> torch.Tensor:
        """Text-prefix K/V reuse.  Rows < soi_idx never attend to the image block
        (training/prompting_utils.py:994-1011), so their hidden states and K/V are identical in every MaskGIT
        step.  `x` holds only the rows [s0, L) (s0 = the common soi_idx): fp32 [B, Ls, H]; `kv_cache` holds the
        K/V of all L rows from the step-0 `forward(..., kv_cache=...)`; its rows [s0, L) are refreshed here.
        Every linear layer runs on Ls = N+2 rows instead of L."""
        cfg = self.cfg
        B, Ls, H = x.shape
        L = kv_cache[0].shape[1]
        assert s0 + Ls == L
        Hq, Hkv, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        M = B * Ls
        x2 = x.view(M, H)
        cos, sin = self._rope(L)
        cos, sin = cos[s0:], sin[s0:]                       # positions s0 .. L-1 (contiguous row slices)
        if D != 128:
            raise ValueError("the fused q/k/v + RoPE kernel needs head_dim 128")
        scale = D ** -0.5
        delta = None
        for li, ly in enumerate(self.layers):
            h = ops.add_rmsnorm(x2, ly["ln1"], cfg.rms_norm_eps, delta)
            qkv = ops.qkv_rope(h, ly["wqkv"], ly["bqkv"], cos, sin, Ls, Hq, Hkv, kv_cache=kv_cache[li], cache_row_offset=s0)
            q = qkv[:, : Hq * D].view(B, Ls, Hq, D)
            k, v = kv_cache[li][:, :, 0], kv_cache[li][:, :, 1]
            a = self._attention(q, k, v, pad_len, soi_idx, scale, s0)
            o = self._plain_linear(a.view(M, Hq * D), ly["wo"])
            h = ops.add_rmsnorm(x2, ly["ln2"], cfg.rms_norm_eps, o)
            act = self._gate_up(h, ly["wgu"])
            delta = self._plain_linear(act, ly["wd"])
        hn = ops.add_rmsnorm(x2, self.final_norm, cfg.rms_norm_eps, delta).view(B, Ls, H)
        return hn if rows is None else hn[:, rows]

    def _plain_linear(self, a, w):
        """o_proj / down_proj: plain GEMMs with nothing to fuse."""
        ev = self.plain_events
        if ev is None:
            return ops.gemm(a, w) if self.own_plain_gemms else F.linear(a, w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ops.gemm(a, w) if self.own_plain_gemms else F.linear(a, w)
        e1.record()
        ev.append((e0, e1, a.shape[0], a.shape[1]))
        return out

    def _gate_up(self, h, wgu):
        ev = self.mlp_events
        if ev is None:
            return ops.gemm(h, wgu, mode=ops.GEMM_SWIGLU)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        act = ops.gemm(h, wgu, mode=ops.GEMM_SWIGLU)
        e1.record()
        ev.append((e0, e1, h.shape[0]))
        return act

    def _attention(self, q, k, v, pad_len, soi_idx, scale, q_offset):
        ev = self.attn_events
        if ev is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        a = ops.omni_attention(q, k, v, pad_len, soi_idx, scale, q_offset=q_offset)
        if ev is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            ev.append((e0, e1, q.shape[1]))
        return a

    def codebook_logits(self, hidden_img: torch.Tensor, text_vocab_size: int, V: int) -> torch.Tensor:
        """bf16 [B,N,H] -> bf16 [B,N,V]: tied LM head restricted to rows [text_vocab, text_vocab+V)."""
        B, N, H = hidden_img.shape
        h2 = hidden_img.reshape(B * N, H)
        return ops.gemm(h2, self.embed[text_vocab_size: text_vocab_size + V]).view(B, N, V)


def _rope_theta(c) -> float:
    t = getattr(c, "rope_theta", None)
    if t is None:
        rp = getattr(c, "rope_parameters", None) or {}
        t = rp.get("rope_theta", 1000000.0)
    return float(t)
