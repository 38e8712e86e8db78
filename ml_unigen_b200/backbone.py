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
* linear layers: ALL five GEMM shapes of a layer and the codebook head run on the hand-written
  tcgen05 2-CTA kernel, each with the elementwise work that follows it in the reference fused into its epilogue:
    q/k/v + bias + RoPE + K/V-cache update  (`ops.qkv_rope`),
    gate|up + SwiGLU                          (`ops.gemm(..., mode=GEMM_SWIGLU)`: the [M, 2I] activation never reaches HBM),
    o_proj / down_proj + residual add + the statistics of the following RMSNorm (`ops.gemm_residual_`);
* RMSNorm folding: Qwen2RMSNorm commutes with the linear layer behind it, (g x / rms) W^T = (1/rms) (x (W diag g)^T).  The gains
  are folded into the q/k/v, gate|up and head weights once at load time (from the fp32 product, one bf16 rounding), the A
  operand of those GEMMs is bf16(x) as the residual epilogue left it, and 1/rms scales the accumulator rows.  A decoder layer
  is 5 launches; no add+RMSNorm pass over HBM and no cuBLAS call remain in the loop (round 1: 7 launches, two of them cuBLAS).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

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


class Residual:
    """The residual stream between layers: x fp32 [M,H] (what the reference's autocast mode keeps), xb = bf16(x) and
    ss fp32 [M, H/128] = per-128-column sums of x^2 -- the last two are what the norm-folded GEMMs consume."""
    __slots__ = ("x", "xb", "ss", "B", "L")

    def __init__(self, x, xb, ss, B, L):
        self.x, self.xb, self.ss, self.B, self.L = x, xb, ss, B, L


class Hidden:
    """Pre-norm hidden states of selected rows, ready for a norm-folded head GEMM: xb bf16 [B,R,H], ss fp32 [B,R,H/128]."""
    __slots__ = ("xb", "ss")

    def __init__(self, xb, ss):
        self.xb, self.ss = xb, ss


class QwenBackboneB200:
    """Inference-only decoder stack.  `weights` uses transformers' Qwen2Model parameter names
    (`embed_tokens.weight`, `layers.{i}.self_attn.q_proj.weight`, ..., `norm.weight`)."""

    def __init__(self, cfg: QwenConfig, weights: Dict[str, torch.Tensor], device, lm_head_weight: Optional[torch.Tensor] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("QwenBackboneB200 needs a CUDA device (no CPU path)")
        if cfg.hidden_size % 256 or cfg.head_dim != 128:
            raise ValueError("QwenBackboneB200 needs hidden_size % 256 == 0 and head_dim 128 (tile shapes of the fused epilogues)")
        self.cfg = cfg
        self.device = torch.device(device)
        dev = self.device
        bf = lambda t: t.detach().to(dev, torch.bfloat16).contiguous()
        # norm gain folded into the weight columns: the product is formed in fp32 from the weights as given and rounded once
        fold = lambda w, g: (w.detach().to(dev, torch.float32) * g.detach().to(dev, torch.float32)[None, :]).to(torch.bfloat16).contiguous()
        self.embed = bf(weights["embed_tokens.weight"])
        # the LM head: tied to the embedding unless the caller hands in an untied lm_head.weight (models/unigen.py:287)
        self._head_src = weights["embed_tokens.weight"] if lm_head_weight is None else lm_head_weight
        self._final_gain = weights["norm.weight"]
        self.final_norm = bf(weights["norm.weight"])
        self._head_cache = {}
        self.layers = []
        for l in range(cfg.num_hidden_layers):
            p = f"layers.{l}."
            ln1, ln2 = weights[p + "input_layernorm.weight"], weights[p + "post_attention_layernorm.weight"]
            wqkv = torch.cat([weights[p + "self_attn.q_proj.weight"], weights[p + "self_attn.k_proj.weight"],
                              weights[p + "self_attn.v_proj.weight"]], 0)
            self.layers.append(dict(
                ln1=bf(ln1), ln2=bf(ln2),                                  # kept for the unfused yardstick (ops.add_rmsnorm)
                wqkv=fold(wqkv, ln1),
                bqkv=bf(torch.cat([weights[p + "self_attn.q_proj.bias"], weights[p + "self_attn.k_proj.bias"],
                                   weights[p + "self_attn.v_proj.bias"]], 0)),
                wo=bf(weights[p + "self_attn.o_proj.weight"]),
                wgu=ops.pack_gate_up(fold(weights[p + "mlp.gate_proj.weight"], ln2), fold(weights[p + "mlp.up_proj.weight"], ln2)),
                wd=bf(weights[p + "mlp.down_proj.weight"]),
            ))
        self._rope_cache = {}
        self.attn_events = None        # bench hook: list receiving (start, end) CUDA events of every attention launch
        self.mlp_events = None         # bench hook: the same for every fused gate|up + SwiGLU GEMM launch
        self.plain_events = None       # bench hook: the same for every o_proj / down_proj (+ residual) launch
        self.launch_count = 0          # own kernel launches issued through this object (bench.py: gpu_launches)

    # -------------------------------------------------------------------------------------------
    @staticmethod
    def random_weights(cfg: QwenConfig, device, seed: int = 0, std: float = 0.02, gain_std: float = 0.0) -> Dict[str, torch.Tensor]:
        """Random-init weights of the named architecture (transformers' Qwen2Model parameter names), generated directly on the
        GPU as bf16 values; `gain_std` > 0 perturbs the RMSNorm gains around 1 (tests of the gain folding)."""
        g = torch.Generator(device=device).manual_seed(seed)
        H, I, D = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
        Hq, Hkv = cfg.num_attention_heads, cfg.num_key_value_heads
        rn = lambda *s: (torch.randn(*s, device=device, generator=g, dtype=torch.float32) * std).bfloat16()
        gain = lambda: torch.ones(H, device=device) if gain_std == 0 else \
            (1.0 + gain_std * torch.randn(H, device=device, generator=g)).bfloat16().float()
        w = {"embed_tokens.weight": rn(cfg.vocab_size, H), "norm.weight": gain()}
        for l in range(cfg.num_hidden_layers):
            p = f"layers.{l}."
            w[p + "input_layernorm.weight"] = gain()
            w[p + "post_attention_layernorm.weight"] = gain()
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
        return w

    @classmethod
    def random_init(cls, cfg: QwenConfig, device, seed: int = 0, std: float = 0.02) -> "QwenBackboneB200":
        return cls(cfg, cls.random_weights(cfg, device, seed, std), device)

    @classmethod
    def from_qwen2_model(cls, qwen2_model, device, lm_head_weight: Optional[torch.Tensor] = None) -> "QwenBackboneB200":
        """From a transformers `Qwen2Model` (e.g. the reference's `UniGen.llm.model`); `lm_head_weight` when the LM head is
        not tied to the embedding."""
        c = qwen2_model.config
        cfg = QwenConfig(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                         num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                         num_key_value_heads=c.num_key_value_heads,
                         head_dim=getattr(c, "head_dim", None) or c.hidden_size // c.num_attention_heads,
                         vocab_size=c.vocab_size, rope_theta=_rope_theta(c), rms_norm_eps=c.rms_norm_eps)
        return cls(cfg, dict(qwen2_model.state_dict()), device, lm_head_weight=lm_head_weight)

    # -------------------------------------------------------------------------------------------
    def _rope(self, L: int):
        if L not in self._rope_cache:
            D = self.cfg.head_dim
            inv = 1.0 / (self.cfg.rope_theta ** (torch.arange(0, D, 2, dtype=torch.float32, device=self.device) / D))
            fr = torch.outer(torch.arange(L, dtype=torch.float32, device=self.device), inv)      # positions arange(L)
            self._rope_cache[L] = (fr.cos().contiguous(), fr.sin().contiguous())
        return self._rope_cache[L]

    def embed_tokens(self, token_ids: torch.Tensor) -> Residual:
        """int64 [B,L] -> the residual stream of the first layer (fp32 x + its bf16 copy and row statistics), one launch."""
        B, L = token_ids.shape
        x, xb, ss = ops.embed_rows(self.embed, token_ids.contiguous(), with_norm_operands=True)
        self.launch_count += 1
        return Residual(x, xb, ss, B, L)

    def residual_from_embeddings(self, emb: torch.Tensor) -> Residual:
        """Caller-provided input embeddings [B,L,H] (t2i_generate's `input_embeddings`, models/unigen.py:370) -> Residual."""
        B, L, H = emb.shape
        x = emb.detach().to(self.device, torch.float32).reshape(B * L, H).contiguous().clone()
        xb, ss = ops.norm_operands(x)
        return Residual(x, xb, ss, B, L)

    def new_kv_cache(self, B: int, L: int) -> list:
        """Per-layer post-RoPE K/V, bf16 [B, L, 2, Hkv, D] (19 MB per layer at B=16, L=1163)."""
        cfg = self.cfg
        return [torch.empty(B, L, 2, cfg.num_key_value_heads, cfg.head_dim, dtype=torch.bfloat16, device=self.device)
                for _ in self.layers]

    def _layers(self, st: Residual, pad_len, soi_idx, cos, sin, q_offset: int, kv_cache: Optional[list]):
        """The decoder layers on the rows held by `st` (row r of a sequence = sequence row q_offset + r).  Five launches per
        layer, all own kernels; `st` is updated in place."""
        cfg = self.cfg
        B, Lr = st.B, st.L
        Hq, Hkv, D = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        M = B * Lr
        eps = cfg.rms_norm_eps
        scale = D ** -0.5
        for li, ly in enumerate(self.layers):
            cache = None if kv_cache is None else kv_cache[li]
            qkv = ops.qkv_rope(st.xb, ly["wqkv"], ly["bqkv"], cos, sin, Lr, Hq, Hkv, kv_cache=cache, cache_row_offset=q_offset,
                               row_ss=st.ss, rms_eps=eps)
            q = qkv[:, : Hq * D].view(B, Lr, Hq, D)
            if cache is not None:
                k, v = cache[:, :, 0], cache[:, :, 1]
            else:
                k = qkv[:, Hq * D: (Hq + Hkv) * D].view(B, Lr, Hkv, D)
                v = qkv[:, (Hq + Hkv) * D:].view(B, Lr, Hkv, D)
            a = self._attention(q, k, v, pad_len, soi_idx, scale, q_offset)
            self._residual_linear(a.view(M, Hq * D), ly["wo"], st)
            act = self._gate_up(st, ly["wgu"])
            self._residual_linear(act, ly["wd"], st)
        self.launch_count += 5 * len(self.layers)

    def forward(self, st: Residual, pad_len: torch.Tensor, soi_idx: torch.Tensor, rows: Optional[slice] = None,
                kv_cache: Optional[list] = None) -> Hidden:
        """st: the residual stream of all L rows (from `embed_tokens`; consumed and modified in place).
        Returns the pre-norm hidden states of `rows` as a `Hidden` (the final RMSNorm is folded into the head GEMM,
        `codebook_logits`; `final_hidden` gives the normalised activations themselves).
        With `kv_cache` (from `new_kv_cache`) the post-RoPE K/V of every layer are kept for `forward_suffix`."""
        cos, sin = self._rope(st.L)
        self._layers(st, pad_len, soi_idx, cos, sin, 0, kv_cache)
        return self._select(st, rows)

    def forward_suffix(self, st: Residual, pad_len: torch.Tensor, soi_idx: torch.Tensor, s0: int, kv_cache: list,
                       rows: Optional[slice] = None) -> Hidden:
        """Text-prefix K/V reuse.  Rows < soi_idx never attend to the image block
        (training/prompting_utils.py:994-1011), so their hidden states and K/V are identical in every MaskGIT
        step.  `st` holds only the rows [s0, L) (s0 = the common soi_idx); `kv_cache` holds the
        K/V of all L rows from the step-0 `forward(..., kv_cache=...)`; its rows [s0, L) are refreshed here.
        Every linear layer runs on Ls = N+2 rows instead of L."""
        L = kv_cache[0].shape[1]
        assert s0 + st.L == L
        cos, sin = self._rope(L)
        self._layers(st, pad_len, soi_idx, cos[s0:], sin[s0:], s0, kv_cache)      # positions s0 .. L-1 (contiguous row slices)
        return self._select(st, rows)

    def _select(self, st: Residual, rows: Optional[slice]) -> Hidden:
        H = self.cfg.hidden_size
        xb, ss = st.xb.view(st.B, st.L, H), st.ss.view(st.B, st.L, H // 128)
        return Hidden(xb, ss) if rows is None else Hidden(xb[:, rows], ss[:, rows])

    def final_hidden(self, st: Residual, rows: Optional[slice] = None) -> torch.Tensor:
        """`norm(x)` itself, bf16 [B,L,H] (or the selected rows): the unfused yardstick of the folded head."""
        hn = ops.add_rmsnorm(st.x, self.final_norm, self.cfg.rms_norm_eps).view(st.B, st.L, self.cfg.hidden_size)
        return hn if rows is None else hn[:, rows]

    def _residual_linear(self, a, w, st: Residual):
        """o_proj / down_proj + residual add + statistics of the next RMSNorm: one launch."""
        ev = self.plain_events
        if ev is None:
            ops.gemm_residual_(a, w, st.x, st.xb, st.ss)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm_residual_(a, w, st.x, st.xb, st.ss)
        e1.record()
        ev.append((e0, e1, a.shape[0], a.shape[1]))

    def _gate_up(self, st: Residual, wgu):
        ev = self.mlp_events
        if ev is None:
            return ops.gemm(st.xb, wgu, mode=ops.GEMM_SWIGLU, row_ss=st.ss, rms_eps=self.cfg.rms_norm_eps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        act = ops.gemm(st.xb, wgu, mode=ops.GEMM_SWIGLU, row_ss=st.ss, rms_eps=self.cfg.rms_norm_eps)
        e1.record()
        ev.append((e0, e1, st.xb.shape[0]))
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

    def _head_weight(self, text_vocab_size: int, V: int) -> torch.Tensor:
        """Rows [text_vocab, text_vocab + V) of the LM head with the final norm's gain folded into the columns (25 MB)."""
        key = (text_vocab_size, V)
        if key not in self._head_cache:
            w = self._head_src[text_vocab_size: text_vocab_size + V].detach().to(self.device, torch.float32)
            g = self._final_gain.detach().to(self.device, torch.float32)
            self._head_cache[key] = (w * g[None, :]).to(torch.bfloat16).contiguous()
        return self._head_cache[key]

    def codebook_logits(self, hidden: Hidden, text_vocab_size: int, V: int) -> torch.Tensor:
        """Hidden (pre-norm rows) -> bf16 [B,R,V]: final RMSNorm + the LM head restricted to rows [text_vocab, text_vocab+V)
        (models/unigen.py:287, 416), one norm-folded GEMM."""
        B, R, H = hidden.xb.shape
        xb = hidden.xb.reshape(B * R, H)
        ss = hidden.ss.reshape(B * R, H // 128).contiguous()
        self.launch_count += 1
        return ops.gemm(xb, self._head_weight(text_vocab_size, V), row_ss=ss, rms_eps=self.cfg.rms_norm_eps).view(B, R, V)


def _rope_theta(c) -> float:
    t = getattr(c, "rope_theta", None)
    if t is None:
        rp = getattr(c, "rope_parameters", None) or {}
        t = rp.get("rope_theta", 1000000.0)
    return float(t)
