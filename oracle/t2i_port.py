"""oracle/t2i_port.py -- TEST INFRASTRUCTURE ONLY (and the CPU arm of bench.py, which may execute oracle/).

CPU restatement ("port") of the whole reference path `UniGen.t2i_generate` (apple/ml-unigen models/unigen.py:344-455) as the
reference executes it on a host without a GPU:

  per MaskGIT step   embed_tokens of the current ids (:399-400, 451-452)
                  -> Qwen2Model.forward in fp32 with the dense omni mask (oracle/backbone_oracle.qwen2_forward; :260-266, 402)
                  -> the FULL-vocabulary LM head the reference computes (:287), then the slice [:, -(N+1):-1, text_vocab:-1] (:416)
                  -> one decode step (oracle/decode_step_oracle.c: softmax, race, confidence, mask_by_random_topk; :418-453)
  with the host-side schedule arithmetic of :425-426, 436, 444.

Pinned against the UNMODIFIED reference end to end by tests/golden/t2i_tiny.npz (oracle/make_golden_backbone.py `make_t2i_tiny`:
`UniGen.t2i_generate` itself on a tiny random-init Qwen2, noise captured by generator replay): final ids and every step's mask equal.
CFG (guidance_scale > 1) is not restated here -- the timed CPU configuration (BASELINE.json configs[0]) runs without it.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import backbone_oracle as BO
from . import oracle as O


def t2i_generate_port(weights: Dict[str, torch.Tensor], cfg: dict, input_ids: torch.Tensor, *, image_token_num_per_image: int,
                      codebook_size: int, text_vocab_size: int, mask_token_id: int, pad_id: int, soi_id: int,
                      timesteps: int = 18, temperature: float = 1.0, seed: int = 0, q_noise=None, u_noise=None,
                      max_steps: Optional[int] = None, on_step: Optional[Callable] = None, full_vocab_head: bool = True):
    """Returns (sampled ids int64 [B,N] of the last executed step, per-step records).  `q_noise` / `u_noise`: optional lists of
    per-step injected noise (parity with the reference's generator); otherwise the oracle's Philox stream (seed, offset=step).
    `max_steps` stops early (bounded timing samples); `on_step(step, seconds)` is called after every step."""
    N, V = image_token_num_per_image, codebook_size
    B, L = input_ids.shape
    emb = weights["embed_tokens.weight"]
    ids_img = input_ids[:, -(N + 1):-1].numpy().astype(np.int64).copy()                    # unigen.py:367
    tokens = input_ids.clone()
    # pad_len / soi_idx of the closed-form omni mask (training/prompting_utils.py:975-1020)
    pos = torch.arange(1, L + 1)
    P = ((tokens == pad_id) * pos).amax(dim=1).to(torch.int32)
    S = (tokens == soi_id).float().argmax(dim=1).to(torch.int32)
    sched = O.step_schedule(timesteps, N, temperature)
    recs = []
    sampled = None
    with torch.no_grad():
        for step in range(timesteps if max_steps is None else min(timesteps, max_steps)):
            t0 = time.perf_counter()
            img_tok = np.where(ids_img == mask_token_id, ids_img, ids_img + text_vocab_size)   # unigen.py:451
            tokens[:, -(N + 1):-1] = torch.from_numpy(img_tok)
            hid = BO.qwen2_forward(weights, cfg, emb[tokens], P, S)                           # [B,L,H] fp32
            if full_vocab_head:
                logits = torch.nn.functional.linear(hid, emb)                                 # [B,L,vocab]: unigen.py:287
                img = logits[:, -(N + 1):-1, text_vocab_size:text_vocab_size + V]            # unigen.py:416 (":-1" == ":+V")
            else:
                img = BO.codebook_logits(weights, hid[:, -(N + 1):-1], text_vocab_size, V)
            cap, temp = sched[step]
            r = O.decode_step(O.f32_to_bf16_bits(img.contiguous().numpy()), ids_img, mask_token_id=mask_token_id, mask_len_cap=cap,
                              temperature=temp, seed=seed, offset=step,
                              q_noise=None if q_noise is None else q_noise[step], u_noise=None if u_noise is None else u_noise[step])
            ids_img, sampled = r["ids"], r["sampled"]
            dt = time.perf_counter() - t0
            recs.append(dict(mask=r["mask"].astype(bool), sampled=sampled.copy(), seconds=dt))
            if on_step is not None:
                on_step(step, dt)
    return sampled, recs


def random_qwen_weights(cfg: dict, vocab_size: int, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """fp32 random-init weights of the named architecture on the host (BASELINE.json configs[0]: "random-init, fp32 on host CPU")."""
    g = torch.Generator().manual_seed(seed)
    H, I, D = cfg["hidden_size"], cfg["intermediate_size"], cfg["head_dim"]
    Hq, Hkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    rn = lambda *s: torch.randn(*s, generator=g) * std
    w = {"embed_tokens.weight": rn(vocab_size, H), "norm.weight": torch.ones(H)}
    for l in range(cfg["num_hidden_layers"]):
        p = f"layers.{l}."
        w[p + "input_layernorm.weight"] = torch.ones(H)
        w[p + "post_attention_layernorm.weight"] = torch.ones(H)
        for name, shape in (("self_attn.q_proj.weight", (Hq * D, H)), ("self_attn.k_proj.weight", (Hkv * D, H)),
                            ("self_attn.v_proj.weight", (Hkv * D, H)), ("self_attn.q_proj.bias", (Hq * D,)),
                            ("self_attn.k_proj.bias", (Hkv * D,)), ("self_attn.v_proj.bias", (Hkv * D,)),
                            ("self_attn.o_proj.weight", (H, Hq * D)), ("mlp.gate_proj.weight", (I, H)),
                            ("mlp.up_proj.weight", (I, H)), ("mlp.down_proj.weight", (H, I))):
            w[p + name] = rn(*shape)
    return w
