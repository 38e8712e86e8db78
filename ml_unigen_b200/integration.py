"""The bindings a maintainer of apple/ml-unigen adds to route the reference through libunigen_b200, as
executable code -- tests/test_integration_gpu.py runs both.

* `enable_b200_t2i(model)`      -- whole-path drop-in: `model.t2i_generate` (models/unigen.py:344-455) becomes
                                   `UniGenB200.t2i_generate`, same keyword arguments, same return value;
* `register_omni_attention()`   -- operator-level: the tcgen05 omni-attention kernel as a transformers attention function
                                   (`AttentionInterface`, selected by `config._attn_implementation`; the reference selects
                                   'sdpa', models/unigen.py:67), the reference's own loop and modules stay in charge.
"""
from __future__ import annotations

import torch

from . import ops, prompting
from .unigen import UniGenB200

OMNI_ATTENTION_NAME = "omni_b200"


def enable_b200_t2i(model, device=None, **kw):
    """Route `model.t2i_generate` of a reference `UniGen` through libunigen_b200 (sm_100a).  Returns the `UniGenB200` that now
    serves the call (bf16 copy of the Qwen2 weights, RMSNorm gains folded)."""
    device = device or next(model.parameters()).device
    fast = UniGenB200.from_reference(model, device=device, **kw)
    model.t2i_generate = fast.t2i_generate        # bound method with the signature of models/unigen.py:344-359
    return fast


def omni_b200(module, query, key, value, attention_mask, dropout=0.0, scaling=None, **kw):
    """transformers attention-function interface: query [B,Hq,L,D], key/value [B,Hkv,L,D] (no repeat_kv needed),
    attention_mask = the dense [B,1,L,L] omni mask of create_attention_mask_predict_next (training/prompting_utils.py:975-1020)
    or a (pad_len, soi_idx) tuple.  Returns (out [B,L,Hq,D], None)."""
    if isinstance(attention_mask, tuple):
        P, S = attention_mask
    else:
        P, S = prompting.omni_mask_params_from_dense(attention_mask)
    if scaling is None:
        scaling = query.shape[-1] ** -0.5
    q, k, v = (t.transpose(1, 2).to(torch.bfloat16) for t in (query, key, value))     # strided views: consumed without a copy
    if k.stride() != v.stride():      # transformers hands over a rotated (new, [B,H,L,D]-contiguous) key and a projection view as value;
        k, v = k.contiguous(), v.contiguous()          # the C ABI takes ONE set of k/v strides
    out = ops.omni_attention(q, k, v, P.contiguous(), S.contiguous(), float(scaling))
    return out.to(query.dtype), None


def register_omni_attention(name: str = OMNI_ATTENTION_NAME) -> str:
    from transformers import AttentionInterface
    AttentionInterface.register(name, omni_b200)
    return name
