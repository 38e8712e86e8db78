"""MaskGIT sampling primitives with the reference's names and call signatures
(apple/ml-unigen models/sampling.py), backed by the CUDA kernels of libunigen_b200.so.

* `mask_by_random_topk(mask_len, probs, temperature=1.0, generator=None)` -- sampling.py:41-46.
  The confidence / k-th-smallest / compare work runs in `ug_mask_by_random_topk`; the uniform
  noise is drawn from `generator` exactly like the reference's `gumbel_noise` does
  (`zeros_like(probs).uniform_(0, 1, generator=generator)`, sampling.py:24-26), so the
  generator advances identically and results are comparable draw for draw.
* the mask schedules (`cosine_schedule`, `linear_schedule`, `pow`, `sigmoid_schedule`,
  `get_mask_chedule` -- sic, the reference's spelling) are host-side scalar functions of a 0-d
  tensor, evaluated in float32 exactly as sampling.py:49-88 evaluates them; `t2i_generate`
  receives one of them as its `noise_schedule` argument.
"""
from __future__ import annotations

import math
from functools import partial

import torch

from . import ops


def log(t, eps=1e-20):
    """sampling.py:20 -- safe log (host/torch helper kept for API parity; the kernels use their own)."""
    return torch.log(t.clamp(min=eps))


def gumbel_noise(t, generator=None):
    """sampling.py:24-26."""
    u = torch.zeros_like(t).uniform_(0, 1, generator=generator)
    return -log(-log(u))


def mask_by_random_topk(mask_len, probs, temperature=1.0, generator=None):
    """Drop-in for models/sampling.py:41.  probs float32 [B,N] on CUDA, mask_len [B,1] (any
    numeric dtype, truncated like `.long()`).  Returns a bool [B,N] tensor: True = re-mask."""
    if not probs.is_cuda:
        raise RuntimeError("ml_unigen_b200.sampling.mask_by_random_topk runs on CUDA only (no CPU fallback)")
    p32 = probs if probs.dtype == torch.float32 else probs.float()
    u = torch.zeros_like(p32).uniform_(0, 1, generator=generator)
    return ops.mask_by_random_topk(p32, mask_len.long().reshape(-1), float(temperature), u_noise=u)


# ---------------------------------------------------------------------------------- schedules
def cosine_schedule(t):
    return torch.cos(t * math.pi * 0.5)


def linear_schedule(t):
    return (1 - t).clamp(min=1e-6, max=1.0)


def pow(t, method):
    exponent = float(method.replace("pow", ""))
    return (1.0 - t ** exponent).clamp(min=1e-6, max=1.0)


def sigmoid_schedule(t, start=-3, end=3, tau=1.0, clip_min=1e-6):
    lo = torch.sigmoid(torch.tensor(start / tau))
    hi = torch.sigmoid(torch.tensor(end / tau))
    cur = torch.sigmoid((t * (end - start) + start) / tau)
    return torch.clip((hi - cur) / (hi - lo), clip_min, 1.0)


def get_mask_chedule(method, **schedule_kwargs):
    if method == "cosine":
        return cosine_schedule
    if method == "linear":
        return linear_schedule
    if "pow" in method:
        return partial(pow, method=method)
    if method == "sigmoid":
        return partial(sigmoid_schedule, **schedule_kwargs)
    raise ValueError("Unknown schedule method: {}".format(method))


get_mask_schedule = get_mask_chedule  # correctly spelled alias
