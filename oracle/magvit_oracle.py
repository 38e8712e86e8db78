"""oracle/magvit_oracle.py -- TEST INFRASTRUCTURE ONLY.

Plain torch fp32 restatement of `MAGVITv2.decode_code` (apple/ml-unigen
models/multimodal_encoder/magvitv2.py:217-230 `LFQuantizer.get_codebook_entry`, :286-408 `VQGANDecoder`, :438-442;
common_modules.py:19-43 swish / GroupNorm(32, eps 1e-6) / nearest-2x Upsample + conv, :171-215 AttnBlock, :301-360
ResnetBlock), functional over a state dict with the reference's parameter names.  Pinned on
tests/golden/magvit_tiny.npz, which is recorded by executing the reference modules themselves
(oracle/make_golden_magvit.py)."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F


def lfq_codebook_entry(ids: torch.Tensor, codebook_dim: int = 13) -> torch.Tensor:
    """int64 [B,N] -> fp32 [B,codebook_dim,h,w], entries +-1: channel c is bit (codebook_dim-1-c) of the index
    (magvitv2.py:194-201, 217-230)."""
    B, N = ids.shape
    h = w = int(round(N ** 0.5))
    assert h * w == N
    shifts = torch.arange(codebook_dim - 1, -1, -1, dtype=torch.long, device=ids.device)
    bits = (ids.reshape(-1, 1) >> shifts) & 1
    z = bits.float() * 2 - 1
    return z.view(B, h, w, codebook_dim).permute(0, 3, 1, 2).contiguous()


def _gn_swish(x, w, b, swish=True):
    y = F.group_norm(x, 32, w, b, eps=1e-6)
    return y * torch.sigmoid(y) if swish else y


def _resblock(sd, p, x):
    h = _gn_swish(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    h = F.conv2d(h, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    h = _gn_swish(h, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    h = F.conv2d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if p + "nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
    return x + h


def _attnblock(sd, p, x):
    h = _gn_swish(x, sd[p + "norm.weight"], sd[p + "norm.bias"], swish=False)
    q = F.conv2d(h, sd[p + "q.weight"], sd[p + "q.bias"])
    k = F.conv2d(h, sd[p + "k.weight"], sd[p + "k.bias"])
    v = F.conv2d(h, sd[p + "v.weight"], sd[p + "v.bias"])
    B, C, H, W = q.shape
    q = q.reshape(B, C, H * W).permute(0, 2, 1)
    k = k.reshape(B, C, H * W)
    a = torch.softmax(torch.bmm(q, k) * (int(C) ** -0.5), dim=2)
    o = torch.bmm(v.reshape(B, C, H * W), a.permute(0, 2, 1)).reshape(B, C, H, W)
    return x + F.conv2d(o, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def decoder_forward(sd: Dict[str, torch.Tensor], num_res_blocks: List[int], z: torch.Tensor) -> torch.Tensor:
    """VQGANDecoder.forward (magvitv2.py:371-408): z fp32 [B,13,h,w] -> [B,3,H,W]."""
    n_levels = len(num_res_blocks)
    h = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(h, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    h = _resblock(sd, "mid.block_1.", h)
    h = _attnblock(sd, "mid.attn_1.", h)
    h = _resblock(sd, "mid.block_2.", h)
    for lvl in reversed(range(n_levels)):
        for blk in range(num_res_blocks[lvl]):
            h = _resblock(sd, f"up.{lvl}.block.{blk}.", h)
            if f"up.{lvl}.attn.{blk}.norm.weight" in sd:
                h = _attnblock(sd, f"up.{lvl}.attn.{blk}.", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"up.{lvl}.upsample.conv.weight"], sd[f"up.{lvl}.upsample.conv.bias"], padding=1)
    h = _gn_swish(h, sd["norm_out.weight"], sd["norm_out.bias"])
    return F.conv2d(h, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)


def decode_code(sd, num_res_blocks, ids: torch.Tensor) -> torch.Tensor:
    """MAGVITv2.decode_code (magvitv2.py:438-442)."""
    return decoder_forward(sd, num_res_blocks, lfq_codebook_entry(ids))
