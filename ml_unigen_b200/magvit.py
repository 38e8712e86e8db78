"""MAGViT-v2 pixel decoder for the token grids the T2I path produces:
`MAGVITv2.decode_code` of apple/ml-unigen (models/multimodal_encoder/magvitv2.py:438-442 =
`LFQuantizer.get_codebook_entry` :217-230 + `VQGANDecoder.forward` :371-408; blocks in common_modules.py:19-43,
171-215, 301-360), as the callers run it right after `t2i_generate` (evaluation/inference_geneval.py:266-267).

Layout for one B200: activations NHWC bf16 (`channels_last`), so that every convolution is a cuDNN implicit GEMM on
tensor cores without layout changes, and everything between the convolutions is one hand-written, single-purpose pass
over HBM (csrc/magvit_ops.cu): the LFQ bit unpack (no embedding table), GroupNorm(32)+swish in one statistics pass and one
apply pass (the reference: group_norm, sigmoid, mul = 5 passes), nearest-2x upsampling.  The convolutions and the single
mid-block attention (1024 tokens, one head of 512 channels: outside the D=128 omni-attention kernel) are library calls
(cuDNN / torch SDPA), like the plain cuBLAS GEMMs of the backbone.  Weights: the reference's `VQGANDecoder` state dict.
There is no CPU path."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import _lib, ops


def lfq_unpack(ids: torch.Tensor, dim: int = 13) -> torch.Tensor:
    """int64 [B,N] -> bf16 NHWC [B,h,w,dim] of +-1 (channel c = bit dim-1-c), h = w = sqrt(N)."""
    ops._need_cuda(ids)
    B, N = ids.shape
    h = int(round(N ** 0.5))
    if h * h != N or ids.dtype != torch.int64:
        raise ValueError("ids must be int64 [B, N] with N a square")
    ids = ids.contiguous()
    z = torch.empty(B, h, h, dim, dtype=torch.bfloat16, device=ids.device)
    _lib.check(_lib.lib().ug_lfq_unpack_bf16(ops._ptr(ids), ops._ptr(z), B * N, dim, ops._stream()))
    return z


class _GroupNormWorkspace:
    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def get(self, B, HW, C, device):
        n = _lib.lib().ug_groupnorm_workspace_floats(B, HW, C)
        if self.buf is None or self.buf.numel() < n or self.buf.device != device:
            self.buf = torch.empty(n, dtype=torch.float32, device=device)
        return self.buf


_gn_ws = _GroupNormWorkspace()


def groupnorm_swish(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6, swish: bool = True,
                    out: Optional[torch.Tensor] = None, in_bias: Optional[torch.Tensor] = None,
                    stats_ready: bool = False) -> torch.Tensor:
    """x bf16 NHWC [B,H,W,C] (contiguous) -> GroupNorm(32, eps)(x + in_bias)*gamma+beta, then y*sigmoid(y) if `swish`.
    `in_bias` (bf16 [C]): the bias of the convolution that produced x, folded into this pass.  `stats_ready`: x was just
    produced by `bias_add(..., emit_stats=True)`, whose partial sums are still in the shared workspace: no statistics pass."""
    ops._need_cuda(x, gamma, beta, in_bias)
    B, H, W, C = x.shape
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or gamma.dtype != torch.bfloat16 or beta.dtype != torch.bfloat16:
        raise ValueError("x must be contiguous bf16 NHWC, gamma/beta bf16 [C]")
    if out is None:
        out = torch.empty_like(x)
    ws = _gn_ws.get(B, H * W, C, x.device)
    if in_bias is not None and (in_bias.dtype != torch.bfloat16 or in_bias.numel() != C or not in_bias.is_contiguous()):
        raise ValueError("in_bias must be contiguous bf16 [C]")
    _lib.check(_lib.lib().ug_groupnorm_nhwc_bf16(ops._ptr(x), ops._ptr(in_bias), ops._ptr(gamma), ops._ptr(beta), ops._ptr(out), ops._ptr(ws),
                                                 B, H * W, C, float(eps), 1 if swish else 0, 1 if stats_ready else 0, ops._stream()))
    return out


def bias_add(a: torch.Tensor, bias_a: Optional[torch.Tensor], r: Optional[torch.Tensor] = None,
             bias_r: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, emit_stats: bool = False) -> torch.Tensor:
    """out = a + bias_a (+ r (+ bias_r)) on bf16 NHWC tensors [.., C] in one pass (default: in place on `a`).  With
    `emit_stats` (a must be [B,H,W,C]) the pass also leaves the GroupNorm partial sums of `out` in the shared workspace for
    an immediately following `groupnorm_swish(out, ..., stats_ready=True)`."""
    ops._need_cuda(a, bias_a, r, bias_r)
    C = a.shape[-1]
    if a.dtype != torch.bfloat16 or not a.is_contiguous() or (r is not None and (r.shape != a.shape or not r.is_contiguous()
                                                                                   or r.dtype != torch.bfloat16)):
        raise ValueError("a (and r) must be contiguous bf16 tensors of the same shape")
    for b in (bias_a, bias_r):
        if b is not None and (b.dtype != torch.bfloat16 or b.numel() != C or not b.is_contiguous()):
            raise ValueError("biases must be contiguous bf16 [C]")
    if out is None:
        out = a
    if emit_stats:
        B, H, W, _ = a.shape
        ws = _gn_ws.get(B, H * W, C, a.device)
        _lib.check(_lib.lib().ug_bias_add_stats_nhwc_bf16(ops._ptr(a), ops._ptr(bias_a), ops._ptr(r), ops._ptr(bias_r), ops._ptr(out),
                                                          ops._ptr(ws), B, H * W, C, ops._stream()))
        return out
    _lib.check(_lib.lib().ug_bias_add_nhwc_bf16(ops._ptr(a), ops._ptr(bias_a), ops._ptr(r), ops._ptr(bias_r), ops._ptr(out),
                                                a.numel() // C, C, ops._stream()))
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    """bf16 NHWC [B,H,W,C] -> [B,2H,2W,C], nearest neighbour."""
    ops._need_cuda(x)
    B, H, W, C = x.shape
    if x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise ValueError("x must be contiguous bf16 NHWC")
    y = torch.empty(B, 2 * H, 2 * W, C, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().ug_upsample2x_nhwc_bf16(ops._ptr(x), ops._ptr(y), B, H, W, C, ops._stream()))
    return y


def _conv(x_nhwc: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], padding: int) -> torch.Tensor:
    """cuDNN convolution on an NHWC bf16 activation; returns NHWC (contiguous).  The big convolutions are called with
    b = None: torch adds a convolution bias in a separate, non-vectorised broadcast kernel (326 us per convolution at
    16 x 512 x 512 x 128 -- a third of the decoder before it was folded into the consuming kernels)."""
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w, b, padding=padding)        # NCHW view of channels_last memory
    return y.permute(0, 2, 3, 1).contiguous()                                # no copy when cuDNN kept channels_last


class MagvitV2DecoderB200:
    """`VQGANDecoder` (default: ch=128, ch_mult=[1,1,2,2,4], num_res_blocks=[4,4,3,4,3], z_channels=13) + LFQ lookup.
    `weights`: the reference decoder's state dict (`conv_in.weight`, `mid.block_1.norm1.weight`, ...,
    `up.{l}.block.{i}.conv1.weight`, `up.{l}.upsample.conv.weight`, `norm_out.weight`, `conv_out.weight`,
    `post_quant_conv.weight`)."""

    def __init__(self, weights: Dict[str, torch.Tensor], num_res_blocks: Sequence[int] = (4, 4, 3, 4, 3), device="cuda",
                 codebook_dim: int = 13):
        if not torch.cuda.is_available():
            raise RuntimeError("MagvitV2DecoderB200 needs a CUDA device (no CPU path)")
        self.device = torch.device(device)
        self.num_res_blocks = list(num_res_blocks)
        self.codebook_dim = codebook_dim
        self.w: Dict[str, torch.Tensor] = {}
        for k, v in weights.items():
            t = v.detach().to(self.device, torch.bfloat16)
            if t.dim() == 4:
                t = t.contiguous(memory_format=torch.channels_last)
            self.w[k] = t.contiguous() if t.dim() != 4 else t

    @classmethod
    def random_init(cls, device, seed: int = 0, ch: int = 128, ch_mult: Sequence[int] = (1, 1, 2, 2, 4),
                    num_res_blocks: Sequence[int] = (4, 4, 3, 4, 3), z_channels: int = 13) -> "MagvitV2DecoderB200":
        """Random weights of the named architecture (PyTorch's default conv init scale), generated on the GPU."""
        g = torch.Generator(device=device).manual_seed(seed)
        w: Dict[str, torch.Tensor] = {}

        def conv(name, cout, cin, k):
            bound = (cin * k * k) ** -0.5
            w[name + ".weight"] = (torch.rand(cout, cin, k, k, device=device, generator=g) * 2 - 1) * bound
            w[name + ".bias"] = (torch.rand(cout, device=device, generator=g) * 2 - 1) * bound

        def norm(name, c):
            w[name + ".weight"] = torch.ones(c, device=device)
            w[name + ".bias"] = torch.zeros(c, device=device)

        def res(name, cin, cout):
            norm(name + ".norm1", cin); conv(name + ".conv1", cout, cin, 3)
            norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout, 3)
            if cin != cout:
                conv(name + ".nin_shortcut", cout, cin, 1)

        n_levels = len(ch_mult)
        block_in = ch * ch_mult[-1]
        conv("post_quant_conv", z_channels, z_channels, 1)
        conv("conv_in", block_in, z_channels, 3)
        res("mid.block_1", block_in, block_in)
        norm("mid.attn_1.norm", block_in)
        for n in ("q", "k", "v", "proj_out"):
            conv("mid.attn_1." + n, block_in, block_in, 1)
        res("mid.block_2", block_in, block_in)
        for lvl in reversed(range(n_levels)):
            block_out = ch * ch_mult[lvl]
            for i in range(num_res_blocks[lvl]):
                res(f"up.{lvl}.block.{i}", block_in, block_out)
                block_in = block_out
            if lvl != 0:
                conv(f"up.{lvl}.upsample.conv", block_in, block_in, 3)
        norm("norm_out", block_in)
        conv("conv_out", 3, block_in, 3)
        return cls(w, num_res_blocks, device, z_channels)

    # ---------------------------------------------------------------------------------------------
    def _res(self, p: str, x: torch.Tensor, x_stats_ready: bool = False, emit_stats: bool = False) -> torch.Tensor:
        """ResnetBlock.  `x_stats_ready`: x comes from a block that ran with `emit_stats` (its residual add already summed
        x for this block's norm1); `emit_stats`: the next consumer of the result is a GroupNorm over the same tensor."""
        w = self.w
        h = groupnorm_swish(x, w[p + "norm1.weight"], w[p + "norm1.bias"], stats_ready=x_stats_ready)
        h = _conv(h, w[p + "conv1.weight"], None, 1)
        h = groupnorm_swish(h, w[p + "norm2.weight"], w[p + "norm2.bias"], out=h, in_bias=w[p + "conv1.bias"])
        h = _conv(h, w[p + "conv2.weight"], None, 1)
        if p + "nin_shortcut.weight" in w:
            xs = _conv(x, w[p + "nin_shortcut.weight"], None, 0)
            return bias_add(h, w[p + "conv2.bias"], xs, w[p + "nin_shortcut.bias"], emit_stats=emit_stats)
        return bias_add(h, w[p + "conv2.bias"], x, emit_stats=emit_stats)

    def _attn(self, p: str, x: torch.Tensor) -> torch.Tensor:
        w = self.w
        B, H, W, C = x.shape
        h = groupnorm_swish(x, w[p + "norm.weight"], w[p + "norm.bias"], swish=False).view(B, H * W, C)
        q = F.linear(h, w[p + "q.weight"].view(C, C), w[p + "q.bias"])
        k = F.linear(h, w[p + "k.weight"].view(C, C), w[p + "k.bias"])
        v = F.linear(h, w[p + "v.weight"].view(C, C), w[p + "v.bias"])
        o = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1), scale=float(C) ** -0.5).squeeze(1)
        o = F.linear(o, w[p + "proj_out.weight"].view(C, C), w[p + "proj_out.bias"])
        return x + o.view(B, H, W, C)

    @torch.no_grad()
    def decode_code(self, codebook_indices: torch.Tensor, shape=None) -> torch.Tensor:
        """int64 [B,N] token grid -> pixels fp32 [B,3,16*sqrt(N)...] in (approximately) [-1,1], NCHW like the reference."""
        if shape is not None and tuple(shape) != (int(round(codebook_indices.shape[1] ** 0.5)),) * 2:
            raise ValueError("only square token grids are supported")
        w = self.w
        z = lfq_unpack(codebook_indices.to(self.device), self.codebook_dim)                      # [B,h,w,13]
        h = _conv(z, w["post_quant_conv.weight"], w["post_quant_conv.bias"], 0)
        h = _conv(h, w["conv_in.weight"], w["conv_in.bias"], 1)
        h = self._res("mid.block_1.", h)
        h = self._attn("mid.attn_1.", h)
        h = self._res("mid.block_2.", h)
        ready = False                      # h was produced by a block that left its GroupNorm partial sums in the workspace
        for lvl in reversed(range(len(self.num_res_blocks))):
            n_blk = self.num_res_blocks[lvl]
            for i in range(n_blk):
                # the result feeds a GroupNorm directly (next block's norm1, or norm_out) unless an upsample follows
                emit = i + 1 < n_blk or lvl == 0
                h = self._res(f"up.{lvl}.block.{i}.", h, x_stats_ready=ready, emit_stats=emit)
                ready = emit
            if lvl != 0:
                h = upsample2x(h)
                h = bias_add(_conv(h, w[f"up.{lvl}.upsample.conv.weight"], None, 1), w[f"up.{lvl}.upsample.conv.bias"], emit_stats=True)
                ready = True                 # ... and feeds norm1 of the next level's first block
        h = groupnorm_swish(h, w["norm_out.weight"], w["norm_out.bias"], out=h, stats_ready=ready)
        h = _conv(h, w["conv_out.weight"], w["conv_out.bias"], 1)
        return h.permute(0, 3, 1, 2).float()
