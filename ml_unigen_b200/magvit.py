"""MAGViT-v2 pixel decoder for the token grids the T2I path produces:
`MAGVITv2.decode_code` of apple/ml-unigen (models/multimodal_encoder/magvitv2.py:438-442 =
`LFQuantizer.get_codebook_entry` :217-230 + `VQGANDecoder.forward` :371-408; blocks in common_modules.py:19-43,
171-215, 301-360), as the callers run it right after `t2i_generate` (evaluation/inference_geneval.py:266-267).

Layout for one B200: activations NHWC bf16, every convolution an IMPLICIT GEMM on the library's own tcgen05 kernel
(`ug_conv_nhwc_bf16`: the A operand of tap (ky, kx) is the output's [4 x 32]-pixel box shifted by the tap, fetched by 4-D TMA;
borders, partial boxes and channel padding are TMA out-of-bounds fill), the single attention block of the decoder (1024 tokens,
one head of 512 channels) two plain GEMMs of the same kernel around a row-softmax kernel, and everything between the
convolutions one hand-written, single-purpose pass over HBM (csrc/magvit_ops.cu): LFQ lookup + post_quant_conv (no embedding
table, no GEMM), GroupNorm(32)+swish in one statistics pass and one apply pass (the reference: group_norm, sigmoid, mul = 5
passes), bias + residual adds, nearest-2x upsampling.  No cuDNN / cuBLAS / SDPA call remains (round 1 ran the convolutions on
cuDNN and the attention on torch SDPA).  Weights: the reference's `VQGANDecoder` state dict.  There is no CPU path."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

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


def _ceil(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def pack_conv(weight: torch.Tensor, bias: Optional[torch.Tensor] = None, cin_pad_to: Optional[int] = None):
    """Conv2d weight [Cout, Cin, k, k] (k = 3 or 1) -> the layout `ug_conv_nhwc_bf16` reads: bf16 [ceil128(Cout), k*k*ceil64(Cin')]
    with column (ky*k + kx)*ceil64(Cin') + c, zero in the padding (Cin' = cin_pad_to when the activation carries zero channels
    beyond Cin, e.g. the 64-channel output of `lfq_postquant`); bias -> bf16 [ceil128(Cout)] (or None)."""
    Cout, Cin, k, _ = weight.shape
    cin = cin_pad_to or Cin
    cp, npad = _ceil(cin, 64), _ceil(Cout, 128)
    w = torch.zeros(npad, k * k, cp, dtype=torch.bfloat16, device=weight.device)
    w[:Cout, :, :Cin] = weight.detach().permute(0, 2, 3, 1).reshape(Cout, k * k, Cin).to(torch.bfloat16)
    b = None
    if bias is not None:
        b = torch.zeros(npad, dtype=torch.bfloat16, device=weight.device)
        b[:Cout] = bias.detach().to(torch.bfloat16)
    return w.reshape(npad, k * k * cp).contiguous(), b


def conv_nhwc(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], cout: int, ksize: int) -> torch.Tensor:
    """k x k convolution (k = 3: zero padding 1; k = 1), stride 1, on an NHWC bf16 activation [B,H,W,Cin] -> [B,H,W,cout], one
    launch of the tcgen05 implicit-GEMM kernel; `w_packed`, `bias` from `pack_conv`."""
    ops._need_cuda(x, w_packed, bias)
    B, H, W, Cin = x.shape
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or w_packed.dtype != torch.bfloat16 or not w_packed.is_contiguous():
        raise ValueError("x must be contiguous bf16 NHWC, w_packed contiguous bf16")
    if tuple(w_packed.shape) != (_ceil(cout, 128), ksize * ksize * _ceil(Cin, 64)):
        raise ValueError(f"w_packed must be [{_ceil(cout, 128)}, {ksize * ksize * _ceil(Cin, 64)}] for Cin={Cin}, cout={cout}, k={ksize}")
    y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().ug_conv_nhwc_bf16(ops._ptr(x), ops._ptr(w_packed), ops._ptr(bias), ops._ptr(y), B, H, W, Cin, cout, ksize,
                                            ops._stream()))
    return y


def lfq_postquant(ids: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], dim: int = 13, cpad: int = 64) -> torch.Tensor:
    """int64 [B,N] -> bf16 NHWC [B,h,w,cpad]: LFQ codebook entry (+-1 bits) through the 1x1 `post_quant_conv` (w fp32 [dim,dim], b fp32
    [dim] or None) in one pass; channels >= dim are zero."""
    ops._need_cuda(ids, w, b)
    B, N = ids.shape
    h = int(round(N ** 0.5))
    if h * h != N or ids.dtype != torch.int64 or w.dtype != torch.float32 or tuple(w.shape) != (dim, dim) or not w.is_contiguous():
        raise ValueError("ids must be int64 [B, N] with N a square, w contiguous fp32 [dim, dim]")
    z = torch.empty(B, h, h, cpad, dtype=torch.bfloat16, device=ids.device)
    _lib.check(_lib.lib().ug_lfq_postquant_bf16(ops._ptr(ids.contiguous()), ops._ptr(w), ops._ptr(b), ops._ptr(z), B * N, dim, cpad, ops._stream()))
    return z


def softmax_rows(x: torch.Tensor, scale: float, cols_valid: Optional[int] = None) -> torch.Tensor:
    """bf16 [R, C] -> softmax(scale * x[:, :cols_valid]) per row (columns beyond cols_valid: zero), bf16."""
    ops._need_cuda(x)
    R, C = x.shape
    if x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise ValueError("x must be contiguous bf16 [R, C]")
    y = torch.empty_like(x)
    _lib.check(_lib.lib().ug_softmax_rows_bf16(ops._ptr(x), ops._ptr(y), R, C, C if cols_valid is None else cols_valid, float(scale), ops._stream()))
    return y


def _pad2d(t: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """zero-pad a 2-D tensor to [rows, cols] (a no-op view at the default architecture's shapes)."""
    if t.shape[0] == rows and t.shape[1] == cols and t.is_contiguous():
        return t
    out = torch.zeros(rows, cols, dtype=t.dtype, device=t.device)
    out[: t.shape[0], : t.shape[1]] = t
    return out


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
        self.w: Dict[str, torch.Tensor] = {}          # norm gains / biases and conv biases, bf16
        self.cw: Dict[str, tuple] = {}                # conv name -> (packed weight, packed bias, Cout, ksize)
        # the parameters as the kernels see them (rounded to bf16), in fp32: what the fp32 oracle of the tests is run with
        self.raw = {k: v.detach().to(self.device, torch.bfloat16).float() for k, v in weights.items()}
        convs = {k[: -len(".weight")] for k, v in weights.items() if v.dim() == 4}
        for k, t in self.raw.items():
            name = k.rsplit(".", 1)[0]
            if name not in convs:                         # GroupNorm gain / bias
                self.w[k] = t.to(torch.bfloat16).contiguous()
            elif k.endswith(".bias"):
                self.w[k] = t.to(torch.bfloat16).contiguous()
            elif name == "post_quant_conv":               # fused with the LFQ lookup (fp32 matrix, lfq_postquant)
                self.pq_w = t.reshape(t.shape[0], t.shape[1]).contiguous()
                self.pq_b = self.raw.get(name + ".bias")
            else:
                cin_pad = 64 if name == "conv_in" else None   # its input is the 64-channel (zero-padded) output of lfq_postquant
                wp, bp = pack_conv(t, self.raw.get(name + ".bias"), cin_pad_to=cin_pad)
                self.cw[name] = (wp, bp, t.shape[0], t.shape[2])
        # AttnBlock: softmax rows sum to one, so P (V + 1 b_v^T) = P V + b_v: the value bias moves behind the attention and is
        # folded into proj_out's bias (b_p + W_p b_v) once here; the v projection itself then needs no bias
        if "mid.attn_1.proj_out" in self.cw:
            wp32 = self.raw["mid.attn_1.proj_out.weight"].reshape(self.raw["mid.attn_1.proj_out.weight"].shape[0], -1)
            b_eff = self.raw["mid.attn_1.proj_out.bias"] + wp32 @ self.raw["mid.attn_1.v.bias"]
            wpk, _, co, ks = self.cw["mid.attn_1.proj_out"]
            bpad = torch.zeros(wpk.shape[0], dtype=torch.bfloat16, device=self.device)
            bpad[:co] = b_eff.to(torch.bfloat16)
            self.cw["mid.attn_1.proj_out"] = (wpk, bpad, co, ks)

    def _conv(self, name: str, x: torch.Tensor, with_bias: bool) -> torch.Tensor:
        wp, bp, cout, ks = self.cw[name]
        c8 = _ceil(cout, 8)                               # TMA needs 16-byte row strides: conv_out's 3 channels are stored as 8
        y = conv_nhwc(x, wp, bp if with_bias else None, c8, ks)
        return y if c8 == cout else y[..., :cout]

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
        """ResnetBlock (`p` ends with a dot).  `x_stats_ready`: x comes from a block that ran with `emit_stats` (its residual add already summed
        x for this block's norm1); `emit_stats`: the next consumer of the result is a GroupNorm over the same tensor."""
        w = self.w
        h = groupnorm_swish(x, w[p + "norm1.weight"], w[p + "norm1.bias"], stats_ready=x_stats_ready)
        h = self._conv(p + "conv1", h, with_bias=True)               # the bias rides in the GEMM epilogue
        h = groupnorm_swish(h, w[p + "norm2.weight"], w[p + "norm2.bias"], out=h)
        h = self._conv(p + "conv2", h, with_bias=False)              # its bias joins the residual add below (one pass, + GroupNorm sums)
        if p + "nin_shortcut" in self.cw:
            xs = self._conv(p + "nin_shortcut", x, with_bias=False)
            return bias_add(h, w[p + "conv2.bias"], xs, w[p + "nin_shortcut.bias"], emit_stats=emit_stats)
        return bias_add(h, w[p + "conv2.bias"], x, emit_stats=emit_stats)

    def _attn(self, p: str, x: torch.Tensor) -> torch.Tensor:
        """AttnBlock (common_modules.py:171-215): GroupNorm, 1x1 q/k/v, softmax(q k^T / sqrt(C)) v, 1x1 proj_out, residual.  One head of
        C channels over H*W tokens: per image two plain tcgen05 GEMMs (scores, P V) around the row-softmax kernel."""
        w = self.w
        B, H, W, C = x.shape
        T = H * W
        h = groupnorm_swish(x, w[p + "norm.weight"], w[p + "norm.bias"], swish=False)
        q = self._conv(p + "q", h, with_bias=True).view(B, T, C)
        k = self._conv(p + "k", h, with_bias=True).view(B, T, C)
        v = self._conv(p + "v", h, with_bias=False).view(B, T, C)          # b_v is folded into proj_out's bias (see __init__)
        Tp, Cp, Cn = _ceil(T, 128), _ceil(C, 64), _ceil(C, 128)            # GEMM shape limits: N % 128, K % 64 (no-ops at C=512, T=1024)
        o = torch.empty(B, T, C, dtype=torch.bfloat16, device=x.device)
        for b in range(B):
            s_b = ops.gemm(_pad2d(q[b], T, Cp), _pad2d(k[b], Tp, Cp))                        # [T, Tp] scores
            p_b = softmax_rows(s_b, float(C) ** -0.5, cols_valid=T)                           # zero beyond T
            o_b = ops.gemm(p_b, _pad2d(v[b].t(), Cn, Tp))                                     # [T, Cn] = P V
            o[b] = o_b[:, :C]
        o = self._conv(p + "proj_out", o.view(B, H, W, C), with_bias=True)
        return bias_add(o, None, x)

    @torch.no_grad()
    def decode_code(self, codebook_indices: torch.Tensor, shape=None) -> torch.Tensor:
        """int64 [B,N] token grid -> pixels fp32 [B,3,16*sqrt(N)...] in (approximately) [-1,1], NCHW like the reference."""
        if shape is not None and tuple(shape) != (int(round(codebook_indices.shape[1] ** 0.5)),) * 2:
            raise ValueError("only square token grids are supported")
        w = self.w
        z = lfq_postquant(codebook_indices.to(self.device), self.pq_w, self.pq_b, self.codebook_dim)   # [B,h,w,64]: 13 channels + zeros
        h = self._conv("conv_in", z, with_bias=True)
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
                h = bias_add(self._conv(f"up.{lvl}.upsample.conv", h, with_bias=False), w[f"up.{lvl}.upsample.conv.bias"], emit_stats=True)
                ready = True                 # ... and feeds norm1 of the next level's first block
        h = groupnorm_swish(h, w["norm_out.weight"], w["norm_out.bias"], out=h, stats_ready=ready)
        h = self._conv("conv_out", h, with_bias=True)                  # 3 output channels: rows of the packed weight beyond 3 are zero,
        return h.permute(0, 3, 1, 2).float()                           # the TMA store clips them
