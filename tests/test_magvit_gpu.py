"""GPU: MAGViT-v2 pixel decoder (ml_unigen_b200.magvit) -- its three hand-written kernels against the plain torch fp32
formulas, and `decode_code` against the golden recorded from the reference's own modules.

Tolerances (bf16 activations between all layers, fp32 accumulation inside the convolutions): GroupNorm+swish max-abs
<= 3e-2 on O(1) outputs (one bf16 rounding of the result); decode_code relative L2 <= 3e-2, max-abs <= 8e-2 * max|ref|."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import magvit_oracle as MO
from oracle.oracle import bf16_bits_to_f32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_lfq_unpack_exact(dev):
    from ml_unigen_b200 import magvit
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 8192, (3, 256), generator=g)
    ids[0, :4] = torch.tensor([0, 8191, 1, 4096])
    z = magvit.lfq_unpack(ids.to(dev))
    ref = MO.lfq_codebook_entry(ids).permute(0, 2, 3, 1)
    assert z.shape == (3, 16, 16, 13) and torch.equal(z.float().cpu(), ref)


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 32), (1, 37, 19, 64), (2, 64, 64, 128), (3, 32, 32, 256), (1, 32, 32, 512), (2, 300, 7, 128)])
@pytest.mark.parametrize("swish", [True, False])
def test_groupnorm_swish(dev, B, H, W, C, swish):
    from ml_unigen_b200 import magvit
    g = torch.Generator(device=dev).manual_seed(C + H)
    x = (torch.randn(B, H, W, C, device=dev, generator=g) * 2 + 0.7).bfloat16()
    gamma = (1 + 0.2 * torch.randn(C, device=dev, generator=g)).bfloat16()
    beta = (0.1 * torch.randn(C, device=dev, generator=g)).bfloat16()
    y = magvit.groupnorm_swish(x, gamma, beta, swish=swish)
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma.float(), beta.float(), eps=1e-6)
    if swish:
        ref = ref * torch.sigmoid(ref)
    ref = ref.permute(0, 2, 3, 1)
    assert (y.float() - ref).abs().max().item() <= 3e-2
    assert ((y.float() - ref).norm() / ref.norm()).item() <= 4e-3
    # in place
    x2 = x.clone()
    magvit.groupnorm_swish(x2, gamma, beta, swish=swish, out=x2)
    assert torch.equal(x2, y)


def test_groupnorm_input_bias_and_bias_add(dev):
    from ml_unigen_b200 import magvit
    g = torch.Generator(device=dev).manual_seed(9)
    B, H, W, C = 2, 40, 24, 128
    x = torch.randn(B, H, W, C, device=dev, generator=g).bfloat16()
    r = torch.randn(B, H, W, C, device=dev, generator=g).bfloat16()
    b1 = torch.randn(C, device=dev, generator=g).bfloat16()
    b2 = torch.randn(C, device=dev, generator=g).bfloat16()
    gamma = (1 + 0.2 * torch.randn(C, device=dev, generator=g)).bfloat16()
    beta = (0.1 * torch.randn(C, device=dev, generator=g)).bfloat16()
    y = magvit.groupnorm_swish(x, gamma, beta, in_bias=b1)
    xin = (x.float() + b1.float()).permute(0, 3, 1, 2)
    ref = F.group_norm(xin, 32, gamma.float(), beta.float(), eps=1e-6)
    ref = (ref * torch.sigmoid(ref)).permute(0, 2, 3, 1)
    assert (y.float() - ref).abs().max().item() <= 3e-2
    # bias + residual (+ shortcut bias): one rounding of the fp32 sum
    out = magvit.bias_add(x.clone(), b1, r, b2)
    assert torch.equal(out, (x.float() + b1.float() + (r.float() + b2.float())).bfloat16())
    assert torch.equal(magvit.bias_add(x.clone(), b1, r), (x.float() + b1.float() + r.float()).bfloat16())
    assert torch.equal(magvit.bias_add(x.clone(), b1), (x.float() + b1.float()).bfloat16())


def test_bias_add_with_fused_groupnorm_statistics(dev):
    """bias_add(..., emit_stats=True) + groupnorm_swish(..., stats_ready=True) == bias_add + groupnorm_swish, bit for bit
    (same values, same thread mapping, same summation order in the fused statistics)."""
    from ml_unigen_b200 import magvit
    g = torch.Generator(device=dev).manual_seed(3)
    for B, H, W, C in ((2, 40, 24, 128), (1, 33, 17, 256), (3, 16, 16, 32), (2, 64, 64, 512)):
        a = torch.randn(B, H, W, C, device=dev, generator=g).bfloat16()
        r = torch.randn(B, H, W, C, device=dev, generator=g).bfloat16()
        b1 = torch.randn(C, device=dev, generator=g).bfloat16()
        b2 = torch.randn(C, device=dev, generator=g).bfloat16()
        gamma = (1 + 0.2 * torch.randn(C, device=dev, generator=g)).bfloat16()
        beta = (0.1 * torch.randn(C, device=dev, generator=g)).bfloat16()
        for rr, bb in ((r, b2), (r, None), (None, None)):
            plain = magvit.bias_add(a.clone(), b1, rr, bb)
            y_plain = magvit.groupnorm_swish(plain, gamma, beta)
            fused = magvit.bias_add(a.clone(), b1, rr, bb, emit_stats=True)
            y_fused = magvit.groupnorm_swish(fused, gamma, beta, stats_ready=True)
            assert torch.equal(fused, plain) and torch.equal(y_fused, y_plain)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 32, 32, 512, 512, 3), (1, 64, 64, 256, 128, 3), (2, 16, 16, 64, 256, 3), (1, 8, 8, 32, 64, 3),
                                              (3, 5, 7, 128, 128, 3), (1, 128, 128, 128, 8, 3), (2, 32, 32, 512, 256, 1), (1, 9, 33, 64, 32, 1),
                                              (1, 32, 32, 64, 512, 3)])
def test_conv_implicit_gemm_matches_torch(dev, B, H, W, Cin, Cout, k):
    """ug_conv_nhwc_bf16 (tcgen05 implicit GEMM: 4-D TMA boxes per tap, borders / partial boxes / channel padding by TMA out-of-bounds
    handling) against torch's fp32 convolution of the same bf16 values.  Tolerance: one bf16 rounding of an fp32-accumulated sum
    (relative L2 <= 4e-3, max-abs <= 2^-6 max|ref|), as for the plain GEMM."""
    from ml_unigen_b200 import magvit
    g = torch.Generator(device=dev).manual_seed(B * 1000 + H + Cin + Cout + k)
    x = torch.randn(B, H, W, Cin, device=dev, generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, device=dev, generator=g) * (Cin * k * k) ** -0.5).bfloat16()
    b = torch.randn(Cout, device=dev, generator=g).bfloat16()
    wp, bp = magvit.pack_conv(w.float(), b.float())
    for with_bias in (True, False):
        y = magvit.conv_nhwc(x, wp, bp if with_bias else None, Cout, k)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b.float() if with_bias else None, padding=k // 2).permute(0, 2, 3, 1)
        d = y.float() - ref
        assert y.shape == ref.shape
        assert (d.norm() / ref.norm()).item() <= 4e-3, (d.norm() / ref.norm()).item()
        assert d.abs().max().item() <= 2 ** -6 * ref.abs().max().item()


def test_lfq_postquant_and_softmax_rows(dev):
    from ml_unigen_b200 import magvit
    g = torch.Generator(device=dev).manual_seed(4)
    ids = torch.randint(0, 8192, (3, 256), device=dev, generator=g)
    w = torch.randn(13, 13, device=dev, generator=g) * 0.3
    b = torch.randn(13, device=dev, generator=g)
    z = magvit.lfq_postquant(ids, w, b, 13, 64)
    bits = MO.lfq_codebook_entry(ids.cpu()).permute(0, 2, 3, 1).to(dev)                  # [B,h,w,13] of +-1
    ref = bits @ w.t() + b
    assert z.shape == (3, 16, 16, 64) and int(z[..., 13:].abs().sum()) == 0
    assert (z[..., :13].float() - ref).abs().max().item() <= 2e-2          # fp32 sums in a different order, one bf16 rounding
    x = (torch.randn(300, 256, device=dev, generator=g) * 3).bfloat16()
    y = magvit.softmax_rows(x, 0.7, cols_valid=200)
    r = torch.softmax(0.7 * x[:, :200].float(), -1)
    assert (y[:, :200].float() - r).abs().max().item() <= 2 ** -8 and int(y[:, 200:].abs().sum()) == 0


def test_upsample2x_exact(dev):
    from ml_unigen_b200 import magvit
    x = torch.randn(2, 5, 7, 64, device=dev).bfloat16()
    y = magvit.upsample2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y.float(), ref)


def test_decode_code_matches_reference_golden(dev, golden_dir):
    from ml_unigen_b200.magvit import MagvitV2DecoderB200
    g = np.load(os.path.join(golden_dir, "magvit_tiny.npz"))
    sd = {k[2:]: torch.from_numpy(bf16_bits_to_f32(g[k])) for k in g.files if k.startswith("w:")}
    dec = MagvitV2DecoderB200(sd, num_res_blocks=[2, 1], device=dev)
    out = dec.decode_code(torch.from_numpy(g["ids"]).to(dev)).cpu()
    ref = torch.from_numpy(g["out"])
    rel = ((out - ref).norm() / ref.norm()).item()
    mx = (out - ref).abs().max().item()
    print(f"decode_code tiny: rel_l2={rel:.3e} max_abs={mx:.3e} max|ref|={ref.abs().max().item():.3f}")
    assert out.shape == ref.shape
    assert rel <= 3e-2 and mx <= 8e-2 * ref.abs().max().item()


def test_decode_code_full_architecture_against_fp32_oracle(dev):
    """Default architecture (ch=128, ch_mult=[1,1,2,2,4], 40 M parameters), random init, 4x4 token grid -> 64x64 pixels:
    the bf16 path against the fp32 restatement run on the same GPU with the same (bf16-representable) weights."""
    from ml_unigen_b200.magvit import MagvitV2DecoderB200
    dec = MagvitV2DecoderB200.random_init(dev, seed=3)
    ids = torch.randint(0, 8192, (2, 16), device=dev)
    out = dec.decode_code(ids)
    ref = MO.decode_code(dec.raw, dec.num_res_blocks, ids)
    assert out.shape == ref.shape == (2, 3, 64, 64)
    rel = ((out - ref).norm() / ref.norm()).item()
    print(f"decode_code full arch: rel_l2={rel:.3e}")
    assert rel <= 5e-2
