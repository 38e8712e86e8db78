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
    sd = {k: v.float().contiguous() for k, v in dec.w.items()}
    ref = MO.decode_code(sd, dec.num_res_blocks, ids)
    assert out.shape == ref.shape == (2, 3, 64, 64)
    rel = ((out - ref).norm() / ref.norm()).item()
    print(f"decode_code full arch: rel_l2={rel:.3e}")
    assert rel <= 5e-2
