"""TEST INFRASTRUCTURE: records tests/golden/magvit_tiny.npz by executing the reference's own LFQuantizer and
VQGANDecoder (models/multimodal_encoder/magvitv2.py, unmodified, loaded by oracle/ref_loader.py) on a small
configuration: ch=32, ch_mult=[1,2], num_res_blocks=[2,1], 8x8 token grids (N=64) -> 16x16 pixels, B=2.
Weights are rounded to bf16-representable values (stored as bf16 bit patterns) so that the bf16 GPU path and the
fp32 oracle start from identical parameters.

    python oracle/make_golden_magvit.py        # needs /root/reference
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.oracle import f32_to_bf16_bits, bf16_bits_to_f32  # noqa: E402

TINY = dict(ch=32, ch_mult=[1, 2], num_res_blocks=[2, 1], resolution=16)
N_TOK, B = 64, 2


def main():
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_loader.load_magvitv2()
        dec = m.VQGANDecoder(**TINY).eval()
        quant = m.LFQuantizer()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in dec.named_parameters():
            if "norm" in name:                       # non-trivial GroupNorm gains / biases
                p.copy_((1.0 + 0.2 * torch.randn(p.shape, generator=g)) if name.endswith("weight") else 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            p.copy_(torch.from_numpy(bf16_bits_to_f32(f32_to_bf16_bits(p.numpy()))))
    ids = torch.randint(0, 8192, (B, N_TOK), generator=g)
    ids[0, :4] = torch.tensor([0, 8191, 1, 4096])    # all bits clear / set, lowest / highest bit
    with torch.no_grad():
        z = quant.get_codebook_entry(ids)
        out = dec(z)["output"]
    arrays = {"ids": ids.numpy(), "z": z.numpy(), "out": out.numpy()}
    for k, v in dec.state_dict().items():
        arrays["w:" + k] = f32_to_bf16_bits(v.numpy())
    path = os.path.join(ROOT, "tests", "golden", "magvit_tiny.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes; out", tuple(out.shape), "absmax", float(out.abs().max()))


if __name__ == "__main__":
    main()
