"""CPU: the fp32 restatement of MAGVITv2.decode_code (oracle/magvit_oracle.py) against the golden recorded from the
reference's own LFQuantizer + VQGANDecoder (tests/golden/magvit_tiny.npz, oracle/make_golden_magvit.py)."""
import os

import numpy as np
import torch

from oracle import magvit_oracle as MO
from oracle.oracle import bf16_bits_to_f32


def _golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "magvit_tiny.npz"))
    sd = {k[2:]: torch.from_numpy(bf16_bits_to_f32(g[k])) for k in g.files if k.startswith("w:")}
    return g, sd


def test_lfq_codebook_entry_matches_reference(golden_dir):
    g, _ = _golden(golden_dir)
    ids = torch.from_numpy(g["ids"])
    z = MO.lfq_codebook_entry(ids)
    assert torch.equal(z, torch.from_numpy(g["z"]))
    # known answers: id 0 -> all -1, id 8191 -> all +1, id 1 -> only the LAST channel +1, id 4096 -> only the FIRST
    zz = z.permute(0, 2, 3, 1).reshape(-1, 13)
    assert (zz[0] == -1).all() and (zz[1] == 1).all()
    assert zz[2].tolist() == [-1.0] * 12 + [1.0] and zz[3].tolist() == [1.0] + [-1.0] * 12


def test_decoder_restatement_matches_reference_golden(golden_dir):
    g, sd = _golden(golden_dir)
    out = MO.decode_code(sd, [2, 1], torch.from_numpy(g["ids"]))
    ref = torch.from_numpy(g["out"])
    assert out.shape == ref.shape == (2, 3, 16, 16)
    assert (out - ref).abs().max().item() <= 1e-5
