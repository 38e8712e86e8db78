"""Shared by the CPU (oracle) and GPU (kernel) tests of tests/golden/decode_full_*.npz: regenerates the logits and the noise of
every step of a full-size golden (N=1024, V=8192, T=18; only seeds, ids and masks are stored -- oracle/make_golden.py
`make_decode_fullsize`) and yields what the reference produced for it."""
import os

import numpy as np

from oracle import oracle as O

FULL_CASES = ["peaked_integer", "flat_integer", "unit_replay"]


class ReplayMismatch(Exception):
    """torch's CPU exponential_/uniform_ stream on this host differs from the recording host's (MKL / libm dispatch)."""


def full_steps(golden_dir, name):
    import torch
    g = np.load(os.path.join(golden_dir, f"decode_full_{name}.npz"))
    T, B, N, V = [int(x) for x in g["shape"]]
    seed, scale, noise = int(g["seed"]), float(g["scale"]), str(g["noise"])
    sched = O.step_schedule(T, N, float(g["temperature0"]))
    replay = torch.Generator().manual_seed(seed + 100) if noise == "replay" else None
    for t in range(T):
        bits = O.synthetic_logits_bits(seed, t, B, N, V, scale)
        if noise == "replay":
            q = torch.empty(B * N, V).exponential_(1, generator=replay).view(B, N, V).numpy()
            u = torch.zeros(B, N).uniform_(0, 1, generator=replay).numpy()
            if O.noise_checksum(q) != int(g["q_sum"][t]) or O.noise_checksum(u) != int(g["u_sum"][t]):
                raise ReplayMismatch(f"step {t}")
        else:
            q, u = O.integer_noise(seed, t, B, N, V)
        yield dict(t=t, bits=bits, q=q, u=u, cap=sched[t][0], temp=sched[t][1], mask_id=int(g["mask_token_id"]),
                   sampled_raw=g["sampled_raw"][t].astype(np.int64), mask_len=g["mask_len"][t],
                   masking=np.unpackbits(g["masking"][t], axis=-1)[:, :N].astype(bool),
                   final_ids=g["final_ids"].astype(np.int64), last=(t == T - 1), shape=(T, B, N, V))
