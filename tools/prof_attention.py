"""A few launches of the per-step (window) omni-attention pass of config 2 for `ncu --set full` (tools/ncu_summary.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_unigen_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, L, S, Hq, Hkv, D = 16, 1163, 137, 12, 2, 128
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(B, L - S, Hq, D, device=dev, generator=g).bfloat16()
k = torch.randn(B, L, Hkv, D, device=dev, generator=g).bfloat16()
v = torch.randn(B, L, Hkv, D, device=dev, generator=g).bfloat16()
P = torch.zeros(B, dtype=torch.int32, device=dev)
St = torch.full((B,), S, dtype=torch.int32, device=dev)
for _ in range(4):
    ops.omni_attention(q, k, v, P, St, D ** -0.5, q_offset=S)
torch.cuda.synchronize()
