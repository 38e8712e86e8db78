"""One launch of each linear-layer shape of config 2 through ug_gemm_bf16, for `ncu --set full` (tools/ncu_summary.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_unigen_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
M = 16 * 1026
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["gate_up", "down"]
for name in which:
    N, K, mode = {"gate_up": (17920, 1536, ops.GEMM_SWIGLU), "down": (1536, 8960, ops.GEMM_PLAIN),
                  "qkv": (2048, 1536, ops.GEMM_PLAIN), "o": (1536, 1536, ops.GEMM_PLAIN),
                  "o_res": (1536, 1536, "residual"), "down_res": (1536, 8960, "residual")}[name]
    x = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (0.02 * torch.randn(N, K, device=dev, generator=g)).bfloat16()
    if mode == "residual":
        xs = torch.randn(M, N, device=dev, generator=g)
        xb = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ss = torch.empty(M, N // 128, device=dev)
    for _ in range(3):
        if mode == "residual":
            ops.gemm_residual_(x, w, xs, xb, ss)
        else:
            ops.gemm(x, w, mode=mode)
    torch.cuda.synchronize()
