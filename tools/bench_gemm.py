"""Isolated check + timing of the tcgen05 linear-layer kernel (ug_gemm_bf16) against cuBLAS (F.linear) on the
layer shapes of config 2 (M = 16 x 1026 rows).  Writes one JSON object per shape.

    python tools/bench_gemm.py [--m 16416] [--iters 20] [--out gpurun_out/gemm_bench.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_unigen_b200 import ops  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=16 * 1026)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/gemm_bench.json")
    ap.add_argument("--shapes", default="gate_up,qkv_rope,head,o_res,down_res,o,down")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    M = a.m
    shapes = {"gate_up": (17920, 1536, "swiglu"), "qkv": (2048, 1536, "bias"), "o": (1536, 1536, "plain"),
              "down": (1536, 8960, "plain"), "head": (8192, 1536, "plain"),
              "qkv_rope": (2048, 1536, "rope"), "o_res": (1536, 1536, "residual"), "down_res": (1536, 8960, "residual")}
    res = []
    for name in a.shapes.split(","):
        N, K, kind = shapes[name]
        ncopies = 4                              # rotate weights so that they come from HBM like inside the layer loop
        x = (torch.randn(M, K, device=dev, generator=g)).bfloat16()
        ws = [(0.02 * torch.randn(N, K, device=dev, generator=g)).bfloat16() for _ in range(ncopies)]
        bias = (0.02 * torch.randn(N, device=dev, generator=g)).bfloat16() if kind == "bias" else None
        flops = 2.0 * M * N * K
        if kind == "swiglu":
            I = N // 2
            packed = [ops.pack_gate_up(w[:I], w[I:]) for w in ws]
            ref = ops.swiglu(F.linear(x, ws[0]))
            got = ops.gemm(x, packed[0], mode=ops.GEMM_SWIGLU)
            ref32 = F.silu(F.linear(x[:512].float(), ws[0][:I].float()).bfloat16().float()) * F.linear(x[:512].float(), ws[0][I:].float()).bfloat16().float()
            out_u = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            out_a = torch.empty(M, I, dtype=torch.bfloat16, device=dev)
            t_ref = timed(lambda i: ops.swiglu(F.linear(x, ws[i % ncopies], out=None), out=out_a), a.iters)
            t_lin = timed(lambda i: F.linear(x, ws[i % ncopies]), a.iters)
            t_own = timed(lambda i: ops.gemm(x, packed[i % ncopies], mode=ops.GEMM_SWIGLU, out=out_a), a.iters)
            t_own_plain = timed(lambda i: ops.gemm(x, ws[i % ncopies], out=out_u), a.iters)
        elif kind == "rope":
            from oracle import backbone_oracle as BO
            L = 1026
            cos, sin = BO.rope_cos_sin(L, 128, 1e6, dev)
            cos, sin = cos[:, :64].contiguous(), sin[:, :64].contiguous()
            bias = (0.02 * torch.randn(N, device=dev, generator=g)).bfloat16()
            caches = [torch.zeros(M // L, L, 2, 2, 128, dtype=torch.bfloat16, device=dev) for _ in range(2)]
            out_p = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            ref = ops.rope_qk_(F.linear(x, ws[0], bias), cos, sin, L, 12, 2, 128, kv_cache=caches[0])
            got = ops.qkv_rope(x, ws[0], bias, cos, sin, L, 12, 2, kv_cache=caches[1])
            ref32 = ref[:512].float()
            t_ref = timed(lambda i: ops.rope_qk_(F.linear(x, ws[i % ncopies], bias), cos, sin, L, 12, 2, 128, kv_cache=caches[0]), a.iters)
            t_lin = timed(lambda i: F.linear(x, ws[i % ncopies], bias), a.iters)
            t_own = timed(lambda i: ops.qkv_rope(x, ws[i % ncopies], bias, cos, sin, L, 12, 2, kv_cache=caches[1], out=out_p), a.iters)
            t_own_plain = timed(lambda i: ops.gemm(x, ws[i % ncopies], bias, out=out_p), a.iters)
        elif kind == "residual":
            # round 2: o_proj / down_proj + residual add + RMSNorm statistics in ONE launch, against what round 1 ran:
            # cuBLAS GEMM + ug_add_rmsnorm (add, write x back, normalise, write h)
            gain = torch.ones(N, dtype=torch.bfloat16, device=dev)
            xs = torch.randn(M, N, device=dev, generator=g)
            x_own, x_ref = xs.clone(), xs.clone()
            xb = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            ss = torch.empty(M, N // 128, dtype=torch.float32, device=dev)
            h = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            ops.gemm_residual_(x, ws[0], x_own, xb, ss)
            ops.add_rmsnorm(x_ref, gain, 1e-6, F.linear(x, ws[0]), out=h)
            got, ref = x_own.bfloat16(), x_ref.bfloat16()
            ref32 = xs[:512] + F.linear(x[:512].float(), ws[0].float())
            t_ref = timed(lambda i: ops.add_rmsnorm(x_ref, gain, 1e-6, F.linear(x, ws[i % ncopies]), out=h), a.iters)
            t_lin = timed(lambda i: F.linear(x, ws[i % ncopies]), a.iters)
            t_own = timed(lambda i: ops.gemm_residual_(x, ws[i % ncopies], x_own, xb, ss), a.iters)
            out_p = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            t_own_plain = timed(lambda i: ops.gemm(x, ws[i % ncopies], out=out_p), a.iters)
        else:
            ref = F.linear(x, ws[0], bias)
            got = ops.gemm(x, ws[0], bias)
            ref32 = F.linear(x[:512].float(), ws[0].float(), None if bias is None else bias.float())
            out_p = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            t_ref = timed(lambda i: F.linear(x, ws[i % ncopies], bias), a.iters)
            t_lin = t_ref
            t_own = timed(lambda i: ops.gemm(x, ws[i % ncopies], bias, out=out_p), a.iters)
            t_own_plain = t_own
        torch.cuda.synchronize()
        d = (got.float() - ref.float())
        d32 = (got[:512].float() - ref32)
        r = dict(shape=name, M=M, N=N, K=K, kind=kind,
                 max_abs_vs_cublas=d.abs().max().item(), rel_l2_vs_cublas=(d.norm() / ref.float().norm()).item(),
                 max_abs_vs_fp32=d32.abs().max().item(), rel_l2_vs_fp32=(d32.norm() / ref32.norm()).item(),
                 cublas_rel_l2_vs_fp32=((ref[:512].float() - ref32).norm() / ref32.norm()).item(),
                 frac_equal=(got == ref).float().mean().item(),
                 ms_cublas_path=t_ref, ms_cublas_gemm_only=t_lin, ms_own=t_own, ms_own_plain_epilogue=t_own_plain,
                 tflops_cublas_path=flops / t_ref / 1e9, tflops_own=flops / t_own / 1e9, speedup=t_ref / t_own)
        print(json.dumps(r), flush=True)
        res.append(r)
        del x, ws
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
