#!/usr/bin/env python
"""bench.py -- headline benchmark of the T2I parallel-decoding hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this framework (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation (oracle port)

Workload (BASELINE.json configs[1]): UniGen-1.5B (Qwen2.5-1.5B backbone + 8192-entry MAGViTv2
codebook), random-init weights, bf16, text-to-image, batch 16 per GPU, 1024 image tokens (L=1163),
18 MaskGIT steps.  One benchmark "step" = one `t2i_generate` call = 16 images per GPU.

Prints ONE JSON line (rank 0).  `value` = images/s with prompts resident in HBM, `e2e` = the same
through the public API from pinned host memory (H2D of the prompts and D2H of the token grid inside
the timed region), `roofline` = the dominant kernel of the step (the fused gate|up + SwiGLU tcgen05 GEMM, tensor bound;
41 % of a layer), `roofline_attention` / `roofline_decode` = the two north-star kernels, all three timed live with
CUDA events inside the timed region, `cpu_baseline` = the oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "T2I images/s (1024 tok, 18 steps)"
UNIT = "images/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="MEASURED_PEAKS.json")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx, samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------------------
def cpu_reference_sample(L: int, N: int, V: int, timesteps: int, threads: int, repeats: int = 1, warmup: int = 0):
    """The reference's CPU path for this workload, restated (oracle port): Qwen2 forward in fp32 with the dense
    omni mask, the FULL-vocabulary LM head the reference computes (models/unigen.py:287), the slice
    (:416) and one MaskGIT decode step.  Sample = 1 prompt x 1 of `timesteps` steps; returns seconds per sample."""
    import numpy as np
    import torch
    from oracle import backbone_oracle as BO
    from oracle import oracle as O
    from ml_unigen_b200 import prompting
    from ml_unigen_b200.backbone import QwenConfig

    torch.set_num_threads(threads)
    cfg = QwenConfig()
    g = torch.Generator().manual_seed(0)
    H, I, D = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    Hq, Hkv = cfg.num_attention_heads, cfg.num_key_value_heads
    rn = lambda *s: torch.randn(*s, generator=g) * 0.02
    w = {"embed_tokens.weight": rn(cfg.vocab_size, H), "norm.weight": torch.ones(H)}
    for l in range(cfg.num_hidden_layers):
        p = f"layers.{l}."
        w[p + "input_layernorm.weight"] = torch.ones(H)
        w[p + "post_attention_layernorm.weight"] = torch.ones(H)
        for name, shape in (("self_attn.q_proj.weight", (Hq * D, H)), ("self_attn.k_proj.weight", (Hkv * D, H)),
                            ("self_attn.v_proj.weight", (Hkv * D, H)), ("self_attn.q_proj.bias", (Hq * D,)),
                            ("self_attn.k_proj.bias", (Hkv * D,)), ("self_attn.v_proj.bias", (Hkv * D,)),
                            ("self_attn.o_proj.weight", (H, Hq * D)), ("mlp.gate_proj.weight", (I, H)),
                            ("mlp.up_proj.weight", (I, H)), ("mlp.down_proj.weight", (H, I))):
            w[p + name] = rn(*shape)
    ids = prompting.synthetic_prompt_batch(1, N, seed=3, mask_token_id=cfg.vocab_size - 1)
    assert ids.shape[1] == L
    P, S = prompting.omni_mask_params(ids, prompting.QWEN_PAD_ID, prompting.QWEN_SOI_ID)
    cfgd = cfg.as_dict()
    times = []
    with torch.no_grad():
        for it in range(warmup + repeats):
            t0 = time.perf_counter()
            x = w["embed_tokens.weight"][ids]
            hid = BO.qwen2_forward(w, cfgd, x, P, S)
            logits = torch.nn.functional.linear(hid, w["embed_tokens.weight"])                 # full vocabulary, as the reference
            img = logits[:, -(N + 1):-1, prompting.QWEN_TEXT_VOCAB:-1].contiguous()            # unigen.py:416
            bits = O.f32_to_bf16_bits(img.numpy())
            cur = np.full((1, N), cfg.vocab_size - 1, np.int64)
            O.decode_step(bits, cur, mask_token_id=cfg.vocab_size - 1, mask_len_cap=N - 4, temperature=1.0, seed=1, offset=it)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    N, V, T, B = args.tokens, 8192, args.timesteps, args.batch
    L = 139 + N
    t = cpu_reference_sample(L, N, V, T, threads, repeats=max(1, args.steps), warmup=min(args.warmup, 1))
    ips = 1.0 / (t * T)
    sample = f"1 of {B} prompts x 1 of {T} MaskGIT steps per timed step (L={L}, fp32, full-vocab head), images/s = 1/(t*{T})"
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * T * B * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "UniGen-1.5B random-init T2I, batch 16/GPU, 1024 image tokens (L=1163), 18 MaskGIT steps",
                   "reference_path": "CPU oracle port of models/unigen.py:344-455 + transformers Qwen2 (torch fp32, all host threads)"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def attention_flops(P, S, L, N, Hq=12, D=128):
    """Algorithmic FLOPs of one omni-attention launch: 4*D*Hq*sum_b pairs_b (allowed (q,k) pairs only)."""
    tot = 0
    for p, s in zip(P, S):
        tot += p * (p + 1) // 2 + (s - p) * (s - p + 1) // 2 + (N + 2) * (L - p)
    return 4 * D * Hq * tot


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ml_unigen_b200 import prompting, ops
    from ml_unigen_b200 import dist as ugd
    from ml_unigen_b200.unigen import UniGenB200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, T, V = args.batch, args.tokens, args.timesteps, 8192
    L = 139 + N
    model = UniGenB200.random_init(dev, seed=0)
    mask_id = model.mask_token_id
    gB = B * world
    all_ids = prompting.synthetic_prompt_batch(gB, N, seed=11, mask_token_id=mask_id)          # [gB, L] on the host
    s, e = ugd.shard_range(gB, rank, world)
    host_ids = all_ids[s:e].contiguous().pin_memory()
    dev_ids = host_ids.to(dev)
    P, S = prompting.omni_mask_params(dev_ids, prompting.QWEN_PAD_ID, prompting.QWEN_SOI_ID)
    P_l, S_l = P.tolist(), S.tolist()
    gen_kw = dict(timesteps=T, image_token_num_per_image=N, text_vocab_size=prompting.QWEN_TEXT_VOCAB,
                  temperature=1.0, guidance_scale=0, validate_mask=False)
    host_out = torch.empty(e - s, N, dtype=torch.int64).pin_memory()

    def one_generation(step_idx, from_host: bool):
        ids_in = host_ids.to(dev, non_blocking=True) if from_host else dev_ids
        out = model.t2i_generate(input_ids=ids_in, philox_seed=1234, philox_offset=step_idx * 64, batch_offset=s, **gen_kw)
        scores = ugd.candidate_scores(model.last_sel_prob)
        if world > 1:
            out_all, _ = ugd.gather_ids_scores(out, scores, gB)
        if from_host:
            host_out.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return out

    def timed(K, from_host, sampler=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(K):
            one_generation(1000 + k, from_host)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks

    for w in range(args.warmup):
        one_generation(w, False)
    torch.cuda.synchronize()

    # instrumented region (not the headline number): the same K generations with the launches of the three roofline kernels
    # bracketed by CUDA events; the events keep the per-step pass out of its CUDA graph, so this region runs eagerly
    model.backbone.attn_events = []
    model.backbone.mlp_events = []
    model.decode_events = []
    t_instr, _ = timed(args.steps, False)
    attn_all = [(a.elapsed_time(b), lq) for a, b, lq in model.backbone.attn_events]
    attn_ms = [t for t, lq in attn_all if lq != L]      # the per-step launches (image-block rows only)
    attn_full_ms = [t for t, lq in attn_all if lq == L] # the step-0 launches over all rows
    dec_ms = [a.elapsed_time(b) for a, b in model.decode_events]
    mlp_all = [(a.elapsed_time(b), m) for a, b, m in model.backbone.mlp_events]
    model.backbone.attn_events = None
    model.backbone.mlp_events = None
    model.decode_events = None
    # timed region 1 (`value`): prompts resident in HBM, no instrumentation
    sampler = ClockSampler(local) if rank == 0 else None
    t_dev, clocks = timed(args.steps, False, sampler)
    # timed region 2 (`e2e`): end to end through the public API from pinned host memory
    t_e2e, _ = timed(args.steps, True)

    # after the measured path: the MAGViT-v2 pixel decoder on the generated token grids (reported
    # separately -- BASELINE.json's images/s counts token grids)
    pix = None
    if world == 1 and not args.no_pixels:           # N = 1 only: one_generation() holds a collective at N > 1
        from ml_unigen_b200.magvit import MagvitV2DecoderB200
        from ml_unigen_b200 import callers
        dec = MagvitV2DecoderB200.random_init(dev, seed=0)
        grid = callers.clamp_token_ids(one_generation(2000, False), V)
        for _ in range(2):
            dec.decode_code(grid)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(3):
            img = dec.decode_code(grid)
        p1.record()
        torch.cuda.synchronize()
        pix_ms = p0.elapsed_time(p1) / 3
        pix = {"ms_per_batch": pix_ms, "images": int(grid.shape[0]), "pixels": f"{img.shape[-2]}x{img.shape[-1]}",
               "images_per_s_decoder_alone": grid.shape[0] / pix_ms * 1e3,
               "note": "MagvitV2DecoderB200.decode_code (NHWC bf16; cuDNN convolutions, hand-written LFQ unpack / GroupNorm+swish / "
                       "upsample kernels), random-init weights of the default architecture; not part of `value`"}
        del dec, img

    if rank == 0:
        peaks = _peaks()
        images = gB * args.steps
        n_layers = model.backbone.cfg.num_hidden_layers
        flops_full = attention_flops(P_l, S_l, L, N)
        flops = 4 * 128 * 12 * sum((N + 2) * (L - p) for p in P_l) if attn_ms else flops_full   # suffix rows only
        if not attn_ms:
            attn_ms = attn_full_ms
        attn_avg = sum(attn_ms) / len(attn_ms) * 1e-3
        dec_avg = sum(dec_ms) / len(dec_ms) * 1e-3
        achieved = flops / attn_avg / 1e12
        attn_share = (sum(attn_ms) + sum(attn_full_ms if attn_full_ms is not attn_ms else [])) * 1e-3 / t_instr
        # own launches: embed, per layer 2 x add+RMSNorm + RoPE + attention + gate|up GEMM, final norm, head GEMM, decode rows + selection
        launches_per_gen = T * (1 + n_layers * 5 + 1 + 1 + 2)
        cfgm = model.backbone.cfg
        m_step = (e - s) * (N + 2)                           # rows of the per-step passes (text-prefix K/V reuse)
        mlp_ms = [t for t, m in mlp_all if m == m_step] or [t for t, m in mlp_all]
        m_rows = m_step if any(m == m_step for _, m in mlp_all) else mlp_all[0][1]
        mlp_avg = sum(mlp_ms) / len(mlp_ms) * 1e-3
        mlp_flops = 2.0 * m_rows * 2 * cfgm.intermediate_size * cfgm.hidden_size
        mlp_tf = mlp_flops / mlp_avg / 1e12
        mlp_share = sum(t for t, _ in mlp_all) * 1e-3 / t_instr
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            tc = cpu_reference_sample(L, N, V, T, threads, repeats=1, warmup=0)
            cpu = {"value": 1.0 / (tc * T), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"1 of {B} prompts x 1 of {T} MaskGIT steps (L={L}, torch fp32 Qwen2 restatement + full-vocab head "
                             f"+ C decode-step oracle): {tc:.2f} s, images/s = 1/(t*{T})"}
        line = {
            "metric": METRIC, "value": images / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"UniGen-1.5B random-init T2I, batch {B}/GPU, {N} image tokens (L={L}), {T} MaskGIT steps, "
                                   "guidance 0, cosine schedule (BASELINE.json configs[1])",
                       "global_batch": gB, "parallelism": f"replicas x{world}, images sharded, 1 all-gather/generation",
                       "l2": "working set (3.1 GB weights + activations) >> 126 MB L2, no flush needed",
                       "launch": "steps 1..17 of a generation replay one CUDA graph of the per-step backbone pass; the roofline objects "
                                 "come from an instrumented eager pass of the same K generations (%.1f ms per generation)" % (t_instr / args.steps * 1e3),
                       "gemms": "gate|up + SwiGLU and the codebook head: own tcgen05 2-CTA kernel (ug_gemm_bf16); q/k/v, o, down: cuBLAS bf16 "
                                "(F.linear); attention, decode step and row kernels hand-written"},
            "roofline": {"kernel": "gemm_tc_kernel, SwiGLU epilogue (tcgen05 cta_group::2): gate_proj | up_proj + silu(g)*u in one launch",
                         "bound": "tensor", "achieved": mlp_tf, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                         "frac": mlp_tf / peaks["tf_sust"], "traffic": 618629376,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                                         "(profiles/r1_ncu_gemm_swiglu.md); algorithmic bytes 400 MB (A 50 + W 55 + act 294)",
                         "peak_source": peaks["source"] + " bf16_tflops_sustained (burst: %.1f)" % peaks["tf_burst"],
                         "flops_per_launch": mlp_flops, "avg_launch_ms": mlp_avg * 1e3, "launches_timed": len(mlp_ms),
                         "share_of_step": mlp_share,
                         "note": "2*M*N*K with M = 16 x 1026 rows, N = 17920, K = 1536; launches of MaskGIT steps 1..17"},
            "roofline_attention": {"kernel": "omni_attention_kernel (tcgen05)", "bound": "tensor", "achieved": achieved,
                         "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sust"], "traffic": 83301888,
                         "traffic_note": "dram__bytes_read.sum (69.6 MB = q + k + v once) + dram__bytes_write.sum (13.7 MB; most of the "
                                         "50 MB output is still in L2 when the capture ends) of one window-pass launch, ncu --set full "
                                         "(profiles/r1_ncu_omni_attention.md)",
                         "peak_source": peaks["source"] + " bf16_tflops_sustained", "flops_per_launch": flops,
                         "avg_launch_ms": attn_avg * 1e3, "launches_timed": len(attn_ms), "share_of_step": attn_share,
                         "note": "launches of MaskGIT steps 1..17 (query rows = the N+2 image-block rows, keys = all rows; "
                                 "text-prefix K/V reuse); the 28 step-0 launches over all L rows are timed separately",
                         "step0_full_avg_launch_ms": (sum(attn_full_ms) / len(attn_full_ms)) if attn_full_ms else None,
                         "step0_full_flops_per_launch": flops_full},
            "roofline_decode": {"kernel": "decode_fast_kernel + select_kernel (the two launches of one decode step)", "bound": "hbm", "achieved": B * N * V * 2 / dec_avg / 1e9,
                                "peak": peaks["hbm"], "unit": "GB/s", "frac": B * N * V * 2 / dec_avg / 1e9 / peaks["hbm"],
                                "avg_launch_ms": dec_avg * 1e3, "launches_timed": len(dec_ms),
                                "note": "average over the 18 mask states of a generation; rows of known tokens are skipped"},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e": {"value": images / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(host_ids.numel() * 8),
                    "d2h_bytes_per_step": int(host_out.numel() * 8)},
            "gpu_launches": launches_per_gen * args.steps,
        }
        if pix is not None:
            pix["images_per_s_tokens_plus_pixels"] = (e - s) / (t_dev / args.steps + pix["ms_per_batch"] * 1e-3)
            line["pixel_decoder"] = pix
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--tokens", type=int, default=1024)
    ap.add_argument("--timesteps", type=int, default=18)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pixels", action="store_true", help="skip the MAGViT-v2 pixel-decoder timing after the measured path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
