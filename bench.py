#!/usr/bin/env python
"""bench.py -- headline benchmark of the T2I parallel-decoding hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this framework (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, config 1, real steps (oracle/t2i_port.py)

Workload (BASELINE.json configs[1]): UniGen-1.5B (Qwen2.5-1.5B backbone + 8192-entry MAGViTv2
codebook), random-init weights, bf16, text-to-image, batch 16 per GPU, 1024 image tokens (L=1163),
18 MaskGIT steps.  One benchmark "step" = one `t2i_generate` call = 16 images per GPU.

Prints ONE JSON line (rank 0).  `value` = images/s with prompts resident in HBM, `e2e` = the same
through the public API from pinned host memory (H2D of the prompts and D2H of the token grid inside
the timed region), `roofline` = the dominant kernel of the step (the fused gate|up + SwiGLU tcgen05 GEMM, tensor bound),
`roofline_attention` / `roofline_decode` = the two north-star kernels, all timed live with CUDA events in an instrumented
pass, `best_of_n` = BASELINE.json configs[4] at this GPU count, `cpu_baseline` = the CPU port on the host cores.
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
# The reference's CPU path (BASELINE.json configs[0] / BASELINE.md section 4 row 1): random-init UniGen-1.5B in fp32 on the host
# cores, ONE prompt, 256 image tokens (L = 395), 18 MaskGIT steps, guidance 0.  /root/reference does not exist on the GPU box, so
# what runs is the port oracle/t2i_port.py -- pinned end to end against the UNMODIFIED reference by tests/golden/t2i_tiny.npz --
# executing every step for real: fp32 Qwen2 forward with the dense omni mask, the full-vocabulary LM head the reference
# computes (models/unigen.py:287), the slice (:416), the decode step, re-embedding.  Nothing is extrapolated unless said so.
CPU_CONFIG = dict(prompts=1, tokens=256, text_len=128, timesteps=18)


class CpuReference:
    def __init__(self, threads: int):
        import torch
        from oracle import t2i_port
        from ml_unigen_b200 import prompting           # synthetic token ids and tokenizer constants only (no kernels)
        from ml_unigen_b200.backbone import QwenConfig
        torch.set_num_threads(threads)
        self.t2i_port, self.prompting, self.torch = t2i_port, prompting, torch
        self.cfg = QwenConfig()
        t0 = time.perf_counter()
        self.w = t2i_port.random_qwen_weights(self.cfg.as_dict(), self.cfg.vocab_size, seed=0)
        self.init_s = time.perf_counter() - t0
        self.threads = threads

    def run(self, N: int, max_steps=None, seed: int = 7, on_step=None):
        """One generation (or its first `max_steps` MaskGIT steps) for ONE prompt; returns the per-step seconds."""
        cfg, P = self.cfg, self.prompting
        ids = P.synthetic_prompt_batch(1, N, pads=[100], seed=3, mask_token_id=cfg.vocab_size - 1)
        _, recs = self.t2i_port.t2i_generate_port(
            self.w, cfg.as_dict(), ids, image_token_num_per_image=N, codebook_size=8192, text_vocab_size=P.QWEN_TEXT_VOCAB,
            mask_token_id=cfg.vocab_size - 1, pad_id=P.QWEN_PAD_ID, soi_id=P.QWEN_SOI_ID, timesteps=CPU_CONFIG["timesteps"],
            temperature=1.0, seed=seed, max_steps=max_steps, on_step=on_step)
        return [r["seconds"] for r in recs]


def reference_decode_ops_cpu(B: int, N: int = 1024, V: int = 8192, threads=None):
    """The reference's decode op sequence (models/unigen.py:418-445 + models/sampling.py:41-46) with torch on the host cores, on
    synthetic bf16 logits.  Returns seconds for one step."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    logits = (0.02 * torch.randn(B, N, V, generator=g)).bfloat16()
    ids = torch.full((B, N), 159866, dtype=torch.int64)
    t0 = time.perf_counter()
    probs = logits.float().softmax(dim=-1)
    sampled = torch.multinomial(probs.reshape(-1, V), 1, generator=g)[:, 0].view(B, N)
    unknown = ids == 159866
    sampled = torch.where(unknown, sampled, ids)
    sel = torch.gather(probs, -1, sampled[..., None]).squeeze(-1)
    sel = torch.where(unknown, sel, torch.finfo(sel.dtype).max)
    mask_len = torch.max(torch.tensor([1]), torch.min(unknown.sum(dim=-1, keepdim=True) - 1, torch.tensor([1020.0])))
    u = torch.zeros_like(sel).uniform_(0, 1, generator=g)
    conf = torch.log(sel.clamp(min=1e-20)) + 1.0 * (-torch.log((-torch.log(u.clamp(min=1e-20))).clamp(min=1e-20)))
    cut = torch.gather(torch.sort(conf, dim=-1).values, 1, mask_len.long())
    masking = conf < cut
    torch.where(masking, 159866, sampled)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_start = time.perf_counter()
    threads = os.cpu_count() or 1
    ref = CpuReference(threads)
    N, T = CPU_CONFIG["tokens"], CPU_CONFIG["timesteps"]
    L = 139 + N
    K, W = max(1, args.steps), max(0, args.warmup)
    # one bench "step" of this arm = ONE real MaskGIT step of consecutive real generations (a bounded sample of the workload, as the
    # contract asks); the first T timed steps are one COMPLETE generation, whose measured wall time is the reported value
    if W:
        ref.run(N, max_steps=min(W, T), seed=5)
    t0 = time.perf_counter()
    secs = []
    gen_index = 0
    while len(secs) < K:
        secs += ref.run(N, max_steps=min(T, K - len(secs)), seed=7 + gen_index)
        gen_index += 1
    wall = time.perf_counter() - t0
    complete = K >= T
    gen_s = sum(secs[:T]) if complete else None
    ips = 1.0 / gen_s if complete else 1.0 / (T * sum(secs) / len(secs))
    sample = (f"config 1 (BASELINE.json configs[0]): 1 prompt, {N} image tokens (L={L}), fp32, {threads} host threads; "
              f"{K} real MaskGIT steps timed one by one; " +
              (f"the first {T} are one complete generation = {gen_s:.2f} s measured" if complete else
               f"fewer than {T} steps requested: images/s = 1 / ({T} x mean step time), EXTRAPOLATED"))
    extra = {}
    if time.perf_counter() - t_start < 110 and not args.no_config2_sample:
        # like-for-like with the GPU arm's shape, clearly an extrapolation: one MaskGIT step of ONE prompt at N=1024 (L=1163)
        s2 = ref.run(1024, max_steps=1, seed=9)[0]
        extra["config2_sample"] = {"what": "1 prompt x 1 of 18 MaskGIT steps at 1024 image tokens (L=1163), fp32 CPU port",
                                   "seconds": s2, "images_per_s_extrapolated": 1.0 / (s2 * T)}
    dec = {f"B{b}": reference_decode_ops_cpu(b, threads=threads) for b in (1, 16)}
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": wall / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE.json configs[0]: UniGen-1.5B random-init, 1 prompt, {N} image tokens (16x16, L={L}), {T} MaskGIT "
                               "steps, fp32 on host CPU -- the reference's own CPU-runnable case (the GPU arm's configs[1] would take hours here)",
                   "step": "one bench step = one real MaskGIT step (1/18 of an image); steps run back to back inside real generations",
                   "reference_path": "oracle/t2i_port.py: CPU port of models/unigen.py:344-455 + transformers Qwen2 (torch fp32, dense omni mask, "
                                     "full-vocabulary head, C decode step), pinned end to end to the unmodified reference by tests/golden/t2i_tiny.npz; "
                                     "/root/reference itself cannot travel to the GPU box",
                   "weights_init_s": ref.init_s},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "seconds_per_image": gen_s, "seconds_per_maskgit_step": sum(secs) / len(secs), "extrapolated": not complete,
                         "reference_decode_ops_s": dec},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def attention_flops(P, S, L, N, Hq=12, D=128):
    """Algorithmic FLOPs of one omni-attention launch: 4*D*Hq*sum_b pairs_b (allowed (q,k) pairs only)."""
    tot = 0
    for p, s in zip(P, S):
        tot += p * (p + 1) // 2 + (s - p) * (s - p + 1) // 2 + (N + 2) * (L - p)
    return 4 * D * Hq * tot


def _traffic(key):
    """DRAM bytes per launch from the committed `ncu --set full` capture of THIS round (profiles/r2_traffic.json, written by
    tools/ncu_summary.py --traffic); null when the capture is missing -- never a constant carried over from an older build."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)
        return t[key]["dram_bytes"], t[key].get("source", "profiles/r2_traffic.json")
    except Exception:
        return None, "no ncu capture of this kernel committed for this round"


def reference_decode_ops_gpu(torch, dev, B, N=1024, V=8192, iters=5):
    """The reference's decode op sequence (models/unigen.py:418-445 + models/sampling.py:41-46) as torch executes it on this GPU
    (softmax, multinomial with its host-syncing checks, gather, where, log, uniform_, sort, gather, <, where), ms per step."""
    g = torch.Generator(device=dev).manual_seed(0)
    logits = (0.02 * torch.randn(B, N, V, device=dev, generator=g)).bfloat16()
    ids = torch.full((B, N), 159866, dtype=torch.int64, device=dev)

    def step():
        probs = logits.float().softmax(dim=-1)
        sampled = torch.multinomial(probs.reshape(-1, V), 1, generator=g)[:, 0].view(B, N)
        unknown = ids == 159866
        sampled = torch.where(unknown, sampled, ids)
        sel = torch.gather(probs, -1, sampled[..., None]).squeeze(-1)
        sel = torch.where(unknown, sel, torch.finfo(sel.dtype).max)
        mask_len = torch.max(torch.tensor([1], device=dev), torch.min(unknown.sum(dim=-1, keepdim=True) - 1, torch.tensor([1020.0]).to(dev)))
        u = torch.zeros_like(sel).uniform_(0, 1, generator=g)
        conf = torch.log(sel.clamp(min=1e-20)) + 1.0 * (-torch.log((-torch.log(u.clamp(min=1e-20))).clamp(min=1e-20)))
        cut = torch.gather(torch.sort(conf, dim=-1).values, 1, mask_len.long())
        return torch.where(conf < cut, 159866, sampled)

    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


def isolated_decode_points(torch, dev, ops, hbm_peak, N=1024, V=8192, batches=(16, 64, 128), iters=20):
    """BASELINE.json configs[2] in brief (the full sweep: tools/bench_decode.py -> profiles/r2_decode_sweep.json): the decode step
    alone on synthetic flat logits (0.02 * randn: what a random-init model produces), all rows masked, in-kernel Philox, CUDA events
    around both launches, logits larger than L2 or L2 flushed between iterations."""
    out = []
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for Bq in batches:
        g = torch.Generator(device=dev).manual_seed(0)
        logits = (0.02 * torch.randn(Bq, N, V, device=dev, generator=g)).bfloat16()
        ws = ops.DecodeWorkspace(Bq, N, V, dev)
        ids0 = torch.full((Bq, N), 159866, dtype=torch.int64, device=dev)
        ids = ids0.clone()
        ms = []
        for it in range(iters + 3):
            ids.copy_(ids0)
            if Bq * N * V * 2 < 300e6:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(300000)              # the GPU is busy while the host enqueues: no launch gap inside the events
            e0.record()
            ops.decode_step(logits, ids, mask_token_id=159866, mask_len_cap=1020, temperature=0.5, workspace=ws, seed=1, offset=it)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ms.append(e0.elapsed_time(e1))
        ms.sort()
        t = ms[len(ms) // 2] * 1e-3
        out.append({"B": Bq, "us": t * 1e6, "GB/s": Bq * N * V * 2 / t / 1e9, "frac": Bq * N * V * 2 / t / 1e9 / hbm_peak})
        del logits, ws
    del flush
    return out


def isolated_attention_point(torch, dev, ops, peak_tf, B=16, S=137, N=1024, Hq=12, Hkv=2, D=128, iters=28):
    """BASELINE.json configs[3] in brief (the full sweep: tools/bench_attention.py -> profiles/r2_attention_sweep.json): the
    per-step window pass of config 2 (N + 2 query rows x all L keys, P = 0) alone, 28 launches back to back, CUDA events."""
    L = S + N + 2
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn(B, L - S, Hq, D, device=dev, generator=g).bfloat16()
    k = torch.randn(B, L, Hkv, D, device=dev, generator=g).bfloat16()
    v = torch.randn(B, L, Hkv, D, device=dev, generator=g).bfloat16()
    P = torch.zeros(B, dtype=torch.int32, device=dev)
    St = torch.full((B,), S, dtype=torch.int32, device=dev)
    o = torch.empty_like(q)
    for _ in range(5):
        ops.omni_attention(q, k, v, P, St, D ** -0.5, out=o, q_offset=S)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.omni_attention(q, k, v, P, St, D ** -0.5, out=o, q_offset=S)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / iters
    fl = 4 * D * Hq * B * (N + 2) * L
    return {"ms": t * 1e3, "TFLOP/s": fl / t / 1e12, "frac": fl / t / 1e12 / peak_tf}


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
    bb = model.backbone
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

    def timed(K, from_host, sampler=None, fn=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(K):
            (fn or one_generation)(1000 + k, from_host)
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

    # instrumented region (not the headline number): the same K generations with the launches of the roofline kernels
    # bracketed by CUDA events; the events keep the per-step pass out of its CUDA graph, so this region runs eagerly
    bb.attn_events, bb.mlp_events, bb.plain_events = [], [], []
    model.decode_events = []
    t_instr, _ = timed(min(args.steps, 3), False)
    n_instr = min(args.steps, 3)
    attn_all = [(a.elapsed_time(b), lq) for a, b, lq in bb.attn_events]
    attn_ms = [t for t, lq in attn_all if lq != L]      # the per-step launches (image-block rows only)
    attn_full_ms = [t for t, lq in attn_all if lq == L] # the step-0 launches over all rows
    dec_ms = [a.elapsed_time(b) for a, b in model.decode_events]
    mlp_all = [(a.elapsed_time(b), m) for a, b, m in bb.mlp_events]
    res_all = [(a.elapsed_time(b), m, k) for a, b, m, k in bb.plain_events]
    bb.attn_events = bb.mlp_events = bb.plain_events = None
    model.decode_events = None
    # timed region 1 (`value`): prompts resident in HBM, no instrumentation; own kernel launches are COUNTED (bb.launch_count)
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = bb.launch_count
    t_dev, clocks = timed(args.steps, False, sampler)
    launches = bb.launch_count - n0
    # timed region 2 (`e2e`): end to end through the public API from pinned host memory
    t_e2e, _ = timed(args.steps, True)

    # BASELINE.json configs[4] (Best-of-N): 64 candidates of ONE prompt split over the ranks (strong scaling: 64 / G per GPU),
    # Philox streams keyed by the global candidate index, one all-gather of ids + scores, top-4 kept
    bon = None
    if not args.no_best_of_n:
        n_cand = 64
        one_prompt = all_ids[0].to(dev)

        def bon_round(k, _):
            ugd.best_of_n(model, one_prompt, n_cand, topk=4, philox_seed=99, philox_offset=k * 64, **gen_kw)

        bon_round(0, False)
        t_bon, _ = timed(2, False, fn=bon_round)
        gather_ms = None
        if world > 1:
            cs, ce = ugd.shard_range(n_cand, rank, world)
            ids_l = torch.zeros(ce - cs, N, dtype=torch.int64, device=dev)
            sc_l = torch.zeros(ce - cs, device=dev)
            ugd.gather_ids_scores(ids_l, sc_l, n_cand)
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(10):
                ugd.gather_ids_scores(ids_l, sc_l, n_cand)
            g1.record()
            torch.cuda.synchronize()
            gather_ms = g0.elapsed_time(g1) / 10
        bon = {"candidates": n_cand, "gpus": world, "candidates_per_gpu": (n_cand + world - 1) // world,
               "seconds_per_prompt": t_bon / 2, "candidates_per_s": n_cand * 2 / t_bon, "all_gather_ms": gather_ms,
               "scaling": "strong", "note": "BASELINE.json configs[4]; score = mean log selected-probability (the reference's CoT-V rater is "
                                            "outside the path); max over ranks, CUDA events"}

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
               "note": "MagvitV2DecoderB200.decode_code (NHWC bf16), random-init weights of the default architecture; not part of `value`"}
        del dec, img

    if rank == 0:
        peaks = _peaks()
        images = gB * args.steps
        flops_full = attention_flops(P_l, S_l, L, N)
        flops = 4 * 128 * 12 * sum((N + 2) * (L - p) for p in P_l) if attn_ms else flops_full   # suffix rows only
        if not attn_ms:
            attn_ms = attn_full_ms
        attn_avg = sum(attn_ms) / len(attn_ms) * 1e-3
        achieved = flops / attn_avg / 1e12
        attn_share = (sum(attn_ms) + sum(attn_full_ms if attn_full_ms is not attn_ms else [])) * 1e-3 / t_instr
        cfgm = bb.cfg
        m_step = (e - s) * (N + 2)                           # rows of the per-step passes (text-prefix K/V reuse)
        mlp_ms = [t for t, m in mlp_all if m == m_step] or [t for t, m in mlp_all]
        m_rows = m_step if any(m == m_step for _, m in mlp_all) else mlp_all[0][1]
        mlp_avg = sum(mlp_ms) / len(mlp_ms) * 1e-3
        mlp_flops = 2.0 * m_rows * 2 * cfgm.intermediate_size * cfgm.hidden_size
        mlp_tf = mlp_flops / mlp_avg / 1e12
        mlp_share = sum(t for t, _ in mlp_all) * 1e-3 / t_instr

        def res_obj(k_dim, name):
            ms = [t for t, m, k in res_all if m == m_rows and k == k_dim]
            if not ms:
                return None
            avg = sum(ms) / len(ms) * 1e-3
            fl = 2.0 * m_rows * cfgm.hidden_size * k_dim
            by = m_rows * (k_dim * 2 + cfgm.hidden_size * (4 + 4 + 2)) + cfgm.hidden_size * k_dim * 2
            return {"kernel": f"gemm_tc_kernel, residual epilogue ({name}: x += bf16(a W^T), xb = bf16(x), sum x^2 per row)",
                    "avg_launch_ms": avg * 1e3, "launches_timed": len(ms), "TFLOP/s": fl / avg / 1e12, "frac_of_sustained_tensor_peak": fl / avg / 1e12 / peaks["tf_sust"],
                    "algorithmic_GB/s": by / avg / 1e9, "frac_of_hbm_peak": by / avg / 1e9 / peaks["hbm"],
                    "share_of_step": sum(t for t, m, k in res_all if k == k_dim) * 1e-3 / t_instr}

        # decode step: the bytes actually streamed (rows of known tokens are never read), per mask state of the schedule
        caps = []
        for st in range(T):
            r = 1.0 * (st + 1) / T
            caps.append(int((N * torch.cos(torch.tensor(r) * 3.141592653589793 * 0.5)).floor().item()))
        unknown = [N] + [max(1, min(N - 1, c)) for c in caps[:-1]]          # tokens still masked when step st starts
        live_bytes = sum(unknown) * (e - s) * V * 2 * n_instr
        dec_total = sum(dec_ms) * 1e-3
        dec_avg = dec_total / len(dec_ms)
        dec0 = [t for i, t in enumerate(dec_ms) if i % T == 0]               # step 0: every row is read
        dec0_avg = sum(dec0) / len(dec0) * 1e-3
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            ref = CpuReference(threads)
            n_s = 4
            secs = ref.run(CPU_CONFIG["tokens"], max_steps=n_s)
            mean = sum(secs) / len(secs)
            cpu = {"value": 1.0 / (mean * CPU_CONFIG["timesteps"]), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"BASELINE.json configs[0] (1 prompt, 256 image tokens, L=395, fp32): the first {n_s} of 18 real MaskGIT steps of one "
                             f"generation through oracle/t2i_port.py, {mean:.2f} s per step; images/s = 1 / (18 x mean step) (bounded sample; "
                             "`bench.py --impl reference` times the complete generation)",
                   "seconds_per_maskgit_step": mean, "weights_init_s": ref.init_s,
                   "reference_decode_ops_s": {f"B{b}": reference_decode_ops_cpu(b, threads=threads) for b in (1, 16)}}
            del ref
        iso_dec = iso_att = None
        if world == 1 and not args.no_isolated:
            from ml_unigen_b200 import ops as _ops
            iso_dec = isolated_decode_points(torch, dev, _ops, peaks["hbm"])
            iso_att = isolated_attention_point(torch, dev, _ops, peaks["tf_sust"])
        tr_mlp, src_mlp = _traffic("gemm_swiglu")
        tr_att, src_att = _traffic("omni_attention")
        tr_dec, src_dec = _traffic("decode_fast")
        line = {
            "metric": METRIC, "value": images / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"UniGen-1.5B random-init T2I, batch {B}/GPU, {N} image tokens (L={L}), {T} MaskGIT steps, "
                                   "guidance 0, cosine schedule (BASELINE.json configs[1])",
                       "global_batch": gB, "parallelism": f"replicas x{world}, images sharded, 1 all-gather/generation",
                       "l2": "working set (3.1 GB weights + activations) >> 126 MB L2, no flush needed",
                       "launch": "steps 1..17 of a generation replay one CUDA graph of the per-step backbone pass; the roofline objects "
                                 "come from an instrumented eager pass of %d generations (%.1f ms per generation)" % (n_instr, t_instr / n_instr * 1e3),
                       "gemms": "all five layer GEMMs and the codebook head on the own tcgen05 2-CTA kernel: q/k/v + bias + RoPE + K/V cache "
                                "(ug_qkv_rope_bf16), gate|up + SwiGLU (ug_gemm_bf16), o_proj / down_proj + residual add + RMSNorm statistics "
                                "(ug_gemm_residual_bf16); RMSNorm folded into the consumer GEMMs; no cuBLAS call and no add+RMSNorm pass in the loop"},
            "roofline": {"kernel": "gemm_tc_kernel, SwiGLU epilogue (tcgen05 cta_group::2): gate_proj | up_proj + folded RMSNorm + silu(g)*u in one launch",
                         "bound": "tensor", "achieved": mlp_tf, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                         "frac": mlp_tf / peaks["tf_sust"], "traffic": tr_mlp, "traffic_note": src_mlp,
                         "peak_source": peaks["source"] + " bf16_tflops_sustained (burst: %.1f)" % peaks["tf_burst"],
                         "flops_per_launch": mlp_flops, "avg_launch_ms": mlp_avg * 1e3, "launches_timed": len(mlp_ms),
                         "share_of_step": mlp_share,
                         "note": "2*M*N*K with M = 16 x 1026 rows, N = 17920, K = 1536; launches of MaskGIT steps 1..17"},
            "roofline_attention": {"kernel": "omni_attention_kernel (tcgen05)", "bound": "tensor", "achieved": achieved,
                         "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sust"], "traffic": tr_att,
                         "traffic_note": src_att,
                         "peak_source": peaks["source"] + " bf16_tflops_sustained", "flops_per_launch": flops,
                         "avg_launch_ms": attn_avg * 1e3, "launches_timed": len(attn_ms), "share_of_step": attn_share,
                         "note": "launches of MaskGIT steps 1..17 (query rows = the N+2 image-block rows, keys = all rows; "
                                 "text-prefix K/V reuse); the 28 step-0 launches over all L rows are timed separately",
                         "step0_full_avg_launch_ms": (sum(attn_full_ms) / len(attn_full_ms)) if attn_full_ms else None,
                         "step0_full_flops_per_launch": flops_full,
                         "isolated_window_pass": iso_att,
                         "isolated_note": "the same launch alone (28 back to back, P = 0): no GEMMs holding the part at its power cap, SM clock "
                                          "near 1.9 GHz instead of ~1.4 GHz; BASELINE.json configs[3]"},
            "roofline_decode": {"kernel": "decode_fast_kernel + select_kernel (one decode step)", "bound": "hbm",
                                "achieved": live_bytes / dec_total / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                "frac": live_bytes / dec_total / 1e9 / peaks["hbm"], "traffic": tr_dec, "traffic_note": src_dec,
                                "avg_launch_ms": dec_avg * 1e3, "launches_timed": len(dec_ms),
                                "step0_GBs": (e - s) * N * V * 2 / dec0_avg / 1e9, "step0_frac": (e - s) * N * V * 2 / dec0_avg / 1e9 / peaks["hbm"],
                                "step0_ms": dec0_avg * 1e3,
                                "algorithmic_GBs_crediting_skipped_rows": (e - s) * N * V * 2 / dec_avg / 1e9,
                                "reference_ops_same_gpu_ms": reference_decode_ops_gpu(torch, dev, e - s),
                                "isolated_flat_step0": iso_dec,
                                "isolated_note": "BASELINE.json configs[2]: the decode step alone, synthetic flat logits, all rows masked; the fixed "
                                                 "~22 us of a launch pair weigh less as B grows (DESIGN.md 4.1)",
                                "note": "`achieved` = bytes actually streamed (rows of still-masked tokens x V x 2 B, summed over the 18 mask states "
                                        "of the schedule) / time of all decode steps; step0_* = the all-rows state alone; reference_ops = the reference's "
                                        "torch op sequence (models/unigen.py:418-445) on this GPU at the same batch, flat logits"},
            "residual_gemms": {"o_proj": res_obj(cfgm.num_attention_heads * cfgm.head_dim, "o_proj"),
                               "down_proj": res_obj(cfgm.intermediate_size, "down_proj")},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e": {"value": images / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(host_ids.numel() * 8),
                    "d2h_bytes_per_step": int(host_out.numel() * 8)},
            "gpu_launches": launches,
            "gpu_launches_note": "counted: every call into libunigen_b200 inside the `value` region increments a counter (graph replays add the "
                                 "launches recorded when the graph's body ran eagerly)",
        }
        if bon is not None:
            line["best_of_n"] = bon
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
    ap.add_argument("--no-best-of-n", action="store_true", help="skip the Best-of-N (configs[4]) measurement after the measured path")
    ap.add_argument("--no-config2-sample", action="store_true", help="reference arm: skip the 1-step sample at the GPU arm's shape")
    ap.add_argument("--no-isolated", action="store_true", help="skip the isolated decode-step / attention points (configs[2] / configs[3])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
