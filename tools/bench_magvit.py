"""MAGViT-v2 pixel decoder at the size that follows config 2: 16 token grids of 32x32 ->
16 images of 512x512, default architecture, random-init weights.  Times `MagvitV2DecoderB200.decode_code` with CUDA
events and, beside it, the same decoder as plain torch eager modules in bf16 NCHW (what the reference executes)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ml_unigen_b200.magvit import MagvitV2DecoderB200  # noqa: E402
from oracle import magvit_oracle as MO  # noqa: E402


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    B, N = int(os.environ.get("B", 16)), int(os.environ.get("N", 1024))
    dec = MagvitV2DecoderB200.random_init(dev, seed=0)
    ids = torch.randint(0, 8192, (B, N), device=dev)
    t_own = timed(lambda: dec.decode_code(ids), 5)
    sd_bf16 = {k: v.bfloat16().contiguous() for k, v in dec.raw.items()}

    def eager():
        with torch.no_grad():
            z = MO.lfq_codebook_entry(ids).bfloat16()
            return MO.decoder_forward(sd_bf16, dec.num_res_blocks, z)
    t_ref = timed(eager, 3)
    out = dec.decode_code(ids)
    ref = eager().float()
    rel = ((out - ref).norm() / ref.norm()).item()
    row = dict(B=B, N=N, pixels=f"{out.shape[-2]}x{out.shape[-1]}", ms_own=t_own, images_per_s_own=B / t_own * 1e3,
               ms_torch_eager_bf16_nchw=t_ref, speedup=t_ref / t_own, rel_l2_vs_eager_bf16=rel)
    print(json.dumps(row))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(row, open(os.path.join(ROOT, "gpurun_out", "magvit_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
