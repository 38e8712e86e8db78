"""BASELINE.json config 5: CoT-V Best-of-N -- one prompt, 64 candidates split across the ranks (torchrun),
N=1024, 18 steps; one all_gather_into_tensor of ids + scores; reports candidates/s and the gather latency."""
import json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ml_unigen_b200 import prompting, dist as ugd
from ml_unigen_b200.unigen import UniGenB200

def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_cand, N, T = int(os.environ.get("NCAND", 64)), 1024, 18
    model = UniGenB200.random_init(dev, seed=0)
    prompt = prompting.synthetic_prompt_batch(1, N, seed=5, pads=[40], mask_token_id=model.mask_token_id).to(dev)[0]
    kw = dict(timesteps=T, image_token_num_per_image=N, text_vocab_size=prompting.QWEN_TEXT_VOCAB, philox_seed=99, validate_mask=False)
    ugd.best_of_n(model, prompt, n_cand, topk=4, **kw)                 # warm-up
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ids, scores, top = ugd.best_of_n(model, prompt, n_cand, topk=4, **kw)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # gather latency alone
    s, e = ugd.shard_range(n_cand, rank, world)
    my_ids, my_sc = ids[s:e].contiguous(), scores[s:e].contiguous()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): ugd.gather_ids_scores(my_ids, my_sc, n_cand)
    torch.cuda.synchronize(); g0.record()
    for _ in range(20): ugd.gather_ids_scores(my_ids, my_sc, n_cand)
    g1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"config": f"Best-of-N N={n_cand}, 1024 tokens, 18 steps, {world} GPU(s)", "candidates_per_s": n_cand / float(t),
                          "seconds": float(t), "gather_ms": g0.elapsed_time(g1) / 20, "top4": top.tolist(),
                          "ids_checksum": int(ids.sum()), "n_gpus": world}), flush=True)
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
