"""Multi-GPU sharding of the T2I path: one process per GPU, replicas of the 3.1 GB model, images
split across ranks, exactly ONE collective per generation round.

Reference behaviour being replaced (apple/ml-unigen):
  * evaluation/inference_geneval.py:190-192 -- `PartialState.split_between_processes`: every rank
    generates its own prompts, no tensor collectives;
  * evaluation/inference_unigen_cot.py:146-243, 417-427, 507-534 -- CoT-V Best-of-N: candidates are
    generated in rounds on one rank, rated, the rating dicts travel through a pickled
    `dist.all_gather_object`, and rank 0 keeps the top-4 by score.
Here the N candidates of a prompt (or the prompts of a batch) are split contiguously across ranks,
the Philox streams are keyed by the GLOBAL image index (`batch_offset`), so the generated images do
not depend on the number of GPUs, and token ids + scores come back through one
`all_gather_into_tensor` (NCCL over NVLink; ~0.5 MB at N=64, latency-bound -- there is no
compute step to fuse it with, the decode loop itself has no exchange).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) of `total` items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def candidate_scores(sel_prob: torch.Tensor) -> torch.Tensor:
    """Stand-in score for Best-of-N: mean log selected-probability of the final step, fp32 [B].
    (The reference's score is the CoT-V yes-ratio produced by the understanding branch, which is
    outside this path; any per-candidate fp32 scalar travels the same way.)"""
    p = sel_prob.float()
    p = torch.where(p >= 3.0e38, torch.ones_like(p), p)           # tokens known before the last step
    return torch.log(p.clamp(min=1e-20)).mean(dim=-1)


def gather_ids_scores(ids: torch.Tensor, scores: torch.Tensor, total: int, group=None):
    """All ranks contribute ids int64 [n_r, N] and scores fp32 [n_r]; every rank receives
    (ids [total, N], scores [total]) in global candidate order.  One collective: ids and scores are
    packed into a single int64 buffer (scores bit-cast), padded to the largest shard."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ids, scores
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    N = ids.shape[1]
    per = (total + world - 1) // world
    buf = torch.zeros(per, N + 1, dtype=torch.int64, device=ids.device)
    n = ids.shape[0]
    buf[:n, :N] = ids
    buf[:n, N] = scores.float().view(torch.int32).to(torch.int64)
    out = torch.empty(world * per, N + 1, dtype=torch.int64, device=ids.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    rows = []
    for r in range(world):
        s, e = shard_range(total, r, world)
        rows.append(out[r * per: r * per + (e - s)])
    full = torch.cat(rows)
    return full[:, :N].contiguous(), full[:, N].to(torch.int32).view(torch.float32)


def select_topk(scores: torch.Tensor, k: int = 4) -> torch.Tensor:
    """Indices of the k best candidates (inference_unigen_cot.py:507-534 keeps topk=4), ties by lower index."""
    order = torch.sort(scores, descending=True, stable=True).indices
    return order[: min(k, scores.numel())]


def best_of_n(model, input_ids_one: torch.Tensor, n_candidates: int, *, topk: int = 4, group=None, **gen_kwargs):
    """Generate `n_candidates` images for ONE prompt, sharded over the ranks of `group`; returns
    (all ids [n, N], all scores [n], indices of the top-k) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if "philox_seed" not in gen_kwargs:
        # Sharding invariance needs ONE Philox key for the whole prompt: rank 0 draws it (from `generator` if given, else from its
        # default CUDA generator, which advances) and every rank adopts it -- a rank-local default generator would give each shard
        # its own stream, and the images would depend on the number of GPUs.
        gen = gen_kwargs.pop("generator", None)
        dev = model.device
        key = torch.zeros(2, dtype=torch.int64, device=dev)
        if rank == 0:
            g = gen if gen is not None else torch.cuda.default_generators[dev.index or 0]
            key = torch.randint(0, 2 ** 62, (2,), generator=g, device=g.device).to(dev)
        if world > 1:
            dist.broadcast(key, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        k = key.tolist()
        gen_kwargs["philox_seed"], gen_kwargs["philox_offset"] = int(k[0]), int(k[1] & 0xFFFFFFFF)
    s, e = shard_range(n_candidates, rank, world)
    batch = input_ids_one.reshape(1, -1).expand(e - s, -1).contiguous()
    ids = model.t2i_generate(input_ids=batch, batch_offset=s, **gen_kwargs)
    scores = candidate_scores(model.last_sel_prob)
    all_ids, all_scores = gather_ids_scores(ids, scores, n_candidates, group)
    return all_ids, all_scores, select_topk(all_scores, topk)
