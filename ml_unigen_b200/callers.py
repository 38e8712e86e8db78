"""Caller-side formats around the T2I path, so that the reference's external scorers
(GenEval, DPG-Bench) and its CoT-V selection consume this path's output unchanged.

Mirrors, without their model / accelerate / PIL plumbing:
* prompt files: one JSON object per line with a `prompt` key (evaluation/inference_geneval.py:152-153,
  inference_dpg.py:145);
* one prompt per process per round: rank r takes metadata index `batch_index + r`
  (evaluation/inference_geneval.py:188-193, `PartialState.split_between_processes`);
* output layout `<out>/t2i_samples_step{T}_scale{g}/{index:05}/metadata.jsonl` and `.../samples/{k:05}.png`
  (evaluation/inference_geneval.py:184-203, 266-276); the token grid that produced sample k is stored beside it as
  `{k:05}.tokens.npy` (the MAGViT-v2 pixel decoder is outside the measured path);
* CoT-V selection: `greedy_samples_byscore(rating_table, topk)` (evaluation/inference_unigen_cot.py:507-534) -- per
  instance (the first component of `image_id`), the `topk` images with the largest fraction of "yes" answers -- and
  its on-device counterpart `select_topk_candidates` for the Best-of-N path, where the scores arrive through the
  all-gather of `ml_unigen_b200.dist` instead of a pickled `all_gather_object` (inference_unigen_cot.py:417-427).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np


def load_prompt_metadata(path: str) -> List[dict]:
    """GenEval / DPG-Bench prompt file: one JSON object per non-empty line, each with a `prompt` key."""
    out = []
    with open(path) as fp:
        for ln, line in enumerate(fp, 1):
            if not line.strip():
                continue
            obj = json.loads(line)
            if "prompt" not in obj:
                raise ValueError(f"{path}:{ln}: no 'prompt' key")
            out.append(obj)
    return out


def output_path_name(generation_timesteps: int, guidance_scale) -> str:
    """Default run directory name (`config.training.img_log_path` overrides it in the reference)."""
    return f"t2i_samples_step{generation_timesteps}_scale{guidance_scale}"


def iter_rank_prompts(metadatas: Sequence[dict], rank: int, world: int) -> Iterator[Tuple[int, dict]]:
    """(index, metadata) pairs of this process: round b hands metadata b*world + r to rank r; ranks past the end of the
    list sit the last round out (they still take part in the barrier after it)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    for batch_index in range(0, len(metadatas), world):
        index = batch_index + rank
        if index < len(metadatas):
            yield index, metadatas[index]


def prepare_sample_dir(output_dir: str, index: int, metadata: dict) -> str:
    """Creates `<output_dir>/{index:05}/samples/` and writes `metadata.jsonl`; returns the samples directory."""
    outpath = os.path.join(output_dir, f"{index:0>5}")
    sample_path = os.path.join(outpath, "samples")
    os.makedirs(sample_path, exist_ok=True)
    with open(os.path.join(outpath, "metadata.jsonl"), "w") as fp:
        json.dump(metadata, fp)
    return sample_path


def clamp_token_ids(gen_token_ids, codebook_size: int):
    """`torch.clamp(gen_token_ids, max=codebook_size - 1, min=0)` (evaluation/inference_geneval.py:266)."""
    return gen_token_ids.clamp(min=0, max=codebook_size - 1)


def save_samples(sample_path: str, token_ids, images=None, first_index: int = 0) -> List[str]:
    """Writes sample k as `{k:05}.tokens.npy` (int64 token grid, always) and, when `images` (uint8 [B,H,W,3] or float
    [B,3,H,W] in [-1,1], the decoder's output) is given, as `{k:05}.png` exactly like the reference
    (`clamp((x+1)/2, 0, 1) * 255`, `astype(uint8)`: truncation, not rounding).  Returns the image ids
    `{index:05}/samples/{k:05}.png` the scorers and `greedy_samples_byscore` use."""
    toks = np.asarray(token_ids.cpu() if hasattr(token_ids, "cpu") else token_ids)
    if toks.ndim != 2:
        raise ValueError("token_ids must be [B, N]")
    arr = None
    if images is not None:
        arr = images.detach().cpu().numpy() if hasattr(images, "detach") else np.asarray(images)
        if arr.dtype != np.uint8:
            arr = np.clip((arr.astype(np.float32) + 1.0) / 2.0, 0.0, 1.0) * 255.0
            arr = arr.transpose(0, 2, 3, 1).astype(np.uint8)
        if arr.shape[0] != toks.shape[0]:
            raise ValueError("images and token_ids disagree on the batch size")
    ids = []
    inst = os.path.basename(os.path.dirname(os.path.normpath(sample_path)))
    for b in range(toks.shape[0]):
        k = first_index + b
        np.save(os.path.join(sample_path, f"{k:05}.tokens.npy"), toks[b].astype(np.int64))
        if arr is not None:
            from PIL import Image
            Image.fromarray(arr[b]).save(os.path.join(sample_path, f"{k:05}.png"))
        ids.append(f"{inst}/samples/{k:05}.png")
    return ids


def greedy_samples_byscore(rating_table: Iterable[Dict[str, str]], topk: int = 4):
    """rating_table rows: {'image_id': '00521/samples/00000.png', 'answer': 'yes' | 'no', ...} (one row per question).
    score(image) = #yes / #questions; per instance (image_id up to the first '/') the `topk` best images, instances in
    ascending order.  Returns (image_ids, stats) with stats = list of dicts (instance_id, image_id, yes_count,
    total_questions, score) ordered by (instance_id, image_id) like the reference's groupby.  Ties in the score keep
    the (instance_id, image_id) order (the reference's pandas `sort_values` uses an unstable sort there, so any order
    of tied images is a valid reference result)."""
    agg: Dict[Tuple[str, str], List[int]] = {}
    for row in rating_table:
        image_id = row["image_id"]
        key = (image_id.split("/")[0], image_id)
        a = agg.setdefault(key, [0, 0])
        a[0] += 1 if row["answer"] == "yes" else 0
        a[1] += 1
    stats = [dict(instance_id=k[0], image_id=k[1], yes_count=v[0], total_questions=v[1], score=v[0] / v[1])
             for k, v in sorted(agg.items())]
    picked: List[str] = []
    order = sorted(range(len(stats)), key=lambda i: (stats[i]["instance_id"], -stats[i]["score"]))    # stable
    taken: Dict[str, int] = {}
    for i in order:
        inst = stats[i]["instance_id"]
        if taken.get(inst, 0) < topk:
            taken[inst] = taken.get(inst, 0) + 1
            picked.append(stats[i]["image_id"])
    return picked, stats


def select_topk_candidates(token_ids, scores, topk: int = 4):
    """Best-of-N on the device: token_ids int64 [n_cand, N], scores fp32 [n_cand] (after `dist.gather_ids_scores`) ->
    (token_ids[topk, N], scores[topk], candidate indices), best first; ties keep the lower candidate index."""
    import torch
    if token_ids.shape[0] != scores.shape[0]:
        raise ValueError("one score per candidate")
    k = min(topk, scores.shape[0])
    order = torch.sort(scores.float(), descending=True, stable=True).indices[:k]
    return token_ids[order], scores[order], order
