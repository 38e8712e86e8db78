"""Caller-side formats: prompt files, per-rank prompt assignment, output layout and the
CoT-V selection, against results recorded from the reference's own `greedy_samples_byscore`
(tests/golden/callers.json, oracle/make_golden_callers.py)."""
import json
import os

import numpy as np
import pytest
import torch

from ml_unigen_b200 import callers


def test_greedy_samples_byscore_matches_reference_golden(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "callers.json")))
    assert len(cases) >= 5
    for c in cases:
        picked, stats = callers.greedy_samples_byscore(c["table"], topk=c["topk"])
        assert picked == c["picked"]
        assert [(s["instance_id"], s["image_id"], s["yes_count"], s["total_questions"]) for s in stats] == \
               [(s["instance_id"], s["image_id"], s["yes_count"], s["total_questions"]) for s in c["stats"]]
        assert np.allclose([s["score"] for s in stats], [s["score"] for s in c["stats"]], rtol=0, atol=1e-12)


def test_greedy_samples_byscore_edge_cases():
    assert callers.greedy_samples_byscore([], topk=4) == ([], [])
    rows = [{"image_id": "00001/samples/00000.png", "answer": "no"}, {"image_id": "00001/samples/00001.png", "answer": "no"}]
    picked, stats = callers.greedy_samples_byscore(rows, topk=1)          # tie: the lower image id is kept
    assert picked == ["00001/samples/00000.png"] and [s["score"] for s in stats] == [0.0, 0.0]
    picked, _ = callers.greedy_samples_byscore(rows, topk=10)              # fewer images than topk
    assert picked == ["00001/samples/00000.png", "00001/samples/00001.png"]


def test_prompt_file_and_rank_assignment(tmp_path):
    p = tmp_path / "prompts.jsonl"
    metas = [{"tag": "single_object", "include": [{"class": "bench", "count": 1}], "prompt": f"a photo of thing {i}"} for i in range(7)]
    p.write_text("\n".join(json.dumps(m) for m in metas) + "\n\n")
    got = callers.load_prompt_metadata(str(p))
    assert got == metas
    # reference: round b gives metadata b*world + r to rank r (inference_geneval.py:188-193)
    world = 3
    seen = {}
    for r in range(world):
        for idx, m in callers.iter_rank_prompts(got, r, world):
            assert idx % world == r and m == metas[idx]
            seen[idx] = r
    assert sorted(seen) == list(range(7))
    (tmp_path / "bad.jsonl").write_text('{"text": "no prompt key"}\n')
    with pytest.raises(ValueError):
        callers.load_prompt_metadata(str(tmp_path / "bad.jsonl"))
    with pytest.raises(ValueError):
        list(callers.iter_rank_prompts(got, 3, 3))


def test_output_layout_and_samples(tmp_path):
    out = tmp_path / callers.output_path_name(50, 6)
    assert out.name == "t2i_samples_step50_scale6"
    meta = {"prompt": "a photo of a bench", "tag": "single_object"}
    sp = callers.prepare_sample_dir(str(out), 521, meta)
    assert sp == str(out / "00521" / "samples")
    assert json.load(open(out / "00521" / "metadata.jsonl")) == meta
    ids = torch.tensor([[5, 9000, -3, 8191], [0, 1, 2, 3]])
    clamped = callers.clamp_token_ids(ids, 8192)
    assert clamped.tolist() == [[5, 8191, 0, 8191], [0, 1, 2, 3]]
    imgs = torch.zeros(2, 3, 4, 4)
    imgs[0] = 1.0
    imgs[1, 0] = -1.0
    imgs[1, 1] = 0.999                       # (0.999 + 1) / 2 * 255 = 254.87 -> 254: truncation like the reference
    image_ids = callers.save_samples(sp, clamped, images=imgs, first_index=4)
    assert image_ids == ["00521/samples/00004.png", "00521/samples/00005.png"]
    assert np.array_equal(np.load(os.path.join(sp, "00005.tokens.npy")), np.array([0, 1, 2, 3]))
    from PIL import Image
    a = np.asarray(Image.open(os.path.join(sp, "00004.png")))
    b = np.asarray(Image.open(os.path.join(sp, "00005.png")))
    assert a.shape == (4, 4, 3) and (a == 255).all()
    assert (b[..., 0] == 0).all() and (b[..., 1] == 254).all() and (b[..., 2] == 127).all()
    # token grids alone (no pixel decoder on the path)
    ids2 = callers.save_samples(sp, clamped)
    assert ids2[0].endswith("00000.png") and os.path.exists(os.path.join(sp, "00000.tokens.npy"))


def test_select_topk_candidates():
    ids = torch.arange(6 * 5).view(6, 5)
    scores = torch.tensor([0.1, 0.9, 0.9, -1.0, 0.5, 0.9])
    t, s, idx = callers.select_topk_candidates(ids, scores, topk=4)
    assert idx.tolist() == [1, 2, 5, 4] and torch.equal(t, ids[idx]) and s.tolist() == scores[idx].tolist()
    t, s, idx = callers.select_topk_candidates(ids[:2], scores[:2], topk=4)
    assert idx.tolist() == [1, 0]


def test_pack_gate_up_layout():
    """Host-side weight packing of the fused gate|up + SwiGLU GEMM (ml_unigen_b200.ops.pack_gate_up): block 2j of 128 rows
    = gate rows [128j, 128j+128), block 2j+1 = the same rows of up (include/unigen_b200.h, UG_GEMM_SWIGLU)."""
    from ml_unigen_b200 import ops
    I, K = 384, 16
    wg = torch.arange(I * K, dtype=torch.float32).view(I, K)
    wu = -torch.arange(I * K, dtype=torch.float32).view(I, K) - 1
    w = ops.pack_gate_up(wg, wu)
    assert w.shape == (2 * I, K) and w.is_contiguous()
    for j in range(I // 128):
        assert torch.equal(w[256 * j: 256 * j + 128], wg[128 * j: 128 * j + 128])
        assert torch.equal(w[256 * j + 128: 256 * j + 256], wu[128 * j: 128 * j + 128])
    with pytest.raises(ValueError):
        ops.pack_gate_up(wg[:100], wu[:100])
    with pytest.raises(ValueError):
        ops.pack_gate_up(wg, wu[:256])


def test_kernels_refuse_cpu_tensors():
    """No CPU fallback: every tensor-level wrapper raises on host tensors instead of computing something else."""
    from ml_unigen_b200 import ops, magvit, _lib
    a = torch.zeros(256, 64, dtype=torch.bfloat16)
    w = torch.zeros(256, 64, dtype=torch.bfloat16)
    with pytest.raises(_lib.UnigenB200Error):
        ops.gemm(a, w)
    with pytest.raises(_lib.UnigenB200Error):
        magvit.lfq_unpack(torch.zeros(1, 16, dtype=torch.int64))
    with pytest.raises(_lib.UnigenB200Error):
        magvit.groupnorm_swish(torch.zeros(1, 4, 4, 32, dtype=torch.bfloat16), torch.ones(32, dtype=torch.bfloat16),
                               torch.zeros(32, dtype=torch.bfloat16))
