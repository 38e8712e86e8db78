"""GPU: the whole caller-visible pipeline of evaluation/inference_geneval.py with the drop-in pieces -- prompt jsonl ->
per-rank prompts -> t2i_gen_prompt -> UniGenB200.t2i_generate (CFG, CUDA-graph step replay) -> clamp ->
MagvitV2DecoderB200.decode_code -> {idx:05}/samples/{k:05}.png + token grids -> rating table -> greedy_samples_byscore.
Small random-init models (the tokenizer and the checkpoints are not available offline); what is checked is the plumbing:
shapes, determinism under a fixed Philox stream, file layout, and that the selected images exist."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.make_golden_backbone import TINY, TEXT_VOCAB, V, PAD, SOI, EOI, tiny_weights

pytestmark = pytest.mark.gpu


def test_geneval_style_pipeline(tmp_path):
    from ml_unigen_b200 import callers, prompting
    from ml_unigen_b200.backbone import QwenBackboneB200, QwenConfig
    from ml_unigen_b200.magvit import MagvitV2DecoderB200
    from ml_unigen_b200.unigen import UniGenB200
    dev = torch.device("cuda:0")
    # prompts file (GenEval format) and the per-rank assignment of a 2-process run
    metas = [{"tag": "single_object", "prompt": f"a photo of object {i}"} for i in range(3)]
    pfile = tmp_path / "prompts.jsonl"
    pfile.write_text("\n".join(json.dumps(m) for m in metas))
    loaded = callers.load_prompt_metadata(str(pfile))
    mine = list(callers.iter_rank_prompts(loaded, rank=0, world=2))
    assert [i for i, _ in mine] == [0, 2]
    # models: tiny Qwen2 backbone (V = 4096 codebook) and a small MAGViT-v2 decoder; 12-bit tokens -> codebook_dim 12
    w = {k: torch.from_numpy(v) for k, v in tiny_weights().items()}
    model = UniGenB200(QwenBackboneB200(QwenConfig(vocab_size=TEXT_VOCAB + V + 1, **TINY), w, dev), codebook_size=V,
                       pad_id=PAD, soi_id=SOI)
    dec = MagvitV2DecoderB200.random_init(dev, seed=1, ch=32, ch_mult=(1, 2), num_res_blocks=(1, 1), z_channels=12)
    N, n_samples, T = 16, 4, 6
    out_dir = tmp_path / callers.output_path_name(T, 2)
    rating_table = []
    rng = np.random.default_rng(0)
    for index, meta in mine:
        # "tokenised" prompt: synthetic text ids (no tokenizer offline), left-padded by t2i_gen_prompt like the reference
        text = [list(rng.integers(4, TEXT_VOCAB, size=5 + index)) for _ in range(n_samples)]
        img = torch.full((n_samples, N), model.mask_token_id, dtype=torch.int64)
        kw = dict(conv_start_ids=[4, 5], conv_end_ids=[6], pad_id=PAD, soi_id=SOI, eoi_id=EOI, max_len=10)
        ids, _ = prompting.t2i_gen_prompt(text, img, **kw)
        uncond, _ = prompting.t2i_gen_prompt([[] for _ in range(n_samples)], img, **kw)
        assert ids.shape == uncond.shape
        gen = dict(input_ids=ids.to(dev), uncond_input_ids=uncond.to(dev), guidance_scale=2, timesteps=T,
                   image_token_num_per_image=N, text_vocab_size=TEXT_VOCAB, philox_seed=11 + index)
        tok = model.t2i_generate(**gen)
        tok2 = model.t2i_generate(**gen)                       # second call: the cached step graph, same Philox stream
        assert torch.equal(tok, tok2)
        tok = callers.clamp_token_ids(tok, V)
        images = dec.decode_code(tok)
        assert images.shape == (n_samples, 3, 8, 8) and torch.isfinite(images).all()
        sample_path = callers.prepare_sample_dir(str(out_dir), index, meta)
        image_ids = callers.save_samples(sample_path, tok, images=images)
        for k, image_id in enumerate(image_ids):
            assert os.path.exists(os.path.join(str(out_dir), image_id))
            assert np.array_equal(np.load(os.path.join(sample_path, f"{k:05}.tokens.npy")), tok[k].cpu().numpy())
            # stand-in for the CoT-V rater: k+1 of 5 questions answered "yes" for sample k
            rating_table += [{"image_id": image_id, "answer": "yes" if q <= k else "no"} for q in range(5)]
        assert json.load(open(os.path.join(str(out_dir), f"{index:05}", "metadata.jsonl"))) == meta
    picked, stats = callers.greedy_samples_byscore(rating_table, topk=2)
    assert picked == ["00000/samples/00003.png", "00000/samples/00002.png", "00002/samples/00003.png", "00002/samples/00002.png"]
    assert all(os.path.exists(os.path.join(str(out_dir), p)) for p in picked)
