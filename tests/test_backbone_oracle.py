"""CPU: the floating-point oracle (oracle/backbone_oracle.py) and the host-side prompt/mask logic
against golden data recorded from the UNMODIFIED reference (tests/golden/omni_mask.npz,
backbone_tiny.npz) and against transformers' own Qwen2Model."""
import os

import numpy as np
import pytest
import torch

from oracle import backbone_oracle as BO
from oracle.make_golden_backbone import TINY, TEXT_VOCAB, V, N_IMG, PAD, SOI, EOI, tiny_weights


def test_closed_form_mask_equals_reference(golden_dir):
    from ml_unigen_b200 import prompting
    g = np.load(os.path.join(golden_dir, "omni_mask.npz"))
    pad_id, soi_id, eoi_id = int(g["pad_id"]), int(g["soi_id"]), int(g["eoi_id"])
    for ci in range(3):
        ids = torch.from_numpy(g[f"ids_{ci}"])
        allowed_ref = torch.from_numpy(g[f"allowed_{ci}"])
        P, S = prompting.omni_mask_params(ids, pad_id, soi_id)
        L = ids.shape[1]
        assert torch.equal(prompting.omni_allowed(P, S, L), allowed_ref)
        assert torch.equal(BO.omni_allowed(P, S, L), allowed_ref)
        dense = prompting.create_attention_mask_predict_next(ids, pad_id=pad_id, soi_id=soi_id, eoi_id=eoi_id,
                                                             rm_pad_in_image=True)
        assert dense.dtype == torch.float32 and dense.shape == (ids.shape[0], 1, L, L)
        assert torch.equal(dense[:, 0] == 0, allowed_ref)
        assert float(dense.min()) == float(g[f"fill_{ci}"])            # -9.223372e18
        # (P,S) recovered from the dense mask a caller hands to t2i_generate
        P2, S2 = prompting.omni_mask_params_from_dense(dense)
        assert torch.equal(prompting.omni_allowed(P2, S2, L), allowed_ref)
        assert bool(prompting.dense_mask_matches(dense, P2, S2))
        bad = dense.clone()
        bad[0, 0, 3, 2] = 0 if bad[0, 0, 3, 2] != 0 else -1.0
        assert not bool(prompting.dense_mask_matches(bad, P2, S2))


def test_t2i_gen_prompt_layout():
    from ml_unigen_b200 import prompting
    img = torch.full((2, 4), 999, dtype=torch.long)
    ids, m = prompting.t2i_gen_prompt([[10, 11, 12], [20]], img, conv_start_ids=[1, 2], conv_end_ids=[3, 4, 5],
                                      pad_id=0, soi_id=7, eoi_id=8, max_len=4)
    # L = 4 + 2 + 3 + 2 + 4 = 15 (prompting_utils.py:130-133)
    assert ids.shape == (2, 15)
    assert ids[0].tolist() == [0, 1, 2, 10, 11, 12, 3, 4, 5, 7, 999, 999, 999, 999, 8]
    assert ids[1].tolist() == [0, 0, 0, 1, 2, 20, 3, 4, 5, 7, 999, 999, 999, 999, 8]
    assert m[1].tolist() == [0, 0, 0] + [1] * 12
    # over-long text is truncated and conv_end re-appended (prompting_utils.py:138-140)
    ids, m = prompting.t2i_gen_prompt([[10, 11, 12, 13, 14, 15]], img[:1], conv_start_ids=[1, 2], conv_end_ids=[3, 4, 5],
                                      pad_id=0, soi_id=7, eoi_id=8, max_len=2)
    assert ids.shape == (1, 13) and ids[0].tolist() == [1, 2, 10, 11, 3, 4, 5, 7, 999, 999, 999, 999, 8]
    P, S = prompting.omni_mask_params(ids, 0, 7)
    assert P.tolist() == [0] and S.tolist() == [7]
    b = prompting.synthetic_prompt_batch(4, 256, pads=[0, 5, 99, 50])
    assert b.shape == (4, 395)                                          # L = 139 + N
    P, S = prompting.omni_mask_params(b, prompting.QWEN_PAD_ID, prompting.QWEN_SOI_ID)
    assert P.tolist() == [0, 5, 99, 50] and S.tolist() == [137] * 4


def test_t2i_gen_prompt_matches_reference_class(golden_dir):
    """tests/golden/prompting.json: recorded from the reference's own UniversalPromptingQwen2 (`__call__(..., 't2i_gen')` ->
    t2i_gen_prompt, training/prompting_utils.py:113-153, 398-402) on a stub tokenizer (oracle/make_golden_prompting.py): padded,
    un-padded (max_len None), truncated-with-conv_end and task_token_first cases."""
    import json
    from ml_unigen_b200 import prompting
    cases = json.load(open(os.path.join(golden_dir, "prompting.json")))
    assert len(cases) >= 5
    for c in cases:
        ids, m = prompting.t2i_gen_prompt(c["text_ids"], torch.tensor(c["image_ids"]), conv_start_ids=c["conv_start_ids"],
                                          conv_end_ids=c["conv_end_ids"], pad_id=c["pad_id"], soi_id=c["soi_id"], eoi_id=c["eoi_id"],
                                          max_len=c["text_len"], max_seq_len=c["max_seq_len"])
        assert ids.tolist() == c["input_ids"]
        assert m.tolist() == c["mask"]
    # the ids the eval drivers see for the added specials (evaluation/inference_geneval.py:49-52)
    assert cases[0]["soi_id"] == prompting.QWEN_SOI_ID and cases[0]["eoi_id"] == prompting.QWEN_EOI_ID


def test_alternate_schedules_match_reference(golden_dir):
    """linear / pow / sigmoid schedules (models/sampling.py:53-88) against the values recorded from the reference's own functions
    (tests/golden/schedule.npz `sched_*`, oracle/make_golden.py `make_schedule`), bit for bit in float32."""
    from ml_unigen_b200 import sampling
    g = np.load(os.path.join(golden_dir, "schedule.npz"))
    names = [k[len("sched_"):] for k in g.files if k.startswith("sched_")]
    assert sorted(names) == ["linear", "pow0.5", "pow2", "sigmoid"]
    for name in names:
        fn = sampling.get_mask_chedule(name)
        got = np.array([float(fn(torch.tensor(r / 10))) for r in range(1, 11)], np.float32)
        assert np.array_equal(got.view(np.uint32), g[f"sched_{name}"].view(np.uint32)), name
    assert sampling.get_mask_chedule("cosine") is sampling.cosine_schedule
    # cosine: the floor(N * ratio) caps of SURVEY section 0-4, incl. the -1 of the last step
    caps = [int((1024 * sampling.cosine_schedule(torch.tensor(1.0 * (s + 1) / 18))).floor().item()) for s in range(18)]
    assert caps == list(g["cap_1024_18"]) and caps[-1] == -1
    with pytest.raises(ValueError):
        sampling.get_mask_chedule("nope")


def _tiny_torch_weights():
    return {k: torch.from_numpy(v) for k, v in tiny_weights().items()}


def test_oracle_forward_equals_reference_golden(golden_dir):
    """fp32 restatement vs the logits the unmodified reference UniGen produced (same weights)."""
    from ml_unigen_b200 import prompting
    g = np.load(os.path.join(golden_dir, "backbone_tiny.npz"))
    w = _tiny_torch_weights()
    for name in ("masked", "half"):
        ids = torch.from_numpy(g[f"ids_{name}"])
        P, S = prompting.omni_mask_params(ids, PAD, SOI)
        x = w["embed_tokens.weight"][ids]
        hid = BO.qwen2_forward(w, TINY, x, P, S)
        logits = BO.codebook_logits(w, hid[:, -(N_IMG + 1):-1], TEXT_VOCAB, V)
        ref = torch.from_numpy(g[f"logits_{name}"])
        assert (logits - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())


def test_oracle_forward_equals_transformers_qwen2():
    """and against transformers' Qwen2Model directly (eager attention, dense additive mask)."""
    from transformers import Qwen2Config, Qwen2Model
    from ml_unigen_b200 import prompting
    cfg = Qwen2Config(vocab_size=TEXT_VOCAB + V + 1, max_position_embeddings=512, tie_word_embeddings=True,
                      **{k: v for k, v in TINY.items() if k != "head_dim"})
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    m = Qwen2Model(cfg).eval()
    w = _tiny_torch_weights()
    missing, unexpected = m.load_state_dict(w, strict=False)
    assert not unexpected
    ids = torch.randint(4, TEXT_VOCAB, (2, 40))
    ids[0, :6] = PAD
    ids[:, 20] = SOI
    P, S = prompting.omni_mask_params(ids, PAD, SOI)
    mask = BO.omni_additive_mask(P, S, 40)
    with torch.no_grad():
        ref = m(inputs_embeds=w["embed_tokens.weight"][ids], attention_mask=mask).last_hidden_state
    got = BO.qwen2_forward(w, TINY, w["embed_tokens.weight"][ids], P, S)
    assert (got - ref).abs().max().item() < 1e-4


def test_t2i_port_reproduces_the_reference_end_to_end(golden_dir):
    """oracle/t2i_port.py (the CPU arm of bench.py) against tests/golden/t2i_tiny.npz: the UNMODIFIED UniGen.t2i_generate
    (models/unigen.py:344-455) on the tiny random-init model, real transformers forward, noise captured by generator replay.
    Every step's sampled ids and mask and the returned ids must be equal."""
    from oracle import t2i_port
    g = np.load(os.path.join(golden_dir, "t2i_tiny.npz"))
    w = _tiny_torch_weights()
    T = int(g["timesteps"])
    ids = torch.from_numpy(g["ids"])
    for full_head in (True, False):
        final, recs = t2i_port.t2i_generate_port(
            w, TINY, ids, image_token_num_per_image=N_IMG, codebook_size=V, text_vocab_size=TEXT_VOCAB,
            mask_token_id=TEXT_VOCAB + V, pad_id=PAD, soi_id=SOI, timesteps=T, temperature=1.0,
            q_noise=list(g["q"]), u_noise=list(g["u"]), full_vocab_head=full_head)
        assert len(recs) == T
        cur = np.full(g["sampled_raw"][0].shape, TEXT_VOCAB + V, np.int64)
        for t in range(T):
            want = np.where(cur == TEXT_VOCAB + V, g["sampled_raw"][t], cur)
            assert np.array_equal(recs[t]["sampled"], want), f"step {t}"
            assert np.array_equal(recs[t]["mask"], g["masking"][t]), f"step {t}"
            cur = np.where(g["masking"][t], TEXT_VOCAB + V, want)
        assert np.array_equal(final, g["final_ids"])
