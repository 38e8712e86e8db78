"""oracle/make_golden_backbone.py -- golden logits of the UNMODIFIED reference model (UniGen wrapping
transformers' Qwen2ForCausalLM, fp32, attn_implementation='sdpa') on a tiny random config.

The weights are generated from a numpy seed (portable) and rounded to bf16-representable values, so
tests can rebuild the very same model without shipping a checkpoint.  Recorded: the codebook logits
`logits[:, -(N+1):-1, text_vocab:-1]` (models/unigen.py:416) for an all-masked image block and for a
half-revealed one, computed by `UniGen.forward` with the dense mask of
`create_attention_mask_predict_next(rm_pad_in_image=True)`."""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.oracle import f32_to_bf16_bits, bf16_bits_to_f32  # noqa: E402

TINY = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
            num_key_value_heads=1, head_dim=128, rope_theta=1000000.0, rms_norm_eps=1e-6)
TEXT_VOCAB, V, N_IMG = 64, 4096, 16
PAD, SOI, EOI = 1, 2, 3


def tiny_weights(seed: int = 1234):
    """name -> float32 numpy (bf16-representable), transformers Qwen2Model parameter names."""
    rng = np.random.default_rng(seed)
    H, I, D = TINY["hidden_size"], TINY["intermediate_size"], TINY["head_dim"]
    Hq, Hkv = TINY["num_attention_heads"], TINY["num_key_value_heads"]
    vocab = TEXT_VOCAB + V + 1

    def rn(*shape, std=0.05):
        return bf16_bits_to_f32(f32_to_bf16_bits((std * rng.standard_normal(shape)).astype(np.float32)))

    w = {"embed_tokens.weight": rn(vocab, H), "norm.weight": 1.0 + rn(H, std=0.1)}
    for l in range(TINY["num_hidden_layers"]):
        p = f"layers.{l}."
        w[p + "input_layernorm.weight"] = 1.0 + rn(H, std=0.1)
        w[p + "post_attention_layernorm.weight"] = 1.0 + rn(H, std=0.1)
        w[p + "self_attn.q_proj.weight"] = rn(Hq * D, H)
        w[p + "self_attn.k_proj.weight"] = rn(Hkv * D, H)
        w[p + "self_attn.v_proj.weight"] = rn(Hkv * D, H)
        w[p + "self_attn.q_proj.bias"] = rn(Hq * D)
        w[p + "self_attn.k_proj.bias"] = rn(Hkv * D)
        w[p + "self_attn.v_proj.bias"] = rn(Hkv * D)
        w[p + "self_attn.o_proj.weight"] = rn(H, Hq * D)
        w[p + "mlp.gate_proj.weight"] = rn(I, H)
        w[p + "mlp.up_proj.weight"] = rn(I, H)
        w[p + "mlp.down_proj.weight"] = rn(H, I)
    for k in list(w):
        w[k] = bf16_bits_to_f32(f32_to_bf16_bits(w[k].astype(np.float32)))
    return w


def tiny_inputs(seed: int = 99):
    rng = np.random.default_rng(seed)
    S, N = 20, N_IMG
    L = S + N + 2
    B = 3
    vocab = TEXT_VOCAB + V + 1
    ids = rng.integers(4, TEXT_VOCAB, size=(B, L)).astype(np.int64)
    for b, p in enumerate([0, 7, 15]):
        ids[b, :p] = PAD
    ids[:, S] = SOI
    ids[:, S + 1:S + 1 + N] = vocab - 1
    ids[:, L - 1] = EOI
    ids_half = ids.copy()
    ids_half[:, S + 1:S + 1 + N:2] = TEXT_VOCAB + rng.integers(0, V, size=(B, N // 2))
    return ids, ids_half


def make_backbone_tiny():
    pu = ref_loader.load_prompting()
    vocab = TEXT_VOCAB + V + 1
    with tempfile.TemporaryDirectory() as d:
        ref_loader.qwen_config_dir(d, vocab_size=vocab, max_position_embeddings=512,
                                   **{k: v for k, v in TINY.items() if k != "head_dim"})
        torch.manual_seed(0)
        model = ref_loader.build_reference_unigen(d, vocab_size=vocab, codebook_size=V, num_vq_tokens=N_IMG)
    w = tiny_weights()
    sd = {("model." + k): torch.from_numpy(v) for k, v in w.items()}
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    missing, unexpected = model.llm.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in m or "lm_head" in m for m in missing), (missing, unexpected)
    out = {}
    for name, ids in zip(("masked", "half"), tiny_inputs()):
        t = torch.from_numpy(ids)
        mask = pu.create_attention_mask_predict_next(t, pad_id=PAD, soi_id=SOI, eoi_id=EOI, rm_pad_in_image=True)
        with torch.no_grad():
            logits = model(input_ids=t, input_embeddings=None, attention_mask=mask)          # UniGen.forward :240
        out[f"ids_{name}"] = ids
        out[f"logits_{name}"] = logits[:, -(N_IMG + 1):-1, TEXT_VOCAB:-1].numpy().astype(np.float32)
    path = os.path.join(ROOT, "tests", "golden", "backbone_tiny.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def make_t2i_tiny(T: int = 5, B: int = 2, seed: int = 17):
    """The UNMODIFIED `UniGen.t2i_generate` (models/unigen.py:344-455) end to end on the tiny random-init model: real Qwen2ForCausalLM
    forward (fp32, sdpa, dense mask), full-vocabulary head, slice, softmax, torch.multinomial, mask_by_random_topk, re-embedding, T
    steps.  Recorded: per-step sampled ids and masks, final ids, and the noise the reference's generator produced (generator replay,
    SURVEY.md section 0-7) -- what oracle/t2i_port.py must reproduce."""
    pu = ref_loader.load_prompting()
    ug = ref_loader.load_unigen()
    smp = sys.modules["models.sampling"]
    vocab = TEXT_VOCAB + V + 1
    with tempfile.TemporaryDirectory() as d:
        ref_loader.qwen_config_dir(d, vocab_size=vocab, max_position_embeddings=512,
                                   **{k: v for k, v in TINY.items() if k != "head_dim"})
        torch.manual_seed(0)
        model = ref_loader.build_reference_unigen(d, vocab_size=vocab, codebook_size=V, num_vq_tokens=N_IMG)
    sd = {("model." + k): torch.from_numpy(v) for k, v in tiny_weights().items()}
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    model.llm.load_state_dict(sd, strict=False)
    ids = torch.from_numpy(tiny_inputs()[0][:B])
    mask = pu.create_attention_mask_predict_next(ids, pad_id=PAD, soi_id=SOI, eoi_id=EOI, rm_pad_in_image=True)
    gen = torch.Generator().manual_seed(seed)
    replay = torch.Generator().manual_seed(seed)
    rec = dict(sampled_raw=[], masking=[], q=[], u=[])
    orig_multinomial, orig_topk = torch.multinomial, ug.mask_by_random_topk

    def multinomial_rec(p, n, generator=None):
        out = orig_multinomial(p, n, generator=generator)
        rec["sampled_raw"].append(out[:, 0].view(B, N_IMG).numpy().copy())
        rec["q"].append(torch.empty(B * N_IMG, V).exponential_(1, generator=replay).view(B, N_IMG, V).numpy().copy())
        return out

    def topk_rec(mask_len, probs, temperature=1.0, generator=None):
        out = orig_topk(mask_len, probs, temperature, generator=generator)
        rec["masking"].append(out.numpy().copy())
        rec["u"].append(torch.zeros(B, N_IMG).uniform_(0, 1, generator=replay).numpy().copy())
        return out

    torch.multinomial = multinomial_rec
    ug.mask_by_random_topk = topk_rec
    try:
        with torch.no_grad():
            final = model.t2i_generate(input_ids=ids, attention_mask=mask, temperature=1.0, timesteps=T, guidance_scale=0,
                                       noise_schedule=smp.get_mask_chedule("cosine"), generator=gen,
                                       image_token_num_per_image=N_IMG, text_vocab_size=TEXT_VOCAB)
    finally:
        torch.multinomial, ug.mask_by_random_topk = orig_multinomial, orig_topk
    assert torch.equal(gen.get_state(), replay.get_state()), "generator replay diverged"
    path = os.path.join(ROOT, "tests", "golden", "t2i_tiny.npz")
    np.savez_compressed(path, ids=ids.numpy(), final_ids=final.numpy(), timesteps=np.int64(T),
                        **{k: np.stack(v) for k, v in rec.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    make_backbone_tiny()
    make_t2i_tiny()
