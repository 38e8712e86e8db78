"""oracle/make_golden.py -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py

decode_*.npz : `UniGen.t2i_generate` (models/unigen.py:344-455) is executed as is on a duck-typed
    `self` whose forward returns prepared logits, so every line of the MaskGIT loop
    (softmax, multinomial, gather, mask_len, mask_by_random_topk, re-masking) is the reference's
    own.  The random stream is captured by generator replay: per step the
    reference draws `exponential_` over [B*N, V] inside multinomial and then `uniform_` over
    [B, N] inside gumbel_noise; a clone of the generator reproduces both tensors, which become
    the injected noise of the oracle / CUDA kernel.  Intermediate results are recorded by
    wrapping (not editing) torch.multinomial and mask_by_random_topk.
omni_mask.npz : `create_attention_mask_predict_next(..., rm_pad_in_image=True)`
    (training/prompting_utils.py:975-1020) on random left-padded sequences.
schedule.npz  : cosine_schedule / mask_len / compounded temperature for (N, T) pairs.
backbone_tiny.npz : the reference UniGen (transformers Qwen2ForCausalLM, fp32, sdpa) on a tiny
    random-init config: logits of the image block and final ids of a full t2i_generate call.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.oracle import f32_to_bf16_bits, bf16_bits_to_f32  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class _FakeLLMModel:
    def __init__(self, hidden=4):
        self.hidden = hidden

    def embed_tokens(self, ids):
        return torch.zeros(*ids.shape, self.hidden)


class _FakeLLM:
    def __init__(self):
        self.model = _FakeLLMModel()


class _FakeUniGen:
    """Stands in for `self` in UniGen.t2i_generate: same attributes the loop touches."""

    def __init__(self, mask_token_id, step_logits_full):
        self.config = ref_loader._Config(mask_token_id=mask_token_id)
        self.llm = _FakeLLM()
        self._logits = step_logits_full
        self._step = 0

    def __call__(self, input_ids=None, input_embeddings=None, attention_mask=None):
        out = self._logits[self._step]
        self._step += 1
        return out


def run_reference_decode(logits_f32, *, text_vocab, temperature, seed, schedule="cosine"):
    """logits_f32: [T,B,N,V] float32 holding bf16-representable values.  Returns recorded dict."""
    ug = ref_loader.load_unigen()
    smp = sys.modules["models.sampling"]
    T, B, N, V = logits_f32.shape
    vocab = text_vocab + V + 1
    mask_id = vocab - 1
    L = 3 + N + 1
    full = []
    for t in range(T):
        f = torch.zeros(B, L, vocab)
        f[:, -(N + 1):-1, text_vocab:-1] = torch.from_numpy(logits_f32[t])
        full.append(f)
    fake = _FakeUniGen(mask_id, full)
    input_ids = torch.zeros(B, L, dtype=torch.long)
    input_ids[:, -(N + 1):-1] = mask_id

    gen = torch.Generator().manual_seed(seed)
    replay = torch.Generator().manual_seed(seed)
    rec = dict(sampled_raw=[], mask_len=[], sel_prob=[], temperature=[], masking=[], q=[], u=[])

    orig_multinomial = torch.multinomial
    orig_topk = ug.mask_by_random_topk

    def multinomial_rec(p, n, generator=None):
        out = orig_multinomial(p, n, generator=generator)
        rec["sampled_raw"].append(out[:, 0].view(B, N).numpy().copy())
        rec["q"].append(torch.empty(B * N, V).exponential_(1, generator=replay).view(B, N, V).numpy().copy())
        return out

    def topk_rec(mask_len, probs, temperature=1.0, generator=None):
        out = orig_topk(mask_len, probs, temperature, generator=generator)
        rec["mask_len"].append(mask_len.long().view(B).numpy().copy())
        rec["sel_prob"].append(probs.numpy().copy())
        rec["temperature"].append(float(temperature))
        rec["masking"].append(out.numpy().copy())
        rec["u"].append(torch.zeros(B, N).uniform_(0, 1, generator=replay).numpy().copy())
        return out

    torch.multinomial = multinomial_rec
    ug.mask_by_random_topk = topk_rec
    try:
        final = ug.UniGen.t2i_generate(
            fake, input_ids=input_ids, attention_mask=None, temperature=temperature, timesteps=T,
            guidance_scale=0, noise_schedule=smp.get_mask_chedule(schedule), generator=gen,
            image_token_num_per_image=N, text_vocab_size=text_vocab)
    finally:
        torch.multinomial = orig_multinomial
        ug.mask_by_random_topk = orig_topk
    # the replayed generator must have consumed exactly what the reference consumed
    assert torch.equal(gen.get_state(), replay.get_state()), "generator replay diverged"
    out = {k: np.stack(v) for k, v in rec.items()}
    out["final_ids"] = final.numpy().copy()
    out["mask_token_id"] = np.int64(mask_id)
    out["text_vocab"] = np.int64(text_vocab)
    return out


def make_decode_case(name, *, T, B, N, V, scale, seed, temperature=1.0):
    rng = np.random.default_rng(seed)
    x = (scale * rng.standard_normal((T, B, N, V))).astype(np.float32)
    bits = f32_to_bf16_bits(x)
    x = bf16_bits_to_f32(bits)
    rec = run_reference_decode(x, text_vocab=8, temperature=temperature, seed=seed + 100)
    path = os.path.join(GOLD, f"decode_{name}.npz")
    np.savez_compressed(path, logits_bits=bits, temperature0=np.float64(temperature), **rec)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def make_decode_fullsize(name, *, scale, seed, noise, B=2, N=1024, V=8192, T=18, temperature=1.0):
    """BASELINE.json's decode shape (N=1024, V=8192, T=18) through the reference's own loop (models/unigen.py:398-455).  Only
    seeds, sampled ids and masks are stored (75 KB): logits come from `oracle.synthetic_logits_bits` (integer RNG, exact ops),
    the noise either
      noise="replay" : from the reference's own generator -- torch.multinomial and gumbel_noise run UNMODIFIED and the test
                       re-draws exponential_/uniform_ from a clone of `torch.Generator().manual_seed(seed)`; a checksum of the
                       stream is stored because torch's CPU exponential_ may differ between host CPUs (MKL / libm dispatch), or
      noise="integer": injected, platform-independent (`oracle.integer_noise`): torch.multinomial(p, 1) is replaced by its
                       definition `(p / q).argmax(-1)` (probed bit-identical) and gumbel_noise's
                       uniform_ draw by the injected u; every other line is the reference's."""
    from oracle import oracle as O
    ug = ref_loader.load_unigen()
    smp = sys.modules["models.sampling"]
    text_vocab = 8
    vocab = text_vocab + V + 1
    mask_id = vocab - 1
    L = 3 + N + 1

    class _Lazy:                                    # per-step full-vocabulary logits, built when the loop asks for them
        def __init__(self):
            self.step = 0

        def __call__(self, input_ids=None, input_embeddings=None, attention_mask=None):
            bits = O.synthetic_logits_bits(seed, self.step, B, N, V, scale)
            f = torch.zeros(B, L, vocab)
            f[:, -(N + 1):-1, text_vocab:-1] = torch.from_numpy(bf16_bits_to_f32(bits))
            self.step += 1
            return f

    fake = _FakeUniGen(mask_id, None)
    lazy = _Lazy()
    fake.__class__ = type("_FakeLazy", (_FakeUniGen,), {"__call__": lambda self, **kw: lazy(**kw)})
    input_ids = torch.zeros(B, L, dtype=torch.long)
    input_ids[:, -(N + 1):-1] = mask_id
    gen = torch.Generator().manual_seed(seed + 100)
    replay = torch.Generator().manual_seed(seed + 100)
    rec = dict(sampled_raw=[], masking=[], mask_len=[], q_sum=[], u_sum=[])
    orig_multinomial, orig_topk, orig_gumbel = torch.multinomial, ug.mask_by_random_topk, smp.gumbel_noise
    state = {"u": None}

    def multinomial_rec(p, n, generator=None):
        if noise == "replay":
            out = orig_multinomial(p, n, generator=generator)
            q = torch.empty(B * N, V).exponential_(1, generator=replay)
            assert torch.equal((p / q).argmax(-1), out[:, 0])                 # SURVEY 0-1 on this very call
            rec["q_sum"].append(O.noise_checksum(q.numpy()))
        else:
            q, state["u"] = O.integer_noise(seed, len(rec["sampled_raw"]), B, N, V)
            out = (p / torch.from_numpy(q).view(B * N, V)).argmax(-1, keepdim=True)
        rec["sampled_raw"].append(out[:, 0].view(B, N).numpy().astype(np.int16))
        return out

    def gumbel_injected(t, generator=None):          # sampling.py:24-26 with the draw replaced by the injected tensor
        return -smp.log(-smp.log(torch.from_numpy(state["u"])))

    def topk_rec(mask_len, probs, temperature=1.0, generator=None):
        out = orig_topk(mask_len, probs, temperature, generator=generator)
        rec["mask_len"].append(mask_len.long().view(B).numpy().copy())
        rec["masking"].append(np.packbits(out.numpy(), axis=-1))
        if noise == "replay":
            rec["u_sum"].append(O.noise_checksum(torch.zeros(B, N).uniform_(0, 1, generator=replay).numpy()))
        return out

    torch.multinomial = multinomial_rec
    ug.mask_by_random_topk = topk_rec
    if noise == "integer":
        smp.gumbel_noise = gumbel_injected
    try:
        final = ug.UniGen.t2i_generate(fake, input_ids=input_ids, attention_mask=None, temperature=temperature, timesteps=T,
                                       guidance_scale=0, noise_schedule=smp.get_mask_chedule("cosine"), generator=gen,
                                       image_token_num_per_image=N, text_vocab_size=text_vocab)
    finally:
        torch.multinomial, ug.mask_by_random_topk, smp.gumbel_noise = orig_multinomial, orig_topk, orig_gumbel
    if noise == "replay":
        assert torch.equal(gen.get_state(), replay.get_state()), "generator replay diverged"
    out = {k: np.stack(v) for k, v in rec.items() if v}
    path = os.path.join(GOLD, f"decode_full_{name}.npz")
    np.savez_compressed(path, final_ids=final.numpy().astype(np.int16), mask_token_id=np.int64(mask_id), seed=np.int64(seed),
                        scale=np.float64(scale), temperature0=np.float64(temperature), shape=np.array([T, B, N, V]),
                        noise=np.array(noise), **{k: (np.array(v, dtype=np.uint64) if k.endswith("_sum") else v) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def make_cfg_case():
    """unigen.py:410 executed by torch on bf16 tensors: guidance_scale * (cond - uncond) + uncond."""
    rng = np.random.default_rng(5)
    B, N, V = 1, 4, 4096
    c = torch.from_numpy((3 * rng.standard_normal((B, N, V))).astype(np.float32)).bfloat16()
    u = torch.from_numpy((3 * rng.standard_normal((B, N, V))).astype(np.float32)).bfloat16()
    out = {}
    for s in (2.0, 6, 1.75):
        mixed = s * (c - u) + u
        assert mixed.dtype == torch.bfloat16
        out[f"mixed_{float(s)}"] = mixed.view(torch.int16).numpy().view(np.uint16)
    path = os.path.join(GOLD, "cfg_mix.npz")
    np.savez_compressed(path, cond=c.view(torch.int16).numpy().view(np.uint16),
                        uncond=u.view(torch.int16).numpy().view(np.uint16), **out)
    print("wrote", path)


def make_omni_mask():
    pu = ref_loader.load_prompting()
    rng = np.random.default_rng(11)
    PAD, SOI, EOI = 151643, 151665, 151666
    seqs, masks = [], []
    cases = [(12, 8), (20, 16), (9, 4)]
    out = {}
    for ci, (S, N) in enumerate(cases):
        L = S + N + 2
        B = 6
        ids = rng.integers(0, 1000, size=(B, L)).astype(np.int64)
        pads = [0, 1, S // 2, S - 1, 3, 0]
        for b, p in enumerate(pads):
            ids[b, :p] = PAD
            ids[b, S] = SOI
            ids[b, S + 1:S + 1 + N] = 159866
            ids[b, L - 1] = EOI
        m = pu.create_attention_mask_predict_next(torch.from_numpy(ids), pad_id=PAD, soi_id=SOI, eoi_id=EOI,
                                                  rm_pad_in_image=True)
        assert m.dtype == torch.float32 and m.shape == (B, 1, L, L)
        out[f"ids_{ci}"] = ids
        out[f"allowed_{ci}"] = (m[:, 0] == 0).numpy()
        out[f"fill_{ci}"] = np.float32(m.min().item())
    path = os.path.join(GOLD, "omni_mask.npz")
    np.savez_compressed(path, pad_id=PAD, soi_id=SOI, eoi_id=EOI, **out)
    print("wrote", path)


def make_schedule():
    smp = ref_loader.load_sampling()
    out = {}
    for (N, T) in [(1024, 18), (256, 18), (256, 16), (1024, 50), (16, 4)]:
        caps, temps, ratios = [], [], []
        temperature = 1.0
        for step in range(T):
            ratio = 1.0 * (step + 1) / T
            mr = smp.cosine_schedule(torch.tensor(ratio))                    # unigen.py:426
            caps.append(int((N * mr).floor().item()))                       # unigen.py:436
            temperature = temperature * (1.0 - ratio)                       # unigen.py:444
            temps.append(temperature)
            ratios.append(float(mr))
        out[f"cap_{N}_{T}"] = np.array(caps, np.int64)
        out[f"temp_{N}_{T}"] = np.array(temps, np.float64)
        out[f"ratio_{N}_{T}"] = np.array(ratios, np.float32)
    for name in ("linear", "pow2", "pow0.5", "sigmoid"):
        fn = smp.get_mask_chedule(name)
        out[f"sched_{name}"] = np.array([float(fn(torch.tensor(r / 10))) for r in range(1, 11)], np.float32)
    path = os.path.join(GOLD, "schedule.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    assert ref_loader.available(), "reference tree not mounted"
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["decode", "cfg", "mask", "schedule", "backbone", "full"]
    if "decode" in which:
        make_decode_case("peaked_v4096", T=4, B=2, N=8, V=4096, scale=4.0, seed=1)
        make_decode_case("flat_v4096", T=4, B=2, N=8, V=4096, scale=0.02, seed=2)
        make_decode_case("peaked_v8192", T=3, B=1, N=16, V=8192, scale=4.0, seed=3)
    if "full" in which:
        make_decode_fullsize("peaked_integer", scale=4.0, seed=21, noise="integer")
        make_decode_fullsize("flat_integer", scale=0.02, seed=22, noise="integer")
        make_decode_fullsize("unit_replay", scale=1.0, seed=23, noise="replay")
    if "cfg" in which:
        make_cfg_case()
    if "mask" in which:
        make_omni_mask()
    if "schedule" in which:
        make_schedule()
    if "backbone" in which:
        from oracle.make_golden_backbone import make_backbone_tiny
        make_backbone_tiny()
