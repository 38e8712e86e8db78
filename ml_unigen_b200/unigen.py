"""`UniGenB200.t2i_generate` -- drop-in for `UniGen.t2i_generate` (apple/ml-unigen
models/unigen.py:344-455): same keyword arguments, same return value (int64 [B,N] codebook ids of the
last step), different machinery:

  reference (per step)                                   this module (per step)
  ----------------------------------------------------   ---------------------------------------------
  cat(prefix, image, suffix) embeddings                   image rows re-embedded in place (ug_embed_rows)
  Qwen2Model with dense fp32 [B,1,L,L] mask, SDPA         QwenBackboneB200: tcgen05 omni attention on (P_b,S_b)
  lm_head over L x 159867, then slice [N x 8192]          codebook-only head GEMM -> dense bf16 [B,N,V]
  CFG mix, softmax, multinomial (host sync), gather,      ONE launch of ug_decode_step_bf16 (CFG, softmax,
  where, log, uniform_, sort, gather, <, where (~12       exponential race, confidence, k-th smallest,
  launches, 8 passes over a 537 MB fp32 tensor)           re-masking; logits read once)
  H2D copies of mask_len / tensor([1]) each step          no host<->device traffic inside the loop

The MaskGIT schedule arithmetic (ratio, `noise_schedule(torch.tensor(ratio))`, floor(N*mask_ratio),
the compounding temperature of unigen.py:444) is evaluated on the host exactly as the reference does.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Optional

import torch

from . import ops, prompting
from .backbone import QwenBackboneB200, QwenConfig
from .sampling import cosine_schedule


class _StepGraph:
    """The per-step backbone pass of the prefix-reuse path (embed the N+2 image-block rows, 28 layers, final norm, codebook
    head) as ONE CUDA graph over static buffers.  Every launch inside it is a plain asynchronous kernel launch, so the
    142 launches of a MaskGIT step replay without CPU work and with the launch-to-launch gaps of a graph."""

    def __init__(self, bb: QwenBackboneB200, n_seq: int, L: int, N: int, s0: int, V: int, text_vocab: int):
        dev = bb.device
        self.kv_cache = bb.new_kv_cache(n_seq, L)
        self.tok = torch.zeros(n_seq, L - s0, dtype=torch.int64, device=dev)
        self.P = torch.zeros(n_seq, dtype=torch.int32, device=dev)
        self.S = torch.zeros(n_seq, dtype=torch.int32, device=dev)
        self.args = (bb, N, s0, V, text_vocab)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.logits: Optional[torch.Tensor] = None
        self.warm = False
        self.launches = 0              # own launches inside one replay (counted during the eager warm-up pass)

    def _body(self):
        bb, N, s0, V, text_vocab = self.args
        st = bb.embed_tokens(self.tok)
        hidden = bb.forward_suffix(st, self.P, self.S, s0, self.kv_cache, rows=slice(1, 1 + N))
        return bb.codebook_logits(hidden, text_vocab, V)

    def run(self) -> torch.Tensor:
        bb = self.args[0]
        if self.graph is not None:
            self.graph.replay()
            bb.launch_count += self.launches
            return self.logits
        if not self.warm:                      # first pass eagerly: function attributes, lazily created tables and workspaces
            self.warm = True
            n0 = bb.launch_count
            out = self._body()
            self.launches = bb.launch_count - n0
            return out
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.logits = self._body()
        self.graph = g
        bb.launch_count -= self.launches       # the capture pass enqueued nothing
        g.replay()
        bb.launch_count += self.launches
        return self.logits


class UniGenB200:
    def __init__(self, backbone: QwenBackboneB200, *, codebook_size: int = 8192, mask_token_id: Optional[int] = None,
                 pad_id: int = prompting.QWEN_PAD_ID, soi_id: int = prompting.QWEN_SOI_ID):
        self.backbone = backbone
        self.codebook_size = codebook_size
        # models/unigen.py:92 -- mask_token_id = vocab_size - 1 when no separate image head exists
        self.mask_token_id = backbone.cfg.vocab_size - 1 if mask_token_id is None else mask_token_id
        self.pad_id, self.soi_id = pad_id, soi_id
        self.device = backbone.device
        self._ws = {}
        self._step_graphs = OrderedDict()     # LRU of captured per-step graphs, one per shape (each owns a K/V cache)
        self.max_step_graphs = 3
        self.use_cuda_graph = True     # replay the per-step backbone pass as a CUDA graph (off while bench hooks record events)
        self.last_sel_prob = None
        self.decode_events = None      # bench hook: (start, end) CUDA events of every decode-step launch

    # reference-style constructors -------------------------------------------------------------
    @classmethod
    def random_init(cls, device, cfg: Optional[QwenConfig] = None, seed: int = 0, **kw) -> "UniGenB200":
        return cls(QwenBackboneB200.random_init(cfg or QwenConfig(), device, seed=seed), **kw)

    @classmethod
    def from_reference(cls, ref_unigen, device, **kw) -> "UniGenB200":
        """From an instance of the reference's `UniGen` (its `llm` is a transformers Qwen2ForCausalLM).
        The reference computes logits with `self.llm.lm_head` (models/unigen.py:287): an untied head is loaded as such.
        `gen_proj_depth > 0` checkpoints (separate `img_head` / `get_gen_embed`, models/unigen.py:372, 401-404) are not
        part of any shipped config and are refused rather than silently mis-decoded."""
        cfg = ref_unigen.config
        depth = cfg.get("gen_proj_depth", 0) if hasattr(cfg, "get") else getattr(cfg, "gen_proj_depth", 0)
        if depth and depth > 0:
            raise NotImplementedError("gen_proj_depth > 0 (separate img_head / gen_embed) is outside the shipped T2I path")
        llm = ref_unigen.llm
        head_w = None
        lm_head = getattr(llm, "lm_head", None)
        if lm_head is not None and lm_head.weight.data_ptr() != llm.model.embed_tokens.weight.data_ptr() \
                and not torch.equal(lm_head.weight, llm.model.embed_tokens.weight):
            head_w = lm_head.weight
        bb = QwenBackboneB200.from_qwen2_model(llm.model, device, lm_head_weight=head_w)
        mask_id = cfg.get("mask_token_id") if hasattr(cfg, "get") else cfg.mask_token_id
        return cls(bb, codebook_size=int(getattr(ref_unigen, "img_output_size", 8192)), mask_token_id=int(mask_id), **kw)

    # ---------------------------------------------------------------------------------------------
    def _workspace(self, B, N, V):
        key = (B, N, V)
        if key not in self._ws:
            self._ws[key] = ops.DecodeWorkspace(B, N, V, self.device)
        return self._ws[key]

    def _mask_params(self, input_ids, attention_mask, validate: bool):
        if attention_mask is not None and attention_mask.dim() == 4:
            P, S = prompting.omni_mask_params_from_dense(attention_mask)
            if validate and not bool(prompting.dense_mask_matches(attention_mask, P, S)):
                raise ValueError("attention_mask is not an omni mask of create_attention_mask_predict_next("
                                 "rm_pad_in_image=True); other masks are outside this path")
            return P.contiguous(), S.contiguous()
        if attention_mask is not None:
            raise NotImplementedError("2-D (use_causal_mask) masks are not part of the T2I decoding path")
        return prompting.omni_mask_params(input_ids, self.pad_id, self.soi_id)

    @torch.no_grad()
    def t2i_generate(
        self,
        input_ids: Optional[torch.LongTensor] = None,
        uncond_input_ids: Optional[torch.LongTensor] = None,
        input_embeddings: Optional[torch.Tensor] = None,
        uncond_input_embeddings: Optional[torch.Tensor] = None,
        attention_mask: Optional[torch.Tensor] = None,
        temperature: float = 1.0,
        timesteps: int = 18,
        guidance_scale: int = 0,
        noise_schedule: Callable = cosine_schedule,
        generator: Optional[torch.Generator] = None,
        image_token_num_per_image: int = 256,
        text_vocab_size: int = 151936,
        **kwargs,
    ):
        """See models/unigen.py:344-359 for the arguments.  Extra keyword arguments understood here:
        `batch_offset` (global index of image 0: sharding-invariant noise), `validate_mask` (default True),
        `philox_seed` / `philox_offset` (override the generator-derived stream)."""
        if input_ids is None:
            raise ValueError("input_ids is required (the image slots and their mask ids are read from it, models/unigen.py:367)")
        if not input_ids.is_cuda:
            input_ids = input_ids.to(self.device, non_blocking=True)
        N, V = image_token_num_per_image, self.codebook_size
        bb, mask_id = self.backbone, self.mask_token_id
        B, L = input_ids.shape
        if text_vocab_size + V > bb.cfg.vocab_size:
            raise ValueError(f"text_vocab_size {text_vocab_size} + codebook {V} exceeds the vocabulary ({bb.cfg.vocab_size})")
        use_cfg = guidance_scale > 1                                    # models/unigen.py:382 (sic: > 1, not > 0)
        ids_img = input_ids[:, -(N + 1):-1].clone().contiguous()        # unigen.py:367
        tokens = input_ids.clone()
        if use_cfg:
            if uncond_input_ids is None and uncond_input_embeddings is None:
                raise ValueError("guidance_scale > 1 needs uncond_input_ids (or uncond_input_embeddings)")
            un = input_ids.clone() if uncond_input_ids is None else uncond_input_ids.to(self.device).clone()
            un[:, -1] = input_ids[:, -1]                                # the suffix embedding is the cond one, doubled (unigen.py:395)
            tokens = torch.cat([tokens, un], 0)
        n_seq = tokens.shape[0]
        if attention_mask is not None and attention_mask.shape[0] != n_seq:
            raise ValueError("attention_mask batch must match cond (+ uncond) batch")
        P, S = self._mask_params(tokens, attention_mask, kwargs.get("validate_mask", True))

        # precomputed embeddings (models/unigen.py:370-371, 389-393): they replace embed_tokens(input_ids) for the prefix rows
        # (incl. <soi>), the suffix row and -- in step 0 only -- the image rows; later steps re-embed the image ids.
        emb_fix = None
        if input_embeddings is not None or uncond_input_embeddings is not None:
            H = bb.cfg.hidden_size
            emb_fix = bb.embed.float()[tokens.clamp(0, bb.cfg.vocab_size - 1)]          # [n_seq, L, H] fp32, setup only
            if input_embeddings is not None:
                if tuple(input_embeddings.shape) != (B, L, H):
                    raise ValueError("input_embeddings must be [B, L, hidden]")
                emb_fix[:B] = input_embeddings.to(self.device, torch.float32)
            if use_cfg:
                if uncond_input_embeddings is not None:
                    emb_fix[B:, : L - N - 1] = uncond_input_embeddings.to(self.device, torch.float32)[:, : L - N - 1]
                emb_fix[B:, L - N - 1:] = emb_fix[:B, L - N - 1:]      # image rows and suffix: the cond ones, repeated

        gen = generator
        if "philox_seed" in kwargs:
            seed, base_offset = int(kwargs["philox_seed"]), int(kwargs.get("philox_offset", 0))
        else:
            if gen is None:
                gen = torch.cuda.default_generators[self.device.index or 0]
            if gen.device.type == "cuda":
                seed, base_offset = gen.initial_seed(), gen.get_offset()
                gen.set_offset(base_offset + 4 * ((timesteps + 3) // 4))   # consume a distinct slice of the stream
            else:
                # a CPU generator has no Philox offset to slice: draw the key from it, which also advances it, so repeated
                # calls with the same generator give different images -- as they do in the reference
                seed, base_offset = int(torch.randint(0, 2 ** 62, (1,), generator=gen).item()), 0
        batch_offset = int(kwargs.get("batch_offset", 0))
        trace = kwargs.get("trace")

        ws = self._workspace(B, N, V)
        sampled = torch.empty(B, N, dtype=torch.int64, device=self.device)
        masking = torch.empty(B, N, dtype=torch.uint8, device=self.device)
        sel_prob = torch.empty(B, N, dtype=torch.float32, device=self.device)
        img_rows = slice(L - N - 1, L - 1)
        temp = float(temperature)
        # one host sync per generation: token-id range (ug_embed_rows clamps for memory safety only) and the soi index
        chk = torch.stack([S.min(), S.max(), tokens.min().to(torch.int32), (tokens.max() - (bb.cfg.vocab_size - 1)).clamp(min=0).to(torch.int32),
                           ((ids_img != mask_id) & ((ids_img < 0) | (ids_img >= V))).any().to(torch.int32)]).tolist()
        smin, smax = int(chk[0]), int(chk[1])
        if chk[2] < 0 or chk[3] > 0:
            raise ValueError(f"input_ids outside [0, {bb.cfg.vocab_size}): wrong tokenizer / vocab_size?")
        if chk[4]:
            raise ValueError(f"known image tokens must be codebook ids in [0, {V}) (or mask_token_id {mask_id})")
        # Text-prefix K/V reuse: when every sequence has its <|soi|> at the same index s0 (always the case for
        # t2i_gen_prompt batches, which are left-padded to one length), rows < s0 are computed once (step 0) and
        # only the N+2 rows of the image block go through the backbone in the later steps.
        s0 = None
        if kwargs.get("prefix_reuse", True) and timesteps > 1 and smin == smax and smin == L - N - 2:
            s0 = smin
        sg = None
        if s0 is not None and self.use_cuda_graph and kwargs.get("cuda_graph", True) and bb.attn_events is None \
                and bb.mlp_events is None and bb.plain_events is None and emb_fix is None and timesteps > 3:
            key = (n_seq, L, N, s0, V, text_vocab_size)
            sg = self._step_graphs.get(key)
            if sg is None:
                while len(self._step_graphs) >= self.max_step_graphs:   # least recently used shape goes (it owns a K/V cache)
                    self._step_graphs.popitem(last=False)
                sg = self._step_graphs[key] = _StepGraph(bb, n_seq, L, N, s0, V, text_vocab_size)
            else:
                self._step_graphs.move_to_end(key)
            sg.P.copy_(P)
            sg.S.copy_(S)
        kv_cache = sg.kv_cache if sg is not None else (bb.new_kv_cache(n_seq, L) if s0 is not None else None)

        def residual(tok, row0, step):
            """The residual stream of rows [row0, L) for this step, with the caller's embeddings where they apply."""
            st = bb.embed_tokens(tok)
            if emb_fix is None:
                return st
            x = st.x.view(n_seq, L - row0, -1)
            if step == 0:
                x.copy_(emb_fix[:, row0:])
            else:
                x[:, : L - N - 1 - row0] = emb_fix[:, row0: L - N - 1]
                x[:, -1] = emb_fix[:, -1]
            return bb.residual_from_embeddings(x)

        for step in range(timesteps):
            # tokens of the image block: mask id stays, codebook id c becomes c + text_vocab (unigen.py:451)
            img_tok = torch.where(ids_img == mask_id, ids_img, ids_img + text_vocab_size)
            tokens[:B, img_rows] = img_tok
            if use_cfg:
                tokens[B:, img_rows] = img_tok
            if s0 is None or step == 0:
                hidden = bb.forward(residual(tokens, 0, step), P, S, rows=img_rows, kv_cache=kv_cache)
                logits = bb.codebook_logits(hidden, text_vocab_size, V)  # bf16 [B(2B), N, V]
            elif sg is not None:
                sg.tok.copy_(tokens[:, s0:])
                logits = sg.run()
            else:
                hidden = bb.forward_suffix(residual(tokens[:, s0:].contiguous(), s0, step), P, S, s0, kv_cache, rows=slice(1, 1 + N))
                logits = bb.codebook_logits(hidden, text_vocab_size, V)
            ratio = 1.0 * (step + 1) / timesteps                        # unigen.py:425
            mask_ratio = noise_schedule(torch.tensor(ratio))             # unigen.py:426 (0-d CPU tensor)
            cap = int((N * mask_ratio).floor().item())                   # unigen.py:436 (host scalar, no device sync)
            temp = temp * (1.0 - ratio)                                  # unigen.py:444
            ev = self.decode_events
            if ev is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            logits_seen = logits.clone() if trace is not None else None     # test hook: the logits before the CFG mix
            if use_cfg:
                # guidance in its own in-place pass (bf16 op by op, models/unigen.py:410); the mixed logits are bf16 again,
                # so the certified fast kernel streams them (passing uncond_logits would select the exact kernel)
                ops.cfg_mix_(logits[:B], logits[B:], float(guidance_scale))
                bb.launch_count += 1
            ops.decode_step(logits[:B], ids_img, mask_token_id=mask_id, mask_len_cap=cap, temperature=temp,
                            workspace=ws, sampled_out=sampled, mask_out=masking, sel_prob_out=sel_prob,
                            seed=seed, offset=base_offset + step, batch_offset=batch_offset)
            bb.launch_count += ops.DECODE_STEP_LAUNCHES
            if ev is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                ev.append((e0, e1))
            if trace is not None:                                        # test hook: what each step saw and produced
                trace.append(dict(logits=logits_seen, cap=cap, temperature=temp, offset=base_offset + step,
                                  ids_after=ids_img.clone(), sampled=sampled.clone(), masking=masking.clone()))
        self.last_sel_prob = sel_prob
        return sampled

    # callers in the reference use model(...) only through t2i_generate on this path
    generate_images = t2i_generate
