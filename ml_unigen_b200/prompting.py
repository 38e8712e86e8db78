"""Prompt assembly and omni-attention mask parameters for the T2I path.

Mirrors the two pieces of apple/ml-unigen training/prompting_utils.py that define the hot path's
inputs:

* `t2i_gen_prompt` (:113-153): `conv_start + text + conv_end` left-padded to a common length,
  then `<|soi|>`, the N image token ids, `<|eoi|>`.  Tokenisation itself is the caller's job
  (the Qwen2.5 tokenizer files are not part of either repo); this module works on token ids.
* `create_attention_mask_predict_next(..., rm_pad_in_image=True)` (:975-1020): the reference
  materialises a float32 [B,1,L,L] additive mask.  Its content is a closed form of two integers
  per sequence -- `pad_len` P_b (number of left pads) and `soi_idx` S_b (index of `<|soi|>`):

      row i <  P_b          : columns j <= i
      P_b <= i < S_b        : columns P_b <= j <= i
      i >= S_b              : columns j >= P_b

  (verified against the reference on recorded cases: tests/golden/omni_mask.npz).  The CUDA
  attention kernel consumes (P_b, S_b) directly; the dense mask is only built when a caller asks
  for it through the reference-compatible function below.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

# Qwen2.5 tokenizer ids used by the reference's eval drivers (evaluation/inference_geneval.py:49-52, 104)
QWEN_PAD_ID = 151643          # <|endoftext|> doubles as pad
QWEN_TEXT_VOCAB = 151674      # len(tokenizer) after the 9 added specials
QWEN_SOI_ID = 151665
QWEN_EOI_ID = 151666
MASK_FILL = float(torch.iinfo(torch.int64).min)   # -9.223372e18, the reference's fill value


def t2i_gen_prompt(text_ids: Sequence[Sequence[int]], image_ids: torch.Tensor, *, conv_start_ids: Sequence[int],
                   conv_end_ids: Sequence[int], pad_id: int, soi_id: int, eoi_id: int,
                   max_len: Optional[int] = None, max_seq_len: int = 1 << 30) -> Tuple[torch.Tensor, torch.Tensor]:
    """Token-id level restatement of prompting_utils.py:113-153.  Returns (input_ids [B,L], mask [B,L])."""
    N = image_ids.shape[-1]
    extra = len(conv_start_ids) + len(conv_end_ids) + 2 + N
    total = (max(len(t) for t in text_ids) if max_len is None else max_len) + extra
    total = min(total, max_seq_len)
    rows, masks = [], []
    dev = image_ids.device
    for b, t in enumerate(text_ids):
        body = list(conv_start_ids) + list(t) + list(conv_end_ids)
        if total >= len(body) + N + 2:
            n_pad = total - len(body) - N - 2
            m = [0] * n_pad + [1] * (len(body) + N + 2)
            body = [pad_id] * n_pad + body
        else:
            m = [1] * total
            body = body[: total - N - 2 - len(conv_end_ids)] + list(conv_end_ids)
        row = torch.cat([torch.tensor(body, dtype=torch.long, device=dev), torch.tensor([soi_id], device=dev),
                         image_ids[b].long(), torch.tensor([eoi_id], device=dev)])
        rows.append(row[None])
        masks.append(torch.tensor(m, device=dev)[None])
    return torch.cat(rows), torch.cat(masks)


def omni_mask_params(sequence: torch.Tensor, pad_id: int, soi_id: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(pad_len int32 [B], soi_idx int32 [B]) from token ids.  pad_len = index after the LAST pad token
    (prompting_utils.py:1004-1007 uses the last pad position), soi_idx = position of `<|soi|>`."""
    L = sequence.shape[1]
    pos = torch.arange(1, L + 1, device=sequence.device)
    pad_len = ((sequence == pad_id) * pos).amax(dim=1)
    is_soi = sequence == soi_id
    soi_idx = torch.where(is_soi.any(dim=1), is_soi.float().argmax(dim=1), torch.full_like(pad_len, L))
    return pad_len.to(torch.int32), soi_idx.to(torch.int32)


def omni_mask_params_from_dense(attention_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Recover (pad_len, soi_idx) from the reference's dense additive [B,1,L,L] mask (0 = attend).

    The last row (an image-block row) attends to every non-pad column, so its first allowed column is
    P_b; rows >= S_b are the only rows before L-1 that attend to column L-1."""
    if attention_mask.dim() != 4 or attention_mask.shape[1] != 1 or attention_mask.shape[2] != attention_mask.shape[3]:
        raise ValueError("expected the [B,1,L,L] additive mask of create_attention_mask_predict_next")
    allowed = attention_mask[:, 0] == 0
    pad_len = allowed[:, -1, :].float().argmax(dim=1)
    soi_idx = allowed[:, :, -1].float().argmax(dim=1)
    return pad_len.to(torch.int32), soi_idx.to(torch.int32)


def omni_allowed(pad_len: torch.Tensor, soi_idx: torch.Tensor, L: int) -> torch.Tensor:
    """bool [B,L,L] closed form (True = attend)."""
    dev = pad_len.device
    i = torch.arange(L, device=dev)[None, :, None]
    j = torch.arange(L, device=dev)[None, None, :]
    P = pad_len.long()[:, None, None]
    S = soi_idx.long()[:, None, None]
    return ((i < P) & (j <= i)) | ((i >= P) & (i < S) & (j >= P) & (j <= i)) | ((i >= S) & (j >= P))


def dense_mask_matches(attention_mask: torch.Tensor, pad_len: torch.Tensor, soi_idx: torch.Tensor) -> torch.Tensor:
    """0-d bool tensor: the dense mask is exactly the closed form of (pad_len, soi_idx)."""
    L = attention_mask.shape[-1]
    return ((attention_mask[:, 0] == 0) == omni_allowed(pad_len, soi_idx, L)).all()


def create_attention_mask_predict_next(sequence, pad_id=128256, soi_id=128257, eoi_id=128258, rm_pad_in_image=False,
                                       return_inverse_mask=True):
    """Reference-compatible signature (prompting_utils.py:975).  Only the configuration the T2I path uses
    (`rm_pad_in_image=True`) is implemented; the result equals the reference's tensor element for element."""
    if not rm_pad_in_image:
        raise NotImplementedError("the T2I decoding path always calls this with rm_pad_in_image=True")
    L = sequence.shape[1]
    pad_len, soi_idx = omni_mask_params(sequence, pad_id, soi_id)
    allowed = omni_allowed(pad_len, soi_idx, L)
    if not return_inverse_mask:
        return allowed[:, None]
    m = torch.zeros(allowed.shape, dtype=torch.float32, device=sequence.device)
    m.masked_fill_(~allowed, MASK_FILL)
    return m[:, None]


def synthetic_prompt_batch(B: int, N: int, *, text_len: int = 128, pads: Optional[List[int]] = None, seed: int = 0,
                           device="cpu", mask_token_id: int = QWEN_TEXT_VOCAB + 8192):
    """Synthetic T2I prompts of the reference's shape: L = text_len + 9 + 2 + N (conv_start 4 + conv_end 5
    tokens), per-sample left padding, random text ids, all image slots masked."""
    g = torch.Generator().manual_seed(seed)
    S = text_len + 9
    L = S + N + 2
    ids = torch.randint(0, QWEN_PAD_ID, (B, L), generator=g)
    if pads is None:
        pads = torch.randint(0, 100, (B,), generator=g).tolist()
    for b in range(B):
        ids[b, : pads[b]] = QWEN_PAD_ID
    ids[:, S] = QWEN_SOI_ID
    ids[:, S + 1: S + 1 + N] = mask_token_id
    ids[:, L - 1] = QWEN_EOI_ID
    return ids.to(device)
