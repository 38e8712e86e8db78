"""oracle/make_golden_prompting.py -- records tests/golden/prompting.json by running the UNMODIFIED reference class
`UniversalPromptingQwen2` (training/prompting_utils.py:13-56, `t2i_gen_prompt` :113-153, dispatch :398-402) on a stub tokenizer
(the Qwen2.5 tokenizer files are in neither repo): whitespace words -> stable ids, the special tokens -> the ids the reference's
eval drivers see (evaluation/inference_geneval.py:49-52).  Run in the build container:  python oracle/make_golden_prompting.py"""
from __future__ import annotations

import json
import os
import re
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


class StubTokenizer:
    """The slice of a HF tokenizer the reference's prompting class touches."""
    pad_token_id = 151643
    eos_token_id = 151645

    def __init__(self):
        self.special = {"<|im_start|>": 151644, "<|im_end|>": 151645, "<|vision_start|>": 151652, "<|vision_end|>": 151653}
        self.next_id = 151665

    def add_tokens(self, toks):
        for t in toks:
            if t not in self.special:
                self.special[t] = self.next_id
                self.next_id += 1

    def convert_tokens_to_ids(self, toks):
        return [self.special[t] for t in toks]

    def _encode(self, text):
        ids = []
        for piece in re.findall(r"<\|[a-z0-9_]+\|>|\n|[^\s<]+", text):
            if piece in self.special:
                ids.append(self.special[piece])
            elif piece == "\n":
                ids.append(198)
            else:
                ids.append(1000 + sum(ord(c) * (i + 7) for i, c in enumerate(piece)) % 90000)
        return ids

    def __call__(self, text, **kw):
        if isinstance(text, str):
            return type("Enc", (), {"input_ids": self._encode(text)})()
        return {"input_ids": [self._encode(t) for t in text]}


def main():
    pu = ref_loader.load_prompting()
    cases = []
    prompts = ["a photo of a red bench", "two dogs", "", "a very long prompt " * 12, "x"]
    for ci, (N, text_len, max_seq, task_first) in enumerate([(16, 20, 4096, False), (8, None, 4096, False), (16, 6, 4096, False),
                                                             (4, 128, 40, True), (256, 128, 1600, False)]):
        tok = StubTokenizer()
        up = pu.UniversalPromptingQwen2(tok, max_seq_len=max_seq, task_token_first=task_first)
        img = torch.full((len(prompts), N), 159866, dtype=torch.long)
        img[0, :3] = torch.tensor([151674, 151675, 159000][:3])
        inp = (prompts, img) if text_len is None else (prompts, img, text_len)
        ids, mask = up(inp, "t2i_gen")
        conv_start = tok("<|t2i|><|im_start|>user\n" if task_first else "<|im_start|><|t2i|>user\n").input_ids
        conv_end = tok("<|im_end|>\n<|im_start|>assistant\n").input_ids
        cases.append(dict(N=N, text_len=text_len, max_seq_len=max_seq, conv_start_ids=conv_start, conv_end_ids=conv_end,
                          text_ids=tok(prompts)["input_ids"], image_ids=img.tolist(), pad_id=tok.pad_token_id,
                          soi_id=int(up.sptids_dict["<|soi|>"]), eoi_id=int(up.sptids_dict["<|eoi|>"]),
                          input_ids=ids.tolist(), mask=mask.tolist()))
        print("case", ci, "ids", tuple(ids.shape))
    path = os.path.join(ROOT, "tests", "golden", "prompting.json")
    with open(path, "w") as f:
        json.dump(cases, f)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    assert ref_loader.available(), "reference tree not mounted"
    main()
