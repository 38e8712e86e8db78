"""Small calls of the attention, decode-step and GEMM kernels for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`
(round 2 at HEAD: 0 errors)."""
import torch, sys, os
sys.path.insert(0, os.getcwd())
from ml_unigen_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
B, L, Hq, Hkv, D = 2, 395, 12, 2, 128
q = torch.randn(B, L, Hq, D, device=dev, generator=g).bfloat16(); k = torch.randn(B, L, Hkv, D, device=dev, generator=g).bfloat16(); v = torch.randn(B, L, Hkv, D, device=dev, generator=g).bfloat16()
P = torch.tensor([0, 57], dtype=torch.int32, device=dev); S = torch.tensor([137, 137], dtype=torch.int32, device=dev)
o = ops.omni_attention(q, k, v, P, S, D ** -0.5)
ow = ops.omni_attention(q[:, 137:], k, v, P, S, D ** -0.5, q_offset=137)
torch.cuda.synchronize()
print("attn ok", bool(torch.equal(o[:, 137:], ow)))
Bd, N, V = 2, 64, 8192
logits = (0.02 * torch.randn(Bd, N, V, device=dev, generator=g)).bfloat16()
ws = ops.DecodeWorkspace(Bd, N, V, dev)
ids = torch.full((Bd, N), 159866, dtype=torch.int64, device=dev)
s, m = ops.decode_step(logits, ids, mask_token_id=159866, mask_len_cap=40, temperature=0.7, workspace=ws, seed=1, offset=0)
x = torch.randn(300, 1536, device=dev, generator=g).bfloat16(); w = (0.02 * torch.randn(512, 1536, device=dev, generator=g)).bfloat16()
y = ops.gemm(x, w)
torch.cuda.synchronize()
print("decode+gemm ok", int(m.sum()))
