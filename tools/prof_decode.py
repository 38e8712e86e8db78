"""One decode-step launch shape for ncu: B=16, N=1024, V=8192, Philox mode, step-0 state."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ml_unigen_b200 import ops
dev = torch.device("cuda:0")
B, N, V = int(os.environ.get("B", 16)), 1024, 8192
g = torch.Generator(device=dev).manual_seed(0)
logits = (float(os.environ.get("SCALE", 4.0)) * torch.randn(B, N, V, device=dev, generator=g)).bfloat16()
ws = ops.DecodeWorkspace(B, N, V, dev)
ids0 = torch.full((B, N), 159866, dtype=torch.int64, device=dev)
for it in range(int(os.environ.get("ITERS", 4))):
    ids = ids0.clone()
    ops.decode_step(logits, ids, mask_token_id=159866, mask_len_cap=1020, temperature=0.5, workspace=ws, seed=1, offset=it)
torch.cuda.synchronize()
print("ok")
