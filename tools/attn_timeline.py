"""Cycle-stamp timeline of one omni-attention CTA: where a kv-tile iteration spends its time.
Needs a library built with -DUG_ATTN_TIMELINE=1 (`bash tools/build_attn_variants.sh "-DUG_ATTN_TIMELINE=1"`, then
`UNIGEN_B200_LIB=build/exp/lib__DUG_ATTN_TIMELINE_1.so python tools/attn_timeline.py`): the stamps are compiled out of the product."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ml_unigen_b200 import ops
dev = torch.device("cuda:0")
B, L, Hq, Hkv, D = 16, 1163, 12, 2, 128
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(B, L, Hq, D, device=dev, generator=g).bfloat16()
k = torch.randn(B, L, Hkv, D, device=dev, generator=g).bfloat16()
v = torch.randn(B, L, Hkv, D, device=dev, generator=g).bfloat16()
P = torch.zeros(B, dtype=torch.int32, device=dev)
S = torch.full((B,), 137, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.omni_attention(q, k, v, P, S, D ** -0.5)
dump = torch.zeros(32768 + 2048, dtype=torch.float32, device=dev)
ops.omni_attention(q, k, v, P, S, D ** -0.5, _debug_dump=dump)
torch.cuda.synchronize()
t = dump[32768:].cpu().view(4, 64, 8)
names = {0: "softmaxA", 1: "softmaxB", 2: "mma"}
for role in (0, 1):
    print(names[role], "events per 64-key half: wait_start, s_full, chunk0_loaded (previous half published), chunk1_loaded, P stored")
    for it in range(22):
        print(it, [int(x) for x in t[role, it, :5].tolist()])
print("mma warp of tile A, per half: wait_start, p_full, pv_issued, next_qk_issued")
for it in range(22):
    print(it, [int(x) for x in t[2, it, :4].tolist()])

ti = dump[32768 + 1536:].cpu().view(4, 16, 8)
print("item boundaries of CTA 0 (persistent grid): producer [start, q_empty ok, loads issued] | mma [start, q_full, first QK issued, done] | softmax A/B [start, wait o_done, o_done, o_free]")
for it in range(8):
    print(it, [int(x) for x in ti[0, it, :3].tolist()], [int(x) for x in ti[1, it, :4].tolist()], [int(x) for x in ti[2, it, :4].tolist()], [int(x) for x in ti[3, it, :4].tolist()])
