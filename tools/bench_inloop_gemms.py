"""In-loop timing of every GEMM and attention launch of the step (CUDA events inside two full generations, eager)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ml_unigen_b200 import prompting
from ml_unigen_b200.unigen import UniGenB200

dev = torch.device("cuda:0")
model = UniGenB200.random_init(dev, seed=0)
model.use_cuda_graph = False
ids = prompting.synthetic_prompt_batch(16, 1024, seed=11, mask_token_id=model.mask_token_id).to(dev)
kw = dict(timesteps=18, image_token_num_per_image=1024, text_vocab_size=prompting.QWEN_TEXT_VOCAB, philox_seed=1, validate_mask=False)
for _ in range(2):
    model.t2i_generate(input_ids=ids, **kw)
bb = model.backbone
bb.attn_events, bb.mlp_events, bb.plain_events = [], [], []
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2):
    model.t2i_generate(input_ids=ids, **kw)
e1.record()
torch.cuda.synchronize()
M = 16 * 1026
avg = lambda xs: sum(xs) / max(len(xs), 1)
res = dict(mode="own kernels only (residual epilogues, folded RMSNorm)", ms_per_generation=e0.elapsed_time(e1) / 2,
           gate_up=avg([a.elapsed_time(b) for a, b, m in bb.mlp_events if m == M]),
           attention=avg([a.elapsed_time(b) for a, b, lq in bb.attn_events if lq == 1026]),
           down=avg([a.elapsed_time(b) for a, b, m, k in bb.plain_events if m == M and k == 8960]),
           o=avg([a.elapsed_time(b) for a, b, m, k in bb.plain_events if m == M and k == 1536]))
print(json.dumps(res))
