"""One decode_code call of the MAGViT-v2 pixel decoder (16 token grids of 32x32 -> 16 x 512x512) inside cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum` (launch list; tools/launch_summary.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_unigen_b200.magvit import MagvitV2DecoderB200  # noqa: E402

dev = torch.device("cuda:0")
dec = MagvitV2DecoderB200.random_init(dev, seed=0)
ids = torch.randint(0, 8192, (int(os.environ.get("B", 16)), int(os.environ.get("N", 1024))), device=dev)
for _ in range(2):
    dec.decode_code(ids)
torch.cuda.synchronize()
torch.cuda.profiler.start()
dec.decode_code(ids)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
