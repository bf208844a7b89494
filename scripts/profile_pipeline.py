"""Per-kernel time breakdown of one hot-path step (torch.profiler / CUPTI), to see Amdahl.
Usage: python scripts/profile_pipeline.py [--precision fp32|tf32] [--batch 8]"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from decnet_b200.synthetic import build_workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp32")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--workload", default="sceneflow")
args = ap.parse_args()
torch.backends.cudnn.benchmark = True
model, left, right, info = build_workload(args.workload, args.batch, precision=args.precision)
print(info)
for _ in range(3):
    model(left, right)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        model(left, right)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
