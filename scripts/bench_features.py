"""Timing of the feature-extractor drop-in at SceneFlow size (both views of B pairs), TF32 tensor-core route vs
the same layers on cuDNN (reference-style modules with folded weights)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200 import model as dm
from decnet_b200.features import FeatExtNetChannelPlus
from decnet_b200.params import make_featext_state

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = FeatExtNetChannelPlus(8); m.load_state_dict(make_featext_state(1)); m = m.cuda()
x = torch.randn(2 * B, 3, 540, 972, device="cuda")
torch.backends.cudnn.benchmark = True

def timeit(iters=10):
    for _ in range(3):
        m(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        m(x)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

t_ours = timeit()
dm.USE_NATIVE_CONV2D = False
dm._reset_folded(m)
t_cudnn = timeit()
dm.USE_NATIVE_CONV2D = True
print(f"feature pyramids of {B} pairs (2x{B} images 540x972): ours {t_ours:.2f} ms, all-cuDNN (TF32) {t_cudnn:.2f} ms")
from torch.profiler import ProfilerActivity, profile
dm._reset_folded(m)
m(x); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m(x); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
