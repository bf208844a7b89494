"""Timing of the device-side image-space detail detector for both views of 8 SceneFlow pairs."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200 import ops
x = torch.rand(16, 3, 540, 972, device="cuda")
for _ in range(3):
    ops.detail_detection(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.detail_detection(x)
e1.record(); torch.cuda.synchronize()
print(f"detail_detection, 16 images 540x972, 3 levels: {e0.elapsed_time(e1) / 20:.3f} ms ({e0.elapsed_time(e1) / 20 / 16 * 1e3:.1f} us per image)")
