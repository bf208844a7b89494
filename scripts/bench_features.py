"""Timing of the feature-extractor drop-in at SceneFlow size (both views of B pairs) in both arithmetic modes, against the
reference-style layers (the oracle's torch restatement: F.conv2d + batch_norm) on cuDNN with and without TF32."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200.features import FeatExtNetChannelPlus
from decnet_b200.model import set_precision
from decnet_b200.params import make_featext_state
from oracle import features as ofe

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sd = make_featext_state(1)
m = FeatExtNetChannelPlus(8); m.load_state_dict(sd); m = m.cuda()
x = torch.randn(2 * B, 3, 540, 972, device="cuda")
torch.backends.cudnn.benchmark = True


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


t32 = timeit(lambda: m(x))
set_precision(m, "tf32")
ttf = timeit(lambda: m(x))
sdc = {k: v.cuda() for k, v in sd.items()}
res = {}
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad():
        res[tf32] = timeit(lambda: ofe.feature_pyramid(x, sdc), iters=5)
print(f"feature pyramids of {B} pairs (2x{B} images 540x972), eager: ours 3xTF32 {t32:.2f} ms, ours TF32 {ttf:.2f} ms, "
      f"cuDNN fp32 {res[False]:.2f} ms, cuDNN TF32 {res[True]:.2f} ms")
