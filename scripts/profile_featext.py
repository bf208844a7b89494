"""One feature-extractor forward (B images 540x972) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from decnet_b200.features import FeatExtNetChannelPlus
from decnet_b200.params import make_featext_state
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
fe = FeatExtNetChannelPlus(8, precision=prec)
fe.load_state_dict(make_featext_state(17))
fe = fe.cuda()
x = torch.randn(B, 3, 540, 972, device="cuda")
for _ in range(3):
    fe(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    fe(x)
    with torch.cuda.graph(g, stream=s):
        fe(x)
torch.cuda.current_stream().wait_stream(s)
g.replay(); torch.cuda.synchronize()
e0.record()
for _ in range(5):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(f"extractor {prec} B={B}: {e0.elapsed_time(e1) / 5:.3f} ms per forward (graph replay)")
torch.cuda.profiler.start()
fe(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
