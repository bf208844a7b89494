cd /root/repo
cat > /tmp/one_h.py <<'PY'
import sys; sys.path.insert(0, "/root/repo")
import torch
import torch.nn.functional as F
from decnet_b200 import ops
B, h, w, ci, co = 8, 180, 324, 81, 81
cp = (ci + 7) // 8 * 8
x = F.pad(torch.randn(B, h, w, cp, device="cuda"), (0, 0, 1, 1, 1, 1)).contiguous()
wt = torch.randn(co, ci, 3, 3, device="cuda") * 0.05
wp, bp, np_ = ops.pack_conv2d_tf32_weights(wt, torch.zeros(co, device="cuda"), cp)
for _ in range(3):
    y = ops.conv2d_tf32_nhwc_halo(x, wp, bp, True)
torch.cuda.synchronize()
PY
timeout 280 ncu --set full --clock-control none --import-source on -k regex:conv2d_nhwc_halo -s 2 -c 1 -o gpurun_out/r01_conv2d_nhwc_halo python /tmp/one_h.py > gpurun_out/ncu_h.log 2>&1
tail -2 gpurun_out/ncu_h.log
