cd /root/repo
cat > /tmp/one3.py <<'PY'
import sys; sys.path.insert(0, "/root/repo")
import torch
from decnet_b200 import ops
B, h, w, ci, co = 8, 180, 324, 81, 81
cp = (ci + 7) // 8 * 8
x = torch.randn(B, h, w, cp, device="cuda")
wt = torch.randn(co, ci, 3, 3, device="cuda") * 0.05
wp, bp, np_ = ops.pack_conv2d_tf32_weights(wt, torch.zeros(co, device="cuda"), cp)
for _ in range(3):
    y = ops.conv2d_tf32_nhwc(x, wp, bp, True)
torch.cuda.synchronize()
PY
timeout 280 ncu --set full --clock-control none --import-source on -k regex:conv3d_tcgen05 -s 2 -c 1 -o gpurun_out/r01_conv2d_nhwc_tf32 python /tmp/one3.py > gpurun_out/ncu_c3.log 2>&1
tail -3 gpurun_out/ncu_c3.log
