cd /root/repo
cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0, "/root/repo")
import torch
from decnet_b200 import ops
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(8, 8, 540, 972, device="cuda", generator=g)
w = torch.randn(8, 8, 3, 3, device="cuda", generator=g) * 0.1
b = torch.zeros(8, device="cuda")
wp, bp = ops.pack_conv2d_tf32_nchw_weights(w, b)
for _ in range(3):
    y = ops.conv2d_tf32_nchw(x, wp, bp, 8, 1, True)
torch.cuda.synchronize()
PY
timeout 280 ncu --set full --clock-control none --import-source on -k regex:conv2d_tcgen05 -s 2 -c 1 -o gpurun_out/r01_conv2d_tc python /tmp/one.py > gpurun_out/ncu_c2.log 2>&1
tail -3 gpurun_out/ncu_c2.log
