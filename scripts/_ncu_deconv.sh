cd /root/repo
cat > /tmp/one_d.py <<'PY'
import sys; sys.path.insert(0, "/root/repo")
import torch
from decnet_b200 import ops
xp = torch.randn(8, 24, 180, 324, device="cuda")
wd = torch.randn(24, 8, 3, 3, device="cuda") * 0.1
bd = torch.zeros(8, device="cuda")
for _ in range(3):
    y = ops.deconv3x3s3(xp, wd, bd, True)
torch.cuda.synchronize()
PY
timeout 280 ncu --set full --clock-control none --import-source on -k regex:deconv3x3s3 -s 2 -c 1 -o gpurun_out/r01_deconv python /tmp/one_d.py > gpurun_out/ncu_d.log 2>&1
tail -2 gpurun_out/ncu_d.log
