"""conv3d stack timing (graph replay) for the default kernel and variant 1 (no tail split), SceneFlow B=8."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from decnet_b200 import _lib, conv3d as c3
from decnet_b200.synthetic import build_workload
model, left, right, info = build_workload("sceneflow", 8, rho=0.1)
pk = {"hbm_gbs": 6538.0, "bf16_tflops": 1619.4, "bf16_tflops_sustained": 1370.2, "source": "measured"}
for v in (0, 1, 0):
    _lib.lib().decnet_conv3d_set_variant(v)
    r = c3.measure_roofline(model, left["stage0"], right["stage0"], 8, pk)
    print("variant", v, round(r["us_per_stack"], 1), "us", round(r["frac"], 4))
_lib.lib().decnet_conv3d_set_variant(0)
