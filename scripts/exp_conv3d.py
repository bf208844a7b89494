import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT))
import torch
from decnet_b200.synthetic import build_workload
from decnet_b200 import conv3d as c3
from bench import peaks
model, left, right, info = build_workload("sceneflow", 8)
print(os.environ.get("DECNET_B200_LIB", "default"), c3.measure_roofline(model, left["stage0"], right["stage0"], 8, peaks()))
