"""CPU: pins oracle/codec.py (the restatement the format kernels are checked against) to the
reference's OWN code.  demo.py cannot be imported (it parses argv and builds a CUDA model at import time), so its helper
functions `padding` / `transform` (demo.py:75-88) and its output block (demo.py:191-197) are extracted from the unmodified
source with `ast` and executed here.  Skipped where the reference tree is absent."""
import ast

import numpy as np
import pytest
import torch

from oracle import codec, ref_loader

pytestmark = pytest.mark.skipif(not (ref_loader.REF / "demo.py").exists(), reason="reference tree not present")


def _demo_namespace():
    import torchvision.transforms as transforms
    src = (ref_loader.REF / "demo.py").read_text()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch, "transforms": transforms}
    funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("padding", "transform")]
    assert sorted(f.name for f in funcs) == ["padding", "transform"]
    exec(compile(ast.Module(body=funcs, type_ignores=[]), "demo.py", "exec"), ns)
    # the statements that turn the prediction into the 16-bit image: `output = pred_list[-1]` ... `pre = output_np[...]`
    block = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], (ast.Name, ast.Subscript)):
            seg = ast.get_source_segment(src, node)
            if seg.startswith(("output = pred_list[-1]", "output = output*256", "output[output", "output_np =", "pre = output_np")):
                block.append(node)
    block.sort(key=lambda n: n.lineno)
    assert len(block) == 7 and block[-1].lineno - block[0].lineno == 6, [ast.get_source_segment(src, b) for b in block]
    return ns, compile(ast.Module(body=block, type_ignores=[]), "demo.py", "exec")


def test_codec_restatement_equals_demo_py_functions():
    ns, out_block = _demo_namespace()
    rng = np.random.default_rng(11)
    for h, w in ((100, 150), (108, 162), (375, 1242), (1, 1)):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        ref_pad = ns["padding"](img)
        assert ref_pad.dtype == np.float32 and np.array_equal(ref_pad, codec.pad_topleft(img))
        ref01 = ref_pad / 255                                            # demo.py:158
        assert np.array_equal(ref01, codec.image01(img))
        ref_t = ns["transform"](ref01)                                   # demo.py:164 (ToTensor + Normalize)
        got_t = torch.from_numpy(codec.normalize(ref01))
        assert ref_t.shape == got_t.shape
        # ToTensor yields float64 here (the padded image / 255 is float32 / int -> float32; Normalize in that dtype):
        assert torch.allclose(ref_t, got_t, atol=1e-6, rtol=0), float((ref_t - got_t).abs().max())
    g = torch.Generator().manual_seed(5)
    pred = torch.rand(1, 108, 162, generator=g) * 300 - 10               # negatives and values beyond 65535/256 included
    env = {"pred_list": [pred.clone()], "ori_h": 100, "ori_w": 150}
    exec(out_block, env)                                                 # demo.py:191-197, unmodified
    assert env["pre"].dtype == np.uint16
    assert np.array_equal(env["pre"], codec.disp_to_u16(pred.numpy(), 100, 150)[0])
