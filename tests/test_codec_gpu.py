"""GPU: the format kernels either side of the path against the line-by-line restatement
of demo.py / modules/loss.py in oracle/codec.py (and torchvision's own transforms where installed)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_image_prepare_u8_pads_scales_and_normalises_bit_exactly():
    from decnet_b200 import ops
    from oracle import codec
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, size=(2, 100, 150, 3), dtype=np.uint8)          # pads to 108 x 162
    o01, onm = ops.image_prepare_u8(torch.from_numpy(imgs).cuda())
    assert tuple(o01.shape) == (2, 3, 108, 162)
    for b in range(2):
        w01 = codec.image01(imgs[b])
        assert np.array_equal(o01[b].cpu().numpy(), w01.transpose(2, 0, 1))     # incl. the zero rows/columns at top/left
        assert np.array_equal(onm[b].cpu().numpy(), codec.normalize(w01)[0])
    try:
        from torchvision import transforms
    except Exception:
        return
    tr = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    want = tr(codec.image01(imgs[0])).float()
    assert torch.allclose(onm[0].cpu(), want, atol=1e-6, rtol=0)


def test_disp_to_u16_and_epe():
    from decnet_b200 import ops
    from oracle import codec
    g = torch.Generator(device="cuda").manual_seed(2)
    pred = (torch.rand(2, 108, 162, device="cuda", generator=g) * 300 - 10)     # negatives and > 255.99 (overflow) included
    got = ops.disp_to_u16(pred, 100, 150).cpu().numpy()
    assert np.array_equal(got, codec.disp_to_u16(pred.cpu().numpy(), 100, 150))
    gt = torch.rand(2, 108, 162, device="cuda", generator=g) * 250 - 20
    epe, l3 = ops.epe_3px(pred, gt, 192.0)
    wepe, wl3 = codec.epe_3px(pred.cpu(), gt.cpu(), 192.0)
    assert abs(float(epe) - float(wepe)) <= 1e-4 * float(wepe) and abs(float(l3) - float(wl3)) <= 1e-3
