"""oracle/codec.py -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's host-side formats around the path
(SURVEY.md section 8f rank 4).  demo.py runs argparse and builds a CUDA model at import time, so its helpers cannot be
imported; they are restated here line by line (parity pinned only by the reference's third-party calls where those
are importable: torchvision's ToTensor / Normalize in tests/test_codec_gpu.py when torchvision is present).

  pad_topleft     demo.py:75-81   `padding`: zeros((h+rh, w+rw, c), float32); padded[rh:, rw:] = img
  image01         demo.py:158     padding(img) / 255                      (float32 / int -> float32)
  normalize       demo.py:82-88   ToTensor (HWC float -> CHW, no rescale) + Normalize(mean, std): (x - mean) / std, float32
  disp_to_u16     demo.py:191-197 out = pred*256; out[out<0] = 0; out[out>65535] = 65535; astype('uint16'); [-ori_h:, -ori_w:]
  epe_3px         modules/loss.py:427-437 `test_loss_func`
"""
import numpy as np
import torch

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def pad_topleft(img, multiple=27):
    h, w, c = img.shape
    rh = int(np.ceil(h / multiple) * multiple) - h
    rw = int(np.ceil(w / multiple) * multiple) - w
    out = np.zeros((h + rh, w + rw, c), dtype=np.float32)
    out[rh:, rw:] = img
    return out


def image01(img_u8):
    return pad_topleft(img_u8) / 255


def normalize(img01):
    x = np.ascontiguousarray(img01.transpose(2, 0, 1)).astype(np.float32)
    return ((x - MEAN[:, None, None]) / STD[:, None, None]).astype(np.float32)[None]


def disp_to_u16(pred, ori_h, ori_w):
    out = pred.astype(np.float32) * 256
    out[out < 0] = 0
    out[out > 65535] = 65535
    return out.astype("uint16")[:, -ori_h:, -ori_w:]


def epe_3px(pred, gt, max_disp):
    mask = (gt < max_disp) & (gt > 0)
    err = torch.abs(pred[mask] - gt[mask])
    ok = ((err < 3) | (err < 0.05 * gt[mask])).float()
    return torch.mean(err), 100 - torch.sum(ok) / torch.sum(mask) * 100
