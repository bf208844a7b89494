"""oracle/detail.py -- TEST INFRASTRUCTURE ONLY (never imported by decnet_b200/).

CPU restatement (numpy, no cv2) of the reference's image-space lost-detail detector
`detailDetection(img, scale=3, downsampling_iteration=3, thold=0.3)` as demo.py:161-162 calls it
(/root/reference/utils/utils.py:447-534), for an image already padded to a multiple of 27:

  per level:  down = resize_linear(GaussianBlur(data, 3x3, sigma 1), 1/3)       utils.py:447-462
              up   = GaussianBlur(resize_linear(down, x3), 5x5, sigma 1)        utils.py:464-480
              r    = sum_c |data - up| ;  mask = (r - min) / (max - min) >= thold ;  data = down     :510-519
  (`cv2.resize(img, dsize, cv2.INTER_AREA)` passes INTER_AREA as the `dst` argument, so the interpolation is the
   default INTER_LINEAR: the x3 reduction samples exactly src[3y+1, 3x+1]; diffusion(iteration=0) is the identity.)

cv2 semantics restated: GaussianBlur = float32 separable filter, rows then columns, BORDER_REFLECT_101, kernel
exp(-x^2/2)/sum in float32; resize INTER_LINEAR: fx = float((dx+0.5)/3 - 0.5), sx = floor(fx), fx -= sx, left clamp
(sx<0 -> sx=0, fx=0), right clamp (sx >= w-1 -> sx=w-1, fx=0), horizontal pass then vertical pass in float32.
Pinned by tests/golden/detail_masks.npz (masks returned by the UNMODIFIED reference function, cv2 4.13).
"""
import numpy as np


def gaussian_kernel(ksize):
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) / 2
    k = np.exp(-x * x / 2.0)
    return (k / k.sum()).astype(np.float32)


def _reflect101(i, n):
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def blur(img, ksize):
    """img [H,W,C] float32."""
    k = gaussian_kernel(ksize)
    r = ksize // 2
    H, W, _ = img.shape
    xs = _reflect101(np.arange(-r, W + r), W)
    tmp = np.zeros_like(img)
    for j in range(ksize):
        tmp += k[j] * img[:, xs[j:j + W]]
    ys = _reflect101(np.arange(-r, H + r), H)
    out = np.zeros_like(img)
    for j in range(ksize):
        out += k[j] * tmp[ys[j:j + H]]
    return out


def down3(img):
    return np.ascontiguousarray(img[1::3, 1::3])


def _lin_coeff(n_dst, n_src):
    d = np.arange(n_dst, dtype=np.float64)
    fx = ((d + 0.5) * (n_src / n_dst) - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx).astype(np.float32)
    lo = sx < 0
    sx[lo] = 0; fx[lo] = 0
    hi = sx >= n_src - 1
    sx[hi] = n_src - 1; fx[hi] = 0
    return sx, np.minimum(sx + 1, n_src - 1), fx


def up3(img):
    h, w, _ = img.shape
    sx, sx1, fx = _lin_coeff(3 * w, w)
    sy, sy1, fy = _lin_coeff(3 * h, h)
    hor = img[:, sx] * (1 - fx)[None, :, None] + img[:, sx1] * fx[None, :, None]
    return (hor[sy] * (1 - fy)[:, None, None] + hor[sy1] * fy[:, None, None]).astype(np.float32)


def detail_masks(img, iters=3, thold=0.3, return_residuals=False):
    """img [H,W,3] in [0,1], H and W multiples of 3**iters -> [mask_full, mask_1/3, mask_1/9] (bool)."""
    data = np.asarray(img, dtype=np.float32)
    masks, res = [], []
    for _ in range(iters):
        dn = down3(blur(data, 3))
        up = blur(up3(dn), 5)
        r = np.abs(data - up).sum(axis=2)
        t = (r - r.min()) / (r.max() - r.min())
        masks.append(t >= thold)
        res.append(r)
        data = dn
    return (masks, res) if return_residuals else masks
