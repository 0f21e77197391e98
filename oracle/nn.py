"""Oracle (test infrastructure): differentiable CPU restatements (torch, fp32) of the
TensorFlow-1.7 kernels and reference loss classes used on the hot path.

torch is used here only as a CPU fp32 array library with autograd, so that the same
restatement also yields reference gradients.  Paths relative to /root/reference/.
"""

import numpy as np
import torch
import torch.nn.functional as TF


def crop_and_resize(image, boxes, box_ind, crop_size, extrapolation_value=0.0):
    """tf.image.crop_and_resize (TF 1.7 core/kernels/crop_and_resize_op.cc), bilinear.

    image [B,H,W,C]; boxes [R,4] normalised (y1,x1,y2,x2) mapped with y*(H-1) (trap T13);
    box_ind [R]; returns [R,ch,cw,C].  Call site: object_detection/meta_architectures/
    faster_rcnn_meta_arch.py:1340-1344.
    """
    B_, H, W, C = image.shape
    ch, cw = crop_size
    boxes = boxes.to(torch.float32)
    y1, x1, y2, x2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    f32 = torch.float32
    if ch > 1:
        hs = (y2 - y1) * (H - 1) / (ch - 1)
        in_y = y1[:, None] * (H - 1) + torch.arange(ch, dtype=f32)[None, :] * hs[:, None]
    else:
        in_y = (0.5 * (y1 + y2) * (H - 1))[:, None]
    if cw > 1:
        ws = (x2 - x1) * (W - 1) / (cw - 1)
        in_x = x1[:, None] * (W - 1) + torch.arange(cw, dtype=f32)[None, :] * ws[:, None]
    else:
        in_x = (0.5 * (x1 + x2) * (W - 1))[:, None]
    in_y = in_y.detach()
    in_x = in_x.detach()
    valid_y = (in_y >= 0) & (in_y <= H - 1)
    valid_x = (in_x >= 0) & (in_x <= W - 1)
    top = torch.floor(in_y).clamp(0, H - 1).long()
    bot = torch.ceil(in_y).clamp(0, H - 1).long()
    yl = (in_y - torch.floor(in_y))
    left = torch.floor(in_x).clamp(0, W - 1).long()
    right = torch.ceil(in_x).clamp(0, W - 1).long()
    xl = (in_x - torch.floor(in_x))
    bi = box_ind.long()[:, None, None]
    T, L = top[:, :, None], left[:, None, :]
    Bo, R = bot[:, :, None], right[:, None, :]
    tl = image[bi, T, L]
    tr = image[bi, T, R]
    bl = image[bi, Bo, L]
    br = image[bi, Bo, R]
    xl_ = xl[:, None, :, None]
    yl_ = yl[:, :, None, None]
    topv = tl + (tr - tl) * xl_
    botv = bl + (br - bl) * xl_
    out = topv + (botv - topv) * yl_
    valid = (valid_y[:, :, None] & valid_x[:, None, :])[..., None]
    return torch.where(valid, out, torch.full_like(out, extrapolation_value))


def resize_bilinear(images, out_hw):
    """tf.image.resize_images -> ResizeBilinear(align_corners=False) of TF 1.7
    (core/kernels/resize_bilinear_op.cc): src = dst * in/out, lower=floor, upper=min(lower+1,in-1).
    images [B,H,W,C] -> [B,oh,ow,C].  Call site: faster_rcnn_meta_arch.py:1870-1871."""
    B_, H, W, C = images.shape
    oh, ow = out_hw
    f32 = torch.float32
    sy, sx = H / oh, W / ow
    in_y = torch.arange(oh, dtype=f32) * np.float32(sy)
    in_x = torch.arange(ow, dtype=f32) * np.float32(sx)
    y0 = torch.floor(in_y).long()
    y1 = torch.clamp(y0 + 1, max=H - 1)
    yl = (in_y - y0.to(f32))[None, :, None, None]
    x0 = torch.floor(in_x).long()
    x1 = torch.clamp(x0 + 1, max=W - 1)
    xl = (in_x - x0.to(f32))[None, None, :, None]
    tl = images[:, y0][:, :, x0]
    tr = images[:, y0][:, :, x1]
    bl = images[:, y1][:, :, x0]
    br = images[:, y1][:, :, x1]
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    return top + (bot - top) * yl


def smooth_l1(pred, target, weights, sigma=1.0):
    """object_detection/core/losses.py:169-196 WeightedSmoothL1LocalizationLoss (anchorwise):
    |d| < 1/sigma^2 ? 0.5 sigma^2 d^2 : |d| - 0.5/sigma^2, summed over the code dim, * weights."""
    d = (pred - target).abs()
    s2 = sigma ** 2
    l = torch.where(d < 1.0 / s2, 0.5 * d * d * s2, d - 0.5 / s2)
    return l.sum(-1) * weights


def softmax_ce(logits, labels, weights=None):
    """object_detection/core/losses.py:285-311 / :326-352 (tf.nn.softmax_cross_entropy_with_logits
    [_v2] with constant labels): -sum_k labels_k * log_softmax(logits)_k per row, * weights."""
    ce = -(labels * TF.log_softmax(logits, dim=-1)).sum(-1)
    return ce if weights is None else ce * weights


def same_pad(in_size, k, stride, rate=1):
    """TensorFlow 'SAME' padding arithmetic: returns (out_size, pad_begin, pad_end)."""
    out = -(-in_size // stride)
    ke = k + (k - 1) * (rate - 1)
    total = max((out - 1) * stride + ke - in_size, 0)
    return out, total // 2, total - total // 2


def conv2d_tf(x, w, stride=1, padding="SAME", rate=1, bias=None):
    """slim.conv2d on NHWC input with HWIO weights (TF SAME/VALID semantics)."""
    xn = x.permute(0, 3, 1, 2)
    wn = w.permute(3, 2, 0, 1)
    if padding == "SAME":
        _, pt, pb = same_pad(x.shape[1], w.shape[0], stride, rate)
        _, pl, pr = same_pad(x.shape[2], w.shape[1], stride, rate)
        xn = TF.pad(xn, (pl, pr, pt, pb))
    y = TF.conv2d(xn, wn, bias=bias, stride=stride, dilation=rate)
    return y.permute(0, 2, 3, 1)


def conv2d_same(x, w, stride, rate=1):
    """slim/nets/resnet_utils.py:77-122: SAME for stride 1, explicit symmetric-ish pad + VALID else."""
    k = w.shape[0]
    if stride == 1:
        return conv2d_tf(x, w, 1, "SAME", rate)
    ke = k + (k - 1) * (rate - 1)
    pad_total = ke - 1
    pb = pad_total // 2
    pe = pad_total - pb
    xp = TF.pad(x, (0, 0, pb, pe, pb, pe))
    return conv2d_tf(xp, w, stride, "VALID", rate)


def max_pool_tf(x, k, stride, padding="SAME"):
    """slim.max_pool2d NHWC (SAME pads with -inf semantics: padded cells never win)."""
    xn = x.permute(0, 3, 1, 2)
    if padding == "SAME":
        _, pt, pb = same_pad(x.shape[1], k, stride)
        _, pl, pr = same_pad(x.shape[2], k, stride)
        xn = TF.pad(xn, (pl, pr, pt, pb), value=float("-inf"))
    y = TF.max_pool2d(xn, k, stride)
    return y.permute(0, 2, 3, 1)
