"""Image resizing used inside DetectionModel.preprocess
(/root/reference/object_detection/core/preprocessor.py:1362-1419 `resize_to_range`,
:1452-1492 `resize_image`).  Data augmentation lives outside the hot path."""
import torch

from .. import ops


def _compute_new_static_size(h, w, min_dimension, max_dimension):
    """preprocessor.py:1272-1307 (static-shape branch): scale the smaller side to min_dimension
    unless that makes the larger side exceed max_dimension."""
    orig_min, orig_max = min(h, w), max(h, w)
    large_scale = min_dimension / float(orig_min)
    large_h, large_w = int(round(h * large_scale)), int(round(w * large_scale))
    new = (large_h, large_w)
    if max_dimension:
        small_scale = max_dimension / float(orig_max)
        small_h, small_w = int(round(h * small_scale)), int(round(w * small_scale))
        if max(large_h, large_w) > max_dimension:
            new = (small_h, small_w)
    return new


def resize_image(image, new_h, new_w):
    """tf.image.resize_images(..., BILINEAR, align_corners=False) on a [B,H,W,C] float32 batch."""
    B, H, W, C = image.shape
    if (H, W) == (new_h, new_w):
        return image
    out = torch.empty(B, new_h, new_w, C, dtype=torch.float32, device=image.device)
    ops.call("mtl_resize_bilinear_f32", image, B, H, W, C, new_h, new_w, out)
    return out


def resize_to_range(image, min_dimension=None, max_dimension=None):
    B, H, W, C = image.shape
    nh, nw = _compute_new_static_size(H, W, min_dimension, max_dimension)
    return resize_image(image, nh, nw)
