"""Oracle (test infrastructure): box geometry, anchors, box coder, IoU — NumPy float32.

Every function restates the cited reference code op-for-op in float32 so that integer /
index results are bit-exact and float results agree to the last ulp where the reference
uses only + - * / (transcendentals: documented tolerance).
All paths are relative to /root/reference/object_detection/.
"""
import numpy as np

F = np.float32


def meshgrid(x, y):
    """utils/ops.py:78-120 `meshgrid(x, y)`: NumPy-style broadcasted grids."""
    x = np.asarray(x)
    y = np.asarray(y)
    xg = np.tile(x.reshape((1,) * y.ndim + x.shape), y.shape + (1,) * x.ndim)
    yg = np.tile(y.reshape(y.shape + (1,) * x.ndim), (1,) * y.ndim + x.shape)
    return xg, yg


def grid_anchors(grid_height, grid_width, scales=(0.5, 1.0, 2.0), aspect_ratios=(0.5, 1.0, 2.0),
                 base_anchor_size=(256, 256), anchor_stride=(16, 16), anchor_offset=(0, 0)):
    """anchor_generators/grid_anchor_generator.py:96-214 (`_generate` + `tile_anchors`).

    Order: (y, x, a) with a = aspect_idx * len(scales) + scale_idx (meshgrid at :125-128).
    Returns [H*W*A, 4] float32 [ymin, xmin, ymax, xmax] in absolute pixels.
    """
    scales_grid, ar_grid = meshgrid(np.asarray(scales, F), np.asarray(aspect_ratios, F))
    scales_f = scales_grid.reshape(-1).astype(F)
    ars = ar_grid.reshape(-1).astype(F)
    ratio_sqrts = np.sqrt(ars).astype(F)
    heights = (scales_f / ratio_sqrts * F(base_anchor_size[0])).astype(F)
    widths = (scales_f * ratio_sqrts * F(base_anchor_size[1])).astype(F)
    y_points = (np.arange(grid_height).astype(F) * F(anchor_stride[0]) + F(anchor_offset[0])).astype(F)
    x_points = (np.arange(grid_width).astype(F) * F(anchor_stride[1]) + F(anchor_offset[1])).astype(F)
    x_points, y_points = meshgrid(x_points, y_points)
    widths_grid, x_points_grid = meshgrid(widths, x_points)
    heights_grid, y_points_grid = meshgrid(heights, y_points)
    centers = np.stack([y_points_grid, x_points_grid], axis=3).reshape(-1, 2)
    sizes = np.stack([heights_grid, widths_grid], axis=3).reshape(-1, 2)
    # _center_size_bbox_to_corners_bbox (:217-230)
    return np.concatenate([centers - F(0.5) * sizes, centers + F(0.5) * sizes], axis=1).astype(F)


def anchor_base_sizes(scales, aspect_ratios, base_anchor_size=(256, 256)):
    """The per-location (height, width) table of `tile_anchors` (:179-184), [A, 2] float32."""
    scales_grid, ar_grid = meshgrid(np.asarray(scales, F), np.asarray(aspect_ratios, F))
    s = scales_grid.reshape(-1).astype(F)
    r = np.sqrt(ar_grid.reshape(-1).astype(F)).astype(F)
    return np.stack([(s / r * F(base_anchor_size[0])).astype(F),
                     (s * r * F(base_anchor_size[1])).astype(F)], axis=1)


def area(boxes):
    """core/box_list_ops.py:50-63."""
    boxes = np.asarray(boxes, F)
    return ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])).astype(F)


def intersection(b1, b2):
    """core/box_list_ops.py:201-224: [N, M] pairwise intersection areas."""
    b1 = np.asarray(b1, F)
    b2 = np.asarray(b2, F)
    ymin1, xmin1, ymax1, xmax1 = [b1[:, i:i + 1] for i in range(4)]
    ymin2, xmin2, ymax2, xmax2 = [b2[:, i:i + 1] for i in range(4)]
    ih = np.maximum(F(0), np.minimum(ymax1, ymax2.T) - np.maximum(ymin1, ymin2.T))
    iw = np.maximum(F(0), np.minimum(xmax1, xmax2.T) - np.maximum(xmin1, xmin2.T))
    return (ih * iw).astype(F)


def iou(b1, b2):
    """core/box_list_ops.py:253-272; exactly 0 where the intersection is 0 (trap T7)."""
    inter = intersection(b1, b2)
    unions = (area(b1)[:, None] + area(b2)[None, :] - inter).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (inter / unions).astype(F)
    return np.where(inter == 0, F(0), q).astype(F)


def ioa(b1, b2):
    """core/box_list_ops.py:296-315: intersection over area of b2."""
    inter = intersection(b1, b2)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / area(b2)[None, :]).astype(F)


def clip_to_window(boxes, window, filter_nonoverlapping=True):
    """core/box_list_ops.py:102-137. Returns (clipped boxes, kept indices)."""
    boxes = np.asarray(boxes, F).reshape(-1, 4)
    wy0, wx0, wy1, wx1 = [F(v) for v in window]
    ymin = np.maximum(np.minimum(boxes[:, 0], wy1), wy0)
    ymax = np.maximum(np.minimum(boxes[:, 2], wy1), wy0)
    xmin = np.maximum(np.minimum(boxes[:, 1], wx1), wx0)
    xmax = np.maximum(np.minimum(boxes[:, 3], wx1), wx0)
    clipped = np.stack([ymin, xmin, ymax, xmax], axis=1).astype(F)
    idx = np.arange(len(clipped))
    if filter_nonoverlapping:
        idx = np.nonzero(area(clipped) > 0)[0]
        clipped = clipped[idx]
    return clipped, idx


def prune_outside_window(boxes, window):
    """core/box_list_ops.py:140-169: keep boxes fully inside the window. -> (boxes, indices)."""
    boxes = np.asarray(boxes, F)
    wy0, wx0, wy1, wx1 = [F(v) for v in window]
    viol = (boxes[:, 0] < wy0) | (boxes[:, 1] < wx0) | (boxes[:, 2] > wy1) | (boxes[:, 3] > wx1)
    idx = np.nonzero(~viol)[0]
    return boxes[idx], idx


def to_normalized_coordinates(boxes, height, width):
    """core/box_list_ops.py:738-771 via `scale` (:73-99): multiply by 1/height, 1/width."""
    boxes = np.asarray(boxes, F)
    ys = F(1.0) / F(height)
    xs = F(1.0) / F(width)
    return np.stack([ys * boxes[:, 0], xs * boxes[:, 1], ys * boxes[:, 2], xs * boxes[:, 3]], axis=1).astype(F)


def to_absolute_coordinates(boxes, height, width):
    """core/box_list_ops.py:774-807 via `scale`."""
    boxes = np.asarray(boxes, F)
    ys, xs = F(height), F(width)
    return np.stack([ys * boxes[:, 0], xs * boxes[:, 1], ys * boxes[:, 2], xs * boxes[:, 3]], axis=1).astype(F)


def center_size(boxes):
    """core/box_list.py:172-183 get_center_coordinates_and_sizes."""
    boxes = np.asarray(boxes, F)
    ymin, xmin, ymax, xmax = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    w = (xmax - xmin).astype(F)
    h = (ymax - ymin).astype(F)
    return (ymin + h / F(2.0)).astype(F), (xmin + w / F(2.0)).astype(F), h, w


EPSILON = F(1e-8)
SCALE_FACTORS = (10.0, 10.0, 5.0, 5.0)


def box_encode(boxes, anchors, scale_factors=SCALE_FACTORS):
    """box_coders/faster_rcnn_box_coder.py:60-90 `_encode` -> [N, 4] (ty, tx, th, tw)."""
    yca, xca, ha, wa = center_size(anchors)
    yc, xc, h, w = center_size(boxes)
    ha = ha + EPSILON
    wa = wa + EPSILON
    h = h + EPSILON
    w = w + EPSILON
    tx = ((xc - xca) / wa).astype(F)
    ty = ((yc - yca) / ha).astype(F)
    tw = np.log((w / wa).astype(F)).astype(F)
    th = np.log((h / ha).astype(F)).astype(F)
    if scale_factors:
        ty = ty * F(scale_factors[0])
        tx = tx * F(scale_factors[1])
        th = th * F(scale_factors[2])
        tw = tw * F(scale_factors[3])
    return np.stack([ty, tx, th, tw], axis=1).astype(F)


def box_decode(rel_codes, anchors, scale_factors=SCALE_FACTORS):
    """box_coders/faster_rcnn_box_coder.py:92-118 `_decode` -> [N, 4] corners."""
    yca, xca, ha, wa = center_size(anchors)
    rel_codes = np.asarray(rel_codes, F)
    ty, tx, th, tw = rel_codes[:, 0], rel_codes[:, 1], rel_codes[:, 2], rel_codes[:, 3]
    if scale_factors:
        ty = ty / F(scale_factors[0])
        tx = tx / F(scale_factors[1])
        th = th / F(scale_factors[2])
        tw = tw / F(scale_factors[3])
    w = (np.exp(tw).astype(F) * wa).astype(F)
    h = (np.exp(th).astype(F) * ha).astype(F)
    yc = (ty * ha + yca).astype(F)
    xc = (tx * wa + xca).astype(F)
    return np.stack([yc - h / F(2.), xc - w / F(2.), yc + h / F(2.), xc + w / F(2.)], axis=1).astype(F)
