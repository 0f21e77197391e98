"""GridAnchorGenerator on the device
(/root/reference/object_detection/anchor_generators/grid_anchor_generator.py:30-214)."""
import torch

from .. import ops


class GridAnchorGenerator(object):
    def __init__(self, scales=(0.5, 1.0, 2.0), aspect_ratios=(0.5, 1.0, 2.0), base_anchor_size=None,
                 anchor_stride=None, anchor_offset=None):
        self._scales = [float(s) for s in scales]
        self._aspect_ratios = [float(a) for a in aspect_ratios]
        self._base_anchor_size = [256.0, 256.0] if base_anchor_size is None else [float(v) for v in base_anchor_size]
        self._anchor_stride = [16.0, 16.0] if anchor_stride is None else [float(v) for v in anchor_stride]
        self._anchor_offset = [0.0, 0.0] if anchor_offset is None else [float(v) for v in anchor_offset]
        self._cache = {}

    def name_scope(self):
        return "GridAnchorGenerator"

    def num_anchors_per_location(self):
        return [len(self._scales) * len(self._aspect_ratios)]

    def generate(self, feature_map_shape_list, device="cuda"):
        """-> float32 [H*W*A, 4] absolute-pixel anchors, order (y, x, a) (cached per shape)."""
        if not (isinstance(feature_map_shape_list, list) and len(feature_map_shape_list) == 1):
            raise ValueError("feature_map_shape_list must be a list of length 1.")
        if not all(isinstance(p, tuple) and len(p) == 2 for p in feature_map_shape_list):
            raise ValueError("feature_map_shape_list must be a list of pairs.")
        h, w = feature_map_shape_list[0]
        key = (int(h), int(w), str(device))
        if key not in self._cache:
            a = self.num_anchors_per_location()[0]
            out = torch.empty(h * w * a, 4, dtype=torch.float32, device=device)
            ops.call("mtl_grid_anchors", h, w, self._scales, len(self._scales), self._aspect_ratios,
                     len(self._aspect_ratios), self._base_anchor_size[0], self._base_anchor_size[1],
                     self._anchor_stride[0], self._anchor_stride[1], self._anchor_offset[0],
                     self._anchor_offset[1], out)
            self._cache[key] = out
        return self._cache[key]
