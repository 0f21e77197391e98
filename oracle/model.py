"""Oracle (TEST INFRASTRUCTURE): CPU fp32 restatement of one Faster R-CNN + aux-heads training
step of the reference, with torch autograd supplying the reference gradients.

Follows, line by line where cited:
  object_detection/meta_architectures/faster_rcnn_meta_arch.py  predict :507-719,
      predict_with_window :721-755, predict_edgemask :757-762, predict_with_mtl_results :764-846,
      _postprocess_rpn :1055-1132, _unpad_proposals_and_sample_box_classifier_batch :1134-1216,
      _sample_box_classifier_minibatch :1268-1302, _compute_second_stage_input_feature_maps
      :1304-1348, loss :1514-1590, _loss_rpn :1591-1668, _loss_box_classifier :1670-1793,
      _loss_refined_classifier :1795-1837, _loss_window_class :1839-1858, _loss_edgemask :1860-1881
  slim/nets/resnet_v1.py :69-130, :133-237; slim/nets/resnet_utils.py :59-200
  object_detection/models/faster_rcnn_resnet_v1_feature_extractor.py :74-185
  object_detection/core/box_predictor.py :430-611, :682-755; core/mask_predictor.py :90-119
  slim/deployment/model_deploy.py :198-307 (total loss = task losses + L2 terms)
The reference cannot run here (Python 2 + TF 1.7); TF kernels are restated in oracle/nn.py.
The aux heads / refine / closeness have no reference TESTS; the graph code around them (proposal sampling, the losses
with their normalisers, the refiner's input assembly, target assignment) is pinned by executing the reference's own
methods on a NumPy TensorFlow stand-in (tests/golden/make_*_golden.py, tests/test_oracle_kats.py).

Weights come as a dict {TF variable name: tensor}, conv / FC weights in [K, R, S, C] layout
(TF's HWIO transposed).  `bf16=True` mirrors the device path's rounding points (weights with the
folded BN scale, and every activation tensor the device stores in bf16) with straight-through
gradients, so that losses can be compared tightly; `bf16=False` is the plain fp32 reference.
"""
import numpy as np
import torch
import torch.nn.functional as TF

from . import assign as OA
from . import boxes as OB
from . import nn as ON
from . import postprocess as OP

BLOCKS = {
    "resnet_v1_50": [("block1", 256, 64, 3, 2), ("block2", 512, 128, 4, 2), ("block3", 1024, 256, 6, 2)],
    "resnet_v1_101": [("block1", 256, 64, 3, 2), ("block2", 512, 128, 4, 2), ("block3", 1024, 256, 23, 2)],
    "resnet_v1_152": [("block1", 256, 64, 3, 2), ("block2", 512, 128, 8, 2), ("block3", 1024, 256, 36, 2)],
}


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


class Oracle(object):
    def __init__(self, params, cfg, bf16=True):
        """params: {name: fp32 tensor} (leaf tensors get requires_grad for trainable names);
        cfg: dict with the model hyper-parameters (see tests/helpers.py: oracle_config)."""
        self.cfg = cfg
        self.bf16 = bf16
        self.p = {}
        for k, v in params.items():
            t = v.detach().clone().float()
            self.p[k] = t
        self.arch = cfg["architecture"]

    def rb(self, x):
        return _RoundBF16.apply(x) if self.bf16 else x

    def require_grad(self, names):
        for n in names:
            self.p[n].requires_grad_(True)

    # ------------------------------------------------------------------ layers
    def conv(self, x, scope, stride=1, rate=1, padding="SAME", bn=True, relu=True, res=None, out_round=True,
             eps=1e-5, weight_name="weights", out_scale=None):
        w = self.p[scope + "/" + weight_name]                # [K,R,S,C]
        if bn:
            gamma = self.p.get(scope + "/BatchNorm/gamma")   # absent when slim.batch_norm(scale=False)
            s = (1.0 if gamma is None else gamma) / torch.sqrt(self.p[scope + "/BatchNorm/moving_variance"] + eps)
            bias = self.p[scope + "/BatchNorm/beta"] - self.p[scope + "/BatchNorm/moving_mean"] * s
            w = w * s[:, None, None, None]
        else:
            bias = self.p.get(scope + "/biases")
            if out_scale is not None:                        # net += scale * (conv + bias)
                w = w * out_scale
                bias = bias * out_scale
        w = self.rb(w)
        whwio = w.permute(1, 2, 3, 0)
        if padding == "EXPLICIT":
            y = ON.conv2d_same(x, whwio, stride, rate)
        else:
            y = ON.conv2d_tf(x, whwio, stride, padding, rate)
        if bias is not None:
            y = y + bias
        if res is not None:
            y = y + res
        if relu:
            y = torch.relu(y)
            if relu == 2:                       # ReLU6 (mobilenet_v1_arg_scope)
                y = torch.clamp(y, max=6.0)
        return self.rb(y) if out_round else y

    def dwconv(self, x, scope, stride, bn=True, act=2, eps=1e-3):
        """Depthwise stage of slim.separable_conv2d (depth_multiplier 1, SAME): weights [C,3,3]."""
        w = self.p[scope + "/depthwise_weights"]
        bias = None
        if bn:
            s = self.p[scope + "/BatchNorm/gamma"] / torch.sqrt(self.p[scope + "/BatchNorm/moving_variance"] + eps)
            bias = self.p[scope + "/BatchNorm/beta"] - self.p[scope + "/BatchNorm/moving_mean"] * s
            w = w * s[:, None, None]
        w = self.rb(w)
        C = x.shape[-1]
        xn = x.permute(0, 3, 1, 2)
        _, pt, pb = ON.same_pad(x.shape[1], 3, stride)
        _, pl, pr = ON.same_pad(x.shape[2], 3, stride)
        xn = TF.pad(xn, (pl, pr, pt, pb))
        y = TF.conv2d(xn, w[:, None], stride=stride, groups=C).permute(0, 2, 3, 1)
        if bias is not None:
            y = y + bias
        if act:
            y = torch.clamp(torch.relu(y), max=6.0)
        return self.rb(y)

    def trunk_mobilenet(self, img, scope):
        """mobilenet_v1_base up to Conv2d_11_pointwise (slim/nets/mobilenet_v1.py:142-266); the device
        stores Conv2d_0 as packed [32, 64] im2col-GEMM weights (27 real columns)."""
        defs = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1), (512, 1), (512, 1),
                (512, 1)]
        x = self.rb((img - 127.5) * (2.0 / 255.0))
        s0 = scope + "/Conv2d_0"
        w0 = self.p[s0 + "/weights"].reshape(32, -1)[:, :27].reshape(32, 3, 3, 3)
        sc = self.p[s0 + "/BatchNorm/gamma"] / torch.sqrt(self.p[s0 + "/BatchNorm/moving_variance"] + 1e-3)
        b0 = self.p[s0 + "/BatchNorm/beta"] - self.p[s0 + "/BatchNorm/moving_mean"] * sc
        w0 = self.rb(w0 * sc[:, None, None, None])
        x = ON.conv2d_tf(x, w0.permute(1, 2, 3, 0), 2, "SAME") + b0
        x = self.rb(torch.clamp(torch.relu(x), max=6.0))
        for i, (depth, stride) in enumerate(defs):
            x = self.dwconv(x, "%s/Conv2d_%d_depthwise" % (scope, i + 1), stride)
            x = self.conv(x, "%s/Conv2d_%d_pointwise" % (scope, i + 1), relu=2, eps=1e-3)
        return x

    def tail_mobilenet(self, x, scope):
        """Conv2d_12_pointwise (stride 2) + Conv2d_13_pointwise fused separable convs (mob fe:148-184)."""
        for name, stride in (("Conv2d_12_pointwise", 2), ("Conv2d_13_pointwise", 1)):
            s = scope + "/" + name
            x = self.dwconv(x, s, stride, bn=False, act=0)
            x = self.conv(x, s, relu=2, eps=1e-3, weight_name="pointwise_weights")
        return x

    def psroi(self, fmap, boxes, box_ind, D, bins, crop):
        """utils/ops.py:462-609 position_sensitive_crop_regions(global_pool=True): fmap [B,H,W,bins*D],
        boxes float32 numpy [R,4] -> [R, D]."""
        F32 = np.float32
        nby, nbx = bins
        gh, gw = crop[0] // nby, crop[1] // nbx
        boxes = np.asarray(boxes, F32)
        ymin, xmin, ymax, xmax = [boxes[:, i] for i in range(4)]
        step_y = ((ymax - ymin) / F32(nby)).astype(F32)
        step_x = ((xmax - xmin) / F32(nbx)).astype(F32)
        crops = []
        for by in range(nby):
            for bx in range(nbx):
                b = np.stack([ymin + F32(by) * step_y, xmin + F32(bx) * step_x, ymin + F32(by + 1) * step_y,
                              xmin + F32(bx + 1) * step_x], 1).astype(F32)
                split = fmap[..., (by * nbx + bx) * D:(by * nbx + bx + 1) * D]
                crops.append(ON.crop_and_resize(split, torch.from_numpy(b), torch.from_numpy(box_ind), (gh, gw)))
        ps = sum(crops) / len(crops)
        return ps.mean(dim=(1, 2))

    def rfcn_head(self, feats, scope, boxes, box_ind, groups):
        """RfcnBoxPredictor (bp:180-337): reduce_depth (+bias, ReLU) -> per-group 1x1 maps -> PS-ROI."""
        r = self.cfg["rfcn"]
        red = self.conv(feats, scope + "/reduce_depth", bn=False, relu=True)
        outs = []
        for name, D in groups:
            m = self.rb(self.conv(red, scope + "/" + name, bn=False, relu=False, out_round=False))
            outs.append(self.psroi(m, boxes, box_ind, D, r["bins"], r["crop"]))
        return outs

    # ---- Inception-ResNet-v2 (slim/nets/inception_resnet_v2.py:33-268; incres fe:112-170)
    def _ic(self, x, scope, stride=1, padding="SAME"):
        return self.conv(x, scope, stride, 1, padding, bn=True, relu=True, eps=1e-3)

    def _ir_block(self, x, scope, kind, scale, act=True):
        c = self._ic
        if kind == 35:
            b = [c(x, scope + "/Branch_0/Conv2d_1x1"),
                 c(c(x, scope + "/Branch_1/Conv2d_0a_1x1"), scope + "/Branch_1/Conv2d_0b_3x3"),
                 c(c(c(x, scope + "/Branch_2/Conv2d_0a_1x1"), scope + "/Branch_2/Conv2d_0b_3x3"),
                   scope + "/Branch_2/Conv2d_0c_3x3")]
        elif kind == 17:
            b = [c(x, scope + "/Branch_0/Conv2d_1x1"),
                 c(c(c(x, scope + "/Branch_1/Conv2d_0a_1x1"), scope + "/Branch_1/Conv2d_0b_1x7"),
                   scope + "/Branch_1/Conv2d_0c_7x1")]
        else:
            b = [c(x, scope + "/Branch_0/Conv2d_1x1"),
                 c(c(c(x, scope + "/Branch_1/Conv2d_0a_1x1"), scope + "/Branch_1/Conv2d_0b_1x3"),
                   scope + "/Branch_1/Conv2d_0c_3x1")]
        mixed = torch.cat(b, 3)
        y = self.conv(mixed, scope + "/Conv2d_1x1", bn=False, relu=False, res=x, out_round=False, out_scale=scale)
        return self.rb(torch.relu(y) if act else y)

    def _avgpool3_same(self, x):
        xn = x.permute(0, 3, 1, 2).contiguous()
        return TF.avg_pool2d(xn, 3, 1, 1, count_include_pad=False).permute(0, 2, 3, 1)

    def trunk_incres(self, img, scope):
        c = self._ic
        x = self.rb((img - 127.5) * (2.0 / 255.0))
        s0 = scope + "/Conv2d_1a_3x3"
        w0 = self.p[s0 + "/weights"].reshape(32, -1)[:, :27].reshape(32, 3, 3, 3)
        sc = 1.0 / torch.sqrt(self.p[s0 + "/BatchNorm/moving_variance"] + 1e-3)
        b0 = self.p[s0 + "/BatchNorm/beta"] - self.p[s0 + "/BatchNorm/moving_mean"] * sc
        x = self.rb(torch.relu(ON.conv2d_tf(x, self.rb(w0 * sc[:, None, None, None]).permute(1, 2, 3, 0), 2, "SAME") + b0))
        x = c(c(x, scope + "/Conv2d_2a_3x3"), scope + "/Conv2d_2b_3x3")
        x = ON.max_pool_tf(x, 3, 2, "SAME")
        x = c(c(x, scope + "/Conv2d_3b_1x1"), scope + "/Conv2d_4a_3x3")
        x = ON.max_pool_tf(x, 3, 2, "SAME")
        m = scope + "/Mixed_5b"
        x = torch.cat([c(x, m + "/Branch_0/Conv2d_1x1"),
                       c(c(x, m + "/Branch_1/Conv2d_0a_1x1"), m + "/Branch_1/Conv2d_0b_5x5"),
                       c(c(c(x, m + "/Branch_2/Conv2d_0a_1x1"), m + "/Branch_2/Conv2d_0b_3x3"),
                         m + "/Branch_2/Conv2d_0c_3x3"),
                       c(self.rb(self._avgpool3_same(x)), m + "/Branch_3/Conv2d_0b_1x1")], 3)
        for i in range(10):
            x = self._ir_block(x, "%s/Repeat/block35_%d" % (scope, i + 1), 35, 0.17)
        m = scope + "/Mixed_6a"
        x = torch.cat([c(x, m + "/Branch_0/Conv2d_1a_3x3", 2),
                       c(c(c(x, m + "/Branch_1/Conv2d_0a_1x1"), m + "/Branch_1/Conv2d_0b_3x3"),
                         m + "/Branch_1/Conv2d_1a_3x3", 2),
                       ON.max_pool_tf(x, 3, 2, "SAME")], 3)
        for i in range(20):
            x = self._ir_block(x, "%s/Repeat_1/block17_%d" % (scope, i + 1), 17, 0.10)
        return x

    def tail_incres(self, x, scope):
        c = self._ic
        m = scope + "/Mixed_7a"
        x = torch.cat([c(c(x, m + "/Branch_0/Conv2d_0a_1x1"), m + "/Branch_0/Conv2d_1a_3x3", 2, "VALID"),
                       c(c(x, m + "/Branch_1/Conv2d_0a_1x1"), m + "/Branch_1/Conv2d_1a_3x3", 2, "VALID"),
                       c(c(c(x, m + "/Branch_2/Conv2d_0a_1x1"), m + "/Branch_2/Conv2d_0b_3x3"),
                         m + "/Branch_2/Conv2d_1a_3x3", 2, "VALID"),
                       ON.max_pool_tf(x, 3, 2, "VALID")], 3)
        for i in range(9):
            x = self._ir_block(x, "%s/Repeat_2/block8_%d" % (scope, i + 1), 8, 0.20)
        x = self._ir_block(x, scope + "/Block8", 8, 1.0, act=False)
        return c(x, scope + "/Conv2d_7b_1x1")

    def bottleneck(self, x, scope, depth, stride, rate=1):
        s = scope + "/bottleneck_v1"
        cin = x.shape[-1]
        if depth == cin:
            sc = x if stride == 1 else ON.max_pool_tf(x, 1, stride)          # resnet_utils.subsample
        else:
            sc = self.conv(x, s + "/shortcut", stride, relu=False)
        r = self.conv(x, s + "/conv1")
        r = self.conv(r, s + "/conv2", stride, rate, padding="SAME" if stride == 1 else "EXPLICIT")
        return self._unit_out(r, s, sc)

    def _unit_out(self, r, s, sc):
        # conv3 (no activation) + shortcut -> relu; the device fuses add+relu in the conv3 epilogue
        w_scope = s + "/conv3"
        y = self.conv(r, w_scope, relu=False, res=sc, out_round=False)
        return self.rb(torch.relu(y))

    def trunk(self, img, scope):
        """fe:92-146 with output_stride 16."""
        means = torch.tensor(self.cfg.get("means", [123.68, 116.779, 103.939]))
        x = self.rb(img - means)
        x = self.conv(x, scope + "/conv1", 2, padding="EXPLICIT")
        x = ON.max_pool_tf(x, 3, 2, "SAME")
        current_stride, rate = 4, 1
        for bname, depth, db, n, bstride in BLOCKS[self.arch]:
            for u in range(n):
                stride = bstride if u == n - 1 else 1
                us = "%s/%s/unit_%d" % (scope, bname, u + 1)
                if current_stride == 16:
                    x = self.bottleneck(x, us, depth, 1, rate)
                    rate *= stride
                else:
                    x = self.bottleneck(x, us, depth, stride, 1)
                    current_stride *= stride
        return x

    def block4(self, x, scope):
        for u in range(3):
            x = self.bottleneck(x, "%s/block4/unit_%d" % (scope, u + 1), 2048, 1)
        return x

    def head(self, feats, scope, names):
        """MaskRCNNBoxPredictor: spatial average -> FC(s) (bp:470-500, :568-602)."""
        pooled = self.rb(feats.mean(dim=(1, 2)))
        outs = []
        for n in names:
            w = self.rb(self.p["%s/%s/weights" % (scope, n)].reshape(-1, pooled.shape[1]))
            outs.append(pooled @ w.t() + self.p["%s/%s/biases" % (scope, n)])
        return outs

    # ------------------------------------------------------------------ forward
    def training_proposals(self, rpn_box, rpn_cls, anchors, gt_abs, gt_cls_bg, key, H, W, decoded=None, scores=None):
        """`_postprocess_rpn` in training mode for one image (fmA:1055-1132): decode / score filter / clip / NMS ->
        `_unpad_proposals_and_sample_box_classifier_batch` (fmA:1134-1216) with `_sample_box_classifier_minibatch`
        (:1268-1302: detector assignment, the all-ignored guard, balanced sampling that keeps score order) -> zero pad to
        second_stage_batch_size -> normalise by the image size (:1123-1131).  Pinned against those methods executed on
        the NumPy TF shim (tests/golden/make_graph_golden.py).
        Returns (normalised boxes [P,4], scores [P], count, (nms boxes, nms scores, nms count))."""
        cfg = self.cfg
        P, M = cfg["second_stage_batch_size"], cfg["first_stage_max_proposals"]
        pb, ps, n = OP.rpn_postprocess_single(rpn_box, rpn_cls, anchors, (H, W), cfg["nms_score_threshold"],
                                              cfg["nms_iou_threshold"], M, decoded=decoded, scores=scores)
        t = OA.assign_detection(pb[:n], gt_abs, gt_cls_bg)
        cls_w = t["cls_weights"] + np.float32(t["cls_weights"].sum() == 0)      # fmA:1296
        pos = t["cls_targets"].argmax(1) > 0
        sel = OA.balanced_subsample(cls_w > 0, P, pos, cfg["second_stage_balance_fraction"], np.asarray(key)[:n])
        idx = np.nonzero(sel)[0][:P]
        boxes = np.zeros((P, 4), np.float32)
        scores = np.zeros((P,), np.float32)
        boxes[:len(idx)] = OB.to_normalized_coordinates(pb[idx], H, W)
        scores[:len(idx)] = ps[idx]
        return boxes, scores, len(idx), (pb, ps, n)

    def forward(self, images, examples, keys, H, W, proposal_inputs=None, inference=False, inference_mtl=False):
        """proposal_inputs: optional (rpn_box [B,N,4], rpn_cls [B,N,2]) numpy arrays used INSTEAD of the
        oracle's own RPN outputs for the (non-differentiable) proposal selection, so that index-level
        parity can be checked on identical inputs.  A 4-tuple (rpn_box, rpn_cls, decoded [B,N,4], scores [B,N])
        additionally fixes the decoded (clipped) boxes and objectness scores the sort / NMS / sampler start from:
        `exp` differs by an ulp between libm and the device, which reorders near-tied scores (the decode itself is
        compared separately, to 2e-6).
        inference=True: the is_training=False graph (fmA:586-590 anchors clipped instead of pruned; fmA:1111-1131 no
        minibatch sampling, max_num_proposals = first_stage_max_proposals, fmA:475-477); returns after the
        second-stage box classifier (what `postprocess` consumes); `examples` / `keys` are not used.
        inference_mtl=True: continue like the reference's evaluator (evaluator.py:145-152) through the closeness head and
        `predict_with_mtl_results` on the inference proposals, so that `postprocess` scores the refined logits
        (fmA:1040-1043); the window / edge-mask heads run only when `examples` is given (they need the ground-truth
        windows)."""
        cfg, p = self.cfg, self.p
        B = images.shape[0]
        K = cfg["num_classes"]
        K1 = K + 1
        A = len(cfg["scales"]) * len(cfg["aspect_ratios"])
        mobile = self.arch == "MobilenetV1"
        fs = "FirstStageFeatureExtractor/" + self.arch
        incres = self.arch == "InceptionResnetV2"
        feat = self.trunk_mobilenet(images, fs) if mobile else (self.trunk_incres(images, fs) if incres
                                                                  else self.trunk(images, fs))
        tail = self.tail_mobilenet if mobile else (self.tail_incres if incres else self.block4)
        _, Hf, Wf, _ = feat.shape
        rpn_feat = self.conv(feat, "FirstStageBoxPredictor/Conv", bn=False, relu=True)
        box = self.conv(rpn_feat, "FirstStageBoxPredictor/BoxEncodingPredictor", bn=False, relu=False, out_round=False)
        cls = self.conv(rpn_feat, "FirstStageBoxPredictor/ClassPredictor", bn=False, relu=False, out_round=False)
        box = box.reshape(B, Hf * Wf * A, 4)
        cls = cls.reshape(B, Hf * Wf * A, 2)
        anchors_all = OB.grid_anchors(Hf, Wf, cfg["scales"], cfg["aspect_ratios"], (256, 256), (16, 16), (0, 0))
        if inference:
            anchors, keep = OB.clip_to_window(anchors_all, (0, 0, H, W))             # fmA:586-590
            assert len(keep) == len(anchors_all), "a grid anchor lies completely outside the image"
        else:
            anchors, keep = OB.prune_outside_window(anchors_all, (0, 0, H, W))        # fmA:930-976
        keep_t = torch.from_numpy(np.asarray(keep)).long()
        rpn_box, rpn_cls = box[:, keep_t], cls[:, keep_t]
        P, M = cfg["second_stage_batch_size"], cfg["first_stage_max_proposals"]
        if inference:
            P = M
        # ---- _postprocess_rpn + minibatch sampling (no gradient: fmA:1109-1110)
        gts = []
        prop_norm = np.zeros((B, P, 4), np.float32)
        prop_abs = np.zeros((B, P, 4), np.float32)
        nprop = np.zeros((B,), np.int64)
        nms_out = []
        for b in range(B):
            if inference:
                src_box = proposal_inputs[0][b] if proposal_inputs is not None else rpn_box[b].detach().numpy()
                src_cls = proposal_inputs[1][b] if proposal_inputs is not None else rpn_cls[b].detach().numpy()
                pb, ps, n = OP.rpn_postprocess_single(src_box, src_cls, anchors, (H, W), cfg["nms_score_threshold"],
                                                      cfg["nms_iou_threshold"], M)
                nms_out.append((pb, ps, n))
                nprop[b] = n
                nb = OB.to_normalized_coordinates(pb[:n], H, W)
                prop_norm[b, :n] = nb
                prop_abs[b, :n] = OB.to_absolute_coordinates(nb, H, W)
                continue
            ex = examples[b]
            gt_abs = OB.to_absolute_coordinates(np.asarray(ex["groundtruth_boxes"], np.float32).reshape(-1, 4), H, W)
            oh = np.asarray(ex["groundtruth_classes"], np.float32)
            gt_cls_bg = np.concatenate([np.zeros((len(oh), 1), np.float32), oh], 1)
            gts.append((gt_abs, gt_cls_bg, np.asarray(ex["groundtruth_closeness"], np.float32)))
            src_box = proposal_inputs[0][b] if proposal_inputs is not None else rpn_box[b].detach().numpy()
            src_cls = proposal_inputs[1][b] if proposal_inputs is not None else rpn_cls[b].detach().numpy()
            fixed = proposal_inputs is not None and len(proposal_inputs) == 4
            nb, _, cnt, nms = self.training_proposals(src_box, src_cls, anchors, gt_abs, gt_cls_bg, keys[1][b], H, W,
                                                      decoded=proposal_inputs[2][b] if fixed else None,
                                                      scores=proposal_inputs[3][b] if fixed else None)
            nms_out.append(nms)
            nprop[b] = cnt
            prop_norm[b] = nb
            prop_abs[b, :cnt] = OB.to_absolute_coordinates(nb[:cnt], H, W)        # fmA:682-683
        c = cfg["initial_crop_size"]
        mk = cfg["maxpool_kernel_size"]

        def crops_of(boxes_flat, box_ind):
            cr = self.rb(ON.crop_and_resize(feat, torch.from_numpy(boxes_flat), torch.from_numpy(box_ind), (c, c)))
            return ON.max_pool_tf(cr, mk, mk, "VALID") if mk > 1 else cr

        bi = np.repeat(np.arange(B), P).astype(np.int64)
        maps = None if cfg.get("rfcn") else crops_of(prop_norm.reshape(-1, 4), bi)
        out = dict(feat=feat, rpn_box=rpn_box, rpn_cls=rpn_cls, anchors=anchors, keep=keep, prop_norm=prop_norm,
                   prop_abs=prop_abs, nprop=nprop, gts=gts, nms=nms_out)
        rfcn = cfg.get("rfcn")
        mtl = cfg["mtl"]
        stop = mtl.get("stop_gradient_for_aux_tasks", False)
        flat_props = prop_norm.reshape(-1, 4)
        if rfcn:
            bx, cl = self.rfcn_head(self.block4(feat, "SecondStageFeatureExtractor/" + self.arch),
                                    "SecondStageBoxPredictor", flat_props, bi,
                                    [("refined_locations", 4 * K), ("class_predictions", K1)])
        else:
            bx, cl = self.head(tail(maps, "SecondStageFeatureExtractor/" + self.arch), "SecondStageBoxPredictor",
                               ["BoxEncodingPredictor", "ClassPredictor"])
        out["refined_box_encodings"] = bx.reshape(B * P, K, 4)
        out["class_predictions_with_background"] = cl
        if inference and not inference_mtl:
            return out
        if mtl.get("closeness"):
            if rfcn:
                f2 = feat.detach() if stop else feat
                out["closeness_predictions"] = self.rfcn_head(self.block4(f2, "ClosenessBoxPredictor/" + self.arch),
                                                              "ClosenessBoxPredictor", flat_props, bi,
                                                              [("class_predictions", K1)])[0]
            else:
                m2 = maps.detach() if stop else maps
                out["closeness_predictions"] = self.head(tail(m2, "ClosenessBoxPredictor/" + self.arch),
                                                         "ClosenessBoxPredictor", ["ClassPredictor"])[0]
        win_feat = None
        if rfcn and mtl.get("window") and mtl.get("refine") and examples is None:
            win_feat = self.block4(feat, "WindowBoxPredictor/" + self.arch)
        if mtl.get("window") and examples is not None:
            wb = np.stack([np.asarray(e["window_boxes"], np.float32) for e in examples])
            nw = wb.shape[1]
            wbi = np.repeat(np.arange(B), nw).astype(np.int64)
            if rfcn:
                f3 = feat.detach() if stop else feat
                win_feat = self.block4(f3, "WindowBoxPredictor/" + self.arch)
                out["window_class_predictions"] = self.rfcn_head(win_feat, "WindowBoxPredictor", wb.reshape(-1, 4), wbi,
                                                                 [("class_predictions", K1)])[0]
            else:
                wm = crops_of(wb.reshape(-1, 4), wbi)
                if stop:
                    wm = wm.detach()
                out["window_class_predictions"] = self.head(tail(wm, "WindowBoxPredictor/" + self.arch),
                                                            "WindowBoxPredictor", ["ClassPredictor"])[0]
        if mtl.get("edgemask"):
            w = p["EdgeMaskPredictor/BoxEncodingPredictor/weights"].reshape(2, -1)
            out["edgemask_predictions"] = torch.tanh(feat @ w.t() + p["EdgeMaskPredictor/BoxEncodingPredictor/biases"])
        if mtl.get("refine"):
            ew = close = None
            if mtl.get("window"):
                exp = self.expanded_windows(prop_norm)                      # [5,B,P,4]
                ebi = np.broadcast_to(np.arange(B)[None, :, None], (5, B, P)).reshape(-1).astype(np.int64)
                with torch.no_grad():
                    if rfcn:
                        ew = self.rfcn_head(win_feat, "WindowBoxPredictor", exp.reshape(-1, 4), ebi,
                                            [("class_predictions", K1)])[0]
                    else:
                        em = crops_of(exp.reshape(-1, 4), ebi)
                        ew = self.head(tail(em, "WindowBoxPredictor/" + self.arch), "WindowBoxPredictor",
                                       ["ClassPredictor"])[0]
            if mtl.get("closeness"):
                close = out["closeness_predictions"]
            ref, net = self.refine_logits(cl, ew, close)
            out["mtl_refined_class_predictions_with_background"] = ref
            out["refine_in"] = net
        return out

    @staticmethod
    def expanded_windows(prop_norm):
        """fmA:783-803: per proposal five boxes growing linearly from the proposal (i = 0) to the whole image (i = 4);
        returned [5, B, P, 4], the order in which `predict_with_mtl_results` flattens them into one window list."""
        F32 = np.float32
        ymin, xmin, ymax, xmax = [np.asarray(prop_norm, F32)[..., i] for i in range(4)]
        exp = []
        for e in range(5):
            exp.append(np.stack([ymin - (ymin / F32(4)) * F32(e), xmin - (xmin / F32(4)) * F32(e),
                                 ymax + ((F32(1) - ymax) / F32(4)) * F32(e),
                                 xmax + ((F32(1) - xmax) / F32(4)) * F32(e)], -1).astype(F32))
        return np.stack(exp)

    def refine_logits(self, cl, ew, close):
        """fmA:805-846: refiner input = [class logits | the five window logit rows of the SAME proposal, window-major
        ([5, B*P, K+1] -> [B*P, 5(K+1)]) | closeness logits averaged over all proposals and tiled], all without
        gradient; one FC layer (`refine_num_fc_layers: 0` in every shipped config) plus the residue.
        `ew`: [5*B*P, K+1] window logits in `expanded_windows` order, or None; `close`: [B*P, K+1] or None."""
        mtl, p = self.cfg["mtl"], self.p
        n, K1 = cl.shape
        src = [cl]
        if ew is not None:
            src.append(ew.reshape(5, n, K1).permute(1, 0, 2).reshape(n, 5 * K1))
        if close is not None:
            src.append(close.detach().mean(0, keepdim=True).expand(n, K1))
        net = torch.cat([s.detach() for s in src], 1)
        ref = net @ p["MTLClassRefiner/fc1/weights"].t() + p["MTLClassRefiner/fc1/biases"]
        if mtl.get("refine_residue"):
            ref = ref + cl
        return ref, net

    # ------------------------------------------------------------------ losses
    def loss_first_stage(self, out, keys):
        """fmA `_loss_rpn` :1591-1668 (T1 sigma 3, T15 per-image normaliser = number of sampled anchors, then batch
        mean).  Pinned against that method executed on the NumPy TF shim (tests/golden/make_graph_golden.py)."""
        cfg = self.cfg
        anchors = out["anchors"]
        B = len(out["gts"])
        losses = {}
        loc_l = obj_l = 0.0
        for b in range(B):
            t = OA.assign_proposal(anchors, out["gts"][b][0])
            s = OA.balanced_subsample(t["cls_weights"] > 0, cfg["first_stage_minibatch_size"],
                                      t["cls_targets"][:, 0] > 0, cfg["first_stage_positive_balance_fraction"],
                                      keys[b])
            sf = torch.from_numpy(s.astype(np.float32))
            norm = sf.sum()
            loc = ON.smooth_l1(out["rpn_box"][b], torch.from_numpy(t["reg_targets"]),
                               sf * torch.from_numpy(t["reg_weights"]), 3.0)
            onehot = TF.one_hot(torch.from_numpy(t["cls_targets"][:, 0]).long(), 2).float()
            obj = ON.softmax_ce(out["rpn_cls"][b], onehot, sf)
            loc_l = loc_l + loc.sum() / norm
            obj_l = obj_l + obj.sum() / norm
        losses["first_stage_localization_loss"] = cfg["first_stage_localization_loss_weight"] * loc_l / B
        losses["first_stage_objectness_loss"] = cfg["first_stage_objectness_loss_weight"] * obj_l / B
        return losses

    def loss(self, out, examples, keys, H, W):
        losses = self.loss_first_stage(out, keys[0])
        losses.update(self.loss_second_stage(out, examples))
        return losses

    def loss_second_stage(self, out, examples):
        """fmA `_loss_box_classifier` :1670-1793 (incl. closeness), `_loss_refined_classifier` :1795-1837,
        `_loss_window_class` :1839-1858, `_loss_edgemask` :1860-1881.  Pinned against those methods executed on the
        NumPy TF shim (tests/golden/make_stage2_loss_golden.py)."""
        cfg, mtl = self.cfg, self.cfg["mtl"]
        B = len(examples)
        K1 = cfg["num_classes"] + 1
        P = cfg["second_stage_batch_size"]
        losses = {}
        cls_t, reg_t, reg_w, cls_w, close_t = [], [], [], [], []
        for b in range(B):
            gt_abs, gt_cls_bg, gt_close = out["gts"][b]
            t = OA.assign_detection(out["prop_abs"][b], gt_abs, gt_cls_bg, gt_close)
            cls_t.append(t["cls_targets"]); reg_t.append(t["reg_targets"]); reg_w.append(t["reg_weights"])
            cls_w.append(t["cls_weights"]); close_t.append(t["closeness_targets"])
        cls_t = torch.from_numpy(np.stack(cls_t)); reg_t = torch.from_numpy(np.stack(reg_t))
        reg_w = torch.from_numpy(np.stack(reg_w)); cls_w = torch.from_numpy(np.stack(cls_w))
        close_t = torch.from_numpy(np.stack(close_t))
        nprop = torch.from_numpy(out["nprop"])
        pad = (torch.arange(P)[None, :] < nprop[:, None]).float()
        norm = (nprop.clamp(min=1).float() * B)[:, None].expand(B, P)
        enc_bg = torch.cat([torch.zeros(B * P, 1, 4), out["refined_box_encodings"]], 1)
        flat_t = cls_t.reshape(B * P, K1)
        sel = enc_bg[flat_t > 0].reshape(B, P, 4)                           # boolean_mask (one-hot, T12)
        loc = ON.smooth_l1(sel, reg_t, reg_w, 1.0) / norm
        cl = ON.softmax_ce(out["class_predictions_with_background"].reshape(B, P, K1), cls_t, cls_w) / norm
        losses["second_stage_localization_loss"] = cfg["second_stage_localization_loss_weight"] * (loc * pad).sum()
        losses["second_stage_classification_loss"] = cfg["second_stage_classification_loss_weight"] * (cl * pad).sum()
        if mtl.get("closeness"):
            norm_reg = reg_w.sum(1, keepdim=True).clamp(min=1.0)
            cp = out["closeness_predictions"].reshape(B, P, K1)[:, :, 1:]
            ct = close_t[:, :, 1:]
            closs = ON.softmax_ce(cp, ct, reg_w) / norm_reg * ct.sum(2)
            losses["closeness_classification_loss"] = mtl["closeness_loss_weight"] * closs.sum()
        if mtl.get("window"):
            wc = torch.from_numpy(np.stack([np.asarray(e["window_classes"], np.float32) for e in examples]))
            wl = ON.softmax_ce(out["window_class_predictions"], wc.reshape(-1, K1))
            losses["window_class_loss"] = mtl["window_class_loss_weight"] * wl.mean()
        if mtl.get("edgemask"):
            em = torch.from_numpy(np.stack([np.asarray(e["groundtruth_edgemask"], np.float32) for e in examples]))
            fg, wt = em[:, 0], em[:, 1]
            tg = torch.stack([1.0 - fg, fg], -1)
            pr = ON.resize_bilinear(out["edgemask_predictions"], (em.shape[2], em.shape[3]))
            losses["edgemask_loss"] = mtl["edgemask_loss_weight"] * (ON.softmax_ce(pr, tg) * wt).mean()
        if mtl.get("refine"):
            rl = ON.softmax_ce(out["mtl_refined_class_predictions_with_background"].reshape(B, P, K1), cls_t,
                               cls_w) / norm
            losses["refined_classification_loss"] = mtl["refined_classification_loss_weight"] * (rl * pad).sum()
        out["targets"] = dict(cls_t=cls_t, reg_t=reg_t, reg_w=reg_w, cls_w=cls_w, close_t=close_t)
        return losses

    def regularization_loss(self, l2_table):
        """sum over variables of weight * 0.5 * sum(w^2) (slim.l2_regularizer; model_deploy.py:296)."""
        tot = 0.0
        for name, l2 in l2_table.items():
            if l2:
                tot = tot + l2 * 0.5 * (self.p[name] ** 2).sum()
        return tot
