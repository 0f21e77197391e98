"""Faster R-CNN meta-architecture with the three auxiliary heads (multi-object soft label
"window", closeness, foreground mask "edgemask") and the MTL class refiner, on B200 kernels.

Mirrors /root/reference/object_detection/meta_architectures/faster_rcnn_meta_arch.py:208-2013:
same constructor arguments, scope names (:431-461), prediction_dict keys (:593-601, :693-699,
:717, :753, :761, :845), loss-dict keys, and the DetectionModel call sequence used by
trainer._create_losses (trainer.py:157-214):

    provide_groundtruth / provide_window / provide_edgemask -> preprocess -> predict ->
    predict_with_window -> predict_edgemask -> predict_with_mtl_results -> loss

The reference builds a TF graph and lets tf.gradients differentiate it; here every stage is a
fixed sequence of kernel launches over persistent device buffers and `backward()` replays the
chain rule explicitly (there is no autograd), so a whole training step contains no host
synchronisation and can be captured into one CUDA graph.  Shapes are static: proposals are
padded to `max_num_proposals` with device-side counts exactly as the reference pads them.

Deliberate, documented deviations: tf.random_shuffle in the samplers is replaced by
explicit per-anchor keys (`provide_sampler_keys`); for per-replica batch > 1 the refine windows
crop their own image (the reference crops image 0, trap T9); the dead stage-1 block4 of the
reference is not computed (its variables still exist and receive their L2 gradient, trap T4).
"""
import numpy as np
import torch

from .. import ops
from .. import ops_conv
from ..core import model
from ..core.standard_fields import (BoxListFields as fields, BOX_ENCODINGS, CLASS_PREDICTIONS,
                                    CLASS_PREDICTIONS_WITH_BACKGROUND, MASK_PREDICTIONS)
from ..nets.layers import Concurrency, Conv2d, max_pool, max_pool_bwd, max_pool_out_hw
from ..runtime import ParamStore, Workspace

LOSS_KEYS = ["first_stage_localization_loss", "first_stage_objectness_loss", "second_stage_localization_loss",
             "second_stage_classification_loss", "closeness_classification_loss", "window_class_loss",
             "edgemask_loss", "refined_classification_loss"]


class _Lanes(object):
    """Named side streams + events: the independent second-stage chains (closeness tail, window tail,
    forward-only refine windows, edge mask) run concurrently with the main chain; cross-chain
    dependencies are events, which CUDA-graph capture turns into graph edges."""

    def __init__(self):
        self.streams, self.events = {}, {}

    def mark(self, name):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.events[name] = ev

    def run(self, lane, after=()):
        import contextlib
        if not Concurrency.enabled:
            return contextlib.nullcontext()
        if lane not in self.streams:
            self.streams[lane] = torch.cuda.Stream()
        s = self.streams[lane]
        for n in after:
            s.wait_event(self.events[n])
        return torch.cuda.stream(s)

    def wait(self, *names):
        if not Concurrency.enabled:
            return
        cur = torch.cuda.current_stream()
        for n in names:
            if n in self.events:
                cur.wait_event(self.events[n])

    def reset(self):
        self.events = {}


class PredictionDict(dict):
    """dict whose values may be thunks: API tensors (reshaped views of the fused kernel outputs)
    are materialised only when somebody reads them, keeping the training hot path copy-free."""

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        if callable(v) and not isinstance(v, torch.Tensor):
            v = v()
            dict.__setitem__(self, k, v)
        return v

    def get(self, k, default=None):
        return self[k] if k in self else default


class FasterRCNNMetaArch(model.DetectionModel):
    def __init__(self, is_training, num_classes, image_resizer_fn, feature_extractor, first_stage_only,
                 first_stage_anchor_generator, first_stage_clip_window, first_stage_atrous_rate,
                 first_stage_box_predictor_trainable, first_stage_box_predictor_arg_scope,
                 first_stage_box_predictor_kernel_size, first_stage_box_predictor_depth,
                 first_stage_minibatch_size, first_stage_positive_balance_fraction,
                 first_stage_nms_score_threshold, first_stage_nms_iou_threshold, first_stage_max_proposals,
                 first_stage_localization_loss_weight, first_stage_objectness_loss_weight, initial_crop_size,
                 maxpool_kernel_size, maxpool_stride, second_stage_mask_rcnn_box_predictor,
                 second_stage_batch_size, second_stage_balance_fraction, second_stage_non_max_suppression_fn,
                 second_stage_score_conversion_fn, second_stage_localization_loss_weight,
                 second_stage_classification_loss_weight, hard_example_miner, mtl_refiner_arg_scope, mtl=None,
                 window_box_predictor=None, closeness_box_predictor=None, edgemask_predictor=None,
                 parallel_iterations=16, device="cuda", seed=0):
        super(FasterRCNNMetaArch, self).__init__(num_classes=num_classes)
        if second_stage_batch_size > first_stage_max_proposals:
            raise ValueError("second_stage_batch_size should be no greater than first_stage_max_proposals.")
        if maxpool_kernel_size != maxpool_stride:
            raise ValueError("B200 path: maxpool_kernel_size must equal maxpool_stride (all shipped configs)")
        if hard_example_miner is not None:
            raise ValueError("hard_example_miner is not supported on the B200 training path")
        if first_stage_atrous_rate != 1:
            raise ValueError("first_stage_atrous_rate != 1 is not supported on the B200 path")
        if first_stage_only:
            raise ValueError("first_stage_only is not supported on the B200 path")
        if mtl is not None and mtl.shared_feature != "proposal_feature_maps":
            raise ValueError("mtl.shared_feature must be 'proposal_feature_maps' (the only value used)")
        if mtl is not None and mtl.refine and mtl.refine_num_fc_layers != 0:
            raise ValueError("refine_num_fc_layers > 0 is not supported (all shipped configs use 0)")
        self._is_training = is_training
        self._image_resizer_fn = image_resizer_fn
        self._feature_extractor = feature_extractor
        self._first_stage_only = first_stage_only
        self._first_stage_anchor_generator = first_stage_anchor_generator
        self._first_stage_clip_window = first_stage_clip_window
        self._first_stage_box_predictor_kernel_size = first_stage_box_predictor_kernel_size
        self._first_stage_box_predictor_depth = first_stage_box_predictor_depth
        self._first_stage_minibatch_size = first_stage_minibatch_size
        self._first_stage_positive_balance_fraction = first_stage_positive_balance_fraction
        self._first_stage_nms_score_threshold = first_stage_nms_score_threshold
        self._first_stage_nms_iou_threshold = first_stage_nms_iou_threshold
        self._first_stage_max_proposals = first_stage_max_proposals
        self._first_stage_loc_loss_weight = first_stage_localization_loss_weight
        self._first_stage_obj_loss_weight = first_stage_objectness_loss_weight
        self._first_stage_sigma = 3.0                                  # fmA:391-392 (trap T1)
        self._initial_crop_size = initial_crop_size
        self._maxpool_kernel_size = maxpool_kernel_size
        self._maxpool_stride = maxpool_stride
        self._mask_rcnn_box_predictor = second_stage_mask_rcnn_box_predictor
        self._second_stage_batch_size = second_stage_batch_size
        self._second_stage_balance_fraction = second_stage_balance_fraction
        self._second_stage_nms_fn = second_stage_non_max_suppression_fn
        self._second_stage_score_conversion_fn = second_stage_score_conversion_fn
        self._second_stage_loc_loss_weight = second_stage_localization_loss_weight
        self._second_stage_cls_loss_weight = second_stage_classification_loss_weight
        self._hard_example_miner = hard_example_miner
        self._mtl = mtl
        self._window_box_predictor = window_box_predictor
        self._closeness_box_predictor = closeness_box_predictor
        self._edgemask_predictor = edgemask_predictor
        self._mtl_refiner_arg_scope = mtl_refiner_arg_scope
        self._parallel_iterations = parallel_iterations
        self.device = device
        self._ws = Workspace(device)
        self._store = ParamStore()
        self._anchor_cache = {}
        self._trunk_wgrads = {}         # workspace id -> ops_conv.WgradCollector (grouped trunk weight gradients)
        self._wgrad_chunk_ctas = int(__import__("os").environ.get("MTL_WGRAD_CHUNK_CTAS", "64"))
        self._head_wgrads = {}          # workspace id -> collector of the second-stage weight-gradient GEMMs
        self.group_head_wgrads = False  # set by the trainer's deferred schedule (plain backward(): immediate launches)
        self.defer_head_wgrads = False
        self.supports_deferred_heads = True
        self._sampler_keys = None
        self._lanes = _Lanes()
        for pred in (second_stage_mask_rcnn_box_predictor, window_box_predictor, closeness_box_predictor):
            if pred is not None and hasattr(pred, "_feature_mask_hi"):
                pred._feature_mask_hi = feature_extractor.feature_mask_hi
        self._create_variables(first_stage_box_predictor_arg_scope, first_stage_box_predictor_trainable)
        if device is not None:          # device=None: variable table only (host-side inspection)
            self._store.finalize(device, seed)

    # ------------------------------------------------------------------ scopes (fmA:431-461)
    first_stage_feature_extractor_scope = "FirstStageFeatureExtractor"
    second_stage_feature_extractor_scope = "SecondStageFeatureExtractor"
    first_stage_box_predictor_scope = "FirstStageBoxPredictor"
    second_stage_box_predictor_scope = "SecondStageBoxPredictor"
    window_box_predictor_scope = "WindowBoxPredictor"
    edgemask_predictor_scope = "EdgeMaskPredictor"
    closeness_box_predictor_scope = "ClosenessBoxPredictor"
    mtl_refiner_scope = "MTLClassRefiner"

    @property
    def max_num_proposals(self):
        if self._is_training and not self._hard_example_miner:
            return self._second_stage_batch_size
        return self._first_stage_max_proposals

    def gradient_buckets(self):
        """(second-stage bucket, trunk+RPN bucket) as views of the flat gradient arena; the dead
        stage-1 block4 copy (no task gradient) is excluded from the exchange."""
        st = self._store
        first = next(p.offset for p in st.params if p.name.startswith(self.second_stage_feature_extractor_scope))
        dead = next((p.offset for p in st.params if "/_dead/" in p.name), st.total)
        return st.g[first:dead], st.g[:first]

    def head_tensor_range(self):
        """[t0, t1): indices (arena order) of the second-stage / aux-head tensors, i.e. the tensors whose
        gradients are final once backward(part="heads") has run."""
        ps = self._store.params
        t0 = next(i for i, p in enumerate(ps) if p.name.startswith(self.second_stage_feature_extractor_scope))
        t1 = next((i for i, p in enumerate(ps) if "/_dead/" in p.name), len(ps))
        return t0, t1

    @property
    def param_store(self):
        return self._store

    @property
    def workspace(self):
        return self._ws

    # ------------------------------------------------------------------ variables
    def _create_variables(self, rpn_hp, rpn_trainable):
        from ..core.box_predictor import ConvolutionalBoxPredictor
        st, fe, mtl = self._store, self._feature_extractor, self._mtl
        fe.create_proposal_variables(st, self.first_stage_feature_extractor_scope)
        k = self._first_stage_box_predictor_kernel_size
        self._rpn_conv = Conv2d(st, self.first_stage_box_predictor_scope + "/Conv", fe.feature_depth,
                                self._first_stage_box_predictor_depth, k, 1, 1, "SAME", bn=False, bias=True,
                                relu=(rpn_hp.activation != "NONE"), l2=rpn_hp.l2_weight,
                                trainable=self._is_training and rpn_trainable, init=rpn_hp.init)
        if rpn_hp.activation == "RELU_6":
            raise ValueError("RELU_6 in the RPN conv is not supported on the B200 path")
        self._first_stage_box_predictor = ConvolutionalBoxPredictor(
            self._is_training, num_classes=1, conv_hyperparams=rpn_hp, kernel_size=1, box_code_size=4)
        A = self._first_stage_anchor_generator.num_anchors_per_location()[0]
        self._first_stage_box_predictor.create_variables(st, self.first_stage_box_predictor_scope,
                                                         self._first_stage_box_predictor_depth, A)
        fe.create_box_classifier_variables(st, self.second_stage_feature_extractor_scope)
        self._mask_rcnn_box_predictor.create_variables(st, self.second_stage_box_predictor_scope,
                                                       fe.classifier_depth)
        if mtl is not None and mtl.closeness:
            fe.create_box_classifier_variables(st, self.closeness_box_predictor_scope)
            self._closeness_box_predictor.create_variables(st, self.closeness_box_predictor_scope,
                                                           fe.classifier_depth, class_only=True)
        if mtl is not None and (mtl.window or mtl.refine):
            fe.create_box_classifier_variables(st, self.window_box_predictor_scope)
            self._window_box_predictor.create_variables(st, self.window_box_predictor_scope, fe.classifier_depth,
                                                        class_only=True)
        if mtl is not None and mtl.edgemask:
            self._edgemask_predictor.create_variables(st, self.edgemask_predictor_scope, fe.feature_depth)
        if mtl is not None and mtl.refine:
            K1 = self.num_classes + 1
            nf = K1 + (5 * K1 if mtl.window else 0) + (K1 if mtl.closeness else 0)
            hp = self._mtl_refiner_arg_scope
            self._refine_w = st.add(self.mtl_refiner_scope + "/fc1/weights", (K1, nf), l2=hp.l2_weight,
                                    trainable=self._is_training, init=hp.init)
            self._refine_b = st.add(self.mtl_refiner_scope + "/fc1/biases", (K1,), trainable=self._is_training)
            self._refine_nf = nf
        fe.create_dead_variables(st, self.first_stage_feature_extractor_scope)

    # ------------------------------------------------------------------ inputs
    def preprocess(self, inputs):
        """fmA:479-505: resize every image with image_resizer_fn then feature_extractor.preprocess.
        inputs: float32 [B,H,W,3] in [0,255] (device)."""
        if inputs.dtype != torch.float32:
            raise ValueError("`preprocess` expects a tf.float32 tensor")
        resized = self._image_resizer_fn(inputs) if self._image_resizer_fn is not None else inputs
        return self._feature_extractor.preprocess(resized)

    def provide_sampler_keys(self, first_stage_keys, second_stage_keys):
        """Explicit shuffle keys replacing tf.random_shuffle (core/minibatch_sampler.py:80-89):
        first_stage_keys [B, num_kept_anchors], second_stage_keys [B, first_stage_max_proposals],
        float32 on the device.  Candidates with the smallest keys are sampled."""
        self._sampler_keys = (first_stage_keys, second_stage_keys)

    def _format_groundtruth_data(self, image_shape):
        """fmA:1218-1266: boxes normalised -> absolute pixels, one-hot classes -> class index with
        background 0.  Packs the per-image lists into padded device tensors (once per provide_*)."""
        if not getattr(self, "_groundtruth_dirty", True) and getattr(self, "_gt_shape", None) == tuple(image_shape):
            return self._gt
        H, W = float(image_shape[1]), float(image_shape[2])
        boxes_list = self.groundtruth_lists(fields.boxes)
        classes_list = self.groundtruth_lists(fields.classes)
        B = len(boxes_list)
        K1 = self.num_classes + 1
        gmax = max(1, max(int(b.shape[0]) for b in boxes_list))
        gmax = (gmax + 7) // 8 * 8
        gt = np.zeros((B, gmax, 4), np.float32)
        ng = np.zeros((B,), np.int32)
        gc = np.zeros((B, gmax), np.int32)
        gclose = np.zeros((B, gmax, K1), np.float32)
        closeness = self._groundtruth_lists.get(fields.closeness)
        for b in range(B):
            bx = _np(boxes_list[b]).astype(np.float32).reshape(-1, 4)
            g = bx.shape[0]
            if g and bx.max() > 1.01:
                raise ValueError("maximum box coordinate value is larger than 1.01")      # blo:797-803
            ng[b] = g
            # to_absolute_coordinates -> scale(y_scale=H, x_scale=W) in float32 (blo:73-99)
            gt[b, :g] = bx * np.array([H, W, H, W], np.float32)
            if g:                                # (an image without boxes has no classes to check)
                oh = _np(classes_list[b]).astype(np.float32).reshape(g, -1)
                if not np.all((oh.sum(1) == 1) & (oh.max(1) == 1)):
                    raise ValueError("groundtruth classes must be one-hot on the B200 path (trap T12)")
                gc[b, :g] = oh.argmax(1) + 1
            if closeness is not None and closeness[b] is not None and g:
                gclose[b, :g] = _np(closeness[b]).astype(np.float32).reshape(g, K1)
        d = self.device
        out = dict(gt=torch.from_numpy(gt).to(d), num_gt=torch.from_numpy(ng).to(d),
                   gt_cls=torch.from_numpy(gc).to(d), gt_close=torch.from_numpy(gclose).to(d), gmax=gmax, B=B)
        if fields.boxes in self._window_lists:
            wb = np.stack([_np(w).astype(np.float32).reshape(-1, 4) for w in self._window_lists[fields.boxes]])
            wc = np.stack([_np(w).astype(np.float32).reshape(-1, K1) for w in self._window_lists[fields.classes]])
            out["win_boxes"] = torch.from_numpy(wb).to(d)
            out["win_cls"] = torch.from_numpy(wc).to(d)
        if fields.edgemask in self._edgemask_lists:
            em = np.stack([_np(e).astype(np.float32) for e in self._edgemask_lists[fields.edgemask]])
            out["edgemask"] = torch.from_numpy(em).to(d)
        self._gt, self._gt_shape, self._groundtruth_dirty = out, tuple(image_shape), False
        return out

    def _anchors(self, Hf, Wf, H, W):
        """Anchor grid + training-time pruning (fmA:930-976), cached per shape (one host read)."""
        key = (Hf, Wf, H, W, self._is_training and not self._first_stage_clip_window)
        if key not in self._anchor_cache:
            allb = self._first_stage_anchor_generator.generate([(Hf, Wf)], device=self.device)
            n = allb.shape[0]
            if key[-1]:
                keep = torch.empty(n, dtype=torch.int32, device=self.device)
                kept = torch.empty(n, 4, dtype=torch.float32, device=self.device)
                num = torch.zeros(1, dtype=torch.int32, device=self.device)
                ops.call("mtl_prune_outside_window", allb, n, 0.0, 0.0, float(H), float(W), keep, kept, num)
                nk = int(num.item())
                self._anchor_cache[key] = (kept[:nk].contiguous(), keep[:nk].contiguous(), nk)
            else:
                # inference (or first_stage_clip_window): clip, keep every anchor (fmA:586-590)
                clipped = torch.empty_like(allb)
                ops.call("mtl_clip_boxes", allb, n, 0.0, 0.0, float(H), float(W), clipped)
                self._anchor_cache[key] = (clipped, None, n)
        return self._anchor_cache[key]

    def num_kept_anchors(self, image_shape):
        """Anchors that survive the window pruning for INPUT images of `image_shape` [B,H,W,3] (the image resizer of
        `preprocess` is applied to the size first): the length of the first-stage sampler keys."""
        H, W = int(image_shape[1]), int(image_shape[2])
        static_size = getattr(self._image_resizer_fn, "static_size", None)
        if static_size is not None:
            H, W = (int(v) for v in static_size(H, W))
        Hf, Wf = self._feature_extractor.feature_map_shape(H, W)
        return self._anchors(Hf, Wf, H, W)[2]

    # ------------------------------------------------------------------ forward
    def frozen_prefix(self, preprocessed_inputs, tag="s1"):
        """Activations of the frozen leading layers of the stage-1 trunk (image-only dependency), or None when the
        feature extractor has no such split; pass the result to predict(prefix=)."""
        fe = self._feature_extractor
        if not hasattr(fe, "extract_frozen_prefix"):
            return None
        return fe.extract_frozen_prefix(preprocessed_inputs, self.first_stage_feature_extractor_scope, self._ws, tag)

    def predict(self, preprocessed_inputs, prefix=None):
        """fmA:507-609.  `prefix`: output of frozen_prefix() for these inputs (computed ahead of time)."""
        return self.predict_second_stage(self.predict_first_stage(preprocessed_inputs, prefix))

    def predict_second_stage(self, pd):
        """Second half of `predict` (fmA:604-609): ROI crops, box-classifier tails and heads.  Split from the first half
        so that a trainer can run work that must precede every use of a second-stage weight (the previous step's
        deferred head update, trainer.py) beside the first half."""
        self._lanes.mark("feat")        # (re-recorded here: the two halves may be captured into different CUDA graphs)
        pd.update(self._predict_second_stage(pd))
        return pd

    def predict_first_stage(self, preprocessed_inputs, prefix=None):
        """First half of `predict`: feature extractor, RPN head, anchors, and -- no second-stage variable is involved
        yet -- `_postprocess_rpn` (proposal decode, NMS, minibatch sampling), whose result `_predict_second_stage`
        picks up from the dictionary."""
        ws, fe = self._ws, self._feature_extractor
        B, H, W, _ = preprocessed_inputs.shape
        image_shape = (B, H, W, 3)
        self._lanes.reset()
        if prefix is not None:
            feat = fe.extract_proposal_features(preprocessed_inputs, self.first_stage_feature_extractor_scope, ws,
                                                prefix=prefix)
        else:
            feat = fe.extract_proposal_features(preprocessed_inputs, self.first_stage_feature_extractor_scope, ws)
        self._lanes.mark("feat")
        _, Hf, Wf, Cf = feat.shape
        anchors, keep_idx, Nk = self._anchors(Hf, Wf, H, W)
        rpn_feat = self._rpn_conv.fwd(feat, ws.get("rpn/conv", (B, Hf, Wf, self._rpn_conv.cout)))
        lay = self._first_stage_box_predictor.layout(self.first_stage_box_predictor_scope)
        rpn_out = ws.get("rpn/out", (B, Hf, Wf, lay["ld"]), torch.float32)
        rp = self._first_stage_box_predictor.predict(rpn_feat, lay["A"], self.first_stage_box_predictor_scope,
                                                     out=rpn_out)
        pd = PredictionDict()
        pd.update({
            "rpn_box_predictor_features": rpn_feat, "rpn_features_to_crop": feat, "image_shape": image_shape,
            "anchors": anchors,
            "rpn_box_encodings": lambda: (rp[BOX_ENCODINGS]().squeeze(2) if keep_idx is None
                                          else rp[BOX_ENCODINGS]().squeeze(2)[:, keep_idx.long()]),
            "rpn_objectness_predictions_with_background":
                lambda: (rp[CLASS_PREDICTIONS_WITH_BACKGROUND]() if keep_idx is None
                         else rp[CLASS_PREDICTIONS_WITH_BACKGROUND]()[:, keep_idx.long()]),
            "_rpn_out": rpn_out, "_rpn_layout": lay, "_keep_idx": keep_idx, "_Nk": Nk, "_feat_hw": (Hf, Wf),
        })
        pd["_proposals"] = self._postprocess_rpn(pd)
        return pd

    def _postprocess_rpn(self, pd):
        """fmA:1055-1132 + 1134-1216: decode, objectness softmax, clip, NMS, then (training) sample
        the box-classifier minibatch and pad; returns normalised proposals + counts on the device."""
        ws = self._ws
        B, H, W, _ = pd["image_shape"]
        Hf, Wf = pd["_feat_hw"]
        lay, Nk, HW = pd["_rpn_layout"], pd["_Nk"], Hf * Wf
        M, P = self._first_stage_max_proposals, self.max_num_proposals
        boxes = ws.get("rpn/dec_boxes", (B, Nk, 4), torch.float32)
        scores = ws.get("rpn/dec_scores", (B, Nk), torch.float32)
        keys = ws.get("rpn/dec_keys", (B, Nk), torch.int64)
        ops.call("mtl_rpn_decode", pd["_rpn_out"], lay["ld"], lay["box_col0"], lay["cls_col0"], lay["A"], HW,
                 pd["_keep_idx"], pd["anchors"], Nk, B, float(H), float(W), self._first_stage_nms_score_threshold,
                 boxes, scores, keys)
        order = ws.get("rpn/order", (B, Nk), torch.int32)
        nvalid = ws.get("rpn/nvalid", (B,), torch.int32)
        ops.call("mtl_rank_sort_desc", keys, B, Nk, order, nvalid, ws.get("rpn/rank_ws", (B, Nk), torch.int32))
        nms_b = ws.get("rpn/nms_boxes", (B, M, 4), torch.float32)
        nms_s = ws.get("rpn/nms_scores", (B, M), torch.float32)
        nms_n = ws.get("rpn/nms_num", (B,), torch.int32)
        ops.call("mtl_nms", boxes, scores, order, nvalid, B, Nk, self._first_stage_nms_iou_threshold, M, nms_b,
                 nms_s, None, nms_n)
        if not (self._is_training and not self._hard_example_miner):
            # inference: the NMS output is the proposal set (fmA:1111-1131), P == first_stage_max_proposals
            prop_abs = ws.get("det/prop_abs", (B, P, 4), torch.float32)
            prop_norm = ws.get("det/prop_norm", (B, P, 4), torch.float32)
            prop_sc = ws.get("det/prop_scores", (B, P), torch.float32)
            nprop = ws.get("det/num_proposals", (B,), torch.int32)
            ops.call("mtl_proposals_from_nms", nms_b, nms_s, nms_n, B, M, float(H), float(W), prop_abs, prop_norm,
                     prop_sc, nprop)
            return prop_norm, prop_abs, prop_sc, nprop
        gt = self._format_groundtruth_data(pd["image_shape"])
        # _sample_box_classifier_minibatch (fmA:1268-1302): detector assignment on the unpadded proposals
        match = ws.get("det/sample_match", (B, M), torch.int32)
        ops.call("mtl_iou_match", gt["gt"], gt["num_gt"], gt["gmax"], nms_b, M, nms_n, B, M, 0.5, 0.5, 1, 0, match,
                 None, None)
        if self._sampler_keys is None:
            raise RuntimeError("provide_sampler_keys() must be called before predict() in training mode")
        sampled = ws.get("det/sampled", (B, M), torch.uint8)
        counts = ws.get("det/sample_counts", (B, 4), torch.int32)
        ops.call("mtl_balanced_sample", match, self._sampler_keys[1], B, M, self._second_stage_batch_size,
                 self._second_stage_balance_fraction, sampled, counts)
        prop_abs = ws.get("det/prop_abs", (B, P, 4), torch.float32)
        prop_norm = ws.get("det/prop_norm", (B, P, 4), torch.float32)
        prop_sc = ws.get("det/prop_scores", (B, P), torch.float32)
        nprop = ws.get("det/num_proposals", (B,), torch.int32)
        ops.call("mtl_gather_sampled", nms_b, nms_s, sampled, B, M, P, float(H), float(W), prop_abs, prop_norm,
                 prop_sc, nprop)
        return prop_norm, prop_abs, prop_sc, nprop

    def _box_ind(self, B, n, tag):
        key = "box_ind/%s" % tag
        if key not in self._ws.bufs or self._ws.bufs[key].shape != (B * n,):
            self._ws.bufs[key] = torch.arange(B, dtype=torch.int32, device=self.device).repeat_interleave(n)
        return self._ws.bufs[key]

    def _compute_second_stage_input_feature_maps(self, feat, boxes_norm, box_ind, tag):
        """fmA:1304-1348: crop_and_resize to initial_crop_size then max pool."""
        ws = self._ws
        B, Hf, Wf, C = feat.shape
        R = boxes_norm.shape[0]
        c = self._initial_crop_size
        crops = ws.get("crops/%s" % tag, (R, c, c, C))
        ops.call("mtl_crop_and_resize_fwd", feat, B, Hf, Wf, C, boxes_norm, box_ind, R, c, c, crops)
        k = self._maxpool_kernel_size
        if k == 1:
            return crops, None
        p, q = max_pool_out_hw(c, c, k, self._maxpool_stride, "VALID")
        pooled = max_pool(crops, ws.get("crops/%s/pool" % tag, (R, p, q, C)), k, self._maxpool_stride, "VALID")
        return pooled, crops

    def _predict_second_stage(self, pd):
        """fmA:611-719."""
        ws, fe, mtl = self._ws, self._feature_extractor, self._mtl
        B = pd["image_shape"][0]
        P = self.max_num_proposals
        prop_norm, prop_abs, prop_sc, nprop = pd["_proposals"] if "_proposals" in pd else self._postprocess_rpn(pd)
        feat = pd["rpn_features_to_crop"]
        maps, pre_pool = self._compute_second_stage_input_feature_maps(
            feat, prop_norm.view(B * P, 4), self._box_ind(B, P, "props"), "props")
        self._lanes.mark("crops")
        cls_feat = fe.extract_box_classifier_features(maps, self.second_stage_feature_extractor_scope, ws, "main")
        bp = self._mask_rcnn_box_predictor.predict(cls_feat, 1, self.second_stage_box_predictor_scope, ws=ws,
                                                   tag="main", boxes_normalized=prop_norm.view(B * P, 4))
        out = {
            "refined_box_encodings": lambda: bp[BOX_ENCODINGS]().squeeze(1),
            "class_predictions_with_background": lambda: bp[CLASS_PREDICTIONS_WITH_BACKGROUND]().squeeze(1),
            "num_proposals": nprop, "proposal_boxes": prop_abs, "proposal_boxes_normalized": prop_norm,
            "_head_out": bp["_raw"], "_proposal_maps": maps, "_proposal_prepool": pre_pool,
        }
        if mtl is not None and mtl.closeness:
            with self._lanes.run("close", after=["crops"]):
                cfeat = fe.extract_box_classifier_features(maps, self.closeness_box_predictor_scope, ws, "close")
                cp = self._closeness_box_predictor.predict_class(cfeat, self.closeness_box_predictor_scope, ws=ws,
                                                                 tag="close")
                self._lanes.mark("close_fwd")
            out["closeness_predictions"] = lambda: cp[CLASS_PREDICTIONS]().squeeze(1)
            out["_close_out"] = cp["_raw"]
        return out

    def predict_with_window(self, prediction_dict, window_boxes_normalized=None, _tag="win", _keep=True,
                            _box_ind=None, _pre=None):
        """fmA:721-755."""
        ws, fe = self._ws, self._feature_extractor
        feat = prediction_dict["rpn_features_to_crop"]
        B = feat.shape[0]
        if window_boxes_normalized is None:
            window_boxes_normalized = self._format_groundtruth_data(prediction_dict["image_shape"])["win_boxes"]
        wb = window_boxes_normalized
        if wb.dim() == 3:
            nw = wb.shape[1]
            box_ind = self._box_ind(B, nw, "win%d" % nw) if _box_ind is None else _box_ind
            wb = wb.reshape(-1, 4)
        else:
            box_ind = _box_ind          # rank-2 boxes: all box_ind = 0 in the reference (fmA:1330-1332)
        lane, after = ("win", ["feat"]) if _tag == "win" else ("ref", ["crops"])
        with self._lanes.run(lane, after=after):
            if _pre is not None:
                _pre()
            maps, pre_pool = self._compute_second_stage_input_feature_maps(feat, wb, box_ind, _tag)
            if (not _keep and getattr(fe, "supports_pooled_tail", False)
                    and getattr(self._window_box_predictor, "accepts_pooled_tail", lambda s: False)(
                        self.window_box_predictor_scope)):
                # forward-only pass (the refiner's expanded windows): the tail's last conv sums the ROI grid itself
                wfeat = fe.extract_box_classifier_features(maps, self.window_box_predictor_scope, ws, _tag, keep=False,
                                                           pool=True)
            else:
                wfeat = fe.extract_box_classifier_features(maps, self.window_box_predictor_scope, ws, _tag, keep=_keep)
            wp = self._window_box_predictor.predict_class(wfeat, self.window_box_predictor_scope, ws=ws, tag=_tag,
                                                          activation_fn=None)
            self._lanes.mark(lane + "_fwd")
        prediction_dict["window_class_predictions"] = lambda: wp[CLASS_PREDICTIONS]().squeeze(1)
        prediction_dict["_%s_out" % _tag] = wp["_raw"]
        prediction_dict["_%s_boxes" % _tag] = (wb, box_ind, pre_pool)
        return prediction_dict

    def predict_edgemask(self, prediction_dict):
        """fmA:757-762."""
        feat = prediction_dict["rpn_features_to_crop"]
        B, Hf, Wf, _ = feat.shape
        act = self._ws.get("edgemask/act", (B, Hf, Wf, 2), torch.float32)
        with self._lanes.run("win", after=["feat"]):
            r = self._edgemask_predictor.predict(feat, self.edgemask_predictor_scope, out=act)
            self._lanes.mark("em_fwd")
        prediction_dict["edgemask_predictions"] = r[MASK_PREDICTIONS]
        return prediction_dict

    def predict_with_mtl_results(self, prediction_dict):
        """fmA:764-846: the class refiner over [original logits | 5 expanded-window logits |
        mean closeness logits] (all behind stop_gradient), plus the residual connection."""
        ws, mtl = self._ws, self._mtl
        if mtl.stop_gradient_for_prediction_org:
            raise ValueError("stop_gradient_for_prediction_org is not supported on the B200 path")
        B = prediction_dict["image_shape"][0]
        P = self.max_num_proposals
        K1 = self.num_classes + 1
        head_lay = self._mask_rcnn_box_predictor.layout(self.second_stage_box_predictor_scope)
        head_out = prediction_dict["_head_out"]
        win_out = None
        E = 0
        if mtl.window:
            E = 5
            exp = ws.get("refine/expand", (E, B, P, 4), torch.float32)
            bi = ws.get("refine/box_ind", (E, B, P), torch.int32)
            pbn = prediction_dict["proposal_boxes_normalized"]
            wpd = PredictionDict()
            wpd["rpn_features_to_crop"] = prediction_dict["rpn_features_to_crop"]
            wpd["image_shape"] = prediction_dict["image_shape"]
            self.predict_with_window(wpd, window_boxes_normalized=exp.view(E * B * P, 4), _tag="refine",
                                     _keep=False, _box_ind=bi.view(-1),
                                     _pre=lambda: ops.call("mtl_expand_windows", pbn, B, P, E - 1, exp, bi))
            win_out = wpd["_refine_out"]
            prediction_dict["expand_window_class_predictions"] = \
                lambda: win_out[:, :K1].reshape(E, B * P, K1).transpose(0, 1)
        close_out = prediction_dict.get("_close_out") if mtl.closeness else None
        if close_out is not None and not mtl.global_closeness:
            raise ValueError("global_closeness: false is not supported on the B200 path")
        self._lanes.wait("ref_fwd", "close_fwd")
        nf = self._refine_nf
        cat = ws.get("refine/in", (B * P, nf), torch.float32)
        wl = self._window_box_predictor.layout(self.window_box_predictor_scope)["ld"] if mtl.window else 0
        cl = self._closeness_box_predictor.layout(self.closeness_box_predictor_scope)["ld"] if mtl.closeness else 0
        ops.call("mtl_refine_concat", head_out, head_lay["ld"], head_lay["cls_col0"], win_out, wl, 0, E, close_out,
                 cl, 0, B * P, K1, cat, nf)
        refined = ws.get("refine/out", (B * P, K1), torch.float32)
        res = head_out[:, head_lay["cls_col0"]:] if mtl.refine_residue else None
        ops.call("mtl_fc_fwd", cat, nf, self._refine_w.w, self._refine_b.w, res, head_lay["ld"], B * P, K1, nf,
                 refined, K1)
        prediction_dict["mtl_refined_class_predictions_with_background"] = refined
        prediction_dict["_refine_in"] = cat
        return prediction_dict

    # ------------------------------------------------------------------ losses (+ head gradients)
    def loss(self, prediction_dict, scope=None):
        """fmA:1514-1590.  Returns {loss name: 0-d device tensor}.  The fused loss kernels also write
        the gradients w.r.t. every head output; `backward()` consumes them."""
        ws, mtl, pd = self._ws, self._mtl, prediction_dict
        B, H, W, _ = pd["image_shape"]
        P, K = self.max_num_proposals, self.num_classes
        K1 = K + 1
        gt = self._format_groundtruth_data(pd["image_shape"])
        self._lanes.wait("win_fwd", "em_fwd", "close_fwd", "ref_fwd")
        losses = ws.get("loss/values", (8,), torch.float32, zero=True)
        Hf, Wf = pd["_feat_hw"]
        lay, Nk, HW = pd["_rpn_layout"], pd["_Nk"], Hf * Wf
        # ---- _loss_rpn (fmA:1591-1668)
        rmatch = ws.get("rpn/match", (B, Nk), torch.int32)
        rbest = ws.get("rpn/row_best", (B, gt["gmax"]), torch.int64)
        ops.call("mtl_iou_match", gt["gt"], gt["num_gt"], gt["gmax"], pd["anchors"], 0, None, B, Nk, 0.7, 0.3, 1, 1,
                 rmatch, None, rbest)
        rsampled = ws.get("rpn/sampled", (B, Nk), torch.uint8)
        rcounts = ws.get("rpn/sample_counts", (B, 4), torch.int32)
        ops.call("mtl_balanced_sample", rmatch, self._sampler_keys[0], B, Nk, self._first_stage_minibatch_size,
                 self._first_stage_positive_balance_fraction, rsampled, rcounts)
        d_rpn = ws.get("rpn/d_out", (B, Hf, Wf, lay["ld"]))
        ops.call("mtl_rpn_loss", pd["_rpn_out"], lay["ld"], lay["box_col0"], lay["cls_col0"], lay["A"], HW,
                 pd["_keep_idx"], pd["anchors"], Nk, gt["gt"], gt["gmax"], rmatch, rsampled, rcounts, B,
                 self._first_stage_loc_loss_weight, self._first_stage_obj_loss_weight, self._first_stage_sigma,
                 losses[0:2], d_rpn)
        # ---- _loss_box_classifier (fmA:1670-1793)
        dmatch = ws.get("det/match", (B, P), torch.int32)
        ops.call("mtl_iou_match", gt["gt"], gt["num_gt"], gt["gmax"], pd["proposal_boxes"], P, None, B, P, 0.5, 0.5,
                 1, 0, dmatch, None, None)
        cls_t = ws.get("det/cls_t", (B, P), torch.int32)
        reg_t = ws.get("det/reg_t", (B, P, 4), torch.float32)
        reg_w = ws.get("det/reg_w", (B, P), torch.float32)
        cls_w = ws.get("det/cls_w", (B, P), torch.float32)
        use_close = mtl is not None and mtl.closeness
        close_t = ws.get("det/close_t", (B, P, K1), torch.float32) if use_close else None
        close_w = ws.get("det/close_w", (B, P), torch.float32) if use_close else None
        ops.call("mtl_detection_targets", dmatch, pd["proposal_boxes"], gt["gt"], gt["gt_cls"],
                 gt["gt_close"] if use_close else None, B, gt["gmax"], P, K1, cls_t, reg_t, reg_w, cls_w, close_t,
                 close_w)
        hl = self._mask_rcnn_box_predictor.layout(self.second_stage_box_predictor_scope)
        head_out = pd["_head_out"]
        d_head = ws.get("det/d_head", head_out.shape, torch.float32)
        ops.call("mtl_box_classifier_loss", head_out, hl["ld"], hl["box_col0"], hl["cls_col0"], K, cls_t, reg_t,
                 reg_w, cls_w, pd["num_proposals"], B, P, self._second_stage_loc_loss_weight,
                 self._second_stage_cls_loss_weight, losses[2:4], d_head, hl["ld"])
        if use_close:
            co = pd["_close_out"]
            d_close = ws.get("det/d_close", co.shape, torch.float32)
            ops.call("mtl_softmax_ce", co, co.shape[1], 1, K, close_t, K1, 1, None, close_w, None, P, 0, B * P,
                     mtl.closeness_loss_weight, losses[4:5], d_close, co.shape[1], 1, 0)
        # ---- _loss_window_class (fmA:1839-1858)
        if mtl is not None and mtl.window:
            wo = pd["_win_out"]
            rows = wo.shape[0]
            d_win = ws.get("det/d_win", wo.shape, torch.float32)
            ops.call("mtl_softmax_ce", wo, wo.shape[1], 0, K1, gt["win_cls"], K1, 0, None, None, None, 1, 0, rows,
                     mtl.window_class_loss_weight / max(rows, 1), losses[5:6], d_win, wo.shape[1], 0, 0)
        # ---- _loss_edgemask (fmA:1860-1881)
        if mtl is not None and mtl.edgemask:
            act = pd["edgemask_predictions"]
            d_act = ws.get("edgemask/d_act", act.shape, torch.float32)
            em = gt["edgemask"]
            ops.call("mtl_edgemask_loss", act, B, Hf, Wf, em, em.shape[2], em.shape[3], mtl.edgemask_loss_weight,
                     losses[6:7], d_act)
        # ---- _loss_refined_classifier (fmA:1795-1837): same assignment, class part only
        if mtl is not None and mtl.refine:
            refined = pd["mtl_refined_class_predictions_with_background"]
            d_ref = ws.get("refine/d_out", refined.shape, torch.float32)
            ops.call("mtl_softmax_ce", refined, K1, 0, K1, None, 0, 0, cls_t, cls_w, pd["num_proposals"], P, 1,
                     B * P, mtl.refined_classification_loss_weight / B, losses[7:8], d_ref, K1, 0, 0)
            # the residual connection feeds the same gradient into the original class logits
            if mtl.refine_residue:
                ops.call("mtl_softmax_ce", refined, K1, 0, K1, None, 0, 0, cls_t, cls_w, pd["num_proposals"], P, 1,
                         B * P, mtl.refined_classification_loss_weight / B, ws.get("loss/scratch", (1,), torch.float32),
                         d_head, hl["ld"], hl["cls_col0"], 1)
        self._last_pd = pd
        loss_dict = {}
        active = [True, True, True, True, use_close, mtl is not None and mtl.window,
                  mtl is not None and mtl.edgemask, mtl is not None and mtl.refine]
        for i, k in enumerate(LOSS_KEYS):
            if active[i]:
                loss_dict[k] = losses[i]
        return loss_dict

    # ------------------------------------------------------------------ detections
    def postprocess(self, prediction_dict):
        """fmA:996-1053 second-stage branch -> `_postprocess_box_classifier` (fmA:1387-1469): decode the refined
        per-class boxes against the proposals, convert the class scores, per-class NMS
        (`second_stage_post_processing.batch_non_max_suppression`), top `max_total_detections`, zero padded.
        Returns detection_boxes [B,T,4] (normalised to the image), detection_scores, detection_classes
        (0-based, float32) [B,T] and num_detections [B] (float32), all device tensors."""
        if self._first_stage_only:
            raise NotImplementedError("postprocess of first_stage_only models (RPN proposals) is not built yet")
        pd, ws = prediction_dict, self._ws
        nms_cfg = self._second_stage_nms_fn
        mode = {"IDENTITY": 0, "SOFTMAX": 1, "SIGMOID": 2}[str(self._second_stage_score_conversion_fn)]
        enc = pd["refined_box_encodings"].contiguous().float()
        if self._mtl is not None and self._mtl.refine and "mtl_refined_class_predictions_with_background" in pd:
            logits = pd["mtl_refined_class_predictions_with_background"].contiguous().float()      # fmA:1040-1043
        else:
            logits = pd["class_predictions_with_background"].contiguous().float()
        props = pd["proposal_boxes"]
        B, P = props.shape[0], props.shape[1]
        K = self.num_classes
        H, W = pd["image_shape"][1], pd["image_shape"][2]
        M = min(int(nms_cfg.max_detections_per_class), P)
        T = int(nms_cfg.max_total_detections)
        f32, i32 = torch.float32, torch.int32
        boxes_n = ws.get("det_pp/boxes", (B, K, P, 4), f32)
        scores = ws.get("det_pp/scores", (B, K, P), f32)
        keys = ws.get("det_pp/keys", (B, K, P), torch.int64)
        ops.call("mtl_detection_decode", enc, logits, props, pd["num_proposals"], B, P, K, float(H), float(W),
                 float(nms_cfg.score_threshold), mode, boxes_n, scores, keys, None)
        order = ws.get("det_pp/order", (B * K, P), i32)
        nvalid = ws.get("det_pp/nvalid", (B * K,), i32)
        ops.call("mtl_rank_sort_desc", keys, B * K, P, order, nvalid, ws.get("det_pp/rank_ws", (B * K, P), i32))
        cls_b = ws.get("det_pp/cls_boxes", (B, K, M, 4), f32)
        cls_s = ws.get("det_pp/cls_scores", (B, K, M), f32)
        cls_n = ws.get("det_pp/cls_num", (B * K,), i32)
        ops.call("mtl_nms", boxes_n, scores, order, nvalid, B * K, P, float(nms_cfg.iou_threshold), M, cls_b, cls_s,
                 None, cls_n)
        keys2 = ws.get("det_pp/keys2", (B, K * M), torch.int64)
        ops.call("mtl_detection_merge_keys", cls_s, cls_n, B, K, M, keys2)
        order2 = ws.get("det_pp/order2", (B, K * M), i32)
        nvalid2 = ws.get("det_pp/nvalid2", (B,), i32)
        ops.call("mtl_rank_sort_desc", keys2, B, K * M, order2, nvalid2, ws.get("det_pp/rank_ws2", (B, K * M), i32))
        det_b = ws.get("det_pp/det_boxes", (B, T, 4), f32)
        det_s = ws.get("det_pp/det_scores", (B, T), f32)
        det_c = ws.get("det_pp/det_classes", (B, T), f32)
        det_n = ws.get("det_pp/num_detections", (B,), f32)
        ops.call("mtl_detection_gather", cls_b, cls_s, order2, nvalid2, B, K, M, T, det_b, det_s, det_c, det_n)
        return {"detection_boxes": det_b, "detection_scores": det_s, "detection_classes": det_c,
                "num_detections": det_n}

    # ------------------------------------------------------------------ backward
    def backward(self, prediction_dict=None, part=None):
        """Explicit reverse pass: accumulates d(sum of task losses)/d(weights) into the gradient
        arena (`param_store.g`).  Regularisation gradients are added by the optimizer kernel.
        part="heads" runs the second-stage / aux-head half (everything whose gradients live after the
        trunk + RPN variables in the arena), part="trunk" the RPN + trunk half; the trainer all-reduces
        the first half's gradient bucket while the second half computes."""
        pd = prediction_dict or self._last_pd
        if part in ("trunk", "trunk_hi", "trunk_lo"):
            return self._backward_trunk(pd, None if part == "trunk" else part[6:])
        ws, fe, mtl = self._ws, self._feature_extractor, self._mtl
        feat = pd["rpn_features_to_crop"]
        B, Hf, Wf, C = feat.shape
        P = self.max_num_proposals
        stop_aux = mtl is not None and mtl.stop_gradient_for_aux_tasks
        dfeat = ws.get("bwd/dfeat_f32", feat.shape, torch.float32, zero=True)
        L = self._lanes
        # The weight-gradient GEMMs of the second-stage tails and heads are collected (ops_conv.WgradCollector) and run
        # as grouped launches: at the end of this half of the backward pass, or -- `defer_head_wgrads` -- when the
        # trainer calls flush_head_wgrads(), which it does underneath the NEXT step's trunk forward pass.
        hcol = self._head_wgrads.setdefault(id(ws), ops_conv.WgradCollector(target_k_iters=56)) \
            if (self.group_head_wgrads and ops_conv.WgradCollector.enabled and part in ("heads", "heads_async")) else None
        if hcol is not None:
            hcol.__enter__()
        try:
            return self._backward_heads(pd, part, ws, fe, mtl, feat, B, Hf, Wf, C, P, stop_aux, dfeat, L)
        finally:
            if hcol is not None:
                hcol.__exit__(None, None, None)
                if not self.defer_head_wgrads:
                    self.flush_head_wgrads()

    def flush_head_wgrads(self, max_ctas=0, piece=None, pieces=1):
        """Run the second-stage weight-gradient GEMMs collected by the last backward(part="heads*"); piece k of
        `pieces`: only those writing into head_exchange_pieces(pieces)[k] (the last piece also runs any leftovers)."""
        col = self._head_wgrads.get(id(self._ws))
        if col is None:
            return
        if piece is None or pieces <= 1:
            col.flush(max_ctas=max_ctas)
            return
        g = self.head_exchange_pieces(pieces)[piece][0]
        lo = g.data_ptr()
        col.flush(max_ctas=max_ctas, key=("piece", piece, pieces), out_range=(lo, lo + g.numel() * g.element_size()))
        if piece == pieces - 1:
            col.flush(max_ctas=max_ctas, key=("piece", "rest", pieces))

    def head_exchange_pieces(self, pieces):
        """The second-stage gradient bucket cut at tensor boundaries into `pieces` contiguous parts of about equal
        size, arena order: [(gradient view, (first tensor index, end tensor index))].  With several replicas each part
        is exchanged as soon as its weight-gradient GEMMs are done, while the next part's still run."""
        cached = getattr(self, "_head_pieces", None)
        if cached is not None and cached[0] == pieces:
            return cached[1]
        st = self._store
        t0, t1 = self.head_tensor_range()
        ps = st.params
        first = ps[t0].offset
        end = ps[t1].offset if t1 < len(ps) else st.total
        cuts, out = [t0], []
        for k in range(1, pieces):
            want = first + (end - first) * k // pieces
            t = min(range(cuts[-1] + 1, t1), key=lambda i: abs(ps[i].offset - want), default=None)
            if t is not None and t > cuts[-1]:
                cuts.append(t)
        cuts.append(t1)
        for a, b in zip(cuts[:-1], cuts[1:]):
            lo = ps[a].offset
            hi = ps[b].offset if b < len(ps) else st.total
            out.append((st.g[lo:min(hi, end)], (a, b)))
        self._head_pieces = (pieces, out)
        return out

    def _backward_heads(self, pd, part, ws, fe, mtl, feat, B, Hf, Wf, C, P, stop_aux, dfeat, L):
        # refiner FC (inputs are behind stop_gradient: weights / bias only)
        if mtl is not None and mtl.refine:
            K1 = self.num_classes + 1
            ops.call("mtl_fc_bwd", pd["_refine_in"], self._refine_nf, self._refine_w.w, ws.bufs["refine/d_out"], K1,
                     B * P, K1, self._refine_nf, self._refine_w.g, self._refine_b.g, None, 0)
        if mtl is not None and mtl.edgemask:     # plain read-modify-write of dfeat: before the atomic scatters start
            self._edgemask_predictor.backward(self.edgemask_predictor_scope, feat, pd["edgemask_predictions"],
                                              ws.bufs["edgemask/d_act"], dfeat)
        L.mark("bwd_start")
        # closeness and window tails run on their own lanes, concurrently with the main tail
        d_extra = None
        if mtl is not None and mtl.closeness:
            with L.run("close", after=["bwd_start"]):
                g = self._closeness_box_predictor.backward(self.closeness_box_predictor_scope, "close",
                                                           ws.bufs["det/d_close"], ws)
                d_extra = fe.backward_box_classifier_features(self.closeness_box_predictor_scope, g, ws, "close",
                                                              need_dx=not stop_aux)
                if d_extra is not None and not fe.supports_dx_extra:
                    self._crop_backward(pd, d_extra, pd["_proposal_maps"], pd["_proposal_prepool"],
                                        pd["proposal_boxes_normalized"].view(B * P, 4), self._box_ind(B, P, "props"),
                                        dfeat, "close")
                    d_extra = None
                L.mark("close_bwd")
        if mtl is not None and mtl.window:
            with L.run("win", after=["bwd_start"]):
                g = self._window_box_predictor.backward(self.window_box_predictor_scope, "win", ws.bufs["det/d_win"],
                                                        ws)
                dwin = fe.backward_box_classifier_features(self.window_box_predictor_scope, g, ws, "win",
                                                           need_dx=not stop_aux)
                if not stop_aux:
                    wb, bi, pre = pd["_win_boxes"]
                    self._crop_backward(pd, dwin, None, pre, wb, bi, dfeat, "win")
                L.mark("win_bwd")
        g = self._mask_rcnn_box_predictor.backward(self.second_stage_box_predictor_scope, "main",
                                                   ws.bufs["det/d_head"], ws)
        # the closeness crop gradient rides on the main tail's last dgrad (residual input of the epilogue)
        dmaps = fe.backward_box_classifier_features(self.second_stage_feature_extractor_scope, g, ws, "main",
                                                    need_dx=True, dx_extra=d_extra,
                                                    pre_unit0=lambda: L.wait("close_bwd"))
        self._crop_backward(pd, dmaps, pd["_proposal_maps"], pd["_proposal_prepool"],
                            pd["proposal_boxes_normalized"].view(B * P, 4), self._box_ind(B, P, "props"), dfeat,
                            "props")
        L.wait("win_bwd", "close_bwd")
        if part == "heads":
            Concurrency.join()
            return
        if part == "heads_async":       # the caller orders its consumers after the side streams itself
            return
        self._backward_trunk(pd)

    def _backward_trunk(self, pd, half=None):
        """half=None: the whole RPN + trunk backward.  "hi" / "lo": its two halves (RPN and the later trunk units, then
        the earlier ones), for a trainer that exchanges the first half's gradient bucket while the second computes."""
        ws, fe = self._ws, self._feature_extractor
        feat = pd["rpn_features_to_crop"]
        dfeat = ws.bufs["bwd/dfeat_f32"]
        # Weight gradients only feed the optimizer: where every gradient tensor of the chain keeps its own buffer until
        # the end of the pass (`deferred_wgrad_safe`), the weight-gradient GEMMs of the RPN + trunk are collected and run
        # as grouped launches (ops_conv.WgradCollector) instead of 81 launches of 8-54 CTAs each.
        import contextlib
        col = None
        if getattr(fe, "deferred_wgrad_safe", False) and ops_conv.WgradCollector.enabled:
            col = self._trunk_wgrads.setdefault(id(ws), ops_conv.WgradCollector())
        if half is not None and col is None:
            raise ValueError("the split trunk backward needs the grouped weight-gradient path")

        def chunk(j):
            # every 6 bottleneck units the collected weight-gradient problems go to a side stream as one grouped
            # launch sized for 64 SMs; the dgrad chain (<= 76 CTAs per layer at batch 1) keeps the rest
            with torch.cuda.stream(Concurrency.fork()):
                col.flush(max_ctas=self._wgrad_chunk_ctas, key=j)

        with (col if col is not None else contextlib.nullcontext()):
            if half != "lo":
                # RPN head and conv; the conv's dgrad epilogue merges the fp32 ROI/edgemask gradient and
                # applies the ReLU mask of the trunk output
                rpn_feat = pd["rpn_box_predictor_features"]
                d_rpn_feat = ws.get("bwd/d_rpn_feat", rpn_feat.shape)
                self._first_stage_box_predictor.backward(self.first_stage_box_predictor_scope, rpn_feat,
                                                         ws.bufs["rpn/d_out"], d_rpn_feat,
                                                         rpn_feat if self._rpn_conv.relu else None)
                self._rpn_conv.wgrad(feat, d_rpn_feat)
                self._g_feat = self._rpn_conv.dgrad(d_rpn_feat, feat.shape, ws.get("bwd/g_feat", feat.shape), res=dfeat,
                                                    mask=feat, mask_hi=fe.feature_mask_hi)
            if col is not None:
                fe.backward_proposal_features(self.first_stage_feature_extractor_scope, self._g_feat, ws, every=6,
                                              checkpoint=chunk, part=half)
            else:
                fe.backward_proposal_features(self.first_stage_feature_extractor_scope, self._g_feat, ws)
        if col is not None:
            if half == "hi":
                chunk(100)                  # everything of this half: its bucket is exchanged next
            else:
                col.flush(key=0)            # the last chunk has the machine to itself
        Concurrency.join()      # side-stream weight-gradient GEMMs must land before the optimizer / the exchange

    def gradient_buckets3(self):
        """(second-stage bucket, trunk "hi" bucket = later trunk units + RPN, trunk "lo" bucket = the rest): contiguous
        views of the gradient arena in the order the backward pass finishes them."""
        st = self._store
        heads, trunk = self.gradient_buckets()
        pre = self._feature_extractor.trunk_split_variable(self.first_stage_feature_extractor_scope)
        cut = next(p.offset for p in st.params if p.name.startswith(pre))
        return heads, trunk[cut:], trunk[:cut]

    def _crop_backward(self, pd, dmaps, maps, pre_pool, boxes, box_ind, dfeat, tag):
        ws = self._ws
        feat = pd["rpn_features_to_crop"]
        B, Hf, Wf, C = feat.shape
        c = self._initial_crop_size
        if pre_pool is not None:
            dcrops = max_pool_bwd(pre_pool, dmaps, ws.get("bwd/dcrops/%s" % tag, pre_pool.shape),
                                  self._maxpool_kernel_size, self._maxpool_stride, "VALID")
        else:
            dcrops = dmaps
        ops.call("mtl_crop_and_resize_bwd", dcrops, B, Hf, Wf, C, boxes, box_ind, boxes.shape[0], c, c, dfeat)

    # ------------------------------------------------------------------ misc API
    def restore_map(self, from_detection_checkpoint=True):
        """fmA:1947-2013: {checkpoint variable name: variable}.  For classification checkpoints the
        stage scopes are stripped so that every block4 copy maps onto the same keys (traps T5, T14)."""
        out = {}
        for p in self._store.params:
            name = p.name
            if "/_pad/" in name:
                continue
            if from_detection_checkpoint:
                if "/_dead/" not in name:
                    out[name] = p
                continue
            for sc in (self.first_stage_feature_extractor_scope + "/_dead/",
                       self.first_stage_feature_extractor_scope + "/", self.second_stage_feature_extractor_scope + "/",
                       self.closeness_box_predictor_scope + "/", self.window_box_predictor_scope + "/"):
                if name.startswith(sc) and "/" + self._feature_extractor._architecture + "/" in "/" + name:
                    out.setdefault(name[len(sc):], []).append(p)
                    break
        return out


def _np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)
