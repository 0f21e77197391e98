"""R-FCN meta-architecture with the auxiliary heads, on B200 kernels.

Mirrors /root/reference/object_detection/meta_architectures/rfcn_meta_arch.py:48-381: the RPN and all
losses are inherited from FasterRCNNMetaArch; the second stage runs block4 on the WHOLE feature map
(once per scope: box classifier, closeness, window) and pools position-sensitive score maps per
proposal (core/box_predictor.py:131-337).  For the refine windows the reference rebuilds the window
scope's block4 + maps on the same features; the values are identical, so the maps are pooled again
instead of being recomputed."""
import torch

from .. import ops
from ..core.standard_fields import (BOX_ENCODINGS, CLASS_PREDICTIONS, CLASS_PREDICTIONS_WITH_BACKGROUND)
from ..nets.layers import Concurrency
from .faster_rcnn_meta_arch import FasterRCNNMetaArch


class RFCNMetaArch(FasterRCNNMetaArch):
    def __init__(self, second_stage_rfcn_box_predictor, **kwargs):
        kwargs.setdefault("initial_crop_size", 1)          # unused by R-FCN (rfcn:48-150)
        kwargs.setdefault("maxpool_kernel_size", 1)
        kwargs.setdefault("maxpool_stride", 1)
        super(RFCNMetaArch, self).__init__(second_stage_mask_rcnn_box_predictor=second_stage_rfcn_box_predictor,
                                           **kwargs)
        self._rfcn_box_predictor = second_stage_rfcn_box_predictor
        self.supports_deferred_heads = False      # its backward() keeps the plain schedule (trainer.py)

    def _predict_second_stage(self, pd):
        """rfcn:208-310."""
        ws, fe, mtl = self._ws, self._feature_extractor, self._mtl
        B = pd["image_shape"][0]
        P = self.max_num_proposals
        prop_norm, prop_abs, prop_sc, nprop = pd["_proposals"] if "_proposals" in pd else self._postprocess_rpn(pd)
        feat = pd["rpn_features_to_crop"]
        boxes, bi = prop_norm.view(B * P, 4), self._box_ind(B, P, "props")
        self._lanes.mark("crops")
        cls_feat = fe.extract_box_classifier_features(feat, self.second_stage_feature_extractor_scope, ws, "main")
        bp = self._rfcn_box_predictor.predict(cls_feat, 1, self.second_stage_box_predictor_scope,
                                              proposal_boxes=boxes, box_ind=bi, ws=ws, tag="main")
        out = {
            "refined_box_encodings": lambda: bp[BOX_ENCODINGS]().squeeze(1),
            "class_predictions_with_background": lambda: bp[CLASS_PREDICTIONS_WITH_BACKGROUND]().squeeze(1),
            "num_proposals": nprop, "proposal_boxes": prop_abs, "proposal_boxes_normalized": prop_norm,
            "_head_out": bp["_raw"],
        }
        if mtl is not None and mtl.closeness:
            with self._lanes.run("close", after=["crops"]):
                cfeat = fe.extract_box_classifier_features(feat, self.closeness_box_predictor_scope, ws, "close")
                cp = self._closeness_box_predictor.predict_class(cfeat, self.closeness_box_predictor_scope,
                                                                 proposal_boxes=boxes, box_ind=bi, ws=ws, tag="close")
                self._lanes.mark("close_fwd")
            out["closeness_predictions"] = lambda: cp[CLASS_PREDICTIONS]().squeeze(1)
            out["_close_out"] = cp["_raw"]
        return out

    def predict_with_window(self, prediction_dict, window_boxes_normalized=None, _tag="win", _keep=True,
                            _box_ind=None, _pre=None):
        """rfcn:312-381."""
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
            box_ind = _box_ind
        scope = self.window_box_predictor_scope
        if _tag == "win":
            with self._lanes.run("win", after=["feat"]):
                wfeat = fe.extract_box_classifier_features(feat, scope, ws, "win")
                wp = self._window_box_predictor.predict_class(wfeat, scope, proposal_boxes=wb, box_ind=box_ind, ws=ws,
                                                              tag="win")
                self._lanes.mark("win_fwd")
        else:
            # refine windows: same features, same scope -> pool the cached position-sensitive maps again
            with self._lanes.run("win", after=["crops", "win_fwd"]):
                if _pre is not None:
                    _pre()
                wp = self._window_box_predictor.predict_class(None, scope, proposal_boxes=wb, box_ind=box_ind, ws=ws,
                                                              tag=_tag, reuse_maps_of="win")
                self._lanes.mark("ref_fwd")
        prediction_dict["window_class_predictions"] = lambda: wp[CLASS_PREDICTIONS]().squeeze(1)
        prediction_dict["_%s_out" % _tag] = wp["_raw"]
        prediction_dict["_%s_boxes" % _tag] = (wb, box_ind, None)
        return prediction_dict

    def backward(self, prediction_dict=None, part=None):
        pd = prediction_dict or self._last_pd
        if part == "trunk":
            return self._backward_trunk(pd)
        ws, fe, mtl, L = self._ws, self._feature_extractor, self._mtl, self._lanes
        feat = pd["rpn_features_to_crop"]
        B = feat.shape[0]
        P = self.max_num_proposals
        stop_aux = mtl is not None and mtl.stop_gradient_for_aux_tasks
        dfeat = ws.get("bwd/dfeat_f32", feat.shape, torch.float32, zero=True)
        if mtl is not None and mtl.refine:
            K1 = self.num_classes + 1
            ops.call("mtl_fc_bwd", pd["_refine_in"], self._refine_nf, self._refine_w.w, ws.bufs["refine/d_out"], K1,
                     B * P, K1, self._refine_nf, self._refine_w.g, self._refine_b.g, None, 0)
        if mtl is not None and mtl.edgemask:
            self._edgemask_predictor.backward(self.edgemask_predictor_scope, feat, pd["edgemask_predictions"],
                                              ws.bufs["edgemask/d_act"], dfeat)
        L.mark("bwd_start")
        d_close = d_win = None
        if mtl is not None and mtl.closeness:
            with L.run("close", after=["bwd_start"]):
                g = self._closeness_box_predictor.backward(self.closeness_box_predictor_scope, "close",
                                                           ws.bufs["det/d_close"], ws)
                d_close = fe.backward_box_classifier_features(self.closeness_box_predictor_scope, g, ws, "close",
                                                              need_dx=not stop_aux)
                L.mark("close_bwd")
        if mtl is not None and mtl.window:
            with L.run("win", after=["bwd_start"]):
                g = self._window_box_predictor.backward(self.window_box_predictor_scope, "win", ws.bufs["det/d_win"],
                                                        ws)
                d_win = fe.backward_box_classifier_features(self.window_box_predictor_scope, g, ws, "win",
                                                            need_dx=not stop_aux)
                L.mark("win_bwd")
        g = self._rfcn_box_predictor.backward(self.second_stage_box_predictor_scope, "main", ws.bufs["det/d_head"],
                                              ws)
        d_main = fe.backward_box_classifier_features(self.second_stage_feature_extractor_scope, g, ws, "main",
                                                     need_dx=True)
        L.wait("win_bwd", "close_bwd")
        # dense bf16 feature gradients of the (up to three) block4 copies -> fp32 accumulator
        for d in (d_main, d_close, d_win):
            if d is not None:
                ops.call("mtl_add_bf16_to_f32", d, d.numel(), dfeat)
        if part == "heads":
            Concurrency.join()
            return
        if part == "heads_async":
            return
        self._backward_trunk(pd)
