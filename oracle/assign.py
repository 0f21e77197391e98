"""Oracle (test infrastructure): ArgMaxMatcher, TargetAssigner, balanced sampler — NumPy.

Paths are relative to /root/reference/object_detection/.
"""
import numpy as np

from . import boxes as B

F = np.float32


def argmax_match(sim, matched_threshold, unmatched_threshold=None, negatives_lower_than_unmatched=True,
                 force_match_for_each_row=False):
    """matchers/argmax_matcher.py:102-175 `_match`.

    sim: [rows (groundtruth), cols (anchors)] float32.  Returns int32 [cols]:
    >=0 matched row, -1 unmatched (negative), -2 ignored.
    """
    sim = np.asarray(sim, F)
    if unmatched_threshold is None:
        unmatched_threshold = matched_threshold
    ncols = sim.shape[1]
    if sim.shape[0] == 0:                      # _match_when_rows_are_empty (:113-122)
        return -np.ones([ncols], np.int32)
    matches = np.argmax(sim, axis=0).astype(np.int64)          # first max wins (tf.argmax)
    if matched_threshold is not None:
        matched_vals = np.max(sim, axis=0)
        below = F(unmatched_threshold) > matched_vals
        between = (matched_vals >= F(unmatched_threshold)) & (F(matched_threshold) > matched_vals)
        if negatives_lower_than_unmatched:
            matches = np.where(below, -1, matches)
            matches = np.where(between, -2, matches)
        else:
            matches = np.where(below, -2, matches)
            matches = np.where(between, -1, matches)
    if force_match_for_each_row:
        # :158-169: dynamic_stitch([forced_ids, keep_ids], [row_range, kept]) — later entries of
        # forced_ids override earlier ones, so for duplicate columns the HIGHER row index wins.
        forced = np.argmax(sim, axis=1)
        for r in range(sim.shape[0]):
            matches[forced[r]] = r
    return matches.astype(np.int32)


def assign_targets(anchors, gt_boxes, gt_labels, unmatched_cls_target, matched_threshold,
                   unmatched_threshold=None, force_match=False, gt_closeness=None):
    """core/target_assigner.py:99-213 `TargetAssigner.assign` (IouSimilarity + ArgMaxMatcher +
    FasterRcnnBoxCoder as built by create_target_assigner :433-447).

    The crowd / ignore re-matching (:186-194) always sees empty box sets (trap T6: the
    `shape[0] is boxes.num_boxes()` identity test at :222 is always False) and therefore never
    changes cls_weights; it is restated as a no-op.

    gt_labels: [G, d...] float32 (None -> ones [G,1], :167-169).  Returns dict with
    cls_targets [N, d...], cls_weights [N], reg_targets [N,4], reg_weights [N], match [N] int32,
    and closeness_targets [N, Kc] when gt_closeness is given (`extension=True`, :283-307).
    """
    anchors = np.asarray(anchors, F).reshape(-1, 4)
    gt_boxes = np.asarray(gt_boxes, F).reshape(-1, 4)
    n = anchors.shape[0]
    if gt_labels is None:
        gt_labels = np.ones([gt_boxes.shape[0], 1], F)
    gt_labels = np.asarray(gt_labels, F)
    unmatched_cls_target = np.asarray(unmatched_cls_target, F)
    sim = B.iou(gt_boxes, anchors)                                  # rsc:57-74
    match = argmax_match(sim, matched_threshold, unmatched_threshold, True, force_match)
    matched = match >= 0
    mcols = np.nonzero(matched)[0]
    mrows = match[mcols]
    # _create_regression_targets (:256-281)
    reg_targets = np.zeros([n, 4], F)
    if len(mcols):
        reg_targets[mcols] = B.box_encode(gt_boxes[mrows], anchors[mcols])
    # _create_classification_targets (:322-353)
    cls_targets = np.tile(unmatched_cls_target[None], [n] + [1] * unmatched_cls_target.ndim).astype(F)
    if len(mcols):
        cls_targets[mcols] = gt_labels[mrows]
    reg_weights = matched.astype(F)                                 # :355-370
    cls_weights = (matched.astype(F) + (match == -1).astype(F)).astype(F)   # :372-402, weights 1/1
    out = dict(cls_targets=cls_targets, cls_weights=cls_weights, reg_targets=reg_targets,
               reg_weights=reg_weights, match=match)
    if gt_closeness is not None:
        gt_closeness = np.asarray(gt_closeness, F)
        ct = np.zeros([n, gt_closeness.shape[-1]], F)
        if len(mcols):
            ct[mcols] = gt_closeness[mrows]
        out["closeness_targets"] = ct
    return out


def assign_proposal(anchors, gt_boxes):
    """create_target_assigner('FasterRCNN','proposal') (:433-440): 0.7 / 0.3 / force-match,
    scalar label 1 for matched, unmatched_cls_target = 0 (:82-87 default)."""
    return assign_targets(anchors, gt_boxes, None, np.zeros([1], F), 0.7, 0.3, True)


def assign_detection(proposals, gt_boxes, gt_classes_with_background, gt_closeness=None):
    """create_target_assigner('FasterRCNN','detection') (:441-447): IoU>=0.5 positive else
    negative; unmatched target = one-hot background (fmA:354-359)."""
    k1 = np.asarray(gt_classes_with_background).shape[-1]
    unmatched = np.zeros([k1], F)
    unmatched[0] = 1
    return assign_targets(proposals, gt_boxes, gt_classes_with_background, unmatched, 0.5, None, False,
                          gt_closeness)


def subsample_indicator(indicator, num_samples, keys):
    """core/minibatch_sampler.py:64-90.  The reference shuffles the candidate indices with
    tf.random_shuffle (unreproducible); the restatement makes the permutation an INPUT:
    candidates are ordered by ascending `keys` (ties: lower index first) and the first
    `num_samples` are kept — the same distribution when keys are i.i.d. uniform."""
    indicator = np.asarray(indicator, bool)
    idx = np.nonzero(indicator)[0]
    order = idx[np.argsort(np.asarray(keys, F)[idx], kind="stable")]
    sel = order[:max(0, min(len(order), int(num_samples)))]
    out = np.zeros(indicator.shape, bool)
    out[sel] = True
    return out


def balanced_subsample(indicator, batch_size, labels, positive_fraction, keys):
    """core/balanced_positive_negative_sampler.py:50-91 `subsample`."""
    indicator = np.asarray(indicator, bool)
    labels = np.asarray(labels, bool)
    positive_idx = labels & indicator
    negative_idx = (~labels) & indicator
    max_num_pos = int(positive_fraction * batch_size)
    sampled_pos = subsample_indicator(positive_idx, max_num_pos, keys)
    max_num_neg = batch_size - int(sampled_pos.sum())
    sampled_neg = subsample_indicator(negative_idx, max_num_neg, keys)
    return sampled_pos | sampled_neg
