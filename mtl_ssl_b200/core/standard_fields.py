"""Names shared across the detection API (subset of
/root/reference/object_detection/core/standard_fields.py used on the training path)."""


class BoxListFields(object):
    boxes = "boxes"
    classes = "classes"
    scores = "scores"
    closeness = "closeness"
    ignore = "ignore"
    masks = "masks"
    keypoints = "keypoints"
    edgemask = "edgemask"


BOX_ENCODINGS = "box_encodings"
CLASS_PREDICTIONS_WITH_BACKGROUND = "class_predictions_with_background"
CLASS_PREDICTIONS = "class_predictions"
MASK_PREDICTIONS = "mask_predictions"
