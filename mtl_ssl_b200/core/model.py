"""DetectionModel interface, mirroring /root/reference/object_detection/core/model.py:54-313
(including the fork's provide_window / provide_edgemask and the closeness list of
provide_groundtruth).  Tensors are torch CUDA tensors used as device buffers."""
from abc import ABCMeta, abstractmethod

from .standard_fields import BoxListFields as fields


class DetectionModel(object):
    __metaclass__ = ABCMeta

    def __init__(self, num_classes):
        self._num_classes = num_classes
        self._groundtruth_lists = {}
        self._window_lists = {}
        self._edgemask_lists = {}

    @property
    def num_classes(self):
        return self._num_classes

    def groundtruth_lists(self, field):
        if field not in self._groundtruth_lists:
            if field in (fields.ignore,):
                return [None] * len(self._groundtruth_lists.get(fields.boxes, []))
            raise RuntimeError("Groundtruth tensor %s has not been provided" % field)
        return self._groundtruth_lists[field]

    def window_lists(self, field):
        if field not in self._window_lists:
            raise RuntimeError("Window tensor %s has not been provided" % field)
        return self._window_lists[field]

    def edgemask_lists(self, field):
        if field not in self._edgemask_lists:
            raise RuntimeError("Edgemask tensor %s has not been provided" % field)
        return self._edgemask_lists[field]

    @abstractmethod
    def preprocess(self, inputs):
        pass

    @abstractmethod
    def predict(self, preprocessed_inputs):
        pass

    @abstractmethod
    def loss(self, prediction_dict):
        pass

    def provide_groundtruth(self, groundtruth_boxes_list, groundtruth_classes_list, groundtruth_closeness_list,
                            groundtruth_ignore=None, groundtruth_masks_list=None, groundtruth_keypoints_list=None):
        """model.py:225-266: normalised [G,4] boxes, one-hot [G,K] classes, [G,K+1] closeness per image."""
        self._groundtruth_lists[fields.boxes] = groundtruth_boxes_list
        self._groundtruth_lists[fields.classes] = groundtruth_classes_list
        self._groundtruth_lists[fields.closeness] = groundtruth_closeness_list
        if groundtruth_ignore:
            self._groundtruth_lists[fields.ignore] = groundtruth_ignore
        if groundtruth_masks_list:
            self._groundtruth_lists[fields.masks] = groundtruth_masks_list
        if groundtruth_keypoints_list:
            self._groundtruth_lists[fields.keypoints] = groundtruth_keypoints_list
        self._groundtruth_dirty = True

    def provide_window(self, window_boxes_list, window_classes_list):
        """model.py:269-283: per image [Nw,4] normalised window boxes and [Nw,K+1] soft labels."""
        self._window_lists[fields.boxes] = window_boxes_list
        self._window_lists[fields.classes] = window_classes_list
        self._groundtruth_dirty = True

    def provide_edgemask(self, groundtruth_edgemask_list):
        """model.py:285-286: per image [2,64,64] = (foreground mask, weight map)."""
        self._edgemask_lists[fields.edgemask] = groundtruth_edgemask_list
        self._groundtruth_dirty = True
