"""Training on images of more than one size with static-shape trainers.

The reference resizes every image to its own (h, w) (`keep_aspect_ratio_resizer`, core/preprocessor.py:1362-1419) and
runs a dynamic-shape graph at batch 1 per clone (trainer.py:157-214).  `Trainer` is static-shape by design: its workspace,
staging buffers, anchor cache and CUDA graphs belong to one (H, W).  `ShapeBucketTrainer` keeps one such trainer PER
image shape -- each with its own `Workspace` -- over ONE model: parameters, gradients, momenta, the optimizer schedule and
`global_step` are shared, so a run that alternates shapes performs exactly the update sequence of the reference's loop.
The least recently used bucket is dropped (its buffers and graphs freed) beyond `max_buckets`.

Written at the end of round 1 after the GPU budget was spent: exercised on a device only by
tests/test_gpu_zz_shape_buckets.py (sorted last on purpose); the host-side bookkeeping is tested on the CPU.
"""
import collections

from .runtime import Workspace
from .trainer import Trainer


class BucketArrays(dict):
    """`Trainer.host_arrays` output that remembers which image shape it was packed for."""
    hw = None


class ShapeBucketTrainer(object):
    def __init__(self, model, train_config=None, batch_size=1, max_buckets=8, trainer_cls=Trainer,
                 workspace_cls=Workspace, **trainer_kwargs):
        if max_buckets < 1:
            raise ValueError("max_buckets must be >= 1")
        self.model, self.train_config, self.B = model, train_config, batch_size
        self.max_buckets = max_buckets
        self._trainer_cls, self._workspace_cls, self._kw = trainer_cls, workspace_cls, trainer_kwargs
        self.world_size = trainer_kwargs.get("world_size", 1)
        self.buckets = collections.OrderedDict()          # (H, W) -> (trainer, workspace), least recently used first
        self.global_step = 0
        self._last = None
        self.evictions = 0

    # ------------------------------------------------------------------ buckets
    def _activate(self, hw):
        """The trainer of shape `hw` with its workspace installed in the model."""
        hw = (int(hw[0]), int(hw[1]))
        if hw in self.buckets:
            self.buckets.move_to_end(hw)
        else:
            while len(self.buckets) >= self.max_buckets:
                _, (old, _ws) = self.buckets.popitem(last=False)
                if old is self._last:          # max_buckets == 1: `_switch` has already flushed its step
                    if old.flush() is not None:
                        raise RuntimeError("a bucket with a step in flight cannot be evicted")
                    self._last = None
                self.evictions += 1
            ws = self._workspace_cls(self.model.device)
            self.model._ws = ws                            # the trainer's first step allocates into ITS workspace
            tr = self._trainer_cls(self.model, self.train_config, hw[0], hw[1], self.B, **self._kw)
            self.buckets[hw] = (tr, ws)
        tr, ws = self.buckets[hw]
        self.model._ws = ws
        return tr

    def trainer_for(self, height, width):
        return self._activate((height, width))

    # ------------------------------------------------------------------ Trainer API
    def host_arrays(self, examples, keys):
        h, w = examples[0]["image"].shape[:2]
        if any(tuple(e["image"].shape[:2]) != (h, w) for e in examples):
            raise ValueError("one batch holds images of different sizes; batch by shape (the reference trains at "
                             "batch 1 per clone)")
        hw = (int(h), int(w))
        tr = self.buckets[hw][0] if hw in self.buckets else None
        if tr is not None:
            arrays = tr.host_arrays(examples, keys)
        else:                                             # packing needs no device state: do not build the bucket here
            from .trainer import pack_groundtruth
            import numpy as np
            static_size = getattr(getattr(self.model, "_image_resizer_fn", None), "static_size", None)
            hr, wr = (h, w) if static_size is None else static_size(h, w)       # as Trainer.host_arrays packs it
            arrays = pack_groundtruth(examples, self.model.num_classes, hr, wr, self._kw.get("gmax", 64))
            arrays["image"] = np.stack([e["image"] for e in examples]).astype(np.float32)
            arrays["keys1"], arrays["keys2"] = keys
        out = BucketArrays(arrays)
        out.hw = hw
        return out

    def _switch(self, arrays):
        """-> (trainer for these arrays, losses handed over from the bucket that was in flight or None)."""
        hw = getattr(arrays, "hw", None)
        if hw is None:
            hw = tuple(arrays["image"].shape[1:3])
        handed = None
        if self._last is not None and self._last is not self.buckets.get(tuple(hw), (None,))[0]:
            handed = self._last.flush()                   # finish the other shape's step before touching the model
        tr = self._activate(hw)
        tr.global_step = self.global_step
        return tr, handed

    def step(self, arrays, read_losses=True):
        tr, handed = self._switch(arrays)
        if handed is not None:
            raise RuntimeError("step() after step_pipelined() on another shape: call flush() first")
        r = tr.step(arrays, read_losses)
        self.global_step, self._last = tr.global_step, tr
        return r

    def step_pipelined(self, arrays):
        """Losses of the PREVIOUS call (None on the very first), whichever shape that call had."""
        tr, handed = self._switch(arrays)
        r = tr.step_pipelined(arrays)
        self.global_step, self._last = tr.global_step, tr
        return handed if handed is not None else r

    def flush(self):
        return self._last.flush() if self._last is not None else None

    def losses_from_host(self, buf=None):
        return self._last.losses_from_host(buf)
