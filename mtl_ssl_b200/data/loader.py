"""Host-side prefetching between the input reader and the trainer: the role of the reference's batch / prefetch
queue threads (/root/reference/object_detection/core/batcher.py, trainer.py:60-97 `create_input_queue`:
batch_queue_capacity, num_batch_queue_threads, prefetch_queue_capacity), as one bounded queue filled by worker
threads.  Decoding (JPEG, text labels), augmentation and the packing of fixed-shape host arrays happen here, off the
thread that drives the GPU; `Trainer.step_pipelined` then overlaps the H2D copy with the running step."""
import queue
import threading

from . import augment, synthetic


class PrefetchLoader(object):
    """Iterates packed host-array dicts (Trainer.host_arrays output), `depth` batches ahead.

    examples: iterable of example dicts; pack: callable(list of examples, sampler keys) -> arrays (Trainer.host_arrays);
    keys_fn: callable(step) -> sampler keys of that step (the explicit shuffle keys that replace tf.random_shuffle).
    Order is preserved (one producer thread; decoding inside `examples` may itself be parallel)."""

    _END = object()

    def __init__(self, examples, pack, keys_fn, batch_size=1, depth=4):
        self._batches = augment.batches(examples, batch_size)
        self._pack, self._keys_fn = pack, keys_fn
        self._q = queue.Queue(maxsize=max(1, depth))
        self._err = None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._fill, daemon=True)
        self._thread.start()

    def _fill(self):
        try:
            for step, group in enumerate(self._batches):
                if self._stop.is_set():
                    return
                item = self._pack(group, self._keys_fn(step))
                while not self._stop.is_set():
                    try:
                        self._q.put(item, timeout=0.1)
                        break
                    except queue.Full:
                        continue
        except BaseException as e:          # surfaced on the consumer side
            self._err = e
        finally:
            while not self._stop.is_set():
                try:
                    self._q.put(self._END, timeout=0.1)
                    break
                except queue.Full:
                    continue

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is self._END:
            if self._err is not None:
                raise self._err
            raise StopIteration
        return item

    def close(self):
        self._stop.set()
        self._thread.join(timeout=5)


def sampler_keys_fn(seed, batch_size, num_anchors, num_proposals, start_step=0, rank=0):
    """Per-step explicit sampler keys (uniform [0,1) per anchor / proposal), seeded by (seed, rank, global step).
    `start_step`: the trainer's global_step when the loader is (re)built -- a run resumed from a checkpoint continues
    the key sequence instead of replaying the keys of steps 0..k; `rank`: data-parallel replicas draw different keys."""
    def fn(step):
        return synthetic.make_sampler_keys((seed * 1000003 + rank) * 1000033 + start_step + step, batch_size,
                                           num_anchors, num_proposals)
    return fn


def train_loop(trainer, loader, num_steps=None, log_every=0, log=print, checkpoint_prefix=None, save_every=0,
               check_numerics=True, jsonl_path=None, images_per_step=None):
    """Minimal caller of the hot path (the loop of trainer.py:379-429 / slim.learning.train without TF summaries): feeds
    the loader's batches to Trainer.step_pipelined and returns the list of per-step loss dicts.
    * `checkpoint_prefix`: the training state (variables, momentum slots, global_step;
      utils/checkpoint_io.save_training_checkpoint) is written as `<prefix>-<global_step>` every `save_every` steps and at
      the end, the role of the Saver in slim.learning.train;
    * `check_numerics`: a non-finite loss raises FloatingPointError('LossTensor is inf or nan.'), the
      `tf.check_numerics` of trainer.py:207-209, on the loss vector the step already copied to the host;
    * `jsonl_path`: one JSON line per step with the reference's loss keys, `sec/batch` and `instances/sec`
      (the console line of learning.py:509-519; `images_per_step` defaults to the trainer's global batch)."""
    import json
    import math
    import time
    out = []
    if images_per_step is None:
        images_per_step = getattr(trainer, "B", 1) * getattr(trainer, "world_size", 1)
    jf = open(jsonl_path, "a") if jsonl_path else None
    last = [time.perf_counter()]

    def record(r):
        now = time.perf_counter()
        dt, last[0] = now - last[0], now
        if check_numerics and not all(math.isfinite(float(v)) for v in r.values()):
            raise FloatingPointError("LossTensor is inf or nan. (step %d: %s)" % (len(out) + 1, r))
        out.append(r)
        if jf is not None:
            row = dict(r, step=len(out), **{"sec/batch": dt, "instances/sec": images_per_step / dt if dt > 0 else None})
            jf.write(json.dumps(row) + "\n")
            jf.flush()
        if log_every and len(out) % log_every == 0:
            log("step %d: total_loss %.4f (%.3f sec/batch)" % (len(out), r["total_loss"], dt))

    def save():
        import torch
        from ..utils import checkpoint_io
        r = trainer.flush()                       # the pipelined step in flight belongs to the saved state
        if r is not None:
            record(r)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        checkpoint_io.save_training_checkpoint(trainer, "%s-%d" % (checkpoint_prefix, trainer.global_step))

    try:
        for step, arrays in enumerate(loader):
            if num_steps is not None and step >= num_steps:
                break
            r = trainer.step_pipelined(arrays)
            if r is not None:
                record(r)
            if checkpoint_prefix and save_every and (step + 1) % save_every == 0:
                save()
        r = trainer.flush()
        if r is not None:
            record(r)
        if checkpoint_prefix:
            save()
    finally:
        if jf is not None:
            jf.close()
    return out
