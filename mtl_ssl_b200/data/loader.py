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


def sampler_keys_fn(seed, batch_size, num_anchors, num_proposals):
    """Per-step explicit sampler keys (uniform [0,1) per anchor / proposal), seeded by (seed, step)."""
    def fn(step):
        return synthetic.make_sampler_keys(seed * 1000003 + step, batch_size, num_anchors, num_proposals)
    return fn


def train_loop(trainer, loader, num_steps=None, log_every=0, log=print, checkpoint_prefix=None, save_every=0):
    """Minimal caller of the hot path (the loop of trainer.py:379-429 without summaries): feeds the loader's batches to
    Trainer.step_pipelined and returns the list of per-step loss dicts.  With `checkpoint_prefix` the training state
    (variables, momentum slots, global_step; utils/checkpoint_io.save_training_checkpoint) is written as
    `<prefix>-<global_step>` every `save_every` steps and at the end, the role of the Saver in slim.learning.train."""
    out = []

    def save():
        import torch
        from ..utils import checkpoint_io
        r = trainer.flush()                       # the pipelined step in flight belongs to the saved state
        if r is not None:
            out.append(r)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        checkpoint_io.save_training_checkpoint(trainer, "%s-%d" % (checkpoint_prefix, trainer.global_step))

    for step, arrays in enumerate(loader):
        if num_steps is not None and step >= num_steps:
            break
        r = trainer.step_pipelined(arrays)
        if r is not None:
            out.append(r)
            if log_every and len(out) % log_every == 0:
                log("step %d: total_loss %.4f" % (len(out), r["total_loss"]))
        if checkpoint_prefix and save_every and (step + 1) % save_every == 0:
            save()
    r = trainer.flush()
    if r is not None:
        out.append(r)
    if checkpoint_prefix:
        save()
    return out
