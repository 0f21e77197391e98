"""Developer experiment (GPU, torchrun): bench.py's data-parallel step with the gradient exchange patched out
(DP_AB=noexchange: same schedule, no NCCL kernels -- timing only, the replicas diverge) to separate what the collectives
cost from what the several-replica schedule costs.  Other variants are plain environment knobs (NCCL_*, MTL_*).

    DP_AB=noexchange python -m torch.distributed.run --nproc-per-node 2 ... tools/dp_ab.py --gpus 2 --skip-cpu
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

if os.environ.get("DP_AB", "").startswith("noexchange"):
    from mtl_ssl_b200 import parallel, trainer

    class _Done(object):
        def wait(self):
            return True

    real = parallel.allreduce_gradients
    which = os.environ.get("DP_AB_ONLY", "")          # "heads" / "trunk": skip only that bucket's exchange
    state = {}

    def no_exchange(flat, world, group=None, async_op=False):
        if which:
            if "lo" not in state:
                import torch  # noqa: F401
                h = state["model"].gradient_buckets()[0]
                state["lo"], state["hi"] = h.data_ptr(), h.data_ptr() + h.numel() * 4
            is_head = state["lo"] <= flat.data_ptr() < state["hi"]
            if is_head != (which == "heads"):
                return real(flat, world, group, async_op)
        return _Done() if async_op else flat

    if which:
        init = trainer.Trainer.__init__

        def init2(self, model, *a, **k):
            state["model"] = model
            init(self, model, *a, **k)
        trainer.Trainer.__init__ = init2

    parallel.allreduce_gradients = no_exchange
    trainer.allreduce_gradients = no_exchange
    if os.environ.get("DP_AB") == "noexchange2":        # ... and no exchange of the logged losses either
        trainer.average_losses = lambda t, w, g=None: t

if __name__ == "__main__":
    bench.main()
