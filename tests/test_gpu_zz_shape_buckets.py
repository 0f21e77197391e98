"""shape_buckets.ShapeBucketTrainer on a device: two image shapes alternate over ONE model (shared parameters, momenta and
global_step; one workspace + CUDA graphs per shape) and reproduce the losses of plain single-shape trainers that hand the
optimizer state from one to the next.

Round-2 note (tools/debug_buckets.py on a B200): the first version of this test initialised the models WITHOUT
normalised batch-norm statistics -- activations overflowed into losses of ~1e5, and the run-to-run noise of the fp32
atomics in the backward pass, amplified by two chaotic update steps, made even two runs of the PLAIN trainer differ by
1.5e-2 in `edgemask_loss` at the third step (0.6647 vs 0.6544), above this test's 2e-3 bar.  The bucket trainer itself
was not at fault.  The models now start from `randomize_bn` statistics (losses O(1), as in tests/test_gpu_train_step.py)
and the tolerance is tied to the measured noise of two plain runs."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]      # first run on a device: never hang the suite


@pytest.mark.parametrize("graph", [False, True])
def test_two_shapes_share_one_model(graph):
    import test_gpu_train_step as T
    from helpers import load_config, randomize_bn
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from mtl_ssl_b200.shape_buckets import ShapeBucketTrainer
    from mtl_ssl_b200.trainer import Trainer
    cfg = load_config("model12.config", T.SMALL)
    K, M = cfg.model.faster_rcnn.num_classes, cfg.model.faster_rcnn.first_stage_max_proposals
    shapes = [(224, 320), (224, 288), (224, 320)]
    # A tiny learning rate: the fp32 atomics of the backward pass make two runs' weights differ in the last bits, and at
    # 1e-5 that was enough to flip a near-tie in the proposal selection of the third step now and then (one of three
    # runs on a B200: `second_stage_localization_loss` 0.893 or 0.948 in the PLAIN path as well as in the bucket path).
    # What the test checks -- which trainer's buffers and whose optimizer state each step used -- does not need big steps.
    kw = dict(gmax=8, learning_rate=1e-7)

    def batch(model, i, hw):
        ex = synthetic.make_batch(60 + i, 1, hw[0], hw[1], K, max_boxes=4, num_windows=16)
        ky = synthetic.make_sampler_keys(70 + i, 1, model.num_kept_anchors((1, hw[0], hw[1], 3)), M)
        return ex, ky

    def fresh_model():
        model = model_builder.build(cfg.model, True, device="cuda", seed=0)
        model.param_store.load_state_dict(randomize_bn(model.param_store.state_dict(), 0))
        return model

    # ---- reference: a fresh model + plain eager trainer per step, parameters and momenta carried over by hand
    def plain():
        out, state = [], None
        for i, hw in enumerate(shapes):
            model = fresh_model()
            st = model.param_store
            if state is not None:
                st.w.copy_(state[0]); st.m.copy_(state[1])
                st.fold()
            tr = Trainer(model, None, hw[0], hw[1], 1, use_cuda_graph=False, **kw)
            tr.global_step = i
            ex, ky = batch(model, i, hw)
            out.append(tr.step(tr.host_arrays(ex, ky)))
            torch.cuda.synchronize()
            state = (st.w.clone(), st.m.clone())
            del tr, model
        return out, state

    want, state = plain()
    again, state2 = plain()             # run-to-run noise of the plain path (fp32 atomics in the backward pass)
    noise = {k: max(abs(a[k] - b[k]) for a, b in zip(want, again)) for k in want[0]}
    # ---- one model, one bucket per shape
    model = fresh_model()
    bt = ShapeBucketTrainer(model, None, batch_size=1, max_buckets=4, use_cuda_graph=graph, **kw)
    got = []
    for i, hw in enumerate(shapes):
        ex, ky = batch(model, i, hw)
        r = bt.step_pipelined(bt.host_arrays(ex, ky))
        if r is not None:
            got.append(r)
    got.append(bt.flush())
    torch.cuda.synchronize()
    assert len(got) == 3 and bt.global_step == 3 and sorted(bt.buckets) == [(224, 288), (224, 320)]
    assert all(abs(v) < 1e3 for v in want[0].values()), want[0]          # a well-conditioned start (see the note above)
    for i, (a, b) in enumerate(zip(want, got)):
        for k in a:
            assert np.isfinite(b[k]) and abs(a[k] - b[k]) <= max(2e-3 * max(1.0, abs(a[k])), 4 * noise[k]), \
                ("step %d" % i, k, a[k], b[k], noise[k], want, got)
    assert abs(got[0]["total_loss"] - got[1]["total_loss"]) > 1e-4
    torch.testing.assert_close(model.param_store.w, state[0], rtol=0, atol=2e-7)
    # (at this learning rate most fp32 weights do not move by an ulp: the momenta below, which accumulate all three
    # steps' clipped gradients across the two trainers, are what shows that the optimizer state was handed over)
    assert float(model.param_store.m.abs().max()) > 1e-3

    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-20))

    # the momenta are the (clipped) gradients: compared as a whole against the noise two plain runs show
    assert rel(model.param_store.m, state[1]) <= max(4 * rel(state2[1], state[1]), 1e-3), \
        (rel(model.param_store.m, state[1]), rel(state2[1], state[1]))
