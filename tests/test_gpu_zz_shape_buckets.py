"""shape_buckets.ShapeBucketTrainer on a device: two image shapes alternate over ONE model (shared parameters, momenta and
global_step; one workspace + CUDA graphs per shape) and reproduce the losses of plain single-shape trainers that hand the
optimizer state from one to the next.  Written at the end of round 1 after the GPU budget was spent: this file sorts last
so that its first run on a device cannot mask the established parity tests."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]      # first run on a device: never hang the suite


@pytest.mark.parametrize("graph", [False, True])
def test_two_shapes_share_one_model(graph):
    import test_gpu_train_step as T
    from helpers import load_config
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from mtl_ssl_b200.shape_buckets import ShapeBucketTrainer
    from mtl_ssl_b200.trainer import Trainer
    cfg = load_config("model12.config", T.SMALL)
    K, M = cfg.model.faster_rcnn.num_classes, cfg.model.faster_rcnn.first_stage_max_proposals
    shapes = [(224, 320), (224, 288), (224, 320)]
    kw = dict(gmax=8, learning_rate=1e-5)

    def batch(model, i, hw):
        ex = synthetic.make_batch(60 + i, 1, hw[0], hw[1], K, max_boxes=4, num_windows=16)
        ky = synthetic.make_sampler_keys(70 + i, 1, model.num_kept_anchors((1, hw[0], hw[1], 3)), M)
        return ex, ky

    # ---- reference: a fresh model + plain eager trainer per step, parameters and momenta carried over by hand
    want, state = [], None
    for i, hw in enumerate(shapes):
        model = model_builder.build(cfg.model, True, device="cuda", seed=0)
        st = model.param_store
        if state is not None:
            st.w.copy_(state[0]); st.m.copy_(state[1])
            st.fold()
        tr = Trainer(model, None, hw[0], hw[1], 1, use_cuda_graph=False, **kw)
        tr.global_step = i
        ex, ky = batch(model, i, hw)
        want.append(tr.step(tr.host_arrays(ex, ky)))
        torch.cuda.synchronize()
        state = (st.w.clone(), st.m.clone())
        del tr, model
    # ---- one model, one bucket per shape
    model = model_builder.build(cfg.model, True, device="cuda", seed=0)
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
    for a, b in zip(want, got):
        for k in a:
            assert np.isfinite(b[k]) and abs(a[k] - b[k]) <= 2e-3 * max(1.0, abs(a[k])), (k, a[k], b[k])
    assert abs(got[0]["total_loss"] - got[1]["total_loss"]) > 1e-4
    torch.testing.assert_close(model.param_store.w, state[0], rtol=0, atol=1e-5)
    torch.testing.assert_close(model.param_store.m, state[1], rtol=1e-2, atol=1e-3)
