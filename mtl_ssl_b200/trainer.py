"""One synchronous data-parallel training step on B200s.

Replaces the per-clone loss graph + optimizer of /root/reference/object_detection/trainer.py:157-214
(`_create_losses`), :379-429 (gradient post-processing + apply) and the clone deployment of
slim/deployment/model_deploy.py:143-307: one process per GPU; every rank runs forward, loss and the
explicit backward on its shard of the batch, NCCL all-reduces sum contiguous buckets of the flat gradient
arena as the backward pass finishes them (replacing the CPU `tf.add_n`, model_deploy.py:414-444), then
every rank applies the identical per-tensor-clip + momentum update.  Loss scaling follows the reference:
task losses / num_replicas, L2 regularisation once (model_deploy.py:221-225, :296).

The step is a fixed list of ~310 kernel launches over persistent buffers; after a warm-up it is captured
into CUDA graphs, one per stage of `_run_step_deferred` (first stage, second stage + its backward, trunk
backward, the deferred second-stage weight gradients, the two optimizer passes) and replayed, so the
launches cost no Python time.  Inputs arrive through pinned host staging buffers (one H2D per step); the
only D2H is the 9-float loss vector.
"""
import numpy as np
import torch

from .meta_architectures.faster_rcnn_meta_arch import LOSS_KEYS
from .parallel import allreduce_gradients, average_losses, data_parallel_scale
from .utils import learning_schedules


def pack_groundtruth(examples, num_classes, height, width, gmax):
    """Host-side packing of a list of examples (data/synthetic.py format) into fixed-shape arrays:
    boxes normalised -> absolute float32 (fmA:1218-1266), class index with background 0."""
    B, K1 = len(examples), num_classes + 1
    out = dict(gt=np.zeros((B, gmax, 4), np.float32), num_gt=np.zeros((B,), np.int32),
               gt_cls=np.zeros((B, gmax), np.int32), gt_close=np.zeros((B, gmax, K1), np.float32))
    scale = np.array([height, width, height, width], np.float32)
    for b, ex in enumerate(examples):
        bx = np.asarray(ex["groundtruth_boxes"], np.float32).reshape(-1, 4)
        g = bx.shape[0]
        if g > gmax:
            raise ValueError("image has %d groundtruth boxes, static limit is %d" % (g, gmax))
        if g and bx.max() > 1.01:
            raise ValueError("maximum box coordinate value is larger than 1.01")
        out["num_gt"][b] = g
        out["gt"][b, :g] = bx * scale
        if g:                                    # (an image without boxes has nothing to check or copy)
            oh = np.asarray(ex["groundtruth_classes"], np.float32).reshape(g, -1)
            if not np.all((oh.sum(1) == 1) & (oh.max(1) == 1)):
                raise ValueError("groundtruth classes must be one-hot on the B200 path")
            out["gt_cls"][b, :g] = oh.argmax(1) + 1
        cl = ex.get("groundtruth_closeness")
        if cl is not None and g:
            out["gt_close"][b, :g] = np.asarray(cl, np.float32).reshape(g, K1)
    if examples and examples[0].get("window_boxes") is not None:
        out["win_boxes"] = np.stack([np.asarray(e["window_boxes"], np.float32).reshape(-1, 4) for e in examples])
        out["win_cls"] = np.stack([np.asarray(e["window_classes"], np.float32).reshape(-1, K1) for e in examples])
    if examples and examples[0].get("groundtruth_edgemask") is not None:
        out["edgemask"] = np.stack([np.asarray(e["groundtruth_edgemask"], np.float32) for e in examples])
    return out


class StaticInputs(object):
    """Pinned host staging + device buffers of fixed shape; `load()` is the step's only H2D.
    Two extra (pinned host, device staging) slots serve the software-pipelined step: the next batch is
    staged and copied on a side stream while the current step computes, then moved device-to-device
    into the fixed buffers the CUDA graphs read."""

    def __init__(self, device, arrays):
        self.host, self.dev = {}, {}
        self.host_slots, self.stage_slots = [{}, {}], [{}, {}]
        cuda = torch.cuda.is_available()
        for k, a in arrays.items():
            t = torch.from_numpy(np.ascontiguousarray(a))
            self.host[k] = t.pin_memory() if cuda else t
            self.dev[k] = torch.empty_like(t, device=device)
        self.nbytes = sum(t.numel() * t.element_size() for t in self.host.values())
        if cuda:        # allocate the pipelining slots up front: cudaHostAlloc costs milliseconds
            for hs, ds in zip(self.host_slots, self.stage_slots):
                for k, t in self.host.items():
                    hs[k] = torch.empty_like(t).pin_memory()
                    ds[k] = torch.empty_like(self.dev[k])

    def load(self, arrays):
        for k, a in arrays.items():
            self.host[k].copy_(torch.from_numpy(np.ascontiguousarray(a)))
            self.dev[k].copy_(self.host[k], non_blocking=True)

    def stage_async(self, slot, arrays, stream):
        """host copy into pinned slot `slot`, then H2D into the device staging slot on `stream`."""
        hs, ds = self.host_slots[slot], self.stage_slots[slot]
        if not hs:
            for k, t in self.host.items():
                hs[k] = torch.empty_like(t).pin_memory()
                ds[k] = torch.empty_like(self.dev[k])
        for k, a in arrays.items():
            hs[k].copy_(torch.from_numpy(np.ascontiguousarray(a)))
        with torch.cuda.stream(stream):
            for k in arrays:
                ds[k].copy_(hs[k], non_blocking=True)

    def commit(self, slot):
        """device staging slot -> the fixed input buffers (current stream)."""
        for k, t in self.stage_slots[slot].items():
            self.dev[k].copy_(t, non_blocking=True)


class Trainer(object):
    def __init__(self, model, train_config=None, height=600, width=1000, batch_size=1, gmax=64,
                 use_cuda_graph=True, world_size=1, process_group=None, learning_rate=None, momentum=0.9,
                 clip_norm=10.0):
        self.model = model
        self.H, self.W, self.B = height, width, batch_size
        # (height, width) is the size of the images as they arrive; `model.preprocess` resizes them on the device with
        # the config's image resizer (fmA:479-505), and everything downstream -- anchors, absolute ground-truth boxes,
        # the clip window -- lives in pixels of the RESIZED image (fmA:1218-1266 scales by the preprocessed shape)
        static_size = getattr(getattr(model, "_image_resizer_fn", None), "static_size", None)
        self.Hr, self.Wr = (height, width) if static_size is None else tuple(int(v) for v in static_size(height, width))
        self.gmax = gmax
        self.world_size = world_size
        self.pg = process_group
        self.use_graph = use_cuda_graph
        self.global_step = 0
        self.device = model.device
        if train_config is not None:
            self._apply_gradient_knobs(train_config)
            self.lr_fn, momentum = learning_schedules.from_optimizer_config(train_config.optimizer)
            clip_norm = train_config.gradient_clipping_by_norm
        else:
            lr = 0.001 if learning_rate is None else learning_rate
            self.lr_fn = lambda step: lr
        self.momentum, self.clip_norm = momentum, clip_norm
        self._hyper_host = torch.zeros(4).pin_memory() if torch.cuda.is_available() else torch.zeros(4)
        self.inputs = None
        self.graph_fb = None
        self.graph_opt = None
        self._loss_host = (torch.zeros(9).pin_memory() if torch.cuda.is_available() else torch.zeros(9))
        self._loss_dev = torch.zeros(9, device=self.device)
        self.launches_per_step = None
        self.overlap_optimizer = True       # update the head bucket under the trunk backward (and its all-reduce)
        # software-pipelined step (step_pipelined): two slots of hyper-parameter / loss staging, a copy stream
        pin = torch.cuda.is_available()
        self._hyper_slots = [torch.zeros(4).pin_memory() if pin else torch.zeros(4) for _ in range(2)]
        self._loss_slots = [torch.zeros(9).pin_memory() if pin else torch.zeros(9) for _ in range(2)]
        self._copy_stream = None
        self._pipe_i = 0
        self._pending = None
        # frozen-prefix pipelining (set up by _setup_prefix once the first batch is bound)
        self.pipeline_prefix = True
        self._prefix_cur = None         # buffer the step graph reads
        self._prefix_next = None        # buffer the look-ahead pass writes
        self._image_next = None
        self._prefix_stream = None
        self._prefix_ready = None       # event: look-ahead prefix computed
        self._prefix_taken = None       # event: it has been copied into _prefix_cur
        self.graph_prefix = None
        self.graph_prefix_next = None
        self._prefix_primed = False
        self.graph_opt_heads = None
        self._opt_stream = None
        self._h2d_done = None           # event: the last step()'s copies out of the pinned staging buffers have run
        # Deferred head update (see _run_step_deferred): the weight-gradient GEMMs of the second-stage tails / heads, the
        # all-reduce of that gradient bucket and its clip + momentum update run underneath the NEXT step's trunk
        # forward pass + proposal chain (a latency-bound stretch that leaves most SMs idle at batch 1).
        import os
        self.defer_heads = os.environ.get("MTL_NO_DEFER_HEADS") is None
        self._deferred_ctas = int(os.environ.get("MTL_DEFERRED_CTAS", "48"))
        # several replicas: the deferred weight gradients run after the trunk update as with one replica (measured on
        # 2 B200s, same box: 7.52-7.60 ms/step against 7.61-7.67 with them right after the second-stage backward,
        # MTL_HW_EARLY=1, which also needs 72 CTAs to finish in time)
        self._hw_late = os.environ.get("MTL_HW_EARLY") is None
        # several replicas, late placement: the head bucket is exchanged in this many pieces, each as soon as its
        # weight-gradient GEMMs are done (the next piece's GEMMs hide the exchange)
        self._head_pieces = max(1, int(os.environ.get("MTL_HEAD_PIECES", "3"))) if world_size > 1 else 1
        self.graph_hw_pieces = None
        self._no_split = os.environ.get("MTL_NO_SPLIT_TRUNK") is not None
        self._heads_pending = False     # a head update is waiting for the next step (or finish())
        self._head_stats_valid = False  # the head tensors' squared norms (regularisation loss) match the current weights
        model.param_store.post_load_hooks.append(self._weights_replaced)
        self._pd = None
        self.graph_fa = None
        self.graph_hw = None
        self.graph_ft = None
        self.graph_ft2 = None

    def _apply_gradient_knobs(self, tc):
        """trainer.py:387-410 of the reference: `grad_multiplier` / `divide_grad_by_batch` scale every gradient,
        `bias_grad_multiplier` additionally those of variables matching '.*/biases', `freeze_variables` (regular
        expressions, `re.match` on the variable name as variables_helper.filter_variables does) drops variables from
        the update; all before the per-tensor clip.  They become per-tensor entries of the optimizer table."""
        import re
        gm = float(getattr(tc, "grad_multiplier", 0.0) or 0.0)
        div = bool(getattr(tc, "divide_grad_by_batch", False))
        bm = float(getattr(tc, "bias_grad_multiplier", 0.0) or 0.0)
        freeze = [r for r in (getattr(tc, "freeze_variables", None) or []) if r]
        if not (gm or div or bm or freeze):
            return
        base = (gm if gm else 1.0) / (float(tc.batch_size) if div else 1.0)

        def mult(name):
            return base * (bm if bm and re.match(".*/biases", name) else 1.0)

        def frozen(name):
            return any(re.match(r, name) for r in freeze)

        self.model.param_store.set_gradient_policy(mult, frozen)

    # ------------------------------------------------------------------ one step
    def _bind(self, arrays):
        if self.inputs is None:
            self.inputs = StaticInputs(self.device, arrays)
        self.inputs.load(arrays)
        return self._attach()

    def _attach(self):
        d = self.inputs.dev
        m = self.model
        gt = dict(gt=d["gt"], num_gt=d["num_gt"], gt_cls=d["gt_cls"], gt_close=d["gt_close"], gmax=self.gmax,
                  B=self.B)
        for k in ("win_boxes", "win_cls", "edgemask"):
            if k in d:
                gt[k] = d[k]
        m._gt, m._gt_shape, m._groundtruth_dirty = gt, (self.B, self.Hr, self.Wr, 3), False
        m.provide_sampler_keys(d["keys1"], d["keys2"])
        return d["image"]

    # The frozen leading layers (conv1 + block1 under freeze_layer 'block1') depend on the image only: their
    # activations for the NEXT batch are computed on a side stream while the current step is in its backward
    # pass ("prefix"), into their own buffers, and handed over by one device-to-device copy at the step start.
    def _prefix(self, image, tag="s1"):
        m = self.model
        if tag == "s1":
            return m.frozen_prefix(m.preprocess(image), tag)
        # look-ahead pass: runs beside the current step's main chain, on a share of the SMs
        from . import ops_conv
        import os
        old, ops_conv.CTA_CAP = ops_conv.CTA_CAP, int(os.environ.get("MTL_PREFIX_CTAS", "0"))
        try:
            return m.frozen_prefix(m.preprocess(image), tag)
        finally:
            ops_conv.CTA_CAP = old

    def _forward_backward(self, image, prefix=None):
        m = self.model
        mtl = m._mtl
        if prefix is None and self._prefix_cur is not None:
            prefix = self._prefix_cur
        pd = m.predict(m.preprocess(image), prefix=prefix) if prefix is not None else m.predict(m.preprocess(image))
        if mtl is not None and mtl.window:
            pd = m.predict_with_window(pd)
        if mtl is not None and mtl.edgemask:
            pd = m.predict_edgemask(pd)
        if mtl is not None and mtl.refine:
            pd = m.predict_with_mtl_results(pd)
        m.loss(pd)
        if self.world_size > 1:
            m.backward(pd, part="heads")
            return pd
        if not self.overlap_optimizer:
            m.backward(pd)
            return pd
        # single replica: the second-stage / aux-head gradients are final here.  Their clip + momentum
        # update (64 % of the parameters, HBM bound) runs on its own stream underneath the trunk backward,
        # which is a chain of short latency-bound GEMMs that leaves HBM and half of the SMs idle.
        from .nets.layers import Concurrency
        m.backward(pd, part="heads_async")
        cur = torch.cuda.current_stream()
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream()
        self._opt_stream.wait_stream(cur)
        for s in (Concurrency._streams or []):
            self._opt_stream.wait_stream(s)          # head weight-gradient GEMMs run on the side streams
        with torch.cuda.stream(self._opt_stream):
            self._optimize_heads()
        m.backward(None, part="trunk")
        cur.wait_stream(self._opt_stream)
        return pd

    # ------------------------------------------------------------------ deferred-head schedule
    def _deferred(self):
        return bool(self.defer_heads and self.overlap_optimizer and
                    getattr(self.model, "supports_deferred_heads", False))

    def _stage_a(self, image, prefix=None):
        """Trunk forward, RPN head, proposal chain: touches first-stage variables only."""
        m = self.model
        if prefix is None and self._prefix_cur is not None:
            prefix = self._prefix_cur
        pre = m.preprocess(image)
        self._pd = m.predict_first_stage(pre, prefix=prefix) if prefix is not None else m.predict_first_stage(pre)

    def _stage_b(self):
        """Second-stage forward, the eight losses, the second-stage half of the backward pass; its weight-gradient
        GEMMs are only collected (model.flush_head_wgrads runs them, see _stage_c)."""
        m = self.model
        mtl = m._mtl
        pd = m.predict_second_stage(self._pd)
        if mtl is not None and mtl.window:
            pd = m.predict_with_window(pd)
        if mtl is not None and mtl.edgemask:
            pd = m.predict_edgemask(pd)
        if mtl is not None and mtl.refine:
            pd = m.predict_with_mtl_results(pd)
        m.loss(pd)
        m.group_head_wgrads = m.defer_head_wgrads = True
        try:
            m.backward(pd, part="heads")
        finally:
            m.group_head_wgrads = m.defer_head_wgrads = False
        self._pd = pd
        return pd

    def _split_trunk(self):
        """Several replicas and a trunk whose backward pass can be cut in two (the ResNets): its gradient bucket is
        exchanged in two pieces, the first while the second half still computes."""
        return (self.world_size > 1 and not self._no_split and
                bool(getattr(self.model._feature_extractor, "deferred_wgrad_safe", False)))

    def _stage_t(self):
        """RPN + trunk backward.  Split mode: the RPN and the later trunk units only (_stage_t2 does the rest)."""
        self.model.backward(None, part="trunk_hi" if self._split_trunk() else "trunk")

    def _stage_t2(self):
        self.model.backward(None, part="trunk_lo")

    def _stage_c(self, piece=None):
        """The collected second-stage weight gradients, as grouped launches sized to leave the latency-bound trunk
        chains beside them their SMs (piece k: those of the k-th piece of the head bucket only)."""
        self.model.flush_head_wgrads(max_ctas=self._deferred_ctas, piece=piece, pieces=self._head_pieces)

    def _piecewise_heads(self):
        return self.world_size > 1 and self._hw_late and self._head_pieces > 1

    def _optimize_heads_deferred(self):
        st = self.model.param_store
        gs = data_parallel_scale(self.world_size)
        t0, _ = self.model.head_tensor_range()
        st.stats_range(t0, st.num_tensors, gs)
        # the regularisation loss of the NEXT step is made of the squared norms of the weights that step computes with:
        # the update pass leaves the head tensors' new norms behind (the trunk's are refreshed by that step's own trunk
        # update); a separate third pass over the weights cost 0.07 ms on the side chain the next second stage waits for
        st.apply_range(t0, st.num_tensors, gs, hyper=st.hyper_heads, refresh_norms=True)

    def _weights_replaced(self):
        self._head_stats_valid = False

    def finish(self):
        """Wait for a head update that is still in flight on the side stream (before reading or saving weights,
        evaluating, or handing the model to another trainer).  No-op when nothing is pending."""
        if self._heads_pending:
            torch.cuda.current_stream().wait_stream(self._opt_stream)
            self._heads_pending = False

    def _run_step_deferred(self, lookahead, defer):
        """main stream:  A trunk forward + proposals -> [wait for the previous head update] -> B second stage, losses,
                         second-stage backward -> T trunk backward (+ exchange of the trunk bucket in two pieces) ->
                         D trunk update
        side stream:     C the second-stage weight-gradient GEMMs (grouped launches on a share of the SMs), the exchange
                         of the second-stage bucket and its update -- after D, i.e. underneath the NEXT step's A, a
                         latency-bound chain that leaves most SMs idle at batch 1 (several replicas: the weight
                         gradients start right after B, underneath T).
        Nothing on the main stream waits for the side stream until the next step needs a second-stage weight."""
        graph = self.use_graph
        cur = torch.cuda.current_stream()
        image = self.inputs.dev["image"]
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream()
        side = self._opt_stream
        if self._prefix_cur is not None and not lookahead:
            self.graph_prefix.replay() if graph else self._prefix(image)
        if not self._heads_pending and not self._head_stats_valid:
            # first step (or the weights were replaced): no head update precedes this step, so nothing has computed
            # the head tensors' squared norms that the regularisation loss sums
            st = self.model.param_store
            st.stats_range(self.model.head_tensor_range()[0], st.num_tensors, data_parallel_scale(self.world_size))
            self._head_stats_valid = True
        self.graph_fa.replay() if graph else self._stage_a(image)
        if self._heads_pending:
            cur.wait_stream(side)               # every second-stage weight is final from here on
            self._heads_pending = False
        self.graph_fb.replay() if graph else self._stage_b()
        early = self.world_size > 1 and not self._hw_late
        if early:
            # several replicas: the side stream also has the head bucket's exchange to fit in before the next step's
            # second stage, so the weight gradients start now, underneath the trunk backward
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self.graph_hw.replay() if graph else self._stage_c()
        self.graph_ft.replay() if graph else self._stage_t()
        if self._split_trunk():
            # trunk bucket in two pieces, in the order the backward pass finishes them: only the second is exposed
            _, b_hi, b_lo = self.model.gradient_buckets3()
            w_hi = allreduce_gradients(b_hi, self.world_size, self.pg, async_op=True)
            self.graph_ft2.replay() if graph else self._stage_t2()
            w_lo = allreduce_gradients(b_lo, self.world_size, self.pg, async_op=True)
            w_hi.wait()
            w_lo.wait()
        elif self.world_size > 1:
            allreduce_gradients(self.model.gradient_buckets()[1], self.world_size, self.pg)
        self.graph_opt.replay() if graph else self._optimize()
        # (the head bucket is exchanged AFTER the trunk pieces were issued: collectives of one communicator run in
        # issue order, and the trunk's are the ones the main stream waits for)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if self._piecewise_heads():
                works = []
                for k, (g, _) in enumerate(self.model.head_exchange_pieces(self._head_pieces)):
                    self.graph_hw_pieces[k].replay() if graph else self._stage_c(k)
                    works.append(allreduce_gradients(g, self.world_size, self.pg, async_op=True))
                for w in works:
                    w.wait()
            elif not early:
                # one replica: measured best with all of it underneath the next step's trunk forward + proposal chain
                # (same box, B200: 7.30 ms/step against 7.58 with the weight gradients under the trunk backward and
                # 7.66-7.69 with the round-1 schedule)
                self.graph_hw.replay() if graph else self._stage_c()
            if self.world_size > 1 and not self._piecewise_heads():
                allreduce_gradients(self.model.gradient_buckets()[0], self.world_size, self.pg)
            if self.world_size > 1:
                average_losses(self._loss_dev[:8], self.world_size, self.pg)      # logged values only: off the main stream
            self.graph_opt_heads.replay() if graph else self._optimize_heads_deferred()
        self._heads_pending = True
        self._head_stats_valid = True
        if not defer:
            self.finish()

    def eager_pass(self, image=None):
        """One whole step body launched eagerly on the current stream(s), no gradient exchange (bench.py: launch
        count and per-launch timings)."""
        image = self.inputs.dev["image"] if image is None else image
        self._prefix(image)
        if self._deferred():
            self._stage_a(image)
            self._stage_b()
            self._stage_c()
            self._stage_t()
            if self._split_trunk():
                self._stage_t2()
            self._optimize()
            self._optimize_heads_deferred()
            return
        self._forward_backward(image)
        if self.world_size > 1:
            if self.overlap_optimizer:
                self._optimize_heads()
            self._backward_trunk()
        self._optimize()

    def _backward_trunk(self):
        self.model.backward(None, part="trunk")

    def _optimize_heads(self):
        """clip + momentum update of the second-stage / aux-head tensors [t0, t1) (their gradients are final
        once backward(part="heads") has run / their bucket has been all-reduced)."""
        st = self.model.param_store
        gs = data_parallel_scale(self.world_size)
        t0, t1 = self.model.head_tensor_range()
        # ... plus the dead stage-1 block4 copy behind them (trap T4: no task gradient, L2 decay only)
        st.stats_range(t0, st.num_tensors, gs)
        st.apply_range(t0, st.num_tensors, gs)

    def _optimize(self):
        st = self.model.param_store
        gs = data_parallel_scale(self.world_size)
        if self.overlap_optimizer:
            t0, _ = self.model.head_tensor_range()        # [t0, T) was updated under the trunk backward
            st.stats_range(0, t0, gs)
            st.reg_loss_from_stats()
            st.apply_range(0, t0, gs)
        else:
            st.stats_and_reg_loss(gs)
            st.apply(gs)
        self._loss_dev[:8].copy_(self.model.workspace.bufs["loss/values"])
        self._loss_dev[8:9].copy_(st.reg_loss)
        st.hyper_heads.copy_(st.hyper)          # what a deferred head update of THIS step will use

    def _allreduce(self):
        allreduce_gradients(self.model.param_store.g, self.world_size, self.pg)

    def _setup_prefix(self, image):
        """Allocate the look-ahead buffers (called once, before graph capture)."""
        if not self.pipeline_prefix or self._prefix_cur is not None:
            return
        cur = self._prefix(image, "s1")
        if cur is None:
            self.pipeline_prefix = False
            return
        self._image_next = torch.empty_like(image)
        self._image_next.copy_(image)
        self._prefix_cur = cur
        self._prefix_next = self._prefix(self._image_next, "s1n")
        self._prefix_stream = torch.cuda.Stream()
        self._prefix_ready = torch.cuda.Event()
        self._prefix_taken = torch.cuda.Event()
        self._prefix_taken.record()

    def _lookahead_prefix(self, image_src, after=None):
        """Side stream: frozen prefix of the NEXT batch (`image_src`: device tensor holding its image)."""
        ps = self._prefix_stream
        ps.wait_event(self._prefix_taken)          # the previous look-ahead result has been consumed
        if after is not None:
            ps.wait_event(after)
        with torch.cuda.stream(ps):
            if image_src is not self._image_next:
                self._image_next.copy_(image_src, non_blocking=True)
            self.graph_prefix_next.replay() if self.use_graph else self._prefix(self._image_next, "s1n")
            self._prefix_ready.record(ps)

    def _take_prefix(self):
        """Main stream: adopt the look-ahead result as this step's prefix."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._prefix_ready)
        self._prefix_cur.copy_(self._prefix_next, non_blocking=True)
        self._prefix_taken.record(cur)

    def run_resident_step(self):
        """One step on the inputs already resident in HBM, with the frozen prefix software-pipelined: the prefix of
        the next step (same resident image) is computed underneath this step.  Requires one earlier step."""
        if self._prefix_cur is None:
            return self._run_step_body(defer=True)
        if not self._prefix_primed:
            self._lookahead_prefix(self.inputs.dev["image"])
            self._prefix_primed = True
        self._take_prefix()
        self._lookahead_prefix(self.inputs.dev["image"])
        self._run_step_body(lookahead=True, defer=True)

    def _run_step_body(self, lookahead=False, defer=False):
        """forward + backward + gradient exchange + optimizer.  With several replicas the backward is cut
        in two: the second-stage bucket is all-reduced (NCCL stream) while the trunk half still computes.
        `defer`: leave the head update pending for the next call (deferred-head schedule only)."""
        if self._deferred():
            return self._run_step_deferred(lookahead, defer)
        graph = self.use_graph
        if self._prefix_cur is not None and not lookahead:
            # synchronous: frozen prefix of this batch, then the rest of the step
            self.graph_prefix.replay() if graph else self._prefix(self.inputs.dev["image"])
        if self.world_size == 1:
            self.graph_fb.replay() if graph else self._forward_backward(self.inputs.dev["image"])
            self.graph_opt.replay() if graph else self._optimize()
            return
        b_heads, b_trunk = self.model.gradient_buckets()
        cur = torch.cuda.current_stream()
        self.graph_fb.replay() if graph else self._forward_backward(self.inputs.dev["image"])
        w1 = allreduce_gradients(b_heads, self.world_size, self.pg, async_op=True)
        if self.overlap_optimizer:
            # the head bucket's update follows its all-reduce on a side stream, underneath the trunk backward
            # and the trunk bucket's all-reduce
            if self._opt_stream is None:
                self._opt_stream = torch.cuda.Stream()
            self._opt_stream.wait_stream(cur)
            with torch.cuda.stream(self._opt_stream):
                w1.wait()
                self.graph_opt_heads.replay() if graph else self._optimize_heads()
        self.graph_fb2.replay() if graph else self._backward_trunk()
        w2 = allreduce_gradients(b_trunk, self.world_size, self.pg, async_op=True)
        if self.overlap_optimizer:
            cur.wait_stream(self._opt_stream)
        else:
            w1.wait()
        w2.wait()
        self.graph_opt.replay() if graph else self._optimize()
        average_losses(self._loss_dev[:8], self.world_size, self.pg)

    def host_arrays(self, examples, keys):
        arrays = pack_groundtruth(examples, self.model.num_classes, self.Hr, self.Wr, self.gmax)
        arrays["image"] = np.stack([e["image"] for e in examples]).astype(np.float32)
        arrays["keys1"], arrays["keys2"] = keys
        return arrays

    def step(self, arrays, read_losses=True):
        """arrays: output of host_arrays().  Returns {loss name: float} (+ 'regularization_loss',
        'total_loss') when read_losses."""
        st = self.model.param_store
        # the pinned staging buffers (inputs, hyper-parameters) are reused by every call: with read_losses=False
        # nothing else stops the host from overwriting them while the previous call's H2D copies are still queued
        if self._h2d_done is not None:
            self._h2d_done.synchronize()
        self._hyper_host[0] = float(self.lr_fn(self.global_step))
        self._hyper_host[1] = self.momentum
        self._hyper_host[2] = self.clip_norm if self.clip_norm else 0.0
        st.hyper.copy_(self._hyper_host, non_blocking=True)
        image = self._bind(arrays)
        if torch.cuda.is_available():
            if self._h2d_done is None:
                self._h2d_done = torch.cuda.Event()
            self._h2d_done.record()
        self._setup_prefix(image)
        if self.use_graph and self.graph_fb is None:
            self._capture(image)
        self._run_step_body()
        self.global_step += 1
        if not read_losses:
            return None
        self._loss_host.copy_(self._loss_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.losses_from_host()

    def step_pipelined(self, arrays):
        """The same training step, software-pipelined against the host: the batch is staged into pinned memory
        and copied H2D on a side stream while the previous step still computes, and the call returns the losses
        of the PREVIOUS call (None on the first one; `flush()` returns the last).  Every step still performs its
        own H2D copy of the inputs and D2H read of its loss vector; only the host no longer idles the GPU."""
        if self.inputs is None or (self.use_graph and self.graph_fb is None):
            self._pending = ("done", self.step(arrays))       # first call: allocations + graph capture, synchronous
            return None
        st = self.model.param_store
        slot = self._pipe_i & 1
        self._pipe_i += 1
        cur = torch.cuda.current_stream()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        # slot reuse is safe: the step that used it two calls ago has finished (its losses were resolved below)
        self.inputs.stage_async(slot, arrays, self._copy_stream)
        staged = torch.cuda.Event()
        staged.record(self._copy_stream)
        hh = self._hyper_slots[slot]
        hh[0] = float(self.lr_fn(self.global_step))
        hh[1] = self.momentum
        hh[2] = self.clip_norm if self.clip_norm else 0.0
        cur.wait_event(staged)
        if self._prefix_cur is not None:
            # frozen prefix of THIS batch on the side stream (it overlaps the tail of the previous step, which is
            # still running on the main stream), then adopt it
            self._lookahead_prefix(self.inputs.stage_slots[slot]["image"], after=staged)
            self._prefix_primed = True
            self._take_prefix()
        self.inputs.commit(slot)
        st.hyper.copy_(hh, non_blocking=True)
        self._attach()
        self._run_step_body(lookahead=self._prefix_cur is not None, defer=True)
        self.global_step += 1
        # (several replicas, deferred schedule: the replica-averaged losses are produced on the side stream)
        ls = self._opt_stream if (self.world_size > 1 and self._deferred()) else cur
        with torch.cuda.stream(ls):
            self._loss_slots[slot].copy_(self._loss_dev, non_blocking=True)
        done = torch.cuda.Event()
        done.record(ls)
        prev, self._pending = self._pending, ("event", done, slot)
        return self._resolve(prev)

    def flush(self):
        """Losses of the last step_pipelined() call (waits for it); a pending deferred head update is applied, so the
        weights are final when this returns."""
        prev, self._pending = self._pending, None
        self.finish()
        return self._resolve(prev)

    def _resolve(self, pending):
        if pending is None:
            return None
        if pending[0] == "done":
            return pending[1]
        pending[1].synchronize()
        return self.losses_from_host(self._loss_slots[pending[2]])

    def losses_from_host(self, buf=None):
        v = (self._loss_host if buf is None else buf).tolist()
        out = {k: v[i] for i, k in enumerate(LOSS_KEYS)}
        out["regularization_loss"] = v[8]
        # model_deploy.py:198-236: sum over the clones of (task losses / num_clones) + regularisation; with several
        # replicas the eight task losses were averaged over the replicas on the device (parallel.average_losses), so
        # every rank reports the same, global value
        out["total_loss"] = sum(v[:8]) + v[8]
        return out

    def _capture(self, image):
        """Warm up eagerly (allocates every workspace buffer, fills the anchor cache), then capture."""
        st = self.model.param_store
        snap = (st.w.clone(), st.m.clone(), st.wb.clone())
        deferred = self._deferred()
        for _ in range(2):
            if deferred:
                self._stage_a(image)
                self._stage_b()
                if self._piecewise_heads():     # (plans the grouped launches: not allowed while capturing)
                    for k in range(len(self.model.head_exchange_pieces(self._head_pieces))):
                        self._stage_c(k)
                else:
                    self._stage_c()
                self._stage_t()
                if self._split_trunk():
                    self._stage_t2()
            else:
                self._forward_backward(image)
                if self.world_size > 1:
                    self._backward_trunk()
            st.g.zero_()
        st.w.copy_(snap[0]); st.m.copy_(snap[1]); st.wb.copy_(snap[2])
        torch.cuda.synchronize()
        if self._prefix_cur is not None:
            self.graph_prefix = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_prefix):
                self._prefix(image, "s1")
            self.graph_prefix_next = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_prefix_next):
                self._prefix(self._image_next, "s1n")
        if deferred:
            self.graph_fa = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_fa):
                self._stage_a(image)
            self.graph_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_fb):
                self._stage_b()
            if self._piecewise_heads():
                self.graph_hw_pieces = []
                for k in range(len(self.model.head_exchange_pieces(self._head_pieces))):
                    gk = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gk):
                        self._stage_c(k)
                    self.graph_hw_pieces.append(gk)
            else:
                self.graph_hw = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_hw):
                    self._stage_c()
            self.graph_ft = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_ft):
                self._stage_t()
            if self._split_trunk():
                self.graph_ft2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_ft2):
                    self._stage_t2()
            self.graph_opt_heads = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt_heads):
                self._optimize_heads_deferred()
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt):
                self._optimize()
            torch.cuda.synchronize()
            return
        self.graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_fb):
            self._forward_backward(image)
        if self.world_size > 1:
            self.graph_fb2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_fb2):
                self._backward_trunk()
        if self.world_size > 1 and self.overlap_optimizer:
            self.graph_opt_heads = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt_heads):
                self._optimize_heads()
        self.graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_opt):
            self._optimize()
        torch.cuda.synchronize()
