"""Device-memory runtime: one flat parameter arena, persistent activation workspaces, and the
fused multi-tensor optimizer step.

PyTorch is used here purely as an allocator / stream / NCCL front end.  Layout decisions
(B200, 180 GB HBM3e): all fp32 master weights, gradients and momenta live in three arenas of
identical layout, so that (a) the data-parallel gradient exchange is ONE NCCL all-reduce over
one contiguous buffer (replacing the CPU `tf.add_n` of slim/deployment/model_deploy.py:414-444)
and (b) the optimizer is two HBM-bound passes over contiguous memory.  A fourth arena holds the
bf16 compute copy of every weight with the frozen batch-norm scale folded in.
Variable names are the reference's TF variable names (scope strings of
faster_rcnn_meta_arch.py:431-461) so checkpoint name maps stay valid.
"""
import ctypes
import math

import numpy as np
import torch

from . import ops

ALIGN = 64


class Param(object):
    """One named tensor inside the arenas; `.w`/.g/.m are fp32 views, `.wb` the bf16 compute view."""

    def __init__(self, name, shape, l2, trainable, init, fold=None, grad_mult=1.0, tf_kind=None):
        self.name = name
        # how the reference's TF graph stores this variable when it differs from what the rank implies (see
        # utils/tf_checkpoint.tf_to_native): "fc" = slim.fully_connected [in, out] kept here as a [out,1,1,in] GEMM
        # operand; ("packed_conv", R, S, C) = conv [R,S,C,K] kept as im2col rows [K,1,1,ld] with R*S*C real columns
        self.tf_kind = tf_kind
        self.shape = tuple(int(s) for s in shape)
        self.numel = int(np.prod(self.shape))
        self.l2 = float(l2)
        self.trainable = bool(trainable)
        self.init = init
        self.fold = fold            # BatchNorm object whose scale is folded into the bf16 copy
        self.grad_mult = grad_mult
        self.offset = None
        self.w = self.g = self.m = self.wb = None


class BatchNorm(object):
    """Frozen (inference-mode) slim.batch_norm: y = gamma * (x - mean) / sqrt(var + eps) + beta
    (slim/nets/resnet_utils.py:229-256; is_training=False at fe:139).  Folded into the conv."""

    def __init__(self, scope, channels, eps, scale=True):
        self.scope = scope
        self.channels = channels
        self.eps = eps
        self.has_gamma = scale
        self.gamma = torch.ones(channels)
        self.beta = torch.zeros(channels)
        self.mean = torch.zeros(channels)
        self.var = torch.ones(channels)
        self.scale_off = None
        self.scale = None           # device fp32 [K] view into fold_scales
        self.bias = None            # device fp32 [K]

    def fold_host(self):
        s = self.gamma / torch.sqrt(self.var + self.eps)
        return s.float(), (self.beta - self.mean * s).float()


class ParamStore(object):
    def __init__(self):
        self.params = []
        self.by_name = {}
        self.groups = []            # lists of params laid out contiguously (fused GEMM operands)
        self.bns = []
        self.finalized = False
        self.post_load_hooks = []   # callables run after weights change outside the optimizer

    def add(self, name, shape, l2=0.0, trainable=True, init=("zeros",), fold=None, grad_mult=1.0, tf_kind=None):
        assert not self.finalized
        if name in self.by_name:
            raise ValueError("duplicate variable %s" % name)
        p = Param(name, shape, l2, trainable, init, fold, grad_mult, tf_kind)
        self.params.append(p)
        self.by_name[name] = p
        self.groups.append([p])
        return p

    def add_group(self, specs):
        """specs: list of dicts for add(); the tensors are packed back to back (no padding) so a
        single GEMM can view them as one operand while the optimizer still clips each separately."""
        ps = []
        for s in specs:
            p = self.add(**s)
            self.groups.pop()
            ps.append(p)
        self.groups.append(ps)
        return ps

    def add_bn(self, scope, channels, eps, scale=True):
        bn = BatchNorm(scope, channels, eps, scale)
        self.bns.append(bn)
        return bn

    # ------------------------------------------------------------------ layout + allocation
    def finalize(self, device, seed=0):
        off = 0
        for grp in self.groups:
            off = (off + ALIGN - 1) // ALIGN * ALIGN
            for p in grp:
                p.offset = off
                off += p.numel
        self.total = (off + ALIGN - 1) // ALIGN * ALIGN
        self.device = device
        host = torch.zeros(self.total, dtype=torch.float32)
        gen = torch.Generator().manual_seed(seed)
        for p in self.params:
            host[p.offset:p.offset + p.numel] = _init_tensor(p, gen).reshape(-1)
        self.w = host.to(device)
        self.g = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.m = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.wb = torch.zeros(self.total, dtype=torch.bfloat16, device=device)
        for p in self.params:
            sl = slice(p.offset, p.offset + p.numel)
            p.w = self.w[sl].view(p.shape)
            p.g = self.g[sl].view(p.shape)
            p.m = self.m[sl].view(p.shape)
            p.wb = self.wb[sl].view(p.shape)
        self._upload_bn()
        self._build_tables()
        self.finalized = True
        self.fold()
        return self

    def host_state_dict(self, seed=0):
        """The state `finalize(device, seed)` + `state_dict()` would produce, built on the host without a device
        (same generator, same draw order): lets a CPU checker start from exactly the weights a device run starts from."""
        gen = torch.Generator().manual_seed(seed)
        out = {}
        for p in self.params:
            out[p.name] = _init_tensor(p, gen).float()
        for b in self.bns:
            if b.scope is None:
                continue
            if b.has_gamma:
                out[b.scope + "/gamma"] = b.gamma.clone()
            out[b.scope + "/beta"] = b.beta.clone()
            out[b.scope + "/moving_mean"] = b.mean.clone()
            out[b.scope + "/moving_variance"] = b.var.clone()
        return out

    def group_view(self, ps, rows, cols, arena="wb"):
        """View a contiguous group as one [rows, cols] matrix (rows may include zero padding)."""
        base = getattr(self, arena)
        return base[ps[0].offset:ps[0].offset + rows * cols].view(rows, cols)

    def _upload_bn(self):
        n = sum(b.channels for b in self.bns)
        scales = torch.ones(max(n, 1), dtype=torch.float32)
        biases = torch.zeros(max(n, 1), dtype=torch.float32)
        o = 0
        for b in self.bns:
            s, bi = b.fold_host()
            scales[o:o + b.channels] = s
            biases[o:o + b.channels] = bi
            b.scale_off = o
            o += b.channels
        self.fold_scales = scales.to(self.device)
        self.fold_biases = biases.to(self.device)
        for b in self.bns:
            b.scale = self.fold_scales[b.scale_off:b.scale_off + b.channels]
            b.bias = self.fold_biases[b.scale_off:b.scale_off + b.channels]

    def _build_tables(self):
        chunk = ops.opt_chunk_size()
        T = len(self.params)
        td = (ops.TensorDesc * T)()
        chunks = []
        for i, p in enumerate(self.params):
            td[i].offset = p.offset
            td[i].numel = p.numel
            if p.fold is not None:
                td[i].row_len = p.numel // p.shape[0]
                td[i].scale_off = p.fold.scale_off
            else:
                td[i].row_len = 0
                td[i].scale_off = -1
            td[i].l2_weight = p.l2
            td[i].grad_mult = p.grad_mult
            td[i].trainable = int(p.trainable)
            for s in range(0, p.numel, chunk):
                chunks.append((i, min(chunk, p.numel - s), s))
        self.chunk_start = [0] * (T + 1)        # first chunk of every tensor (chunks are in tensor order)
        for j, (t, _, _) in enumerate(chunks):
            self.chunk_start[t + 1] = j + 1
        for t in range(1, T + 1):
            self.chunk_start[t] = max(self.chunk_start[t], self.chunk_start[t - 1])
        cd = (ops.ChunkDesc * len(chunks))()
        for j, (t, ln, st) in enumerate(chunks):
            cd[j].tensor, cd[j].len, cd[j].start = t, ln, st
        self.num_tensors = T
        self.num_chunks = len(chunks)
        self.td = torch.frombuffer(bytearray(bytes(td)), dtype=torch.uint8).to(self.device)
        self.cd = torch.frombuffer(bytearray(bytes(cd)), dtype=torch.uint8).to(self.device)
        self.stats = torch.zeros(T * 2, dtype=torch.float32, device=self.device)
        # per-chunk partial sums + the chunk table on the device: the per-tensor norms are reduced in a fixed order
        self.partials = torch.zeros(max(len(chunks), 1) * 2, dtype=torch.float32, device=self.device)
        self.chunk_start_dev = torch.tensor(self.chunk_start, dtype=torch.int32, device=self.device)
        self.reg_loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.hyper = torch.zeros(4, dtype=torch.float32, device=self.device)
        # learning rate / momentum / clip norm of the step whose head-bucket update is still pending (trainer.py runs
        # that update underneath the next step's trunk forward pass, after `hyper` already holds the next step's values)
        self.hyper_heads = torch.zeros(4, dtype=torch.float32, device=self.device)

    def set_gradient_policy(self, multiplier=None, frozen=None):
        """Per-variable gradient post-processing of the reference trainer (object_detection/trainer.py:387-410,
        utils/variables_helper.py:58-118), applied by the optimizer kernels from the per-tensor table:
        `multiplier(name) -> float` scales the gradient of the total loss (task + L2 term) before the per-tensor clip;
        `frozen(name) -> bool` removes the variable from the update altogether (no momentum, no decay; it still
        counts in the regularisation loss, trap T4).  Rebuilds the device tables: call before the first step."""
        for p in self.params:
            if multiplier is not None:
                p.grad_mult = float(multiplier(p.name))
            if frozen is not None and frozen(p.name):
                p.trainable = False
        if self.finalized:
            self._build_tables()

    # ------------------------------------------------------------------ kernels
    def fold(self):
        """(Re)build the bf16 compute copy from the fp32 masters (after init / load)."""
        ops.call("mtl_opt_fold", self.td, self.cd, self.num_chunks, self.w, self.wb, self.fold_scales)

    def stats_and_reg_loss(self, grad_scale=1.0):
        ops.call("mtl_opt_stats", self.td, self.num_tensors, self.cd, self.num_chunks, self.w, self.g,
                 grad_scale, self.stats, self.reg_loss, self.chunk_start_dev, self.partials)
        return self.reg_loss

    def apply(self, grad_scale=1.0):
        ops.call("mtl_opt_apply", self.td, self.cd, self.num_chunks, self.w, self.g, self.m, self.wb,
                 self.fold_scales, self.stats, self.hyper, grad_scale)

    # the same two passes restricted to the tensors [t0, t1): the update of a finished gradient bucket can
    # run while the rest of the backward pass still computes (trainer.py)
    def _chunk_range(self, t0, t1):
        c0, c1 = self.chunk_start[t0], self.chunk_start[t1]
        return self.cd[c0 * ctypes.sizeof(ops.ChunkDesc):], c1 - c0

    def stats_range(self, t0, t1, grad_scale=1.0):
        cd, n = self._chunk_range(t0, t1)
        if n:
            ops.call("mtl_opt_stats_range", self.td, t0, t1, cd, n, self.w, self.g, grad_scale, self.stats,
                     self.chunk_start_dev, self.chunk_start[t0], self.partials)

    def apply_range(self, t0, t1, grad_scale=1.0, hyper=None, refresh_norms=False):
        """refresh_norms: the pass also leaves the squared norms of the UPDATED weights in `stats` (what a following
        stats_range would compute for the regularisation loss of the next step) -- no third pass over the weights."""
        cd, n = self._chunk_range(t0, t1)
        if n and refresh_norms and self.partials is not None:
            ops.call("mtl_opt_apply_norms", self.td, t0, t1, cd, n, self.w, self.g, self.m, self.wb, self.fold_scales,
                     self.stats, self.hyper if hyper is None else hyper, grad_scale, self.chunk_start_dev,
                     self.chunk_start[t0], self.partials)
            return
        if n:
            ops.call("mtl_opt_apply", self.td, cd, n, self.w, self.g, self.m, self.wb, self.fold_scales,
                     self.stats, self.hyper if hyper is None else hyper, grad_scale)
            if refresh_norms:
                self.stats_range(t0, t1, grad_scale)

    def reg_loss_from_stats(self):
        ops.call("mtl_opt_reg_loss", self.td, self.num_tensors, self.stats, self.reg_loss)
        return self.reg_loss

    def set_hyper(self, lr, momentum, clip_norm):
        self.hyper.copy_(torch.tensor([lr, momentum, clip_norm, 0.0], dtype=torch.float32))

    # ------------------------------------------------------------------ host import / export
    def state_dict(self):
        """name -> fp32 CPU tensor (weights in this framework's [K,R,S,C] layout) + BN statistics."""
        out = {}
        for p in self.params:
            out[p.name] = p.w.detach().cpu().clone()
        for b in self.bns:
            if b.scope is None:
                continue
            if b.has_gamma:
                out[b.scope + "/gamma"] = b.gamma.clone()
            out[b.scope + "/beta"] = b.beta.clone()
            out[b.scope + "/moving_mean"] = b.mean.clone()
            out[b.scope + "/moving_variance"] = b.var.clone()
        return out

    def load_state_dict(self, sd, strict=True):
        missing = []
        for p in self.params:
            if p.name in sd:
                p.w.copy_(sd[p.name].reshape(p.shape).to(self.device))
            else:
                missing.append(p.name)
        for b in self.bns:
            if b.scope is None:
                continue
            for attr, key in ((("gamma", "gamma"),) if b.has_gamma else ()) + (("beta", "beta"), ("mean", "moving_mean"),
                                                                                   ("var", "moving_variance")):
                k = b.scope + "/" + key
                if k in sd:
                    setattr(b, attr, sd[k].clone().float())
                else:
                    missing.append(k)
        if strict and missing:
            raise KeyError("missing variables: %s" % missing[:8])
        self._upload_bn_inplace()
        self.fold()
        for h in self.post_load_hooks:
            h()
        return missing

    def _upload_bn_inplace(self):
        for b in self.bns:
            s, bi = b.fold_host()
            b.scale.copy_(s.to(self.device))
            b.bias.copy_(bi.to(self.device))


def _fans(p):
    """(fan_in, fan_out) as slim.variance_scaling_initializer computes them from the TF variable shape
    ([R,S,C,K] conv / [in,out] FC), expressed on this framework's [K,R,S,C] / [out,in] layout."""
    fan_in = p.numel // p.shape[0]
    fan_out = p.shape[0] * (p.numel // (p.shape[0] * p.shape[-1])) if len(p.shape) == 4 else p.shape[0]
    return fan_in, fan_out


def _init_tensor(p, gen):
    kind = p.init[0]
    if kind == "zeros":
        return torch.zeros(p.shape)
    if kind == "const":
        return torch.full(p.shape, float(p.init[1]))
    if kind == "truncated_normal":          # tf.truncated_normal_initializer(mean, stddev): resample |z| > 2
        mean = float(p.init[2]) if len(p.init) > 2 else 0.0
        return _trunc_normal(p.shape, float(p.init[1]), gen) + mean
    if kind == "variance_scaling":
        # slim.variance_scaling_initializer(factor=2.0, mode='FAN_IN', uniform=False) (tf.contrib.layers
        # initializers.py): n = fan_in | fan_out | their mean; uniform: U(-l, l) with l = sqrt(3 * factor / n);
        # otherwise truncated normal with stddev sqrt(1.3 * factor / n)
        factor = float(p.init[1]) if len(p.init) > 1 else 2.0
        mode = p.init[2] if len(p.init) > 2 else "FAN_IN"
        uniform = bool(p.init[3]) if len(p.init) > 3 else False
        fan_in, fan_out = _fans(p)
        if mode not in ("FAN_IN", "FAN_OUT", "FAN_AVG"):
            raise TypeError("Unknown mode %s [FAN_IN, FAN_OUT, FAN_AVG]" % (mode,))
        n = {"FAN_IN": float(fan_in), "FAN_OUT": float(fan_out), "FAN_AVG": (fan_in + fan_out) / 2.0}[mode]
        if uniform:
            lim = math.sqrt(3.0 * factor / n)
            return (torch.rand(p.shape, generator=gen) * 2 - 1) * lim
        return _trunc_normal(p.shape, math.sqrt(1.3 * factor / n), gen)
    if kind == "packed_conv":           # [K, ld] rows holding `real` truncated-normal columns, zero padding
        real, std = int(p.init[1]), float(p.init[2])
        t = torch.zeros(p.shape)
        t.view(p.shape[0], -1)[:, :real] = _trunc_normal((p.shape[0], real), std, gen)
        return t
    if kind == "xavier":                # slim default: xavier_initializer (uniform) == factor 1.0, FAN_AVG, uniform
        fan_in, fan_out = _fans(p)
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(p.shape, generator=gen) * 2 - 1) * lim
    if kind == "normal":
        mean = float(p.init[2]) if len(p.init) > 2 else 0.0
        return torch.randn(p.shape, generator=gen) * float(p.init[1]) + mean
    raise ValueError(kind)


def _trunc_normal(shape, std, gen):
    t = torch.randn(shape, generator=gen)
    bad = t.abs() > 2
    while bad.any():
        t[bad] = torch.randn(int(bad.sum()), generator=gen)
        bad = t.abs() > 2
    return t * std


class Workspace(object):
    """Persistent, shape-keyed device buffers: allocated on first use, reused every step, so a
    whole training step issues no allocator calls and can be captured into one CUDA graph."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, key, shape, dtype=torch.bfloat16, zero=False):
        shape = tuple(int(s) for s in shape)
        t = self.bufs.get(key)
        if t is None or t.shape != shape or t.dtype != dtype:
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self.bufs[key] = t
        elif zero:
            t.zero_()
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())
