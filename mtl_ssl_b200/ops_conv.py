"""Python binding of the tcgen05 convolution engine (csrc/gemm_tc.cu, mtl_conv_tc).

Layouts: activations NHWC bf16, weights [K, R, S, C] bf16 (the reference's HWIO weights
of slim.conv2d, /root/reference/slim/nets/resnet_utils.py:111-122, are transposed once at
load time).  All functions enqueue on the current torch stream and never synchronise.
"""
import ctypes

import torch

from ._lib import lib, check, ptr, cur_stream

FPROP, DGRAD, WGRAD = 0, 1, 2

# bench.py sets this to a list to time every tcgen05 launch with CUDA events on the launching
# stream: entries are (mode, algorithmic flops, start event, end event).
PROFILE = None


# tools/timeline.py sets this to a list: (label, stream handle, start event, end event) per launch, any stream
TIMELINE = None


def _launch(a, what):
    if TIMELINE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.current_stream()
        e0.record(s)
        check(lib().mtl_conv_tc(ctypes.byref(a), cur_stream()), what)
        e1.record(s)
        TIMELINE.append(("conv%d N%d %dx%d C%d K%d R%d s%d" % (a.mode, a.N, a.H, a.W, a.C, a.K, a.R, a.stride),
                         s.cuda_stream, e0, e1, 2.0 * a.N * a.P * a.Q * a.K * a.R * a.S * a.C))
        return
    if PROFILE is None:
        check(lib().mtl_conv_tc(ctypes.byref(a), cur_stream()), what)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib().mtl_conv_tc(ctypes.byref(a), cur_stream()), what)
    e1.record()
    flops = 2.0 * a.N * a.P * a.Q * a.K * a.R * a.S * a.C
    PROFILE.append((a.mode, flops, e0, e1, (a.N, a.H, a.W, a.C, a.K, a.R, a.stride, a.P, a.Q)))


def _log_group(grp, launch):
    """PROFILE / TIMELINE records of one grouped launch (one kernel, the group's summed algorithmic FLOPs)."""
    if TIMELINE is None and PROFILE is None:
        return launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.current_stream()
    e0.record(s)
    launch()
    e1.record(s)
    if TIMELINE is not None:
        TIMELINE.append(("conv%d group of %d" % (grp.mode, grp.n), s.cuda_stream, e0, e1, grp.flops))
    if PROFILE is not None:
        PROFILE.append((grp.mode, grp.flops, e0, e1, (grp.n, 0, 0, 0, 0, 0, 0, 0, 0)))


class ConvArgs(ctypes.Structure):
    _fields_ = [
        ("mode", ctypes.c_int),
        ("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("C", ctypes.c_int),
        ("K", ctypes.c_int),
        ("R", ctypes.c_int), ("S", ctypes.c_int), ("stride", ctypes.c_int),
        ("pad_h", ctypes.c_int), ("pad_w", ctypes.c_int), ("dil", ctypes.c_int),
        ("P", ctypes.c_int), ("Q", ctypes.c_int),
        ("x", ctypes.c_void_p), ("w", ctypes.c_void_p), ("dy", ctypes.c_void_p),
        ("out", ctypes.c_void_p), ("out_fp32", ctypes.c_int),
        ("bias", ctypes.c_void_p), ("rowscale", ctypes.c_void_p),
        ("res", ctypes.c_void_p), ("res_fp32", ctypes.c_int),
        ("mask", ctypes.c_void_p),
        ("relu", ctypes.c_int), ("alpha", ctypes.c_float),
        ("force_bn", ctypes.c_int), ("force_splits", ctypes.c_int), ("mask_hi", ctypes.c_float),
        ("dy_ld", ctypes.c_longlong), ("out_ld", ctypes.c_longlong), ("res_ld", ctypes.c_longlong),
        ("mask_ld", ctypes.c_longlong), ("bias_scale", ctypes.c_float),
        ("force_stages", ctypes.c_int), ("ws", ctypes.c_void_p), ("ws_bytes", ctypes.c_longlong),
        ("force_cluster", ctypes.c_int), ("max_ctas", ctypes.c_int),
        ("pool_out", ctypes.c_void_p), ("pool_hw", ctypes.c_int),
    ]


# Grid cap applied to every conv launched while it is set (0 = whole GPU): side-stream work that overlaps a
# latency-bound chain (the frozen-prefix look-ahead of the trainer) is sized to leave that chain its SMs.
CTA_CAP = 0



# Split-K workspaces (fprop / dgrad): one zero-filled fp32 buffer per CUDA stream -- launches on one stream
# are ordered, and the kernel leaves the buffer zeroed.  Outgrown buffers are kept alive because CUDA graphs
# captured earlier still point at them.
SPLIT_K = False     # off by default: fp32-atomic partial sums make the forward pass run-to-run non-deterministic (gain ~1 %)
_ws_by_stream = {}
_ws_retired = []


def _attach_ws(a):
    if not SPLIT_K and a.force_splits <= 1:      # an explicit force_splits always gets its workspace
        return
    f = lib().mtl_conv_tc_ws_bytes
    f.restype = ctypes.c_longlong
    need = int(f(ctypes.byref(a)))
    if need <= 0:
        return
    key = torch.cuda.current_stream().cuda_stream
    buf = _ws_by_stream.get(key)
    if buf is None or buf.numel() * 4 < need:
        if buf is not None:
            _ws_retired.append(buf)
        buf = torch.zeros((max(need, 8 << 20) + 3) // 4, device="cuda", dtype=torch.float32)
        _ws_by_stream[key] = buf
    a.ws, a.ws_bytes = buf.data_ptr(), buf.numel() * 4


class ConvGroup(object):
    """Independent problems of ONE kernel instance (same mtl_conv_tc_group_key) as one persistent launch: the CTAs walk
    the concatenated tile space.  Built once from the problems' ConvArgs (the table of tensor maps + parameters is planned
    on the host and copied to the device); `launch()` only enqueues the kernel, so it can be captured into a CUDA graph.
    The buffers the arguments point at must stay where they are (the workspace / parameter arenas do)."""

    def __init__(self, arg_list, target_k_iters=0):
        n = len(arg_list)
        f = lib().mtl_conv_tc_group_entry_bytes
        f.restype = ctypes.c_longlong
        eb = int(f())
        arr = (ConvArgs * n)()
        for i, a in enumerate(arg_list):
            ctypes.memmove(ctypes.byref(arr, i * ctypes.sizeof(ConvArgs)), ctypes.byref(a), ctypes.sizeof(ConvArgs))
        self.n = n
        self.mode = int(arg_list[0].mode)
        self.flops = sum(2.0 * a.N * a.P * a.Q * a.K * a.R * a.S * a.C for a in arg_list)
        self.signature = tuple((a.x, a.w, a.dy, a.out, a.N, a.H, a.W, a.C, a.K, a.R, a.stride) for a in arg_list)
        self.host = torch.zeros(n * eb + 128, dtype=torch.uint8)
        off = (-self.host.data_ptr()) % 128
        self.host_ptr = self.host.data_ptr() + off
        self.info = (ctypes.c_int * 4)()
        check(lib().mtl_conv_tc_group_build(arr, n, int(target_k_iters), ctypes.c_void_p(self.host_ptr), self.info),
              "mtl_conv_tc_group_build")
        self.total_tiles = int(self.info[0])
        dev = torch.empty(n * eb + 128, dtype=torch.uint8, device="cuda" if torch.cuda.is_available() else "cpu")
        doff = (-dev.data_ptr()) % 128
        dev[doff:doff + n * eb].copy_(self.host[off:off + n * eb])
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()    # (the host table is pageable memory)
        self.dev, self.dev_ptr = dev, dev.data_ptr() + doff

    def launch(self, max_ctas=0):
        _log_group(self, lambda: check(lib().mtl_conv_tc_group_launch(
            ctypes.c_void_p(self.host_ptr), ctypes.c_void_p(self.dev_ptr), self.n, self.info, int(max_ctas),
            cur_stream()), "mtl_conv_tc_group_launch"))


def group_key(a):
    f = lib().mtl_conv_tc_group_key
    f.restype = ctypes.c_longlong
    return int(f(ctypes.byref(a)))


class WgradCollector(object):
    """Defers the weight-gradient GEMMs issued while it is active and runs them as grouped launches at `flush()`.
    Weight gradients feed only the optimizer, so a backward chain can run first and undisturbed (at batch 1 the trunk's
    dgrad chain is latency bound and its 81 wgrad launches used to compete with it for SMs), then all of its weight
    gradients run as a few machine-filling launches.  Groups are planned at the first flush and reused while the
    problems (pointers + geometry) stay the same -- they do: every operand lives in a persistent arena."""
    active = None
    enabled = not __import__("os").environ.get("MTL_NO_WGRAD_GROUP")

    def __init__(self, target_k_iters=48):
        self.target = target_k_iters
        self.pending = []
        self.groups = None

    def __enter__(self):
        if WgradCollector.enabled:
            assert WgradCollector.active is None
            WgradCollector.active = self        # (what an earlier, unflushed pass collected stays pending)
        return self

    def __exit__(self, *exc):
        if WgradCollector.active is self:
            WgradCollector.active = None
        return False

    def add(self, a):
        self.pending.append(a)

    def flush(self, max_ctas=0, key=0, out_range=None):
        """Run everything collected since the last flush.  `key` names the flush point: a pass that flushes several
        times (chunks of a backward chain) plans one set of groups per point.  out_range=(lo, hi): only the GEMMs whose
        output (device address) lies in [lo, hi) -- a piece of the gradient arena -- run now, the others stay pending."""
        if out_range is None:
            args, self.pending = self.pending, []
        else:
            lo, hi = out_range
            args = [a for a in self.pending if lo <= a.out < hi]
            self.pending = [a for a in self.pending if not lo <= a.out < hi]
        if not args:
            return
        if self.groups is None:
            self.groups = {}
        sig = tuple((a.x, a.dy, a.out, a.N, a.H, a.W, a.C, a.K, a.R, a.stride) for a in args)
        cur = self.groups.get(key)
        if cur is None or cur[0] != sig:
            buckets = {}
            for a in args:
                buckets.setdefault(group_key(a), []).append(a)
            cur = (sig, [ConvGroup(v, self.target) for _, v in sorted(buckets.items())])
            self.groups[key] = cur
        for g in cur[1]:
            g.launch(max_ctas)


def out_size(h, k, stride, pad_beg, pad_end, dil=1):
    return (h + pad_beg + pad_end - dil * (k - 1) - 1) // stride + 1


def _dp(t):
    return t.data_ptr() if t is not None else None


def _pitch(t):
    """Pixel pitch (elements) of an NHWC tensor that may be a channel slice of a wider buffer."""
    N, H, W, C = t.shape
    ld = t.stride(2) if W > 1 else (t.stride(1) if H > 1 else (t.stride(0) if N > 1 else C))
    ok = t.stride(3) == 1 and (W == 1 or t.stride(2) == ld) and (H == 1 or t.stride(1) == W * ld) and \
        (N == 1 or t.stride(0) == H * W * ld) and ld >= C
    if not ok:
        raise ValueError("tensor must be NHWC-dense or a channel slice of an NHWC-dense buffer (strides %s)"
                         % (t.stride(),))
    return ld


def _geom(a, xshape, wshape, stride, pad, dil, P, Q):
    a.N, a.H, a.W, a.C = xshape
    a.K, a.R, a.S, _ = wshape
    a.stride, a.pad_h, a.pad_w, a.dil = stride, pad[0], pad[1], dil
    a.P, a.Q = P, Q


def conv_fprop(x, w, stride=1, pad=(0, 0), dil=1, out_hw=None, bias=None, res=None, relu=False,
               out=None, out_dtype=torch.bfloat16, force_bn=0, bias_scale=1.0, force_splits=0, force_stages=0,
               force_cluster=0, pool_out=None, pool_hw=0):
    """y = relu?(conv(x, w) + bias + res).  x [N,H,W,C] bf16, w [K,R,S,C] bf16.
    pool_out (fp32 [2 * ceil(N*P*Q / 32), K], with res and relu=1): y is not stored; the kernel writes per 32-row group
    the partial sums of y over the group's first / second window of `pool_hw` rows (pool_partial_rows; the spatial mean
    that follows a forward-only tail is finished by mtl_head_fwd_pooled)."""
    N, H, W, C = x.shape
    K, R, S, C2 = w.shape
    assert C == C2 and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert x.is_contiguous() and w.is_contiguous()
    if out_hw is None:
        out_hw = (out_size(H, R, stride, pad[0], pad[0], dil), out_size(W, S, stride, pad[1], pad[1], dil))
    P, Q = out_hw
    if out is None:
        out = torch.empty((N, P, Q, K), device=x.device, dtype=out_dtype)
    a = ConvArgs()
    a.mode = FPROP
    _geom(a, (N, H, W, C), (K, R, S, C), stride, pad, dil, P, Q)
    a.x, a.w, a.out = _dp(x), _dp(w), _dp(out)
    a.out_fp32 = int(out.dtype == torch.float32)
    a.bias = _dp(bias)
    a.out_ld = _pitch(out)
    if res is not None:
        assert res.shape == out.shape
        a.res, a.res_fp32, a.res_ld = _dp(res), int(res.dtype == torch.float32), _pitch(res)
    a.relu = int(relu)
    a.alpha = 1.0
    a.bias_scale = float(bias_scale)
    a.force_bn, a.force_splits, a.force_stages = force_bn, force_splits, force_stages
    a.force_cluster, a.max_ctas = force_cluster, CTA_CAP
    if pool_out is not None:
        assert pool_out.dtype == torch.float32 and pool_out.is_contiguous()
        assert tuple(pool_out.shape) == (pool_partial_rows(N * P * Q), K), pool_out.shape
        a.pool_out, a.pool_hw = _dp(pool_out), int(pool_hw)
    _attach_ws(a)
    _launch(a, "mtl_conv_tc(fprop)")
    return out


def pool_partial_rows(rows):
    """Rows of the partial-sum buffer of conv_fprop(pool_out=...): two per 32-row group."""
    return 2 * ((rows + 31) // 32)


def conv_dgrad(dy, w, x_shape, stride=1, pad=(0, 0), dil=1, res=None, mask=None, out=None,
               out_dtype=torch.bfloat16, force_bn=0, mask_hi=0.0, force_splits=0, force_stages=0, force_cluster=0):
    """dx = mask>0 ? (conv_transpose(dy, w) + res) : 0.  dy [N,P,Q,K], w [K,R,S,C]."""
    N, P, Q, K = dy.shape
    K2, R, S, C = w.shape
    assert K == K2 and w.is_contiguous()
    _, H, W, C2 = x_shape
    assert C2 == C
    if out is None:
        out = torch.empty((N, H, W, C), device=dy.device, dtype=out_dtype)
    a = ConvArgs()
    a.mode = DGRAD
    _geom(a, (N, H, W, C), (K, R, S, C), stride, pad, dil, P, Q)
    a.dy, a.w, a.out = _dp(dy), _dp(w), _dp(out)
    a.dy_ld, a.out_ld = _pitch(dy), _pitch(out)
    a.out_fp32 = int(out.dtype == torch.float32)
    if res is not None:
        assert res.shape == out.shape
        a.res, a.res_fp32, a.res_ld = _dp(res), int(res.dtype == torch.float32), _pitch(res)
    if mask is not None:
        assert mask.shape == out.shape and mask.dtype == torch.bfloat16
        a.mask, a.mask_ld = _dp(mask), _pitch(mask)
        a.mask_hi = float(mask_hi)
    a.alpha = 1.0
    a.force_bn, a.force_splits, a.force_stages = force_bn, force_splits, force_stages
    a.force_cluster, a.max_ctas = force_cluster, CTA_CAP
    _attach_ws(a)
    _launch(a, "mtl_conv_tc(dgrad)")
    return out


def conv_wgrad(dy, x, dw, stride=1, pad=(0, 0), dil=1, rowscale=None, alpha=1.0, force_bn=0,
               force_splits=0):
    """dw[K,R,S,C] (fp32) += alpha * rowscale[k] * sum_pixels dy[p,k] * im2col(x)[p,(r,s,c)]."""
    N, P, Q, K = dy.shape
    N2, H, W, C = x.shape
    K2, R, S, C2 = dw.shape
    assert N == N2 and K == K2 and C == C2 and dw.dtype == torch.float32
    assert x.is_contiguous() and dw.is_contiguous()
    a = ConvArgs()
    a.mode = WGRAD
    _geom(a, (N, H, W, C), (K, R, S, C), stride, pad, dil, P, Q)
    a.dy, a.x, a.out = _dp(dy), _dp(x), _dp(dw)
    a.dy_ld = _pitch(dy)
    a.rowscale = _dp(rowscale)
    a.alpha = float(alpha)
    a.force_bn = force_bn
    a.force_splits = force_splits
    a.max_ctas = CTA_CAP
    if WgradCollector.active is not None and force_splits == 0 and force_bn == 0:
        WgradCollector.active.add(a)
        return dw
    _launch(a, "mtl_conv_tc(wgrad)")
    return dw
