"""torch-facing wrappers of the C ABI: index structures + autograd Functions.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all
the arithmetic of the hot path happens in libspt_b200.so.  Every function
refuses non-CUDA tensors: there is no CPU fallback.
"""
import ctypes
import os
import weakref

import torch

from . import _lib

REDUCE = {"sum": 0, "add": 0, "mean": 1, "max": 2, "min": 3}
SCALE_D_TIMES_G, SCALE_D_PLUS_G, SCALE_D, SCALE_G, SCALE_CONST = range(5)

_LAUNCHES = 0  # kernels-launching ABI calls issued (bench.py reads it)


def launch_count():
    return _LAUNCHES


def _count(n=1):
    global _LAUNCHES
    _LAUNCHES += n


# optional per-kernel CUDA-event timing (bench.py roofline): list of
# (tag, meta, start_event, end_event); events are recorded on the launch stream.
_TIMING = None


def enable_event_timing(on=True):
    global _TIMING
    _TIMING = [] if on else None


def timing_records():
    return _TIMING


class _timed:
    def __init__(self, tag, **meta):
        self.tag, self.meta = tag, meta

    def __enter__(self):
        if _TIMING is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream())
        return self

    def __exit__(self, *exc):
        if _TIMING is not None:
            self.e.record(torch.cuda.current_stream())
            _TIMING.append((self.tag, self.meta, self.s, self.e))
        return False


class _ZeroPool:
    """Per-step pool of pre-zeroed fp32 scratch for accumulate-style outputs (the dW / dbias
    targets of `spt_gemm_tn_acc`, the packed RPE weight gradients): ONE memset per step
    instead of one fill kernel per buffer (~280 per cfg-2 step).  Opt-in: a training loop
    brackets its backward pass with begin_step() / end_step() (FlatGradients.release() /
    .collect() do); buffers handed out are only valid until the next begin_step(), which is
    fine for gradients that are packed into the flat buffer right after backward.  Outside
    such a bracket `take` is torch.zeros."""

    def __init__(self):
        self.buf, self.off, self.need, self.cur, self.on = None, 0, 0, 0, False

    def begin_step(self, device):
        if self.buf is None or self.buf.device != device or self.buf.numel() < self.need:
            self.buf = torch.empty(max(self.need, 1 << 18), dtype=torch.float32, device=device)
        self.buf.zero_()
        self.off = self.cur = 0
        self.on = True

    def end_step(self):
        self.need = max(self.need, self.cur)
        self.on = False

    def take(self, shape, device):
        n = 1
        for d in shape:
            n *= int(d)
        n_al = (n + 3) // 4 * 4          # keep every buffer 16-byte aligned
        self.cur += n_al
        if (not self.on or self.buf is None or self.buf.device != device
                or self.off + n_al > self.buf.numel()):
            return torch.zeros(shape, dtype=torch.float32, device=device)
        v = self.buf[self.off:self.off + n].view(shape)
        self.off += n_al
        return v


zero_pool = _ZeroPool()


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "superpoint_transformer_b200 ops run on CUDA tensors only "
                "(no CPU fallback); got a tensor on " + str(t.device))


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


# A second stream per device for launches that are independent of what follows them on the
# main stream (the d[Wq;Wk] product of the attention backward runs beside the targets pass).
# Fork / join with wait_stream, so the pattern is CUDA-graph capturable.
_SIDE_STREAMS = {}
ATTN_DW_SIDE_STREAM = os.environ.get('SPT_ATTN_DW_SIDE_STREAM', '0') != '0'   # measured: 13.4-13.6 vs 13.6 ms, within noise -> off


def _side_stream(device):
    key = torch.device(device).index
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32 tensor, got {t.dtype}")
    return t.contiguous()


def _i64c(t):
    if t is None:
        return None
    if t.dtype != torch.int64:
        t = t.long()
    return t.contiguous()


# ---------------------------------------------------------------------------
# index structures
# ---------------------------------------------------------------------------
class SegmentIndex:
    """Stable CSR of `n` items grouped by `key` in [0, num_groups).

    ptr  int32 [num_groups+1], perm int32 [n] (= stable argsort(key)),
    other_sorted int32 [n] (optional payload gathered through perm).
    With key=super_index this equals the reference `Cluster.pointers/points`
    (src/data/cluster.py:19-77).
    """

    __slots__ = ("ptr", "perm", "other_sorted", "n", "num_groups", "_err", "__weakref__")

    def __init__(self, ptr, perm, other_sorted, n, num_groups, err):
        self.ptr, self.perm, self.other_sorted = ptr, perm, other_sorted
        self.n, self.num_groups, self._err = n, num_groups, err

    def num_invalid_keys(self):
        """Host sync: number of keys that were outside [0, num_groups)."""
        return int(self._err[0].item())


_DEBUG_INDEX = bool(int(__import__('os').environ.get('SPT_DEBUG_INDEX', '0')))


def set_debug_index(on=True):
    """validate every index build (host sync): out-of-range super_index / edge_index entries
    raise IndexError instead of being skipped"""
    global _DEBUG_INDEX
    _DEBUG_INDEX = bool(on)


def group_index(key, num_groups, other=None):
    lib = _lib.load()
    _require_cuda(key, other)
    key = _i64c(key)
    other = _i64c(other)
    n = key.numel()
    dev = key.device
    ptr = torch.empty(num_groups + 1, dtype=torch.int32, device=dev)
    # out-of-range keys are skipped by the kernel: with SPT_DEBUG_INDEX=1 their slots are
    # zero-initialised and the call syncs and raises like the reference's index error
    # (default: no host sync on the hot path; valid keys fill every slot)
    alloc = torch.zeros if _DEBUG_INDEX else torch.empty
    perm = alloc(n, dtype=torch.int32, device=dev)
    osort = alloc(n, dtype=torch.int32, device=dev) if other is not None else None
    nbytes = lib.spt_group_index_workspace_bytes(n, num_groups)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev), _timed('group_index', n=n, G=num_groups, other=other is not None):
        rc = lib.spt_group_index(_p(key), _p(other), n, num_groups, _p(ptr), _p(perm),
                                 _p(osort), _p(ws), nbytes, _stream())
    _lib.check(rc, "spt_group_index")
    _count(8 if other is not None else 7)
    err = ws[:4].view(torch.int32)
    seg = SegmentIndex(ptr, perm, osort, n, num_groups, err)
    if _DEBUG_INDEX:
        bad = seg.num_invalid_keys()
        if bad:
            raise IndexError(f"group_index: {bad} of {n} keys outside [0, {num_groups})")
    return seg


class GraphIndex:
    """CSR (by source) + CSC (by target) of an attention graph.

    rowptr/col/perm: edges grouped by source row (stable), col = target of each
    CSR slot, perm = original edge id of each CSR slot.
    csc_ptr/csc_src/csc2csr: edges grouped by target; for each CSC slot the
    source row and the CSR slot of the same edge.
    """

    __slots__ = ("rowptr", "col", "perm", "csc_ptr", "csc_src", "csc2csr",
                 "num_rows", "num_targets", "E", "_edge_row", "__weakref__")

    @property
    def edge_row(self):
        """int32 [E]: the row of every CSR slot (input of the edge-parallel attention kernels);
        built on first use, once per graph."""
        er = getattr(self, "_edge_row", None)
        if er is None:
            lib = _lib.load()
            er = torch.empty(max(self.E, 1), dtype=torch.int32, device=self.rowptr.device)
            with torch.cuda.device(self.rowptr.device):
                _lib.check(lib.spt_expand_pointers_i32(_p(self.rowptr), self.num_rows, _p(er),
                                                       _stream()), "spt_expand_pointers_i32")
            _count()
            self._edge_row = er
        return er


# edge_index tensors whose columns are already grouped by source row in CSR order (stable):
# the attention blocks then read `edge_attr` as it is (no per-stage permutation, no inverse
# permutation of its gradient).  Registered by OnTheFlyHorizontalEdgeFeatures(csr_order=True).
_csr_ordered = {}


def mark_csr_ordered(edge_index):
    import weakref
    key = id(edge_index)
    _csr_ordered[key] = (weakref.ref(edge_index, lambda _r, k=key: _csr_ordered.pop(k, None)),
                         edge_index._version)


def is_csr_ordered(edge_index):
    hit = _csr_ordered.get(id(edge_index))
    return hit is not None and hit[0]() is edge_index and hit[1] == edge_index._version


def build_graph_index(edge_index, num_rows, num_targets=None):
    lib = _lib.load()
    _require_cuda(edge_index)
    presorted = is_csr_ordered(edge_index)
    num_targets = num_rows if num_targets is None else num_targets
    edge_index = _i64c(edge_index)
    src, dst = edge_index[0], edge_index[1]
    E = src.numel()
    csr = group_index(src, num_rows, other=dst)
    csc = group_index(dst, num_targets, other=src)
    g = GraphIndex()
    g.rowptr, g.col, g.perm = csr.ptr, csr.other_sorted, csr.perm
    g.csc_ptr, g.csc_src = csc.ptr, csc.other_sorted
    if presorted:  # CSR slot == edge id
        g.perm = None
        g.csc2csr = csc.perm
        g.num_rows, g.num_targets, g.E = num_rows, num_targets, E
        return g
    inv = torch.empty(E, dtype=torch.int32, device=src.device)
    c2c = torch.empty(E, dtype=torch.int32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(lib.spt_invert_permutation(_p(csr.perm), E, _p(inv), _stream()),
                   "spt_invert_permutation")
        _lib.check(lib.spt_gather_i32(_p(inv), _p(csc.perm), E, _p(c2c), _stream()),
                   "spt_gather_i32")
    _count(2)
    g.csc2csr = c2c
    g.num_rows, g.num_targets, g.E = num_rows, num_targets, E
    return g


class _IdentityCache:
    """Cache keyed by tensor identity (id + version + weakref check).  The
    reference rebuilds nothing because it has no index structure; here the CSR
    of a level is shared by every block of a stage, forward and backward."""

    def __init__(self, maxsize=64):
        self._d = {}
        self._maxsize = maxsize

    def get(self, tensor, extra, builder):
        key = (id(tensor), tensor._version, tuple(tensor.shape), extra)
        hit = self._d.get(key)
        if hit is not None:
            ref, val = hit
            if ref() is tensor:
                return val
        val = builder()
        # drop entries whose key tensor died (frees cached device buffers), then
        # the oldest if still full
        for k in [k for k, (r, _) in self._d.items() if r() is None]:
            del self._d[k]
        while len(self._d) >= self._maxsize:
            del self._d[next(iter(self._d))]
        self._d[key] = (weakref.ref(tensor), val)
        return val

    def put(self, tensor, extra, val):
        key = (id(tensor), tensor._version, tuple(tensor.shape), extra)
        self._d[key] = (weakref.ref(tensor), val)

    def clear(self):
        self._d.clear()


_graph_cache = _IdentityCache()
_segment_cache = _IdentityCache()
_numseg_cache = _IdentityCache(256)
_permuted_cache = _IdentityCache()


def clear_caches():
    for c in (_graph_cache, _segment_cache, _numseg_cache, _permuted_cache, _bf16_cache):
        c.clear()


def graph_index(edge_index, num_rows, num_targets=None):
    return _graph_cache.get(edge_index, ("g", num_rows, num_targets),
                            lambda: build_graph_index(edge_index, num_rows, num_targets))


def segment_index(index, num_groups):
    return _segment_cache.get(index, ("s", num_groups),
                              lambda: group_index(index, num_groups))


def register_segment_index(index, num_groups, seg):
    _segment_cache.put(index, ("s", num_groups), seg)


def num_segments(batch):
    """int(batch.max()) + 1 with one host sync per distinct tensor (the PyG
    GraphNorm the reference uses syncs on every call)."""
    if batch is None:
        return 1
    return _numseg_cache.get(batch, "n",
                             lambda: (int(batch.max().item()) + 1) if batch.numel() else 1)


def register_num_segments(batch, n):
    if batch is not None:
        _numseg_cache.put(batch, "n", int(n))


# ---------------------------------------------------------------------------
# row gathers
# ---------------------------------------------------------------------------
def _gather_rows(x, idx):
    lib = _lib.load()
    n_out, C = idx.numel(), x.shape[1]
    out = torch.empty((n_out, C), dtype=x.dtype, device=x.device)
    fn = lib.spt_gather_rows_i32 if idx.dtype == torch.int32 else lib.spt_gather_rows_i64
    with torch.cuda.device(x.device), _timed('gather_rows', n=n_out, C=C):
        _lib.check(fn(_p(x), _p(idx), n_out, C, _p(out), _stream()), "spt_gather_rows")
    _count()
    return out


class _PermuteRows(torch.autograd.Function):
    """out[j] = x[perm[j]] with perm a permutation; backward scatters back through
    the inverse permutation (a gather, deterministic)."""

    @staticmethod
    def forward(ctx, x, perm):
        ctx.save_for_backward(perm)
        return _gather_rows(x, perm)

    @staticmethod
    def backward(ctx, g):
        (perm,) = ctx.saved_tensors
        lib = _lib.load()
        g = _f32c(g)
        inv = torch.empty_like(perm)
        with torch.cuda.device(g.device):
            _lib.check(lib.spt_invert_permutation(_p(perm), perm.numel(), _p(inv), _stream()),
                       "spt_invert_permutation")
        _count()
        return _gather_rows(g, inv), None


def permute_rows(x, perm):
    _require_cuda(x, perm)
    return _PermuteRows.apply(_f32c(x), perm)


def permute_rows_cached(x, perm):
    """edge_attr is shared by all blocks of a stage (src/nn/stage.py:277-280): the
    CSR-ordered copy is made once and its gradient accumulates in CSR order."""
    if perm is None:   # edges already in CSR order (mark_csr_ordered)
        return _f32c(x)
    if x.requires_grad and torch.is_grad_enabled():
        # a cached autograd output would be backpropagated through twice by a second forward
        # (and would keep the upstream graph alive): the blocks of ONE forward share the copy
        # through the `edge_attr` object identity + version below, nothing survives a backward
        hit = _permuted_grad.get('v')
        if (hit is not None and hit[0]() is x and hit[1] == x._version and hit[2]() is perm
                and hit[3].grad_fn is not None and not getattr(hit[3], '_spt_used', False)):
            return hit[3]
        out = permute_rows(x, perm)
        if out.grad_fn is not None:
            out.register_hook(_mark_used(out))
        _permuted_grad['v'] = (weakref.ref(x), x._version, weakref.ref(perm), out)
        return out
    return _permuted_cache.get(x, ("p", id(perm), perm.data_ptr()),
                               lambda: permute_rows(x, perm))


_permuted_grad = {}


def _mark_used(t):
    """once a gradient reaches the shared CSR copy its graph is being consumed: do not hand
    the same tensor to a later forward"""
    ref = weakref.ref(t)

    def hook(g):
        tt = ref()
        if tt is not None:
            tt._spt_used = True
            _permuted_grad.pop('v', None)
        return g
    return hook


class _IndexUnpool(torch.autograd.Function):
    """x_parent.index_select(0, idx) (src/nn/unpool.py:12-13); backward is a CSR
    segment-sum instead of an atomic index_add."""

    @staticmethod
    def forward(ctx, x, idx):
        ctx.num_parents = x.shape[0]
        ctx.save_for_backward(idx)
        return _gather_rows(x, idx)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        seg = segment_index(idx, ctx.num_parents)
        out, _ = _segment_pool_fwd(_f32c(g), seg, "sum")
        return out, None


def index_unpool(x, idx):
    _require_cuda(x, idx)
    return _IndexUnpool.apply(_f32c(x), _i64c(idx))


# ---------------------------------------------------------------------------
# dense projections: fp32-accurate "3xTF32" GEMMs on the tensor cores
# ---------------------------------------------------------------------------
LINEAR_TC_MIN_ROWS = int(__import__('os').environ.get('SPT_LINEAR_TC_MIN_ROWS', 2048))   # below this the plain library GEMM is as fast
LINEAR_TC = True            # fused single-pass 3xTF32 tensor-core GEMMs (csrc/gemm.cu)


def _split_tf32(x):
    """(hi, lo) with x == hi + lo exactly, hi a tf32 number (spt_split_tf32)."""
    lib = _lib.load()
    x = x.contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.spt_split_tf32(_p(x), x.numel(), _p(hi), _p(lo), _stream()),
                   "spt_split_tf32")
    _count()
    return hi, lo


def _gemm_nt(a, b, bias=None):
    """a [M,K] @ b[N,K]^T (+ bias) on the tensor cores, fp32-accurate."""
    lib = _lib.load()
    M, K = a.shape
    N = b.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device), _timed('gemm_nt', M=M, N=N, K=K):
        _lib.check(lib.spt_gemm_nt(_p(a), M, K, a.stride(0), _p(b), N, b.stride(0), _p(bias),
                                   _p(out), N, _stream()), "spt_gemm_nt")
    _count()
    return out


class _LinearTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return _gemm_nt(x, W, b)

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        lib = _lib.load()
        g = g.contiguous()
        M, N = g.shape
        K = x.shape[1]
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = _gemm_nt(g, W.t().contiguous())
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dW = zero_pool.take((N, K), g.device)
            db = zero_pool.take((N,), g.device) if ctx.has_bias else None
            with torch.cuda.device(g.device), _timed('gemm_tn', M=M, N=N, K=K):
                _lib.check(lib.spt_gemm_tn_acc(_p(g), M, N, g.stride(0), _p(x), K, x.stride(0),
                                               _p(dW), K, _p(db), _stream()),
                           "spt_gemm_tn_acc")
            _count()
        return dx, dW, db


def linear(x, weight, bias=None):
    """y = x W^T + b.  Large row counts run the fused 3xTF32 tensor-core kernels
    (fp32-accurate, operand streamed once); tiny or oddly-shaped ones (K or N not a
    multiple of 4) use the plain fp32 library GEMM."""
    if (LINEAR_TC and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
            and x.shape[0] >= LINEAR_TC_MIN_ROWS):
        K, N = x.shape[1], weight.shape[0]
        if N % 4 != 0:
            # e.g. a 13-class head: zero-pad the output features, slice the result
            padn = 4 - N % 4
            wp = torch.nn.functional.pad(weight, (0, 0, 0, padn))
            bp = torch.nn.functional.pad(bias, (0, padn)) if bias is not None else None
            return linear(x, wp, bp)[:, :N]
        if K % 4 != 0:
            # e.g. the 18 raw edge features: zero-pad K to a multiple of 4 (one extra
            # pass over x, still far cheaper than the SIMT library GEMM it replaces)
            pad = 4 - K % 4
            x = torch.nn.functional.pad(x, (0, pad))
            weight = torch.nn.functional.pad(weight, (0, pad))
        x, weight = x.contiguous(), weight.contiguous()
        if x.data_ptr() % 16 == 0 and weight.data_ptr() % 16 == 0:
            return _LinearTC.apply(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)


# ---------------------------------------------------------------------------
# segment pooling
# ---------------------------------------------------------------------------
def _segment_pool_fwd(x, seg, reduce):
    lib = _lib.load()
    r = REDUCE[reduce]
    Np, C = seg.num_groups, x.shape[1]
    out = torch.empty((Np, C), dtype=torch.float32, device=x.device)
    arg = torch.empty((Np, C), dtype=torch.int32, device=x.device) if r >= 2 else None
    with torch.cuda.device(x.device), _timed('segment_pool_fwd', Nc=x.shape[0], Np=Np, C=C, r=r):
        _lib.check(lib.spt_segment_pool_fwd(_p(x), _p(seg.ptr), _p(seg.perm), Np, C, r,
                                            _p(out), _p(arg), _stream()),
                   "spt_segment_pool_fwd")
    _count()
    return out, arg


class _SegmentPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, index, seg, reduce):
        out, arg = _segment_pool_fwd(x, seg, reduce)
        ctx.reduce = reduce
        ctx.seg = seg
        ctx.shape = x.shape
        ctx.save_for_backward(index, arg if arg is not None else index)
        ctx.has_arg = arg is not None
        return out

    @staticmethod
    def backward(ctx, g):
        index, arg = ctx.saved_tensors
        lib = _lib.load()
        g = _f32c(g)
        Nc, C = ctx.shape
        dx = torch.empty((Nc, C), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device), _timed('segment_pool_bwd', Nc=Nc, Np=g.shape[0], C=C,
                                                 r=REDUCE[ctx.reduce]):
            _lib.check(lib.spt_segment_pool_bwd(_p(g), _p(index), _p(ctx.seg.ptr),
                                                _p(arg) if ctx.has_arg else None, Nc, C,
                                                REDUCE[ctx.reduce], _p(dx), _stream()),
                       "spt_segment_pool_bwd")
        _count()
        return dx, None, None, None


def segment_pool(x, index, num_pool, reduce="max", seg=None):
    """reduce over children of each parent; PyG *Aggregation semantics
    (src/nn/pool.py:44-82): empty parents -> 0, max/min gradient to one arg."""
    _require_cuda(x, index)
    index = _i64c(index)
    if seg is None:
        seg = segment_index(index, num_pool)
    return _SegmentPool.apply(_f32c(x), index, seg, reduce)


def segment_mean_std(x, index, num_segments_, want_mean=True, want_std=True, seg=None):
    """(scatter_mean, scatter_std) of torch_scatter over `index` (reference call sites
    src/transforms/graph.py:266-285, 1025-1044): mean = sum / max(count, 1), unbiased std with
    the `+ 1e-6` of torch_scatter.  No gradient (preprocessing-time features)."""
    lib = _lib.load()
    _require_cuda(x, index)
    x = _f32c(x.detach())
    squeeze = x.dim() == 1
    if squeeze:
        x = x.view(-1, 1)
    index = _i64c(index)
    if seg is None:
        seg = segment_index(index, num_segments_)
    C = x.shape[1]
    mean = torch.empty((num_segments_, C), dtype=torch.float32, device=x.device) if want_mean else None
    std = torch.empty((num_segments_, C), dtype=torch.float32, device=x.device) if want_std else None
    with torch.cuda.device(x.device):
        _lib.check(lib.spt_segment_mean_std_fwd(_p(x), _p(seg.ptr), _p(seg.perm), num_segments_,
                                                C, _p(mean), _p(std), _stream()),
                   "spt_segment_mean_std_fwd")
    _count()
    if squeeze:
        mean = None if mean is None else mean.view(-1)
        std = None if std is None else std.view(-1)
    return mean, std


def superedge_features(points, se_point_index, se_id, num_superedges):
    """edge_attr [num_superedges, 7] = [mean_off(3) | std_off(3) | mean_dist(1)] of
    _minimalistic_horizontal_edge_features (reference src/transforms/graph.py:950-1060)."""
    lib = _lib.load()
    _require_cuda(points, se_point_index, se_id)
    points = _f32c(points.detach())
    spi = _i64c(se_point_index)
    seg = group_index(_i64c(se_id), num_superedges)
    out = torch.empty((num_superedges, 7), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.check(lib.spt_superedge_features_fwd(_p(points), _p(spi[0]), _p(spi[1]), _p(seg.ptr),
                                                  _p(seg.perm), num_superedges, _p(out),
                                                  _stream()), "spt_superedge_features_fwd")
    _count()
    return out


def concat_offset(tensors, offsets=None, skip_first=False):
    """torch.cat([t + off for t, off in zip(tensors, offsets)]) for 1-D int64 CUDA tensors in
    ONE launch (csrc/batch.cu): the integer work of Batch.from_data_list / CSRBatch.from_list
    (reference src/data/data.py:1154-1242, src/data/csr.py:676-757).
    skip_first: drop element 0 of every tensor but the first (CSR pointer concatenation)."""
    lib = _lib.load()
    _require_cuda(*tensors)
    ts = [_i64c(t).view(-1) for t in tensors]
    dev = ts[0].device
    lens = [t.numel() - (1 if (skip_first and i > 0) else 0) for i, t in enumerate(ts)]
    prefix = [0]
    for n in lens:
        prefix.append(prefix[-1] + n)
    total = prefix[-1]
    out = torch.empty(total, dtype=torch.int64, device=dev)
    if total == 0:
        return out
    # one small H2D copy of the segment table (pointers, prefix, offsets)
    table = [t.data_ptr() for t in ts] + prefix + [int(o) for o in (offsets or [0] * len(ts))]
    tab = torch.tensor(table, dtype=torch.int64).to(dev, non_blocking=True)
    S = len(ts)
    with torch.cuda.device(dev):
        _lib.check(lib.spt_concat_offset_i64(tab.data_ptr(), tab.data_ptr() + 8 * S,
                                             tab.data_ptr() + 8 * (2 * S + 1), S, total,
                                             1 if skip_first else 0, _p(out), _stream()),
                   "spt_concat_offset_i64")
    _count()
    return out


def segment_ids(sizes, device):
    """repeat_interleave(arange(len(sizes)), sizes) on the device (the `batch` vector)."""
    lib = _lib.load()
    prefix = [0]
    for n in sizes:
        prefix.append(prefix[-1] + int(n))
    total = prefix[-1]
    out = torch.empty(total, dtype=torch.int64, device=device)
    if total == 0:
        return out
    tab = torch.tensor(prefix, dtype=torch.int64).to(device, non_blocking=True)
    with torch.cuda.device(device):
        _lib.check(lib.spt_concat_offset_i64(None, tab.data_ptr(), None, len(sizes), total, 0,
                                             _p(out), _stream()), "spt_concat_offset_i64")
    _count()
    return out


# ---------------------------------------------------------------------------
# node selection (csrc/select.cu): the integer primitives of NAG.select
# ---------------------------------------------------------------------------
def _ws_bytes(n, device):
    return torch.empty(max(int(n), 1), dtype=torch.uint8, device=device)


def _read_counts(counts, what):
    """The one host read of a two-phase selection primitive: (count, invalid entries)."""
    count, bad = counts.tolist()
    if bad:
        raise IndexError(f"{what}: {bad} index entries are out of range or repeated")
    return count


def relabel_consecutive(ids, num_ids, payload=None):
    """`consecutive_cluster(ids)` of torch_geometric.nn.pool.consecutive for ids in
    [0, num_ids) (reference call sites src/data/cluster.py:131, src/data/data.py:405) without
    the sort: returns (new_ids, unique_ids) with new_ids[i] the rank of ids[i] among the distinct
    values and unique_ids the distinct values in ascending order — the reference's
    `ids[perm]`.  With `payload` also returns payload_by_new (payload_by_new[new_ids[i]] =
    payload[i]; the ids must then be distinct)."""
    lib = _lib.load()
    _require_cuda(ids, payload)
    ids = _i64c(ids).view(-1)
    n, dev = ids.numel(), ids.device
    cap = min(n, int(num_ids))
    new_ids = torch.empty(n, dtype=torch.int64, device=dev)
    uniq = torch.empty(cap, dtype=torch.int64, device=dev)
    counts = torch.empty(2, dtype=torch.int64, device=dev)
    by_new = None
    if payload is not None:
        payload = _i64c(payload).view(-1)
        by_new = torch.empty(cap, dtype=torch.int64, device=dev)
    nb = lib.spt_relabel_consecutive_workspace_bytes(int(num_ids))
    ws = _ws_bytes(nb, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.spt_relabel_consecutive_i64(
            _p(ids), n, int(num_ids), _p(new_ids), _p(uniq), _p(counts), _p(payload),
            _p(by_new), _p(ws), nb, _stream()), "spt_relabel_consecutive_i64")
    _count(6)
    num_unique = _read_counts(counts, "relabel_consecutive")
    if payload is not None:
        if num_unique != n:
            raise IndexError("relabel_consecutive: payload given but the ids are not distinct")
        return new_ids, uniq[:num_unique], by_new[:num_unique]
    return new_ids, uniq[:num_unique]


def select_edges(edge_index, idx, num_nodes):
    """Edge update of Data.select (reference src/data/data.py:356-371): returns
    (edge_index', idx_edge) with edge_index' = reindex[edge_index[:, idx_edge]], reindex the
    old-id -> position-in-idx table, idx_edge the edges whose two end points are selected, in
    their original order.  Also validates `idx` (range, duplicates).  edge_index None: only the
    validation."""
    lib = _lib.load()
    _require_cuda(edge_index, idx)
    idx = _i64c(idx).view(-1)
    dev = idx.device
    K = idx.numel()
    E = 0 if edge_index is None else int(edge_index.shape[1])
    ei = _i64c(edge_index) if E > 0 else None
    reindex = torch.empty(int(num_nodes), dtype=torch.int64, device=dev)
    slot = torch.empty(E + 1, dtype=torch.int32, device=dev) if E > 0 else None
    counts = torch.empty(2, dtype=torch.int64, device=dev)
    nb = lib.spt_select_edges_workspace_bytes(E) if E > 0 else 0
    ws = _ws_bytes(nb, dev) if E > 0 else None
    with torch.cuda.device(dev):
        _lib.check(lib.spt_select_edges_mark(
            _p(ei), E, _p(idx), K, int(num_nodes), _p(reindex), _p(slot), _p(counts), _p(ws),
            nb, _stream()), "spt_select_edges_mark")
        _count(7 if E > 0 else 2)
        kept = _read_counts(counts, "select: idx")
        if edge_index is None:
            return None, None
        out = torch.empty((2, kept), dtype=torch.int64, device=dev)
        idx_edge = torch.empty(kept, dtype=torch.int64, device=dev)
        if kept > 0:
            _lib.check(lib.spt_select_edges_write(
                _p(ei), E, _p(reindex), _p(slot), kept, _p(out), _p(idx_edge), _stream()),
                "spt_select_edges_write")
            _count()
    return out, idx_edge


def csr_select(pointers, values, idx, want_group=False):
    """CSRData.__getitem__ (reference src/data/csr.py:328-393) for one int64 value tensor:
    (new_pointers, values[val_idx][, group of every selected item])."""
    lib = _lib.load()
    _require_cuda(pointers, values, idx)
    pointers, values, idx = _i64c(pointers), _i64c(values).view(-1), _i64c(idx).view(-1)
    dev = pointers.device
    G, K = pointers.numel() - 1, idx.numel()
    new_ptr = torch.empty(K + 1, dtype=torch.int64, device=dev)
    counts = torch.empty(2, dtype=torch.int64, device=dev)
    nb = lib.spt_csr_select_workspace_bytes(K)
    ws = _ws_bytes(nb, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.spt_csr_select_pointers(
            _p(pointers), G, values.numel(), _p(idx), K, _p(new_ptr), _p(counts), _p(ws), nb,
            _stream()), "spt_csr_select_pointers")
        _count(5)
        M = _read_counts(counts, "csr_select: idx")
        out = torch.empty(M, dtype=torch.int64, device=dev)
        group = torch.empty(M, dtype=torch.int64, device=dev) if want_group else None
        if M > 0:
            _lib.check(lib.spt_csr_select_values_i64(
                _p(pointers), _p(idx), K, _p(new_ptr), _p(values), M, _p(out), _p(group),
                _stream()), "spt_csr_select_values_i64")
            _count()
    return (new_ptr, out, group) if want_group else (new_ptr, out)


def sampling_counts(size, n_max, n_min):
    """Number of elements `sparse_sample` draws from segments of `size` elements: the
    reference's fp32 heuristic, written with the same tensor ops (src/utils/sparse.py:176-189)
    so that it rounds the same way."""
    if n_max > 0:
        n_samples = (n_max * torch.tanh(size / n_max)).floor().long()
    else:
        n_samples = size.sqrt().round().long()
    return n_samples.clamp(min=n_min).clamp(max=size)


def sparse_sample(idx, n_max=32, n_min=1, mask=None, return_pointers=False, num_segments=None,
                  seed=None):
    """Indices of elements sampled without replacement from every segment of `idx` — at least
    `n_min`, at most `n_max` per segment, within its size (reference
    src/utils/sparse.py:142-243).  Returns idx_samples (segments ascending) and, with
    `return_pointers`, the [G+1] pointers of the segments in it.

    Device path (csrc/sample.cu): the elements are grouped once (spt_group_index, cached for a
    `super_index`) and every segment draws its own random subset — no global shuffle, no sort.
    `seed`: 64-bit seed of the counter-based generator (default: drawn from torch's global CPU
    generator, so torch.manual_seed makes the call reproducible).  `num_segments` saves the
    idx.max() host read."""
    from .utils.tensor import tensor_idx, sizes_to_pointers
    assert 0 <= n_min <= n_max
    lib = _lib.load()
    _require_cuda(idx)
    idx = _i64c(idx).view(-1)
    dev = idx.device
    G = int(idx.max()) + 1 if num_segments is None else int(num_segments)
    seg = segment_index(idx, G)
    size = (seg.ptr[1:] - seg.ptr[:-1]).long()
    n_samples = sampling_counts(size, n_max, n_min)
    elem_ids = None
    mask = tensor_idx(mask, device=dev)
    if mask is not None:                       # sparse.py:199-205
        seg = group_index(idx[mask], G)
        size = (seg.ptr[1:] - seg.ptr[:-1]).long()
        n_samples = n_samples.clamp(max=size)
        elem_ids = mask.contiguous()
    ptr_samples = sizes_to_pointers(n_samples)
    total = int(ptr_samples[-1])
    out = torch.empty(total, dtype=torch.int64, device=dev)
    if seed is None:
        seed = int(torch.empty((), dtype=torch.int64).random_())
    if total > 0:
        nb = lib.spt_sparse_sample_workspace_bytes(G)
        ws = _ws_bytes(nb, dev)
        with torch.cuda.device(dev):
            _lib.check(lib.spt_sparse_sample(
                _p(seg.ptr), _p(seg.perm), G, _p(n_samples), _p(ptr_samples), _p(elem_ids),
                int(seed) & 0xFFFFFFFFFFFFFFFF, _p(out), _p(ws), nb, _stream()),
                "spt_sparse_sample")
        _count(2)
    return (out, ptr_samples.contiguous()) if return_pointers else out


def _flags_to_index(flags, n, extra=None):
    """Ascending positions of the non-zero entries of an int32 flag vector [n+1] (two-phase
    `where`, csrc/select.cu).  `extra`: a small int tensor read back in the same host read."""
    lib = _lib.load()
    dev = flags.device
    slot = torch.empty(n + 1, dtype=torch.int32, device=dev)
    counts = torch.empty(2, dtype=torch.int64, device=dev)
    nb = lib.spt_where_workspace_bytes(n)
    ws = _ws_bytes(nb, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.spt_where_count(_p(flags), n, _p(slot), _p(counts), _p(ws), nb,
                                       _stream()), "spt_where_count")
        _count(4)
        if extra is None:
            total, extra_host = _read_counts(counts, "where"), None
        else:
            host = torch.cat((counts, extra.long().view(-1))).tolist()
            total, extra_host = host[0], host[2:]
        out = torch.empty(total, dtype=torch.int64, device=dev)
        if total > 0:
            _lib.check(lib.spt_where_write(_p(slot), n, _p(out), _stream()), "spt_where_write")
            _count()
    return out, extra_host


def radius_nodes(pos, seeds, r, k_max=10000, batch=None, cylindrical=False):
    """Sorted ids of the nodes within `r` of one of `seeds` (same `batch` item; sphere, or
    cylinder around z), at most the `k_max` closest per seed: the neighbour search of
    SampleRadiusSubgraphs (reference src/transforms/sampling.py:1196-1231, which sorts every
    seed's distances to all nodes with knn_brute_force, src/utils/neighbors.py:245-295).  One
    pass over the nodes (csrc/select.cu) + an ordered compaction."""
    lib = _lib.load()
    _require_cuda(pos, seeds, batch)
    pos = _f32c(pos)
    seeds = _i64c(seeds).view(-1)
    batch = _i64c(batch)
    N, S, dev = pos.shape[0], seeds.numel(), pos.device
    if pos.dim() != 2 or pos.shape[1] != 3:
        raise ValueError("radius_nodes: pos must be [N, 3]")
    z_offset = None
    if batch is not None:                       # neighbors.py:272-275, same tensor ops
        z = pos[:, 2] * (0 if cylindrical else 1)
        z_offset = (z.max() - z.min() + r + 1).float().contiguous()
    flags, within = None, []
    with torch.cuda.device(dev):
        for s0 in range(0, max(S, 1), 64):          # the kernel keeps <= 64 seeds in shared memory
            chunk = seeds[s0:s0 + 64]
            f = torch.empty(N + 1, dtype=torch.int32, device=dev)
            w = torch.empty(max(chunk.numel(), 1), dtype=torch.int32, device=dev)
            _lib.check(lib.spt_radius_flags(_p(pos), N, _p(batch), _p(chunk), chunk.numel(),
                                            float(r), 1 if cylindrical else 0, _p(z_offset),
                                            _p(f), _p(w), _stream()), "spt_radius_flags")
            _count()
            flags = f if flags is None else flags | f
            within.append(w)
    within = torch.cat(within)
    idx, per_seed = _flags_to_index(flags, N, extra=within)
    if max(per_seed[:S], default=0) > k_max:
        # more than k_max nodes inside the radius of a seed: the reference keeps the k_max
        # closest.  Rare (k_max defaults to 10000): done with its own tensor expression.
        mask = torch.tensor([[1, 1, 0 if cylindrical else 1]], device=dev)
        xs, xq = pos * mask, pos[seeds] * mask
        if batch is not None:
            off_s, off_q = torch.zeros_like(xs), torch.zeros_like(xq)
            off_s[:, 2] = batch * z_offset
            off_q[:, 2] = batch[seeds] * z_offset
            xs, xq = xs + off_s, xq + off_q
        d = (xs.unsqueeze(0) - xq.unsqueeze(1)).norm(dim=2)
        d, nb = d.sort(dim=1)
        d, nb = d[:, :k_max], nb[:, :k_max]
        idx = nb[d <= r].unique()
    return idx


def khop_nodes(edge_index, seeds, hops, num_nodes):
    """Sorted ids of the nodes at most `hops` edges away from `seeds`, edges taken in both
    directions: torch_geometric.utils.k_hop_subgraph(seeds, hops, to_undirected(edge_index))[0]
    (reference src/transforms/sampling.py:1080-1091); one edge pass per hop."""
    lib = _lib.load()
    _require_cuda(edge_index, seeds)
    ei = _i64c(edge_index)
    seeds = _i64c(seeds).view(-1)
    N, E, dev = int(num_nodes), int(ei.shape[1]), ei.device
    flags = torch.zeros(N + 1, dtype=torch.int32, device=dev)
    flags[seeds] = 1
    other = torch.empty_like(flags)
    with torch.cuda.device(dev):
        for _ in range(int(hops)):
            _lib.check(lib.spt_khop_expand(_p(ei), E, N, _p(flags), _p(other), _stream()),
                       "spt_khop_expand")
            flags, other = other, flags
            _count(2)
    return _flags_to_index(flags, N)[0]


SELECT_FUSED = os.environ.get('SPT_SELECT_FUSED', '1') != '0'


def set_select_fused(on=True):
    """Data.select on CUDA tensors: one native call per level (default) or the primitives
    above one by one (the form the CPU host-logic tests exercise)."""
    global SELECT_FUSED
    SELECT_FUSED = bool(on)


def data_select(num_nodes, idx, edge_index=None, sub=None, num_sub=None, update_sub=True,
                super_index=None, num_super=None, update_super=True, node_rows=(),
                edge_rows=()):
    """One level of Data.select (reference src/data/data.py:286-470) in one native call
    (`spt_data_select`): all the primitives of this section with a single host read, outputs
    carved from one arena allocation.  `sub` = (pointers, points) or None; `node_rows` /
    `edge_rows`: the attribute tensors to gather with idx / with the kept edges.
    Returns a dict: edge_index, idx_edge, sub (pointers, points), idx_sub, sub_super,
    super_index, idx_super, super_sub (pointers, points), node_rows, edge_rows (lists)."""
    import ctypes
    lib = _lib.load()
    idx = _i64c(idx).view(-1)
    dev = idx.device
    K = idx.numel()
    E = 0 if edge_index is None else int(edge_index.shape[1])
    ei = _i64c(edge_index) if E > 0 else None
    node_rows = [t.contiguous() for t in node_rows]
    edge_rows = [t.contiguous() for t in edge_rows] if E > 0 else []
    _require_cuda(idx, ei, super_index, *node_rows, *edge_rows)

    def table(ts):
        n = len(ts)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[t.data_ptr() for t in ts])
        rbs = (ctypes.c_int64 * max(n, 1))(
            *[(t.numel() // t.shape[0] if t.shape[0] > 0 else 0) * t.element_size() for t in ts])
        return ptrs, rbs

    nptr, nrb = table(node_rows)
    eptr, erb = table(edge_rows)
    a = _lib.SelectLevel()
    a.num_nodes, a.idx, a.num_selected = int(num_nodes), _p(idx), K
    a.edge_index, a.num_edges = _p(ei), E
    keep = [idx, ei, nptr, nrb, eptr, erb]
    if sub is not None:
        sp, spts = _i64c(sub[0]), _i64c(sub[1]).view(-1)
        _require_cuda(sp, spts)
        keep += [sp, spts]
        a.sub_pointers, a.sub_points, a.sub_items = _p(sp), _p(spts), spts.numel()
        a.num_sub = spts.numel() if num_sub is None else int(num_sub)
        a.update_sub = 1 if update_sub else 0
    if super_index is not None:
        si = _i64c(super_index).view(-1)
        keep.append(si)
        a.super_index, a.num_super = _p(si), int(num_super)
        a.update_super = 1 if update_super else 0
    a.num_node_rows, a.node_src, a.node_row_bytes = len(node_rows), \
        ctypes.cast(nptr, ctypes.c_void_p), ctypes.cast(nrb, ctypes.c_void_p)
    a.num_edge_rows, a.edge_src, a.edge_row_bytes = len(edge_rows), \
        ctypes.cast(eptr, ctypes.c_void_p), ctypes.cast(erb, ctypes.c_void_p)
    nb = lib.spt_data_select_arena_bytes(ctypes.addressof(a))
    arena = torch.empty(max(nb, 256), dtype=torch.uint8, device=dev)
    n_rows = len(node_rows) + len(edge_rows)
    layout = (ctypes.c_int64 * (_lib.SEL_ROWS + n_rows))()
    with torch.cuda.device(dev):
        _lib.check(lib.spt_data_select(ctypes.addressof(a), _p(arena), nb,
                                       ctypes.cast(layout, ctypes.c_void_p), _stream()),
                   "spt_data_select")
    _count(20)
    L = list(layout)

    def i64(off, n):
        return arena[off:off + 8 * n].view(torch.int64)

    def rows(off, n, like):
        nbytes = n * (like.numel() // like.shape[0] if like.shape[0] > 0 else 0) \
            * like.element_size()
        return arena[off:off + nbytes].view(like.dtype).view((n,) + tuple(like.shape[1:]))

    kept, items, parents = L[_lib.SEL_NUM_EDGES], L[_lib.SEL_NUM_ITEMS], L[_lib.SEL_NUM_PARENTS]
    out = {'node_rows': [rows(L[_lib.SEL_ROWS + i], K, t) for i, t in enumerate(node_rows)],
           'edge_rows': [rows(L[_lib.SEL_ROWS + len(node_rows) + i], kept, t)
                         for i, t in enumerate(edge_rows)]}
    if E > 0:
        out['edge_index'] = i64(L[_lib.SEL_EDGE_INDEX], 2 * kept).view(2, kept)
        out['idx_edge'] = i64(L[_lib.SEL_IDX_EDGE], kept)
    if sub is not None:
        out['sub'] = (i64(L[_lib.SEL_SUB_POINTERS], K + 1), i64(L[_lib.SEL_SUB_POINTS], items))
        if update_sub:
            out['idx_sub'] = i64(L[_lib.SEL_IDX_SUB], items)
            out['sub_super'] = i64(L[_lib.SEL_SUB_SUPER], items)
            if _DEBUG_INDEX:
                distinct, bad = i64(L[_lib.SEL_SUB_COUNTS], 2).tolist()
                if bad or distinct != items:
                    raise IndexError("select: `sub` points are not distinct ids below num_sub")
    if super_index is not None:
        out['super_index'] = i64(L[_lib.SEL_SUPER_INDEX], K)
        if update_super:
            out['idx_super'] = i64(L[_lib.SEL_IDX_SUPER], parents)
            out['super_sub'] = (i64(L[_lib.SEL_SUPER_SUB_POINTERS], parents + 1),
                                i64(L[_lib.SEL_SUPER_SUB_POINTS], K))
    return out


def take_rows_multi(tensors, idx):
    """[t[idx] for t in tensors] in one launch (per 16 tensors): all node-level or all
    edge-level attributes of a Data object (reference src/data/data.py:420-463)."""
    import ctypes
    lib = _lib.load()
    if len(tensors) == 0:
        return []
    _require_cuda(idx, *tensors)
    tensors = [t.contiguous() for t in tensors]
    idx = _i64c(idx).view(-1)
    K, n = idx.numel(), len(tensors)
    outs = [torch.empty((K,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            for t in tensors]
    row_bytes = [(t.numel() // t.shape[0] if t.shape[0] > 0 else 0) * t.element_size()
                 for t in tensors]
    if K > 0:
        srcs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
        dsts = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
        rbs = (ctypes.c_int64 * n)(*row_bytes)
        with torch.cuda.device(idx.device):
            _lib.check(lib.spt_gather_rows_multi(
                ctypes.cast(srcs, ctypes.c_void_p), ctypes.cast(dsts, ctypes.c_void_p),
                ctypes.cast(rbs, ctypes.c_void_p), n, _p(idx), K, _stream()),
                "spt_gather_rows_multi")
        _count((n + 15) // 16)
    return outs


def take_rows(t, idx):
    """t[idx] along dim 0 for a CUDA tensor of any dtype (`item[idx]`, reference
    src/data/data.py:447-459); idx int64, in range (Data.select validates it first)."""
    lib = _lib.load()
    _require_cuda(t, idx)
    t = t.contiguous()
    idx = _i64c(idx).view(-1)
    K = idx.numel()
    out = torch.empty((K,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    row_bytes = (t.numel() // t.shape[0] if t.shape[0] > 0 else 0) * t.element_size()
    if K > 0 and row_bytes > 0:
        with torch.cuda.device(t.device):
            _lib.check(lib.spt_gather_rows_bytes(_p(t), row_bytes, _p(idx), K, _p(out),
                                                 _stream()), "spt_gather_rows_bytes")
        _count()
    return out


def node_size(super_index, num_parents, child_size=None):
    """NAG.get_sub_size step (src/data/nag.py:59-110): exact int64 sums."""
    lib = _lib.load()
    _require_cuda(super_index, child_size)
    super_index = _i64c(super_index)
    seg = segment_index(super_index, num_parents)
    out = torch.empty(num_parents, dtype=torch.int64, device=super_index.device)
    vals = _i64c(child_size)
    with torch.cuda.device(super_index.device):
        _lib.check(lib.spt_segment_sum_i64(_p(vals), _p(seg.ptr), _p(seg.perm), num_parents,
                                           _p(out), _stream()), "spt_segment_sum_i64")
    _count()
    return out


# ---------------------------------------------------------------------------
# UnitSphereNorm
# ---------------------------------------------------------------------------
def unit_sphere_norm(pos, idx=None, w=None, num_super=None):
    """(pos_normalised [N,3], diameter [Np,1]); src/nn/norm.py:67-138. No grad
    (pos never requires grad on the reference path)."""
    lib = _lib.load()
    _require_cuda(pos, idx, w)
    pos = _f32c(pos.detach())
    N = pos.shape[0]
    dev = pos.device
    wf = None if w is None else w.detach().float().contiguous()
    if idx is None:
        Np = 1
        # built on the device (no H2D copy: keeps the call CUDA-graph capturable)
        ptr = torch.arange(2, dtype=torch.int32, device=dev) * N
        points, parent = None, None
    else:
        idx = _i64c(idx)
        Np = int(num_super) if num_super is not None else num_segments(idx)
        seg = segment_index(idx, Np)
        ptr, points, parent = seg.ptr, seg.perm, idx
    out = torch.empty_like(pos)
    diam = torch.empty((Np, 1), dtype=torch.float32, device=dev)
    nbytes = lib.spt_unitsphere_workspace_bytes(Np)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev), _timed('unitsphere', N=N, Np=Np):
        _lib.check(lib.spt_unitsphere_fwd(_p(pos), _p(parent), _p(ptr), _p(points), _p(wf),
                                          N, Np, _p(out), _p(diam), _p(ws), nbytes,
                                          _stream()), "spt_unitsphere_fwd")
    _count(2)
    return out, diam


# ---------------------------------------------------------------------------
# GraphNorm
# ---------------------------------------------------------------------------
class _GraphNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, mean_scale, batch, B, eps, act_slope):
        lib = _lib.load()
        N, C = x.shape
        dev = x.device
        y = torch.empty_like(x)
        mean = torch.empty((B, C), dtype=torch.float32, device=dev)
        rstd = torch.empty((B, C), dtype=torch.float32, device=dev)
        nbytes = lib.spt_graphnorm_workspace_bytes(B, C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev), _timed('graphnorm_fwd', N=N, C=C, B=B):
            _lib.check(lib.spt_graphnorm_fwd(_p(x), _p(batch), N, C, B, _p(weight), _p(bias),
                                             _p(mean_scale), eps, act_slope, _p(y), _p(mean),
                                             _p(rstd), _p(ws), nbytes, _stream()),
                       "spt_graphnorm_fwd")
        _count(4)
        ctx.B = B
        ctx.has_batch = batch is not None
        ctx.act_slope = act_slope
        ctx.save_for_backward(x, weight, mean_scale, mean, rstd,
                              batch if batch is not None else mean,
                              y if act_slope != 1.0 else mean)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean_scale, mean, rstd, batch, yact = ctx.saved_tensors
        if not ctx.has_batch:
            batch = None
        if ctx.act_slope == 1.0:
            yact = None
        lib = _lib.load()
        dy = _f32c(dy)
        N, C = x.shape
        dev = x.device
        dx = torch.empty_like(x)
        dw = torch.empty(C, dtype=torch.float32, device=dev)
        db = torch.empty(C, dtype=torch.float32, device=dev)
        dms = torch.empty(C, dtype=torch.float32, device=dev)
        nbytes = lib.spt_graphnorm_workspace_bytes(ctx.B, C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev), _timed('graphnorm_bwd', N=N, C=C, B=ctx.B,
                                            act=yact is not None):
            _lib.check(lib.spt_graphnorm_bwd(_p(x), _p(dy), _p(batch), N, C, ctx.B, _p(weight),
                                             _p(mean_scale), _p(mean), _p(rstd), _p(yact),
                                             ctx.act_slope, _p(dx), _p(dw), _p(db), _p(dms),
                                             _p(ws), nbytes, _stream()), "spt_graphnorm_bwd")
        _count(3)
        return dx, dw, db, dms, None, None, None, None


class _GroupNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, batch, B, G, eps, eps_outside):
        lib = _lib.load()
        N, C = x.shape
        dev = x.device
        y = torch.empty_like(x)
        mean = torch.empty((B, C), dtype=torch.float32, device=dev)
        rstd = torch.empty((B, C), dtype=torch.float32, device=dev)
        nbytes = lib.spt_graphnorm_workspace_bytes(B, C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.spt_groupnorm_fwd(_p(x), _p(batch), N, C, B, G, _p(weight), _p(bias),
                                             eps, eps_outside, _p(y), _p(mean), _p(rstd),
                                             _p(ws), nbytes, _stream()), "spt_groupnorm_fwd")
        _count(6)
        ctx.cfg = (B, G, eps, eps_outside, batch is not None, weight is not None)
        ctx.save_for_backward(x, weight if weight is not None else mean, mean, rstd,
                              batch if batch is not None else mean)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd, batch = ctx.saved_tensors
        B, G, eps, eps_outside, has_batch, affine = ctx.cfg
        if not has_batch:
            batch = None
        if not affine:
            weight = None
        lib = _lib.load()
        dy = _f32c(dy)
        N, C = x.shape
        dev = x.device
        dx = torch.empty_like(x)
        dw = torch.empty(C, dtype=torch.float32, device=dev) if affine else None
        db = torch.empty(C, dtype=torch.float32, device=dev) if affine else None
        nbytes = lib.spt_graphnorm_workspace_bytes(B, C)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.spt_groupnorm_bwd(_p(x), _p(dy), _p(batch), N, C, B, G, _p(weight),
                                             _p(mean), _p(rstd), eps, eps_outside, _p(dx),
                                             _p(dw), _p(db), _p(ws), nbytes, _stream()),
                       "spt_groupnorm_bwd")
        _count(3)
        return dx, dw, db, None, None, None, None, None


def group_norm(x, weight=None, bias=None, batch=None, batch_size=None, num_groups=1, eps=1e-5,
               eps_outside=False):
    """Graph-wise GroupNorm (reference src/nn/norm.py:181-218); `num_groups=1` is PyG
    LayerNorm(mode='graph').  `eps_outside` = PyG LayerNorm called without `batch`."""
    _require_cuda(x, weight, bias, batch)
    batch = _i64c(batch)
    B = int(batch_size) if batch_size is not None else num_segments(batch)
    assert (weight is None) == (bias is None), "affine needs both weight and bias"
    return _GroupNorm.apply(_f32c(x), None if weight is None else _f32c(weight),
                            None if bias is None else _f32c(bias), batch, B, int(num_groups),
                            float(eps), int(bool(eps_outside)))


def graph_norm(x, weight, bias, mean_scale, batch=None, batch_size=None, eps=1e-5,
               act_slope=1.0):
    """PyG GraphNorm semantics; `act_slope != 1` fuses the
    LeakyReLU that follows the norm in the reference MLP (src/nn/mlp.py:41-55)."""
    _require_cuda(x, weight, bias, mean_scale, batch)
    batch = _i64c(batch)
    B = int(batch_size) if batch_size is not None else num_segments(batch)
    return _GraphNorm.apply(_f32c(x), _f32c(weight), _f32c(bias), _f32c(mean_scale), batch,
                            B, float(eps), float(act_slope))


# ---------------------------------------------------------------------------
# fused attention core
# ---------------------------------------------------------------------------
def _split_shape(H, D, Dv, F):
    """shape families of the split attention kernels: the benchmark family (4 heads of 32
    channels) and the shipped head layout (16 heads, C = 64 or 128)"""
    return (H, D, Dv, F) == (4, 4, 32, 32) or (H == 16 and D == 4 and Dv in (4, 8) and F == 32)


class _AttnCore(torch.autograd.Function):
    """q/k/v come either fused as qkv [N, 2HD+C] (kv=None) or as q [R,HD] and
    kv [T, HD+C].  Returns (agg_v [R,C], abar [R,H,F] or None, sump [R,H])."""

    @staticmethod
    def forward(ctx, qsrc, kv, a, Wq, bq, Wk, bk, g, H, D, scale_mode, scale_value,
                want_abar, q_row_add=None, q_tgt_add=None, k_row_add=None, drop_mask=None,
                v_bf16=None):
        lib = _lib.load()
        dev = qsrc.device
        HD = H * D
        fused = kv is None
        if fused:
            ld = qsrc.shape[1]
            C = ld - 2 * HD
            qp, kp, vp = qsrc.data_ptr(), qsrc.data_ptr() + 4 * HD, qsrc.data_ptr() + 8 * HD
            ldq = ldk = ldv = ld
        else:
            C = kv.shape[1] - HD
            qp, kp, vp = qsrc.data_ptr(), kv.data_ptr(), kv.data_ptr() + 4 * HD
            ldq, ldk, ldv = qsrc.shape[1], kv.shape[1], kv.shape[1]
        Dv = C // H
        F = a.shape[1] if a is not None else 0
        R = g.num_rows
        agg = torch.empty((R, C), dtype=torch.float32, device=dev)
        abar = (torch.empty((R, H, F), dtype=torch.float32, device=dev)
                if (want_abar and a is not None) else None)
        sump = torch.empty((R, H), dtype=torch.float32, device=dev)
        m = torch.empty((R, H), dtype=torch.float32, device=dev)
        z = torch.empty((R, H), dtype=torch.float32, device=dev)
        has_ex = any(t is not None for t in (q_row_add, q_tgt_add, k_row_add, drop_mask))
        ex = None
        logits = None
        if has_ex:
            ex = _lib.AttnExtras(_p(q_row_add), _p(q_tgt_add), _p(k_row_add), _p(drop_mask),
                                 None, None, None, None)
        elif ATTN_SPLIT and a is not None and _split_shape(H, D, Dv, F) and g.E > 0:
            # workspace of the split kernels: base-2 logits [E, H] (kept for the backward)
            logits = torch.empty((g.E, H), dtype=torch.float32, device=dev)
            ex = _lib.AttnExtras(None, None, None, None, None, None, None, None,
                                 _p(logits), _p(g.edge_row), None)
            if v_bf16 is not None and fused:
                # bf16 storage: the row passes gather the value rows from the bf16 copy of qkv
                ex.v_bf16 = v_bf16.data_ptr() + 2 * (2 * HD)
                ex.ldv_bf16 = v_bf16.shape[1]
        with torch.cuda.device(dev), _timed('attn_fwd', R=R, E=g.E, H=H, D=D, Dv=Dv, F=F,
                                            abar=abar is not None):
            _lib.check(lib.spt_attn_fwd_ex(qp, ldq, kp, ldk, vp, ldv, _p(a), _p(g.rowptr),
                                           _p(g.col), R, g.E, H, D, Dv, F, _p(Wq), _p(bq),
                                           _p(Wk), _p(bk), scale_mode, scale_value, _p(agg),
                                           _p(abar), _p(sump), _p(m), _p(z),
                                           ctypes.byref(ex) if ex is not None else None,
                                           _stream()), "spt_attn_fwd")
        _count(2 if logits is not None else 1)        # edge pass + row pass, or one fused kernel
        ctx.g, ctx.H, ctx.D, ctx.Dv, ctx.F = g, H, D, Dv, F
        ctx.scale_mode, ctx.scale_value = scale_mode, scale_value
        ctx.fused = fused
        ctx.logits = logits          # split kernels: the backward row pass reads them back
        ctx.v_bf16 = v_bf16 if (logits is not None and fused) else None
        ctx.opt = (kv is not None, a is not None, Wq is not None, bq is not None,
                   Wk is not None, bk is not None, abar is not None)
        ctx.ex = (q_row_add is not None, q_tgt_add is not None, k_row_add is not None,
                  drop_mask is not None)
        dummy = m
        ctx.save_for_backward(qsrc, kv if kv is not None else dummy,
                              a if a is not None else dummy,
                              Wq if Wq is not None else dummy,
                              bq if bq is not None else dummy,
                              Wk if Wk is not None else dummy,
                              bk if bk is not None else dummy, m, z, agg,
                              abar if abar is not None else dummy,
                              q_row_add if q_row_add is not None else dummy,
                              q_tgt_add if q_tgt_add is not None else dummy,
                              k_row_add if k_row_add is not None else dummy,
                              drop_mask if drop_mask is not None else dummy)
        if drop_mask is None:
            # sum_e p_e is the constant 1 (0 for an empty row): no gradient.  Under attention
            # dropout sum_e p_e mask_e is not, and the caller's v-RPE bias multiplies it.
            ctx.mark_non_differentiable(sump)
        ctx.sump = sump if drop_mask is not None else None
        if abar is None:
            return agg, None, sump
        return agg, abar, sump

    @staticmethod
    def backward(ctx, d_agg, d_abar, _d_sump):
        lib = _lib.load()
        (qsrc, kv, a, Wq, bq, Wk, bk, m, z, agg, abar, q_row_add, q_tgt_add, k_row_add,
         drop_mask) = ctx.saved_tensors
        has_kv, has_a, has_Wq, has_bq, has_Wk, has_bk, has_abar = ctx.opt
        has_qr, has_qt, has_kr, has_dm = ctx.ex
        q_row_add = q_row_add if has_qr else None
        q_tgt_add = q_tgt_add if has_qt else None
        k_row_add = k_row_add if has_kr else None
        drop_mask = drop_mask if has_dm else None
        kv = kv if has_kv else None
        a = a if has_a else None
        Wq = Wq if has_Wq else None
        bq = bq if has_bq else None
        Wk = Wk if has_Wk else None
        bk = bk if has_bk else None
        abar = abar if has_abar else None
        g, H, D, Dv, F = ctx.g, ctx.H, ctx.D, ctx.Dv, ctx.F
        HD, C = H * D, H * Dv
        dev = qsrc.device
        d_agg = _f32c(d_agg) if d_agg is not None else torch.zeros_like(agg)
        d_abar = _f32c(d_abar) if (d_abar is not None and abar is not None) else None
        if ctx.fused:
            dqkv = torch.empty_like(qsrc)
            ld = qsrc.shape[1]
            qp, kp, vp = qsrc.data_ptr(), qsrc.data_ptr() + 4 * HD, qsrc.data_ptr() + 8 * HD
            dqp, dkp, dvp = dqkv.data_ptr(), dqkv.data_ptr() + 4 * HD, dqkv.data_ptr() + 8 * HD
            ldq = ldk = ldv = lddq = lddk = lddv = ld
            dkv = None
        else:
            dqkv = torch.empty_like(qsrc)
            dkv = torch.empty_like(kv)
            qp, kp, vp = qsrc.data_ptr(), kv.data_ptr(), kv.data_ptr() + 4 * HD
            dqp, dkp, dvp = dqkv.data_ptr(), dkv.data_ptr(), dkv.data_ptr() + 4 * HD
            ldq, ldk, ldv = qsrc.shape[1], kv.shape[1], kv.shape[1]
            lddq, lddk, lddv = ldq, ldk, ldv
        need = ctx.needs_input_grad
        da = torch.empty_like(a) if (a is not None and need[2]) else None
        packed_dw = (a is not None and Wq is not None and Wk is not None
                     and Wq.shape == Wk.shape and (bq is None) == (bk is None))
        dW2 = db2 = None
        if packed_dw:
            # one contiguous [2HD, F] / [2HD] pair: lets the library run d[Wq;Wk] = G^T a as a
            # single tensor-core gemm_tn
            dW2 = zero_pool.take((2 * Wq.shape[0], Wq.shape[1]), dev)
            dWq, dWk = dW2[:Wq.shape[0]], dW2[Wq.shape[0]:]
            if bq is not None:
                db2 = zero_pool.take((2 * Wq.shape[0],), dev)
                dbq, dbk = db2[:Wq.shape[0]], db2[Wq.shape[0]:]
            else:
                dbq = dbk = None
        else:
            dWq = torch.zeros_like(Wq) if (Wq is not None and a is not None) else None
            dbq = torch.zeros_like(bq) if (bq is not None and dWq is not None) else None
            dWk = torch.zeros_like(Wk) if (Wk is not None and a is not None) else None
            dbk = torch.zeros_like(bk) if (bk is not None and dWk is not None) else None
        E = g.E
        Pb = torch.empty((max(E, 1), H), dtype=torch.float32, device=dev)
        G = torch.empty((max(E, 1), 2 * HD), dtype=torch.float32, device=dev)
        meta = dict(R=g.num_rows, T=g.num_targets, E=E, H=H, D=D, Dv=Dv, F=F,
                    abar=abar is not None, da=da is not None)
        has_ex = has_qr or has_qt or has_kr or has_dm
        d_qr = torch.empty_like(q_row_add) if has_qr else None
        d_kr = torch.empty_like(k_row_add) if has_kr else None
        d_qt = torch.empty_like(q_tgt_add) if has_qt else None
        ex = None
        if has_ex:
            d_sump = None
            if has_dm and _d_sump is not None and ctx.sump is not None:
                d_sump = _f32c(_d_sump)
            ex = _lib.AttnExtras(_p(q_row_add), _p(q_tgt_add), _p(k_row_add), _p(drop_mask),
                                 _p(d_qr), _p(d_kr), _p(d_sump),
                                 _p(ctx.sump) if d_sump is not None else None)
        elif ctx.logits is not None and ATTN_SPLIT:
            ws_ds = torch.empty((E, H), dtype=torch.float32, device=dev)
            ex = _lib.AttnExtras(None, None, None, None, None, None, None, None,
                                 _p(ctx.logits), _p(g.edge_row), _p(ws_ds))
            if ctx.v_bf16 is not None:
                ex.v_bf16 = ctx.v_bf16.data_ptr() + 2 * (2 * HD)
                ex.ldv_bf16 = ctx.v_bf16.shape[1]
        # d[Wq;Wk] = G^T a (and the bias sums) beside the targets pass: both only read G, so the
        # product goes to a side stream (fork / join).  Not while per-launch events are being
        # recorded (bench --kernels-out): the table wants serialised times.
        side_dw = (ATTN_DW_SIDE_STREAM and _TIMING is None and packed_dw and not has_ex
                   and E >= 2048 and F % 4 == 0)
        with torch.cuda.device(dev):
            with _timed('attn_bwd_rows', **meta):
                _lib.check(lib.spt_attn_bwd_rows_ex(
                    qp, ldq, kp, ldk, vp, ldv, _p(a), _p(g.rowptr), _p(g.col), g.num_rows, E,
                    H, D, Dv, F, _p(Wq), _p(bq), _p(Wk), _p(bk), ctx.scale_mode,
                    ctx.scale_value, _p(m), _p(z), _p(agg), _p(abar), _p(d_agg), _p(d_abar),
                    dqp, lddq, _p(da), None if side_dw else _p(dWq), None if side_dw else _p(dbq),
                    None if side_dw else _p(dWk), None if side_dw else _p(dbk), _p(Pb), _p(G),
                    ctypes.byref(ex) if ex is not None else None, _stream()),
                    "spt_attn_bwd_rows")
            if side_dw:
                cur, side = torch.cuda.current_stream(dev), _side_stream(dev)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    _lib.check(lib.spt_gemm_tn_acc(_p(G), E, 2 * HD, 2 * HD, _p(a), F, F,
                                                   _p(dW2), F, _p(db2) if bq is not None else None,
                                                   side.cuda_stream), "spt_gemm_tn_acc")
            with _timed('attn_bwd_targets', **meta):
                _lib.check(lib.spt_attn_bwd_targets_ex(
                    _p(g.csc_ptr), _p(g.csc_src), _p(g.csc2csr), g.num_targets, E, H, D, Dv,
                    _p(Pb), _p(G), _p(d_agg), dkp, lddk, dvp, lddv, _p(d_qt), _stream()),
                    "spt_attn_bwd_targets")
            if side_dw:
                cur.wait_stream(side)
            # rows (row pass + edge pass, or the fused kernel) + d[Wq;Wk] product + targets
            _count(4 if (ex is not None and not has_ex) else 3)
        if bq is not None and dbq is None:
            dbq = torch.zeros_like(bq)
        if bk is not None and dbk is None:
            dbk = torch.zeros_like(bk)
        if Wq is not None and dWq is None:
            dWq = torch.zeros_like(Wq)
        if Wk is not None and dWk is None:
            dWk = torch.zeros_like(Wk)
        return (dqkv, dkv, da, dWq, dbq, dWk, dbk, None, None, None, None, None, None,
                d_qr, d_qt, d_kr, None, None)


# ---------------------------------------------------------------------------
# bf16 storage of the attention operands (BASELINE cfg 3)
# ---------------------------------------------------------------------------
ATTN_STORAGE = 'fp32'
_bf16_cache = _IdentityCache()
# split attention kernels (csrc/attention_split.cuh, csrc/attention_umma.cuh): one edge-parallel
# pass on the tensor cores + one row-parallel pass; SPT_ATTN_SPLIT=0 keeps the fused row-tile
# kernels (the A/B of tools/run_attn.py and the parity tests use the switch)
ATTN_SPLIT = os.environ.get('SPT_ATTN_SPLIT', '1') != '0'


def set_attention_split(on):
    global ATTN_SPLIT
    ATTN_SPLIT = bool(on)


def set_attention_storage(mode):
    """'fp32' (default) or 'bf16' (BASELINE cfg 3), for the shape family H=4, D=4, Dv=32, F=32 with
    k and q RPE; other shapes keep the fp32 kernels.  fp32 accumulation, outputs and gradients
    in both modes.  With the split kernels (default) 'bf16' makes the row passes gather the
    value rows from a bf16 copy of the fused projections; with the fused row-tile kernels
    (set_attention_split(False)) qkv and the CSR-ordered edge features are both stored as bf16."""
    global ATTN_STORAGE
    if mode not in ('fp32', 'bf16'):
        raise ValueError(mode)
    ATTN_STORAGE = mode


def cast_bf16(x):
    """fp32 -> bf16 copy (round to nearest even) as a torch.bfloat16 tensor; no gradient."""
    lib = _lib.load()
    x = _f32c(x.detach())
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.spt_cast_bf16(_p(x), x.numel(), _p(out), _stream()), "spt_cast_bf16")
    _count()
    return out


class _AttnCoreBf16(torch.autograd.Function):
    """_AttnCore with q / k / v / edge features stored as bf16 (see spt_attn_fwd_bf16)."""

    @staticmethod
    def forward(ctx, qkv, a, Wq, bq, Wk, bk, g, H, D, scale_mode, scale_value, want_abar):
        lib = _lib.load()
        dev = qkv.device
        HD = H * D
        ld = qkv.shape[1]
        C = ld - 2 * HD
        Dv, F, R = C // H, a.shape[1], g.num_rows
        qb = cast_bf16(qkv)
        # the blocks of a stage share the edge features: one bf16 copy per tensor
        ab = _bf16_cache.get(a, "bf16", lambda: cast_bf16(a))
        agg = torch.empty((R, C), dtype=torch.float32, device=dev)
        abar = torch.empty((R, H, F), dtype=torch.float32, device=dev) if want_abar else None
        sump = torch.empty((R, H), dtype=torch.float32, device=dev)
        m = torch.empty((R, H), dtype=torch.float32, device=dev)
        z = torch.empty((R, H), dtype=torch.float32, device=dev)
        base = qb.data_ptr()
        with torch.cuda.device(dev), _timed('attn_fwd', R=R, E=g.E, H=H, D=D, Dv=Dv, F=F,
                                            abar=abar is not None, bf16=True):
            _lib.check(lib.spt_attn_fwd_bf16(base, ld, base + 2 * HD, ld, base + 4 * HD, ld,
                                             _p(ab), _p(g.rowptr), _p(g.col), R, g.E, H, D, Dv, F,
                                             _p(Wq), _p(bq), _p(Wk), _p(bk), scale_mode,
                                             scale_value, _p(agg), _p(abar), _p(sump), _p(m),
                                             _p(z), _stream()), "spt_attn_fwd_bf16")
        _count()
        ctx.g, ctx.H, ctx.D, ctx.Dv, ctx.F = g, H, D, Dv, F
        ctx.scale_mode, ctx.scale_value = scale_mode, scale_value
        ctx.has_abar = abar is not None
        ctx.save_for_backward(qb, ab, a, Wq, bq, Wk, bk, m, z, agg,
                              abar if abar is not None else m)
        ctx.mark_non_differentiable(sump)
        return agg, abar, sump

    @staticmethod
    def backward(ctx, d_agg, d_abar, _d_sump):
        lib = _lib.load()
        qb, ab, a, Wq, bq, Wk, bk, m, z, agg, abar = ctx.saved_tensors
        abar = abar if ctx.has_abar else None
        g, H, D, Dv, F = ctx.g, ctx.H, ctx.D, ctx.Dv, ctx.F
        HD = H * D
        dev = qb.device
        ld = qb.shape[1]
        d_agg = _f32c(d_agg) if d_agg is not None else torch.zeros_like(agg)
        d_abar = _f32c(d_abar) if (d_abar is not None and abar is not None) else None
        dqkv = torch.empty(qb.shape, dtype=torch.float32, device=dev)
        da = torch.empty(a.shape, dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] \
            else None
        E = g.E
        Pb = torch.empty((max(E, 1), H), dtype=torch.float32, device=dev)
        G = torch.empty((max(E, 1), 2 * HD), dtype=torch.float32, device=dev)
        dW2 = zero_pool.take((2 * HD, F), dev)
        db2 = zero_pool.take((2 * HD,), dev)
        base = qb.data_ptr()
        dq_p = dqkv.data_ptr()
        meta = dict(R=g.num_rows, T=g.num_targets, E=E, H=H, D=D, Dv=Dv, F=F,
                    abar=abar is not None, da=da is not None, bf16=True)
        with torch.cuda.device(dev):
            with _timed('attn_bwd_rows', **meta):
                _lib.check(lib.spt_attn_bwd_rows_bf16(
                    base, ld, base + 2 * HD, ld, base + 4 * HD, ld, _p(ab), _p(g.rowptr),
                    _p(g.col), g.num_rows, E, H, D, Dv, F, _p(Wq), _p(bq), _p(Wk), _p(bk),
                    ctx.scale_mode, ctx.scale_value, _p(m), _p(z), _p(agg), _p(abar), _p(d_agg),
                    _p(d_abar), dq_p, ld, _p(da), _p(Pb), _p(G), _stream()),
                    "spt_attn_bwd_rows_bf16")
                # d[Wq;Wk] = G^T a, d[bq;bk] = colsum(G): tcgen05 gemm_tn on the fp32 features
                if E > 0:
                    _lib.check(lib.spt_gemm_tn_acc(_p(G), E, 2 * HD, 2 * HD, _p(a), F, F,
                                                   _p(dW2), F, _p(db2), _stream()),
                               "spt_gemm_tn_acc")
            with _timed('attn_bwd_targets', **meta):
                _lib.check(lib.spt_attn_bwd_targets(
                    _p(g.csc_ptr), _p(g.csc_src), _p(g.csc2csr), g.num_targets, E, H, D, Dv,
                    _p(Pb), _p(G), _p(d_agg), dq_p + 4 * HD, ld, dq_p + 8 * HD, ld, _stream()),
                    "spt_attn_bwd_targets")
            _count(3)
        return (dqkv, da, dW2[:HD], db2[:HD], dW2[HD:], db2[HD:], None, None, None, None, None,
                None)


def _bf16_ok(qsrc, kv, a, Wq, bq, Wk, bk, H, D, extras):
    if ATTN_STORAGE != 'bf16' or kv is not None or a is None or extras:
        return False
    if Wq is None or Wk is None or bq is None or bk is None:
        return False
    C = qsrc.shape[1] - 2 * H * D
    return (H == 4 and D == 4 and C == 128 and a.shape[1] == 32 and qsrc.shape[1] % 4 == 0
            and tuple(Wq.shape) == (16, 32) and tuple(Wk.shape) == (16, 32)
            and a.shape[0] > 0)


def attention_core(qsrc, kv, a_csr, Wq, bq, Wk, bk, graph, num_heads, qk_dim,
                   scale_mode=SCALE_D_TIMES_G, scale_value=1.0, want_abar=True,
                   q_row_add=None, q_tgt_add=None, k_row_add=None, drop_mask=None):
    """See include/spt_b200.h:spt_attn_fwd(_ex).  `a_csr` and `drop_mask` must already be in
    CSR order (permute_rows(edge_attr, graph.perm)).  The optional addends / mask are the
    `spt_attn_extras` (node-difference RPE, attention dropout)."""
    _require_cuda(qsrc, kv, a_csr, Wq, bq, Wk, bk, q_row_add, q_tgt_add, k_row_add, drop_mask)
    extras = any(t is not None for t in (q_row_add, q_tgt_add, k_row_add, drop_mask))
    if _bf16_ok(qsrc, kv, a_csr, Wq, bq, Wk, bk, int(num_heads), int(qk_dim), extras):
        if ATTN_SPLIT:
            # split kernels: the gathered value rows (the dominant stream) come from a bf16 copy
            # of the fused projections; q, k and the edge features stay fp32
            qf = _f32c(qsrc)
            return _AttnCore.apply(qf, None, _f32c(a_csr), _f32c(Wq), _f32c(bq), _f32c(Wk),
                                   _f32c(bk), graph, int(num_heads), int(qk_dim),
                                   int(scale_mode), float(scale_value), bool(want_abar),
                                   None, None, None, None, cast_bf16(qf))
        return _AttnCoreBf16.apply(_f32c(qsrc), _f32c(a_csr), _f32c(Wq), _f32c(bq), _f32c(Wk),
                                   _f32c(bk), graph, int(num_heads), int(qk_dim),
                                   int(scale_mode), float(scale_value), bool(want_abar))
    return _AttnCore.apply(_f32c(qsrc), _f32c(kv), _f32c(a_csr), _f32c(Wq), _f32c(bq),
                           _f32c(Wk), _f32c(bk), graph, int(num_heads), int(qk_dim),
                           int(scale_mode), float(scale_value), bool(want_abar),
                           _f32c(q_row_add), _f32c(q_tgt_add), _f32c(k_row_add),
                           _f32c(drop_mask))


class _ValueRpe(torch.autograd.Function):
    """y = agg + blockdiag(Wv) . abar + bv (x) sump  — the value RPE of SelfAttentionBlock
    (reference src/nn/attention.py:294-301) applied to the per-row sums the attention kernel
    emits; one tensor-core GEMM + 2 small kernels forward, GEMM dX + GEMM dW + 2 small kernels
    backward (csrc/vrpe.cu)."""

    @staticmethod
    def forward(ctx, agg, abar, sump, Wv, bv, H, share):
        lib = _lib.load()
        N, C = agg.shape
        F = abar.shape[2]
        Dv = C // H
        dev = agg.device
        Wbd = torch.empty((C, H * F), dtype=torch.float32, device=dev)
        y = torch.empty_like(agg)
        with torch.cuda.device(dev):
            _lib.check(lib.spt_vrpe_blockdiag(_p(Wv), H, Dv, F, share, 0, _p(Wbd), _stream()),
                       "spt_vrpe_blockdiag")
            # small levels: the plain library GEMM (same rule as ops.linear)
            rv = _gemm_nt(abar.view(N, H * F), Wbd) if N >= LINEAR_TC_MIN_ROWS \
                else abar.view(N, H * F) @ Wbd.t()
            with _timed('vrpe_epilogue', N=N, C=C):
                _lib.check(lib.spt_vrpe_epilogue(_p(agg), _p(rv), _p(sump), _p(bv), N, H, Dv, share,
                                                 _p(y), _stream()), "spt_vrpe_epilogue")
        _count(2)
        ctx.save_for_backward(abar, sump, Wv, bv if bv is not None else sump)
        ctx.cfg = (H, share, bv is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        abar, sump, Wv, bv = ctx.saved_tensors
        H, share, has_b = ctx.cfg
        lib = _lib.load()
        dy = _f32c(dy)
        N, C = dy.shape
        F = abar.shape[2]
        Dv = C // H
        dev = dy.device
        need = ctx.needs_input_grad
        d_abar = dWv = dbv = d_sump = None
        if need[2] and has_b:   # only under attention dropout (sump is non-differentiable else)
            bfull = bv.repeat(H) if share else bv
            d_sump = (dy.view(N, H, Dv) * bfull.view(1, H, Dv)).sum(-1)
        with torch.cuda.device(dev):
            if need[1]:
                WbdT = torch.empty((H * F, C), dtype=torch.float32, device=dev)
                _lib.check(lib.spt_vrpe_blockdiag(_p(Wv), H, Dv, F, share, 1, _p(WbdT),
                                                  _stream()), "spt_vrpe_blockdiag")
                d_abar = (_gemm_nt(dy, WbdT) if N >= LINEAR_TC_MIN_ROWS
                          else dy @ WbdT.t()).view(N, H, F)
                _count()
            if need[3] or (has_b and need[4]):
                a2 = abar.view(N, H * F)
                if N >= LINEAR_TC_MIN_ROWS:
                    dWbd = zero_pool.take((C, H * F), dev)
                    with _timed('gemm_tn', M=N, N=C, K=H * F):
                        _lib.check(lib.spt_gemm_tn_acc(_p(dy), N, C, dy.stride(0), _p(a2), H * F,
                                                       a2.stride(0), _p(dWbd), H * F, None,
                                                       _stream()), "spt_gemm_tn_acc")
                else:
                    dWbd = (dy.t() @ a2).contiguous()
                dWv = zero_pool.take(tuple(Wv.shape), dev)
                dbv = zero_pool.take(tuple(bv.shape), dev) if has_b else None
                _lib.check(lib.spt_vrpe_bwd_params(_p(dy), _p(sump), _p(dWbd), N, H, Dv, F, share,
                                                   _p(dWv), _p(dbv), _stream()),
                           "spt_vrpe_bwd_params")
                _count(3)
        return dy, d_abar, d_sump, dWv, dbv, None, None


def value_rpe(agg, abar, sump, Wv, bv, num_heads, heads_share):
    """agg + v-RPE contribution; see _ValueRpe.  Requires C % 4 == 0, C <= 256 and the fused
    GEMM preconditions; callers fall back to the plain composition otherwise."""
    _require_cuda(agg, abar, sump, Wv, bv)
    return _ValueRpe.apply(_f32c(agg), _f32c(abar), _f32c(sump), _f32c(Wv), _f32c(bv),
                           int(num_heads), int(bool(heads_share)))


# ---------------------------------------------------------------------------
# on-the-fly edge features
# ---------------------------------------------------------------------------
def horizontal_edge_features(se, edge_attr7, pos, normal, log_length, log_surface,
                             log_volume, log_size, num_nodes, add_self_loops=False):
    """(edge_index [2, 2Eh(+N)], edge_attr [.., 18]); see spt_edge_features_fwd."""
    lib = _lib.load()
    _require_cuda(se, edge_attr7, pos, normal, log_length, log_surface, log_volume, log_size)
    se = _i64c(se)
    Eh = se.shape[1]
    dev = se.device
    f = lambda t: t.detach().float().contiguous()  # noqa: E731  (fp16 inputs allowed)
    ea, pos, normal = f(edge_attr7), f(pos), f(normal)
    ll, ls, lv, lz = f(log_length).view(-1), f(log_surface).view(-1), f(log_volume).view(-1), \
        f(log_size).view(-1)
    E_out = 2 * Eh + (num_nodes if add_self_loops else 0)
    ei = torch.empty((2, E_out), dtype=torch.int64, device=dev)
    out = torch.empty((E_out, 18), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _timed('edge_features', Eh=Eh, N=num_nodes,
                                        loops=bool(add_self_loops)):
        _lib.check(lib.spt_edge_features_fwd(_p(se), _p(ea), _p(pos), _p(normal), _p(ll), _p(ls),
                                             _p(lv), _p(lz), Eh, num_nodes,
                                             1 if add_self_loops else 0, _p(ei), _p(out),
                                             _stream()), "spt_edge_features_fwd")
    _count()
    return ei, out


def vertical_edge_features(child, parent, normal_key='normal'):
    """v_edge_attr [Nc, 9] of the child level (see spt_vertical_edge_features_fwd);
    `child` / `parent` are Data-like objects with pos, normal, log_* and
    child.super_index."""
    lib = _lib.load()
    idx = _i64c(child.super_index)
    f = lambda t: t.detach().float().contiguous()  # noqa: E731
    cpos, ppos = f(child.pos), f(parent.pos)
    cn, pn = f(child[normal_key]), f(parent[normal_key])
    _require_cuda(idx, cpos, ppos, cn, pn)
    keys = ('log_length', 'log_surface', 'log_volume', 'log_size')
    clog = torch.stack([f(child[k]).view(-1) for k in keys]).contiguous()
    plog = torch.stack([f(parent[k]).view(-1) for k in keys]).contiguous()
    Nc, Np = cpos.shape[0], ppos.shape[0]
    out = torch.empty((Nc, 9), dtype=torch.float32, device=cpos.device)
    with torch.cuda.device(cpos.device):
        _lib.check(lib.spt_vertical_edge_features_fwd(_p(cpos), _p(ppos), _p(cn), _p(pn),
                                                      _p(clog), _p(plog), _p(idx), Nc, Np,
                                                      _p(out), _stream()),
                   "spt_vertical_edge_features_fwd")
    _count()
    return out
