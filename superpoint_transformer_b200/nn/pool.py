"""Child -> parent pooling (API of reference src/nn/pool.py).

Sum/Mean/Max/Min pools run the CSR segment kernel (csrc/segment.cu): one warp
per parent gathers its children rows, no atomics; max/min carry the arg-index so
the gradient goes to a single child (torch_scatter semantics).  The attentive
pools reuse the fused attention core with rows = parents, edges = children.
"""
import torch
from torch import nn

from .linear import Linear

from .. import ops
from ..utils.nn import build_qk_scale, LearnableParameter

__all__ = ['pool_factory', 'SumPool', 'MeanPool', 'MaxPool', 'MinPool', 'StdPool',
           'AttentivePool', 'AttentivePoolWithLearntQueries', 'BaseAttentivePool',
           'AggregationPoolMixIn']


class AggregationPoolMixIn:
    """Common call signature `(x_child, x_parent, index, edge_attr=None,
    num_pool=None)` (reference src/nn/pool.py:44-62)."""
    _reduce = None


class _SegmentPool(AggregationPoolMixIn, nn.Module):
    def forward(self, x_child, x_parent, index, edge_attr=None, num_pool=None):
        if num_pool is None:
            num_pool = ops.num_segments(index)
        return ops.segment_pool(x_child, index, num_pool, reduce=self._reduce)


class SumPool(_SegmentPool):
    _reduce = 'sum'


class MeanPool(_SegmentPool):
    _reduce = 'mean'


class MaxPool(_SegmentPool):
    _reduce = 'max'


class MinPool(_SegmentPool):
    _reduce = 'min'


class StdPool(AggregationPoolMixIn, nn.Module):
    """PyG StdAggregation as used by reference src/nn/pool.py:81-82: biased std from two
    segment means, `sqrt(clamp(E[x^2] - E[x]^2, 1e-5))`, values at the clamp floor set
    to 0.  Both segment reductions run on the CSR pool kernel (csrc/segment.cu)."""
    _reduce = 'std'

    def forward(self, x_child, x_parent, index, edge_attr=None, num_pool=None):
        if num_pool is None:
            num_pool = ops.num_segments(index)
        mean = ops.segment_pool(x_child, index, num_pool, reduce='mean')
        mean2 = ops.segment_pool(x_child * x_child, index, num_pool, reduce='mean')
        out = (mean2 - mean * mean).clamp(min=1e-5).sqrt()
        return out.masked_fill(out <= (1e-5) ** 0.5, 0.0)


class BaseAttentivePool(nn.Module):
    """QK-softmax-V pooling with queries from the parent and keys/values from the
    children (reference src/nn/pool.py:85-243)."""

    def __init__(self, dim=None, num_heads=1, in_dim=None, out_dim=None, qkv_bias=True,
                 qk_dim=8, qk_scale=None, attn_drop=None, drop=None, in_rpe_dim=9,
                 k_rpe=False, q_rpe=False, v_rpe=False, heads_share_rpe=False):
        super().__init__()
        assert dim % num_heads == 0, "dim must be a multiple of num_heads"
        if v_rpe:
            raise NotImplementedError
        if attn_drop is not None and attn_drop > 0:
            raise NotImplementedError("attention dropout is not built in the fused kernel")
        self.dim, self.num_heads, self.qk_dim = dim, num_heads, qk_dim
        self.qk_scale = build_qk_scale(dim, num_heads, qk_scale)
        self.heads_share_rpe = heads_share_rpe
        self.kv = Linear(dim, qk_dim * num_heads + dim, bias=qkv_bias)
        rpe_dim = qk_dim if heads_share_rpe else qk_dim * num_heads
        self.k_rpe = k_rpe if not isinstance(k_rpe, bool) else \
            (Linear(in_rpe_dim, rpe_dim) if k_rpe else None)
        self.q_rpe = q_rpe if not isinstance(q_rpe, bool) else \
            (Linear(in_rpe_dim, rpe_dim) if q_rpe else None)
        self.in_proj = Linear(in_dim, dim) if in_dim is not None else None
        self.out_proj = Linear(dim, out_dim) if out_dim is not None else None
        self.out_drop = nn.Dropout(drop) if drop is not None and drop > 0 else None

    def _rpe_weights(self, lin):
        if lin is None:
            return None, None
        if not isinstance(lin, nn.Linear):
            raise NotImplementedError("RPE encoders must be nn.Linear in the fused kernel")
        W, b = lin.weight, lin.bias
        if self.heads_share_rpe:
            W = W.repeat(self.num_heads, 1)
            b = b.repeat(self.num_heads) if b is not None else None
        return W, b

    def forward(self, x_child, x_parent, index, edge_attr=None, num_pool=None):
        Nc = x_child.shape[0]
        Np = x_parent.shape[0] if num_pool is None else num_pool
        if self.in_proj is not None:
            x_child = self.in_proj(x_child)
        q = self._get_query(x_parent)[:Np] if x_parent is not None else None
        kv = self.kv(x_child)
        # bipartite graph: row = parent, edge = child, target = the child itself
        g = ops._graph_cache.get(
            index, ("apool", Np, Nc),
            lambda: ops.build_graph_index(
                torch.stack((index, torch.arange(Nc, device=index.device))), Np, Nc))
        a = None
        if edge_attr is not None and (self.k_rpe is not None or self.q_rpe is not None):
            a = ops.permute_rows(edge_attr, g.perm)
        Wq, bq = self._rpe_weights(self.q_rpe)
        Wk, bk = self._rpe_weights(self.k_rpe)
        mode, value = self.qk_scale
        agg, _, _ = ops.attention_core(q.contiguous(), kv, a, Wq, bq, Wk, bk, g,
                                       self.num_heads, self.qk_dim, mode, value,
                                       want_abar=False)
        x = agg
        if self.out_proj is not None:
            x = self.out_proj(x)
        if self.out_drop is not None:
            x = self.out_drop(x)
        return x

    def _get_query(self, x_parent):
        raise NotImplementedError

    def extra_repr(self):
        return f'dim={self.dim}, num_heads={self.num_heads}'


class AttentivePool(BaseAttentivePool):
    """Queries = Linear(parent features) (reference src/nn/pool.py:259-304)."""

    def __init__(self, dim=None, q_in_dim=None, num_heads=1, in_dim=None, out_dim=None,
                 qkv_bias=True, qk_dim=8, qk_scale=None, attn_drop=None, drop=None,
                 in_rpe_dim=9, k_rpe=False, q_rpe=False, v_rpe=False,
                 heads_share_rpe=False):
        super().__init__(dim=dim, num_heads=num_heads, in_dim=in_dim, out_dim=out_dim,
                         qkv_bias=qkv_bias, qk_dim=qk_dim, qk_scale=qk_scale,
                         attn_drop=attn_drop, drop=drop, in_rpe_dim=in_rpe_dim, k_rpe=k_rpe,
                         q_rpe=q_rpe, v_rpe=v_rpe, heads_share_rpe=heads_share_rpe)
        self.q = Linear(q_in_dim, qk_dim * num_heads, bias=qkv_bias)

    def _get_query(self, x_parent):
        return self.q(x_parent)


class AttentivePoolWithLearntQueries(BaseAttentivePool):
    """One learnt query per head shared by all parents
    (reference src/nn/pool.py:308-360)."""

    def __init__(self, dim=None, num_heads=1, in_dim=None, out_dim=None, qkv_bias=True,
                 qk_dim=8, qk_scale=None, attn_drop=None, drop=None, in_rpe_dim=18,
                 k_rpe=False, q_rpe=False, v_rpe=False, heads_share_rpe=False):
        super().__init__(dim=dim, num_heads=num_heads, in_dim=in_dim, out_dim=out_dim,
                         qkv_bias=qkv_bias, qk_dim=qk_dim, qk_scale=qk_scale,
                         attn_drop=attn_drop, drop=drop, in_rpe_dim=in_rpe_dim, k_rpe=k_rpe,
                         q_rpe=q_rpe, v_rpe=v_rpe, heads_share_rpe=heads_share_rpe)
        self.q = LearnableParameter(torch.zeros(qk_dim * num_heads))
        nn.init.trunc_normal_(self.q, std=0.02)

    def _get_query(self, x_parent):
        return self.q.repeat(x_parent.shape[0], 1)


def pool_factory(pool, *args, **kwargs):
    """String / module -> pool module (reference src/nn/pool.py:24-41)."""
    if isinstance(pool, (AggregationPoolMixIn, BaseAttentivePool)):
        return pool
    table = {'max': MaxPool, 'min': MinPool, 'mean': MeanPool, 'sum': SumPool, 'std': StdPool}
    if isinstance(pool, str):
        if pool in table:
            return table[pool]()
    return pool(*args, **kwargs)
