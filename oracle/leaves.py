"""Restated third-party leaf ops (NOT in /root/reference — un-vendored wheels):
torch_scatter (for torch 2.2.0, install.sh:99) and torch_geometric==2.3.0
(install.sh:100).  Semantics per SURVEY.md Appendix A; each function names the
reference call sites that rely on it.  Pure torch, CPU, any float dtype.
"""
import torch

__all__ = ['scatter_sum', 'scatter_mean', 'scatter_min', 'scatter_max', 'scatter_std',
           'scatter', 'segment_softmax', 'degree', 'graph_norm', 'layer_norm_graph',
           'aggregate', 'add_self_loops']


def _dim_size(index, dim_size):
    if dim_size is not None:
        return int(dim_size)
    return int(index.max()) + 1 if index.numel() > 0 else 0


def _expand(index, src):
    shape = [index.shape[0]] + [1] * (src.dim() - 1)
    return index.view(shape).expand_as(src)


def scatter_sum(src, index, dim=0, dim_size=None):
    """torch_scatter.scatter_sum — out[index[e]] += src[e]
    (src/nn/attention.py:315, src/nn/pool.py:233, src/utils/scatter.py:30,
    src/data/nag.py:97,108)."""
    assert dim == 0
    n = _dim_size(index, dim_size)
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def _count(index, n, dtype):
    return torch.zeros(n, dtype=dtype).index_add_(0, index, torch.ones(index.shape[0], dtype=dtype))


def scatter_mean(src, index, dim=0, dim_size=None):
    """sum / clamp(count, 1); empty segment -> 0
    (src/transforms/graph.py:276,1025,1044)."""
    n = _dim_size(index, dim_size)
    s = scatter_sum(src, index, 0, n)
    c = _count(index, n, src.dtype).clamp(min=1)
    return s / c.view([-1] + [1] * (src.dim() - 1))


def _scatter_extreme(src, index, dim_size, largest):
    n = _dim_size(index, dim_size)
    M = src.shape[0]
    flat = src.reshape(M, -1)
    idx = index.view(-1, 1).expand_as(flat)
    red = 'amax' if largest else 'amin'
    init = torch.zeros((n, flat.shape[1]), dtype=src.dtype)
    ext = init.scatter_reduce(0, idx, flat.detach(), reduce=red, include_self=False)
    # arg = FIRST item attaining the extreme (CPU torch_scatter loop uses a strict compare)
    hit = flat.detach() == ext[index]
    pos = torch.arange(M).view(-1, 1).expand_as(flat)
    cand = torch.where(hit, pos, torch.full_like(pos, M))
    arg = torch.full((n, flat.shape[1]), M, dtype=torch.long).scatter_reduce(
        0, idx, cand, reduce='amin', include_self=True)
    valid = arg < M
    vals = torch.gather(flat, 0, arg.clamp(max=max(M - 1, 0))) if M > 0 else init
    vals = torch.where(valid, vals, torch.zeros_like(vals))  # empty segment -> 0
    shape = (n,) + tuple(src.shape[1:])
    return vals.reshape(shape), arg.reshape(shape)


def scatter_max(src, index, dim=0, dim_size=None):
    """(values, argindex); empty -> (0, M); gradient to the single arg element."""
    return _scatter_extreme(src, index, dim_size, True)


def scatter_min(src, index, dim=0, dim_size=None):
    return _scatter_extreme(src, index, dim_size, False)


def scatter_std(src, index, dim=0, dim_size=None):
    """torch_scatter.scatter_std: unbiased, sqrt(sum_sq / (clamp(count-1,1) + 1e-6))
    (src/transforms/graph.py:285,1036)."""
    n = _dim_size(index, dim_size)
    mean = scatter_mean(src, index, 0, n)
    var_sum = scatter_sum((src - mean[index]) ** 2, index, 0, n)
    c = _count(index, n, src.dtype)
    den = (c - 1).clamp(min=1) + 1e-6
    return (var_sum / den.view([-1] + [1] * (src.dim() - 1))).sqrt()


def scatter(src, index, dim=0, dim_size=None, reduce='sum'):
    """torch_scatter.scatter(reduce=...) (src/nn/norm.py:118-126,201-211)."""
    if reduce in ('sum', 'add'):
        return scatter_sum(src, index, dim, dim_size)
    if reduce == 'mean':
        return scatter_mean(src, index, dim, dim_size)
    if reduce == 'min':
        return scatter_min(src, index, dim, dim_size)[0]
    if reduce == 'max':
        return scatter_max(src, index, dim, dim_size)[0]
    raise ValueError(reduce)


def segment_softmax(src, index, num_nodes=None):
    """torch_geometric.utils.softmax(src, index, dim=0, num_nodes): max taken on
    src.detach(); + 1e-16 added AFTER the sum (src/nn/attention.py:307, pool.py:225)."""
    n = _dim_size(index, num_nodes)
    m = scatter_max(src.detach(), index, 0, n)[0]
    e = (src - m[index]).exp()
    z = scatter_sum(e, index, 0, n) + 1e-16
    return e / z[index]


def degree(index, num_nodes=None, dtype=torch.float):
    """torch_geometric.utils.degree (src/nn/norm.py:197)."""
    n = _dim_size(index, num_nodes)
    return torch.zeros(n, dtype=dtype).index_add_(0, index, torch.ones(index.shape[0], dtype=dtype))


def graph_norm(x, batch, weight, bias, mean_scale, eps=1e-5, batch_size=None):
    """torch_geometric.nn.norm.GraphNorm.forward (PyG 2.3.0)."""
    if batch is None:
        batch = torch.zeros(x.shape[0], dtype=torch.long)
    n = _dim_size(batch, batch_size)
    mean = scatter_mean(x, batch, 0, n)
    out = x - mean.index_select(0, batch) * mean_scale
    var = scatter_mean(out.pow(2), batch, 0, n)
    std = (var + eps).sqrt().index_select(0, batch)
    return weight * out / std + bias


def layer_norm_graph(x, batch, weight, bias, eps=1e-5, batch_size=None):
    """torch_geometric.nn.norm.LayerNorm(mode='graph') (PyG 2.3.0)."""
    if batch is None:
        x = x - x.mean()
        out = x / (x.std(unbiased=False) + eps)
    else:
        n = _dim_size(batch, batch_size)
        norm = degree(batch, n, dtype=x.dtype).clamp_(min=1).mul_(x.shape[-1]).view(-1, 1)
        mean = scatter_sum(x, batch, 0, n).sum(dim=-1, keepdim=True) / norm
        x = x - mean.index_select(0, batch)
        var = scatter_sum(x * x, batch, 0, n).sum(dim=-1, keepdim=True) / norm
        out = x / (var + eps).sqrt().index_select(0, batch)
    if weight is not None:
        out = out * weight + bias
    return out


def aggregate(x, index, dim_size=None, reduce='max'):
    """torch_geometric.nn.aggr.{Sum,Mean,Max,Min}Aggregation (src/nn/pool.py:61-62):
    scatter(reduce); max/min with grad go through torch_scatter (single-arg grad)."""
    return scatter(x, index, 0, dim_size, 'sum' if reduce == 'add' else reduce)


def add_self_loops(edge_index, edge_attr=None, fill_value=0., num_nodes=None):
    """torch_geometric.utils.add_self_loops: APPENDS [[0..N-1],[0..N-1]] and
    fill_value rows (src/transforms/graph.py:1442-1446)."""
    n = int(num_nodes)
    loop = torch.arange(n, dtype=edge_index.dtype)
    ei = torch.cat((edge_index, torch.stack((loop, loop))), dim=1)
    if edge_attr is not None:
        pad = torch.full((n,) + tuple(edge_attr.shape[1:]), fill_value, dtype=edge_attr.dtype)
        edge_attr = torch.cat((edge_attr, pad), dim=0)
    return ei, edge_attr
