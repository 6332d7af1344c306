"""CPU restatement of the reference hot path (TEST INFRASTRUCTURE — see
oracle/__init__.py).  Functional style: every function takes a `state_dict`-like
mapping `sd` whose keys are the REFERENCE parameter names, a key prefix, and the
tensors; dtype follows the inputs (fp32 for parity, fp64 for "truth").

Each function cites the reference lines it restates (paths relative to
/root/reference).  The leaf arithmetic comes from oracle/leaves.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import leaves as L

LEAKY_SLOPE = 0.01  # nn.LeakyReLU() default, configs/model/semantic/_attention.yaml:7-8


# --------------------------------------------------------------------------- #
#  integer index structures (bit-exact oracles, numpy)
# --------------------------------------------------------------------------- #
def group_index_np(key, num_groups):
    """Stable grouping: (ptr, perm) with perm = stable argsort(key).  This is the
    canonical CSR the product builds; `indices_to_pointers`
    (src/utils/sparse.py:23-41) is the reference's (non-stable) counterpart and
    `Cluster.pointers/points` (src/data/cluster.py:19-77) its stored form."""
    key = np.asarray(key, dtype=np.int64)
    perm = np.argsort(key, kind='stable')
    counts = np.bincount(key, minlength=num_groups)
    ptr = np.zeros(num_groups + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(counts)
    return ptr, perm


def csr_from_coo_np(edge_index, num_nodes):
    """COO -> (rowptr, col, perm) grouped by source, original order inside rows."""
    src = np.asarray(edge_index[0], dtype=np.int64)
    dst = np.asarray(edge_index[1], dtype=np.int64)
    ptr, perm = group_index_np(src, num_nodes)
    return ptr, dst[perm], perm


def node_size_np(super_index, num_parents, child_size=None):
    """One bottom-up step of NAG.get_sub_size (src/data/nag.py:94-110)."""
    w = None if child_size is None else np.asarray(child_size, dtype=np.int64)
    out = np.zeros(num_parents, dtype=np.int64)
    np.add.at(out, np.asarray(super_index, dtype=np.int64), 1 if w is None else w)
    return out


# --------------------------------------------------------------------------- #
#  small helpers
# --------------------------------------------------------------------------- #
def _has(sd, key):
    return key in sd


def linear(sd, prefix, x):
    w = sd[prefix + '.weight'].to(x.dtype)
    b = sd.get(prefix + '.bias')
    return F.linear(x, w, None if b is None else b.to(x.dtype))


NORM_KIND = {'kind': None}   # None: GraphNorm if `mean_scale` exists, else per-node LayerNorm;
                             # 'layer_graph': PyG LayerNorm(mode='graph') (the code default of
                             # src/nn/transformer.py:137); ('group', G): GroupNorm(mode='graph')


def norm_apply(sd, prefix, x, index):
    """Index-based norm dispatch (src/nn/transformer.py:258-265, src/nn/mlp.py:89-94).
    GraphNorm is recognised by its `mean_scale` parameter; LayerNorm(graph) / GroupNorm share
    their parameter names with the per-node LayerNorm and are selected through NORM_KIND."""
    w = sd[prefix + '.weight'].to(x.dtype)
    b = sd[prefix + '.bias'].to(x.dtype)
    if _has(sd, prefix + '.mean_scale'):
        return L.graph_norm(x, index, w, b, sd[prefix + '.mean_scale'].to(x.dtype))
    kind = NORM_KIND['kind']
    if kind == 'layer_graph':
        return L.layer_norm_graph(x, index, w, b, 1e-5)
    if isinstance(kind, tuple) and kind[0] == 'group':
        return group_norm(x, index, w, b, kind[1])
    return F.layer_norm(x, (x.shape[1],), w, b, 1e-5)


def group_norm(x, index, weight, bias, num_groups, eps=1e-5):
    """GroupNorm.forward, mode='graph' (src/nn/norm.py:181-218): mean / variance over the
    nodes of each graph x the channels of each group; eps inside the sqrt."""
    if index is None:
        index = torch.zeros(x.shape[0], dtype=torch.long)
    B = int(index.max()) + 1
    G = num_groups
    xg = x.view(-1, G, x.shape[1] // G)
    norm = (L.degree(index, B, dtype=x.dtype).clamp(min=1) * (x.shape[1] // G)).view(-1, 1, 1)
    mean = L.scatter_sum(xg, index, 0, B).sum(dim=-1, keepdim=True) / norm
    xc = xg - mean.index_select(0, index)
    var = L.scatter_sum(xc * xc, index, 0, B).sum(dim=-1, keepdim=True) / norm
    out = (xc / (var + eps).sqrt().index_select(0, index)).view(-1, x.shape[1])
    if weight is not None:
        out = out * weight + bias
    return out


def std_pool(x, index, num_pool):
    """StdPool (src/nn/pool.py:81-82) -> PyG StdAggregation: biased std from two segment
    means, clamp(min=1e-5).sqrt(), floor values zeroed."""
    mean = L.scatter_mean(x, index, 0, num_pool)
    mean2 = L.scatter_mean(x * x, index, 0, num_pool)
    out = (mean2 - mean * mean).clamp(min=1e-5).sqrt()
    return out.masked_fill(out <= (1e-5) ** 0.5, 0.0)


def mlp(sd, prefix, x, batch=None, last_activation=True):
    """MLP.forward (src/nn/mlp.py:84-94) over the ModuleList built by mlp()
    (:8-57): [Linear(bias iff no norm), norm?, act?]*.  The layout is read back
    from the parameter names: 2-D weight = Linear, 1-D weight = norm, an index
    without parameters = the LeakyReLU (assumes drop=None: no Dropout modules)."""
    idx = {int(k[len(prefix) + 5:].split('.')[0]) for k in sd
           if k.startswith(prefix + '.mlp.')}
    for i in range(max(idx) + 1):
        p = f'{prefix}.mlp.{i}'
        if i not in idx:
            x = F.leaky_relu(x, LEAKY_SLOPE)
        elif sd[p + '.weight'].dim() == 2:
            x = linear(sd, p, x)
        else:
            x = norm_apply(sd, p, x, batch)
    if last_activation:
        x = F.leaky_relu(x, LEAKY_SLOPE)
    return x


# --------------------------------------------------------------------------- #
#  attention (src/nn/attention.py:167-325)
# --------------------------------------------------------------------------- #
def qk_scale(dim, num_heads, qk_scale, s, dtype=torch.float32):
    """build_qk_scale_func (src/utils/nn.py:75-127).  As in the reference,
    `s.bincount() ** -0.5` is an int64 ** float op, i.e. evaluated in float32
    whatever the feature dtype; the product with q then promotes."""
    D = (dim // num_heads) ** -0.5

    def G():
        return (s.bincount() ** -0.5)[s].view(-1, 1, 1)

    if qk_scale is None:
        return D * G()
    if not isinstance(qk_scale, str):
        return qk_scale
    key = qk_scale.lower().replace(' ', '')
    if key in ('dg', 'gd', 'd*g', 'g*d', 'd.g', 'g.d'):
        return D * G()
    if key in ('d+g', 'g+d'):
        return D + G()
    if key == 'd':
        return D
    if key == 'g':
        return G()
    raise ValueError(qk_scale)


def self_attention(sd, prefix, x, edge_index, edge_attr=None, *, num_heads, qk_dim,
                   qk_scale_mode=None, heads_share_rpe=False, qk_share_rpe=False,
                   q_on_minus_rpe=False, attn_drop_mask=None):
    """SelfAttentionBlock.forward.  Which RPE encoders exist is read from the
    parameter names ({k,q,v}_rpe.weight, {k,q}_delta_rpe.weight).  `attn_drop_mask` [E, H]:
    the multipliers nn.Dropout applied to the attention weights (:310-311), None in eval."""
    N, E = x.shape[0], edge_index.shape[1]
    H, D = num_heads, qk_dim
    DH = D * H
    if _has(sd, prefix + '.in_proj.weight'):
        x = linear(sd, prefix + '.in_proj', x)                       # :186-187
    dim = x.shape[1]
    qkv = linear(sd, prefix + '.qkv', x)                              # :191
    q = qkv[:, :DH].view(N, H, D)                                     # :202-204
    k = qkv[:, DH:2 * DH].view(N, H, D)
    v = qkv[:, 2 * DH:].view(N, H, -1)
    s, t = edge_index[0], edge_index[1]
    q, k, v = q[s], k[t], v[t]                                        # :207-211
    q = q * qk_scale(dim, H, qk_scale_mode, s, q.dtype)               # :214

    def rpe(name, a):
        r = linear(sd, prefix + '.' + name, a)
        if heads_share_rpe:
            r = r.repeat(1, H)                                        # :229-230
        return r.view(E, H, -1)

    has_k = _has(sd, prefix + '.k_rpe.weight')
    has_q = _has(sd, prefix + '.q_rpe.weight')
    has_v = _has(sd, prefix + '.v_rpe.weight')
    if has_k and edge_attr is not None:                               # :225-232
        k = k + rpe('k_rpe', edge_attr)
    if has_q and edge_attr is not None:                               # :235-245
        q = q + rpe('q_rpe', -edge_attr if q_on_minus_rpe else edge_attr)
    elif has_k and qk_share_rpe and edge_attr is not None:            # :246-256
        q = q + rpe('k_rpe', -edge_attr if q_on_minus_rpe else edge_attr)
    has_kd = _has(sd, prefix + '.k_delta_rpe.weight')
    has_qd = _has(sd, prefix + '.q_delta_rpe.weight')
    if has_kd:                                                        # :258-265
        k = k + rpe('k_delta_rpe', x[t] - x[s])
    if has_qd:                                                        # :269-279
        q = q + rpe('q_delta_rpe', x[s] - x[t] if q_on_minus_rpe else x[t] - x[s])
    elif has_kd and qk_share_rpe and edge_attr is not None:           # :280-291
        q = q + rpe('k_delta_rpe', x[s] - x[t] if q_on_minus_rpe else x[t] - x[s])
    if has_v and edge_attr is not None:                               # :294-301
        v = v + rpe('v_rpe', edge_attr)
    compat = torch.einsum('ehd,ehd->eh', q, k)                        # :304
    attn = L.segment_softmax(compat, s, num_nodes=N)                  # :307
    if attn_drop_mask is not None:                                    # :310-311
        attn = attn * attn_drop_mask.to(attn.dtype)
    out = (v * attn.unsqueeze(-1)).reshape(E, dim)                    # :314
    out = L.scatter_sum(out, s, 0, N)                                 # :315
    if _has(sd, prefix + '.out_proj.weight'):
        out = linear(sd, prefix + '.out_proj', out)                   # :318-319
    return out


def transformer_block(sd, prefix, x, norm_index, edge_index, edge_attr, *, pre_norm=True,
                      version=(3, 0), **attn_kw):
    """TransformerBlock.forward (src/nn/transformer.py:195-256), DropPath/dropout
    in eval mode (identity)."""
    shortcut = x
    has_sa = _has(sd, prefix + '.sa.qkv.weight')
    if not has_sa or edge_index is None or edge_index.shape[1] == 0:  # :229
        pass
    elif pre_norm:
        x = norm_apply(sd, prefix + '.sa_norm', x, norm_index)
        x = self_attention(sd, prefix + '.sa', x, edge_index, edge_attr, **attn_kw)
        x = shortcut + x
    else:
        x = self_attention(sd, prefix + '.sa', x, edge_index, edge_attr, **attn_kw)
        x = norm_apply(sd, prefix + '.sa_norm', shortcut + x, norm_index)
    if version[0] >= 3 or (version[0] == 2 and version[1] >= 2):      # :240-244
        shortcut = x
    if _has(sd, prefix + '.ffn.mlp.0.weight'):
        if pre_norm:
            x = norm_apply(sd, prefix + '.ffn_norm', x, norm_index)
            x = shortcut + mlp(sd, prefix + '.ffn', x, last_activation=False)
        else:
            x = mlp(sd, prefix + '.ffn', x, last_activation=False)
            x = norm_apply(sd, prefix + '.ffn_norm', shortcut + x, norm_index)
    return x


# --------------------------------------------------------------------------- #
#  pooling / unpooling / position norm
# --------------------------------------------------------------------------- #
def pool(x_child, index, num_pool, reduce='max'):
    """Sum/Mean/Max/MinPool.__call__ (src/nn/pool.py:44-82)."""
    return L.aggregate(x_child, index, dim_size=num_pool, reduce=reduce)


def attentive_pool(sd, prefix, x_child, x_parent, index, edge_attr=None, num_pool=None, *,
                   num_heads, qk_dim, qk_scale_mode=None, heads_share_rpe=False,
                   learnt_queries=False):
    """BaseAttentivePool.forward (src/nn/pool.py:156-243)."""
    Nc = x_child.shape[0]
    Np = x_parent.shape[0] if num_pool is None else num_pool
    H, D = num_heads, qk_dim
    DH = D * H
    if _has(sd, prefix + '.in_proj.weight'):
        x_child = linear(sd, prefix + '.in_proj', x_child)
    dim = x_child.shape[1]
    if learnt_queries:
        q = sd[prefix + '.q'].to(x_child.dtype).repeat(x_parent.shape[0], 1)  # :348-360
    else:
        q = linear(sd, prefix + '.q', x_parent)                       # :295-304
    kv = linear(sd, prefix + '.kv', x_child)
    q = q[index].view(Nc, H, D)
    k = kv[:, :DH].view(Nc, H, D)
    v = kv[:, DH:].view(Nc, H, -1)
    q = q * qk_scale(dim, H, qk_scale_mode, index, q.dtype)

    def rpe(name):
        r = linear(sd, prefix + '.' + name, edge_attr)
        if heads_share_rpe:
            r = r.repeat(1, H)
        return r.view(Nc, H, -1)

    if _has(sd, prefix + '.k_rpe.weight'):
        k = k + rpe('k_rpe')
    if _has(sd, prefix + '.q_rpe.weight'):
        q = q + rpe('q_rpe')
    compat = torch.einsum('nhd,nhd->nh', q, k)
    attn = L.segment_softmax(compat, index, num_nodes=Np)
    out = L.scatter_sum((v * attn.unsqueeze(-1)).reshape(Nc, dim), index, 0, Np)
    if _has(sd, prefix + '.out_proj.weight'):
        out = linear(sd, prefix + '.out_proj', out)
    return out


def unpool(x_parent, idx):
    """IndexUnpool.forward (src/nn/unpool.py:12-13)."""
    return x_parent.index_select(0, idx)


def scatter_mean_weighted(x, idx, w, dim_size=None):
    """src/utils/scatter.py:17-38."""
    w = w.view(-1, 1).to(x.dtype)
    wx = torch.cat((w, x * w), dim=1)
    seg = L.scatter_sum(wx, idx, 0, dim_size)
    ws = seg[:, 0].clone()
    ws[ws == 0] = 1
    return seg[:, 1:] / ws.view(-1, 1)


def unit_sphere_norm(pos, idx, w=None, num_super=None):
    """UnitSphereNorm.forward (src/nn/norm.py:67-138), log_diameter=False."""
    if idx is None:                                                   # :86-110
        mn, mx = pos.min(dim=0).values, pos.max(dim=0).values
        diameter = (mx - mn).max()
        if w is None:
            center = pos.mean(dim=0)
        else:
            ws = w.to(pos.dtype).sum()
            ws = 1 if ws == 0 else ws
            center = (pos * w.view(-1, 1).to(pos.dtype)).sum(dim=0) / ws
        return (pos - center.view(1, -1)) / (diameter + 1e-2), diameter.view(1, 1)
    mn = L.scatter(pos, idx, 0, num_super, 'min')                     # :118-119
    mx = L.scatter(pos, idx, 0, num_super, 'max')
    diameter = (mx - mn).max(dim=1).values
    if w is None:
        center = L.scatter(pos, idx, 0, num_super, 'mean')
    else:
        center = scatter_mean_weighted(pos, idx, w, num_super)
    out = (pos - center[idx]) / (diameter[idx].view(-1, 1) + 1e-2)    # :136
    return out, diameter.view(-1, 1)


# --------------------------------------------------------------------------- #
#  stages (src/nn/stage.py) and SPT.forward (src/models/components/spt.py:760-944)
# --------------------------------------------------------------------------- #
def _cat(a, b):
    """CatFusion (src/nn/fusion.py:25-41)."""
    if a is None:
        return b
    if b is None:
        return a
    return torch.cat((a, b), dim=1)


def stage(sd, prefix, x, norm_index, pos=None, diameter=None, node_size=None,
          super_index=None, edge_index=None, edge_attr=None, *, use_pos=True,
          use_diameter=False, use_diameter_parent=False, block_kw=None):
    """Stage.forward (src/nn/stage.py:215-286)."""
    block_kw = block_kw or {}
    ref = next(t for t in (x, pos, diameter, super_index) if t is not None)
    N = ref.shape[0]
    dtype = ref.dtype if ref.is_floating_point() else torch.float32
    diameter_parent = None
    if pos is not None:                                               # :249-254
        npos, diameter_parent = unit_sphere_norm(pos, super_index, w=node_size)
        if use_pos:
            x = _cat(npos, x)
    if use_diameter:                                                  # :258-261
        diam = diameter if diameter is not None else torch.zeros((N, 1), dtype=dtype)
        x = _cat(diam, x)
    if use_diameter_parent:                                           # :263-270
        if diameter_parent is None:
            diam = torch.zeros((N, 1), dtype=dtype)
        elif super_index is None:
            diam = diameter_parent.repeat(N, 1)
        else:
            diam = diameter_parent[super_index]
        x = _cat(diam, x)
    if any(k.startswith(prefix + '.in_mlp.') for k in sd):            # :273-274
        x = mlp(sd, prefix + '.in_mlp', x, norm_index)
    j = 0
    while any(k.startswith(f'{prefix}.transformer_blocks.{j}.') for k in sd):  # :277-280
        x = transformer_block(sd, f'{prefix}.transformer_blocks.{j}', x, norm_index,
                              edge_index, edge_attr, **block_kw)
        j += 1
    if any(k.startswith(prefix + '.out_mlp.') for k in sd):           # :283-284
        x = mlp(sd, prefix + '.out_mlp', x, norm_index)
    return x, diameter_parent


def down_stage(sd, prefix, x_parent, x_child, norm_index, pool_index, num_super,
               pool_reduce='max', **kw):
    """DownNFuseStage.forward (src/nn/stage.py:413-444); cat fusion."""
    x_pooled = pool(x_child, pool_index, num_super, pool_reduce)
    kw.pop('diameter', None)  # received but not forwarded by the reference (:437-444)
    return stage(sd, prefix, _cat(x_parent, x_pooled), norm_index, **kw)


def up_stage(sd, prefix, x_child, x_parent, norm_index, unpool_index, **kw):
    """UpNFuseStage.forward (src/nn/stage.py:545-571); cat fusion."""
    kw.pop('diameter', None)
    return stage(sd, prefix, _cat(x_child, unpool(x_parent, unpool_index)), norm_index, **kw)


def spt_forward(sd, nag, *, num_heads, qk_dim, nano=True, num_down, num_up,
                segment_hf=('hf',), use_pos=True, use_diameter=False,
                use_diameter_parent=True, pool_reduce='max', use_node_hf=True,
                block_kw=None, output_stage_wise=False):
    """SPT.forward for nano (level-1 start) models, norm_mode='graph'
    (src/models/components/spt.py:760-879, 915-944).  `nag` is any object with the
    reference NAG/Data attribute names; it is NOT modified (features are read, the
    per-level x / edge_attr updates are kept local)."""
    assert nano, "the oracle restates the nano (no level-0) walk only"
    block_kw = dict(block_kw or {})
    heads = num_heads if isinstance(num_heads, (list, tuple)) else None
    stage_kw = dict(use_pos=use_pos, use_diameter=use_diameter,
                    use_diameter_parent=use_diameter_parent)

    def bkw(i):
        d = dict(block_kw)
        d['num_heads'] = heads[i] if heads else num_heads
        d['qk_dim'] = qk_dim
        return d

    start = nag.start_i_level
    levels = list(range(start, nag.absolute_num_levels))
    X, EA = {}, {}
    for l in levels:                                                  # :783-785 add_keys_to
        d = nag[l]
        feats = [getattr(d, k) for k in segment_hf]
        X[l] = torch.cat([f.unsqueeze(-1) if f.dim() == 1 else f for f in feats], dim=1) \
            if feats else None
        EA[l] = d.edge_attr

    def norm_index(l):
        d = nag[l]
        b = d.batch
        return b if b is not None else torch.zeros(d.num_nodes, dtype=torch.long)

    def encode(i_mlp, l):                                             # :788-796 / :826-835
        ni = norm_index(l)
        if any(k.startswith(f'node_mlps.{i_mlp}.') for k in sd) and X[l] is not None:
            X[l] = mlp(sd, f'node_mlps.{i_mlp}', X[l], ni)
        if any(k.startswith(f'h_edge_mlps.{i_mlp}.') for k in sd) and EA[l] is not None:
            EA[l] = mlp(sd, f'h_edge_mlps.{i_mlp}', EA[l], ni[nag[l].edge_index[0]])

    encode(0, start)
    d0 = nag[start]
    x, diameter = stage(sd, 'first_stage', X[start] if use_node_hf else None,
                        norm_index(start), pos=d0.pos, node_size=d0.node_size,
                        super_index=d0.super_index, edge_index=d0.edge_index,
                        edge_attr=EA[start], block_kw=bkw(0), **stage_kw)   # :881-913
    down_outputs = [x]
    for i_stage in range(num_down):                                   # :817-855
        l = i_stage + 2
        encode(i_stage + 1, l)
        d = nag[l]
        is_last = l == nag.end_i_level
        x, diameter = down_stage(
            sd, f'down_stages.{i_stage}', X[l] if use_node_hf else None, x, norm_index(l),
            nag[l - 1].super_index, d.num_nodes, pool_reduce, pos=d.pos,
            node_size=d.node_size, super_index=None if is_last else d.super_index,
            edge_index=d.edge_index, edge_attr=EA[l], block_kw=bkw(i_stage + 1), **stage_kw)
        down_outputs.append(x)
    up_outputs = []
    for i_stage in range(num_up):                                     # :860-868, :932-944
        l = num_down - i_stage - 1 + 1
        d = nag[l]
        x_skip = down_outputs[-(2 + i_stage)]
        x, _ = up_stage(
            sd, f'up_stages.{i_stage}', _cat(x_skip, X[l] if use_node_hf else None), x,
            norm_index(l), d.super_index, pos=d.pos, node_size=d.node_size,
            super_index=d.super_index, edge_index=d.edge_index, edge_attr=EA[l],
            block_kw=bkw(0), **stage_kw)
        up_outputs.append(x)
    if output_stage_wise:
        return [x] + up_outputs[::-1][1:] + [down_outputs[-1]]
    return x


# --------------------------------------------------------------------------- #
#  on-the-fly transforms (src/transforms/graph.py)
# --------------------------------------------------------------------------- #
H_KEYS = ('mean_off', 'std_off', 'mean_dist', 'angle_source', 'angle_target', 'normal_angle',
          'log_length', 'log_surface', 'log_volume', 'log_size', 'centroid_dir', 'centroid_dist')
_H_WIDTH = {'mean_off': 3, 'std_off': 3, 'centroid_dir': 3}


def horizontal_edge_features(se, ea, pos, normal, log_length, log_surface, log_volume,
                             log_size, keys=None):
    """_on_the_fly_horizontal_edge_features (src/transforms/graph.py:1137-1277).  Returns
    (edge_index [2,2Eh], edge_attr [2Eh,18]); with a `keys` subset only the columns of those
    keys, in the reference's assembly order (mean_off first: it is PREPENDED at :1214-1218,
    every other key appended in the order of the `if` chain)."""
    if keys is not None:
        ei, full = horizontal_edge_features(se, ea, pos, normal, log_length, log_surface,
                                            log_volume, log_size)
        cols, c = [], 0
        for k in H_KEYS:
            w = _H_WIDTH.get(k, 1)
            if k in keys:
                cols += list(range(c, c + w))
            c += w
        return ei, (full[:, cols] if cols else None)
    f_list = []
    f = ea[:, 3:6].float()                                            # std_off :1187-1191
    f_list.append(torch.cat((f, f), dim=0))
    f = ea[:, 6].float().view(-1, 1)                                  # mean_dist :1193-1197
    f_list.append(torch.cat((f, f), dim=0))
    mean_off = ea[:, :3].float()                                      # :1199-1212
    direction = mean_off / mean_off.norm(dim=1).view(-1, 1)
    direction[direction.isnan()] = 0
    direction = direction.clip(-1, 1)
    f_list = [torch.cat((mean_off, -mean_off), dim=0)] + f_list      # :1214-1218
    f = (direction * normal[se[0]]).sum(dim=1).abs()                  # angle_source :1220-1223
    f_list.append(torch.cat((f, f), dim=0).view(-1, 1))
    f = (direction * normal[se[1]]).sum(dim=1).abs()                  # angle_target :1225-1228
    f_list.append(torch.cat((f, f), dim=0).view(-1, 1))
    f = (normal[se[0]] * normal[se[1]]).sum(dim=1).abs()              # normal_angle :1230-1233
    f_list.append(torch.cat((f, f), dim=0).view(-1, 1))
    for t in (log_length, log_surface, log_volume, log_size):        # :1235-1249
        f = t[se[0]] - t[se[1]]
        f_list.append(torch.cat((f, -f), dim=0).view(-1, 1))
    cdir = pos[se[1]] - pos[se[0]]                                    # :1251-1267
    cdist = cdir.norm(dim=1).view(-1, 1)
    cdir = cdir / cdist.view(-1, 1)
    cdist = cdist.sqrt()
    cdir[cdir.isnan()] = 0
    cdir = cdir.clip(-1, 1)
    f_list.append(torch.cat((cdir, -cdir), dim=0))
    f_list.append(torch.cat((cdist, cdist), dim=0))
    edge_index = torch.cat((se, se.flip(0)), dim=1)                   # :1270
    return edge_index, torch.cat(f_list, dim=1)                       # :1277


def add_self_loops(edge_index, edge_attr, num_nodes):
    """NAGAddSelfLoops (src/transforms/graph.py:1419-1452)."""
    return L.add_self_loops(edge_index, edge_attr, fill_value=0., num_nodes=num_nodes)


def vertical_edge_features(child_pos, parent_pos, child_normal, parent_normal, child_logs,
                           parent_logs, idx):
    """_on_the_fly_vertical_edge_features with the default 7 keys
    (src/transforms/graph.py:1336-1416).  *_logs = (log_length, log_surface, log_volume,
    log_size) tensors.  Returns v_edge_attr [Nc, 9]."""
    f_list = []
    d = parent_pos[idx] - child_pos                                   # :1373-1376
    dist = d.norm(dim=1)
    d = d / dist.view(-1, 1)
    dist = dist.sqrt()
    d[d.isnan()] = 0                                                  # :1379-1380
    d = d.clip(-1, 1)
    f_list.append(d)
    f_list.append(dist.view(-1, 1))
    f = (child_normal * parent_normal[idx]).sum(dim=1).abs()          # :1388-1392
    f_list.append(f.view(-1, 1))
    for c, p_ in zip(child_logs, parent_logs):                        # :1394-1408
        f_list.append((p_[idx] - c).view(-1, 1))
    return torch.cat(f_list, dim=1)                                   # :1413


# --------------------------------------------------------------------------- #
#  superedge features from level-0 sub-edges (preprocessing)
# --------------------------------------------------------------------------- #
def base_vectors_3d(x):
    """src/utils/geometry.py:42-77.  NB: like the reference, zero rows of `x` are overwritten
    IN PLACE with (1, 0, 0) (`a = x` aliases the argument)."""
    a = x
    a[torch.where(a.norm(dim=1) == 0)[0]] = torch.tensor([[1, 0, 0]], dtype=x.dtype)
    a = a / a.norm(dim=1).view(-1, 1)
    b = torch.vstack((a[:, 1] - a[:, 2], a[:, 2] - a[:, 0], a[:, 0] - a[:, 1])).T
    b[torch.where(b.norm(dim=1) == 0)[0]] = torch.tensor([[2, 1, -1]], dtype=x.dtype)
    b = b / b.norm(dim=1).view(-1, 1)
    c = torch.linalg.cross(a, b)
    return torch.cat((a.unsqueeze(1), b.unsqueeze(1), c.unsqueeze(1)), dim=1)


def minimalistic_horizontal_edge_features(points, se_point_index, se_id, num_superedges):
    """_minimalistic_horizontal_edge_features (src/transforms/graph.py:1007-1058):
    [mean_off | std_off (clipped to [-2, 2]) | sqrt(mean_dist)] per superedge."""
    offset = points[se_point_index[1]] - points[se_point_index[0]]    # :1007
    dist = offset.norm(dim=1)                                         # :1017
    mean_off = L.scatter_mean(offset, se_id, 0, num_superedges)       # :1024
    base = base_vectors_3d(mean_off)[se_id]                           # :1031
    u = (offset * base[:, 0]).sum(dim=1).view(-1, 1)
    v = (offset * base[:, 1]).sum(dim=1).view(-1, 1)
    w = (offset * base[:, 2]).sum(dim=1).view(-1, 1)
    std_off = L.scatter_std(torch.cat((u, v, w), dim=1), se_id, 0, num_superedges)   # :1035
    std_off = std_off.clip(-2, 2)
    mean_dist = L.scatter_mean(dist, se_id, 0, num_superedges).sqrt()  # :1043
    return torch.cat((mean_off, std_off, mean_dist.view(-1, 1)), dim=1)


def cluster_mean_std(f, super_index, num_clusters):
    """scatter parts of _compute_cluster_features (src/transforms/graph.py:266-285)."""
    return (L.scatter_mean(f, super_index, 0, num_clusters),
            L.scatter_std(f, super_index, 0, num_clusters))
