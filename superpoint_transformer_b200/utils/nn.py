"""Host-side helpers mirrored from reference src/utils/nn.py, src/utils/parameter.py
and src/utils/version.py (only what the hot path needs)."""
import torch
from torch import nn

from .. import ops

__all__ = ['build_qk_scale', 'LearnableParameter', 'init_weights', 'VersionHolder',
           'listify_with_reference']


class LearnableParameter(nn.Parameter):
    """Marker subclass so `init_weights` can trunc-normal it
    (reference src/utils/parameter.py:7)."""


def build_qk_scale(dim, num_heads, qk_scale):
    """(mode, value) consumed by the attention kernel; semantics of
    build_qk_scale_func (reference src/utils/nn.py:75-127).  NB the 'd' factor is
    (dim // num_heads)^-1/2 — the VALUE head width, not qk_dim — as in the
    reference; 'g' = out-degree of the query row (self-loops included)."""
    d = float((dim // num_heads) ** -0.5)
    if qk_scale is None:
        return ops.SCALE_D_TIMES_G, d
    if not isinstance(qk_scale, str):
        return ops.SCALE_CONST, float(qk_scale)
    key = qk_scale.lower().replace(' ', '')
    if key in ('d+g', 'g+d'):
        return ops.SCALE_D_PLUS_G, d
    if key in ('dg', 'gd', 'd*g', 'g*d', 'd.g', 'g.d'):
        return ops.SCALE_D_TIMES_G, d
    if key == 'd':
        return ops.SCALE_D, d
    if key == 'g':
        return ops.SCALE_G, d
    raise ValueError(f"Unable to build QK scaling scheme for qk_scale='{qk_scale}'")


def _linear_init(m, method, activation):
    if m.bias is not None:
        nn.init.zeros_(m.bias)
    gain = nn.init.calculate_gain(activation)
    if method == 'xavier_uniform':
        nn.init.xavier_uniform_(m.weight, gain=gain)
    elif method == 'xavier_normal':
        nn.init.xavier_normal_(m.weight, gain=gain)
    elif method == 'kaiming_uniform':
        nn.init.kaiming_uniform_(m.weight, nonlinearity=activation)
    elif method == 'kaiming_normal':
        nn.init.kaiming_normal_(m.weight, nonlinearity=activation)
    elif method == 'trunc_normal':
        nn.init.trunc_normal_(m.weight, std=0.02)
    else:
        raise NotImplementedError(f"Unknown initialization method: {method}")


def init_weights(m, linear=None, rpe=None, activation='leaky_relu'):
    """`module.apply(init_weights)` initialiser (reference src/utils/nn.py:8-52):
    Xavier-uniform (leaky-relu gain) Linear layers, zero biases; k_rpe/q_rpe of
    attention blocks use the `rpe` scheme."""
    from ..nn.attention import SelfAttentionBlock
    linear = linear or 'xavier_uniform'
    rpe = rpe or linear
    if isinstance(m, LearnableParameter):
        nn.init.trunc_normal_(m, std=0.02)
    elif isinstance(m, nn.LayerNorm):
        nn.init.zeros_(m.bias)
        nn.init.ones_(m.weight)
    elif isinstance(m, nn.Linear):
        _linear_init(m, linear, activation)
    elif isinstance(m, SelfAttentionBlock):
        for enc in (m.k_rpe, m.q_rpe):
            if isinstance(enc, nn.Linear):
                _linear_init(enc, rpe, activation)


class VersionHolder:
    """Shared mutable version string; selects the FFN residual definition
    (reference src/utils/version.py:20-100, src/nn/transformer.py:240-244)."""

    def __init__(self, version='3.0.0'):
        self.value = version

    @property
    def parsed(self):
        major, minor, patch = (int(v) for v in str(self.value).split('.')[:3])
        return {'major': major, 'minor': minor, 'patch': patch}

    @property
    def major(self):
        return self.parsed['major']

    @property
    def minor(self):
        return self.parsed['minor']

    @property
    def patch(self):
        return self.parsed['patch']


def _listify(obj):
    if obj is None or isinstance(obj, str) or not hasattr(obj, '__len__'):
        return obj
    if hasattr(obj, 'dim') and obj.dim() == 0:
        return obj
    if len(obj) == 0:
        return obj
    return [_listify(x) for x in obj]


def listify_with_reference(arg_ref, *args):
    """Broadcast scalar constructor arguments against a reference list
    (reference src/utils/list.py:20-44)."""
    arg_ref = _listify(arg_ref)
    out = [_listify(a) for a in args]
    if arg_ref is None:
        return ([],) + tuple([] for _ in args)
    if not isinstance(arg_ref, list):
        return ([arg_ref],) + tuple([a] for a in out)
    if len(arg_ref) == 0:
        return ([],) + tuple([] for _ in args)
    res = []
    for a in out:
        if not isinstance(a, list):
            a = [a]
        if len(a) != len(arg_ref):
            a = a * len(arg_ref)
        res.append(a)
    return (arg_ref,) + tuple(res)
