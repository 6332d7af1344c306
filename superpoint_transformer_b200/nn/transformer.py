"""TransformerBlock: x + DropPath(SA(Norm(x))) [+ FFN branch]
(API of reference src/nn/transformer.py:133-265)."""
from torch import nn

from .attention import SelfAttentionBlock
from .dropout import DropPath
from .mlp import FFN
from .norm import LayerNorm, INDEX_BASED_NORMS
from ..utils.nn import VersionHolder

__all__ = ['TransformerBlock']


class TransformerBlock(nn.Module):
    def __init__(self, dim, num_heads=1, qkv_bias=True, qk_dim=8, qk_scale=None,
                 in_rpe_dim=18, ffn_ratio=4, attn_drop=None, residual_drop=None,
                 drop_path=None, activation=nn.LeakyReLU(), norm=LayerNorm, pre_norm=True,
                 no_sa=False, no_ffn=False, k_rpe=False, q_rpe=False, v_rpe=False,
                 k_delta_rpe=False, q_delta_rpe=False, qk_share_rpe=False,
                 q_on_minus_rpe=False, heads_share_rpe=False, version_holder=None):
        super().__init__()
        self.dim = dim
        self.pre_norm = pre_norm
        self.version_holder = version_holder if version_holder is not None \
            else VersionHolder()
        self.no_sa = no_sa
        if not no_sa:
            self.sa_norm = norm(dim)
            self.sa = SelfAttentionBlock(
                dim, num_heads=num_heads, in_dim=None, out_dim=dim, qkv_bias=qkv_bias,
                qk_dim=qk_dim, qk_scale=qk_scale, in_rpe_dim=in_rpe_dim,
                attn_drop=attn_drop, drop=residual_drop, k_rpe=k_rpe, q_rpe=q_rpe,
                v_rpe=v_rpe, k_delta_rpe=k_delta_rpe, q_delta_rpe=q_delta_rpe,
                qk_share_rpe=qk_share_rpe, q_on_minus_rpe=q_on_minus_rpe,
                heads_share_rpe=heads_share_rpe)
        self.no_ffn = no_ffn
        if not no_ffn:
            self.ffn_norm = norm(dim)
            self.ffn_ratio = ffn_ratio
            self.ffn = FFN(dim, hidden_dim=int(dim * ffn_ratio), activation=activation,
                           drop=residual_drop)
        self.drop_path = DropPath(drop_path) \
            if drop_path is not None and drop_path > 0 else nn.Identity()

    @staticmethod
    def _forward_norm(norm, x, norm_index):
        if isinstance(norm, INDEX_BASED_NORMS):
            return norm(x, batch=norm_index)
        return norm(x)

    def forward(self, x, norm_index, edge_index=None, edge_attr=None):
        assert x.dim() == 2 and x.is_floating_point(), 'x should be a 2D FloatTensor'
        assert norm_index.dim() == 1 and norm_index.shape[0] == x.shape[0], \
            'norm_index should be a 1D LongTensor'
        assert edge_index is None or \
            (edge_index.dim() == 2 and not edge_index.is_floating_point()), \
            'edge_index should be a 2D LongTensor'
        assert edge_attr is None or \
            (edge_attr.dim() == 2 and edge_attr.shape[0] == edge_index.shape[1]), \
            'edge_attr should be a 2D FloatTensor with one row per edge'

        shortcut = x
        # the SA branch is skipped when there are no edges (reference :229)
        if self.no_sa or edge_index is None or edge_index.shape[1] == 0:
            pass
        elif self.pre_norm:
            x = self._forward_norm(self.sa_norm, x, norm_index)
            x = self.sa(x, edge_index, edge_attr=edge_attr)
            x = shortcut + self.drop_path(x)
        else:
            x = self.drop_path(self.sa(x, edge_index, edge_attr=edge_attr))
            x = self._forward_norm(self.sa_norm, shortcut + x, norm_index)

        # version >= 2.2.0: the FFN residual starts from the SA output (reference :240-244)
        vh = self.version_holder
        if vh.major >= 3 or (vh.major == 2 and vh.minor >= 2):
            shortcut = x

        if not self.no_ffn:
            if self.pre_norm:
                x = self._forward_norm(self.ffn_norm, x, norm_index)
                x = shortcut + self.drop_path(self.ffn(x))
            else:
                x = self.drop_path(self.ffn(x))
                x = self._forward_norm(self.ffn_norm, shortcut + x, norm_index)
        return x, norm_index, edge_index
