"""Stage family: pos injection -> in_MLP -> N x TransformerBlock -> out_MLP, with
pool+fuse (Down) or unpool+fuse (Up) in front
(API of reference src/nn/stage.py:18-571; PointStage = level-0 MLP-only variant
of :574-806 without the sparse-CNN branch)."""
import torch
from torch import nn

from .linear import Linear

from .fusion import CatFusion, fusion_factory
from .mlp import MLP
from .norm import BatchNorm, UnitSphereNorm
from .pool import pool_factory
from .transformer import TransformerBlock
from .unpool import IndexUnpool

__all__ = ['Stage', 'DownNFuseStage', 'UpNFuseStage', 'PointStage']


def _shared_rpe(rpe, num_blocks, num_heads, in_dim, out_dim, blocks_share, heads_share):
    """Per-block list of RPE specs; one shared Linear when `blocks_share`
    (reference src/nn/stage.py:289-313 — including its hard-coded input width)."""
    if not isinstance(rpe, bool):
        assert blocks_share, \
            "A prebuilt RPE encoder is passed to all blocks: set blocks_share_rpe=True"
        return [rpe] * num_blocks
    if not heads_share:
        out_dim = out_dim * num_heads
    if blocks_share and rpe:
        return [Linear(in_dim, out_dim)] * num_blocks
    return [rpe] * num_blocks


class Stage(nn.Module):
    def __init__(self, dim, num_blocks=1, num_heads=1, in_mlp=None, out_mlp=None,
                 mlp_activation=nn.LeakyReLU(), mlp_norm=BatchNorm, mlp_drop=None,
                 use_pos=True, use_diameter=False, use_diameter_parent=False, qk_dim=8,
                 k_rpe=False, q_rpe=False, k_delta_rpe=False, q_delta_rpe=False,
                 qk_share_rpe=False, q_on_minus_rpe=False, blocks_share_rpe=False,
                 heads_share_rpe=False, version_holder=None, **transformer_kwargs):
        super().__init__()
        self.version_holder = version_holder
        self.dim = dim
        self.num_blocks = num_blocks
        self.num_heads = num_heads

        self.in_mlp = None
        if in_mlp is not None:
            assert in_mlp[-1] == dim
            self.in_mlp = MLP(in_mlp, activation=mlp_activation, norm=mlp_norm,
                              drop=mlp_drop)
        self.out_mlp = None
        if out_mlp is not None:
            assert out_mlp[0] == dim
            self.out_mlp = MLP(out_mlp, activation=mlp_activation, norm=mlp_norm,
                               drop=mlp_drop)

        self.transformer_blocks = None
        if num_blocks > 0:
            k_list = _shared_rpe(k_rpe, num_blocks, num_heads, 18, qk_dim,
                                 blocks_share_rpe, heads_share_rpe)
            kd_list = _shared_rpe(k_delta_rpe, num_blocks, num_heads, dim, qk_dim,
                                  blocks_share_rpe, heads_share_rpe)
            k_on = (not isinstance(k_rpe, bool)) or k_rpe
            kd_on = (not isinstance(k_delta_rpe, bool)) or k_delta_rpe
            q_spec = q_rpe if not isinstance(q_rpe, bool) else \
                (q_rpe and not (k_on and qk_share_rpe))
            qd_spec = q_delta_rpe if not isinstance(q_delta_rpe, bool) else \
                (q_delta_rpe and not (kd_on and qk_share_rpe))
            q_list = _shared_rpe(q_spec, num_blocks, num_heads, 18, qk_dim,
                                 blocks_share_rpe, heads_share_rpe)
            qd_list = _shared_rpe(qd_spec, num_blocks, num_heads, dim, qk_dim,
                                  blocks_share_rpe, heads_share_rpe)
            self.transformer_blocks = nn.ModuleList(
                TransformerBlock(
                    dim, num_heads=num_heads, qk_dim=qk_dim, k_rpe=k, q_rpe=q,
                    k_delta_rpe=kd, q_delta_rpe=qd, qk_share_rpe=qk_share_rpe,
                    q_on_minus_rpe=q_on_minus_rpe, heads_share_rpe=heads_share_rpe,
                    version_holder=self.version_holder, **transformer_kwargs)
                for k, q, kd, qd in zip(k_list, q_list, kd_list, qd_list))

        self.pos_norm = UnitSphereNorm()
        self.feature_fusion = CatFusion()
        self.use_pos = use_pos
        self.use_diameter = use_diameter
        self.use_diameter_parent = use_diameter_parent

    @property
    def out_dim(self):
        if self.out_mlp is not None:
            return self.out_mlp.out_dim
        if self.transformer_blocks is not None:
            return self.transformer_blocks[-1].dim
        if self.in_mlp is not None:
            return self.in_mlp.out_dim
        return self.dim

    def forward(self, x, norm_index, pos=None, diameter=None, node_size=None,
                super_index=None, edge_index=None, edge_attr=None, *args, **kwargs):
        ref = next(t for t in (x, pos, diameter, super_index) if t is not None)
        N, device = ref.shape[0], ref.device
        dtype = ref.dtype if ref.is_floating_point() else \
            (edge_attr.dtype if edge_attr is not None else torch.float)

        # segment-relative coordinates + parent diameter (reference :249-270)
        diameter_parent = None
        if pos is not None:
            normalized_pos, diameter_parent = self.pos_norm(pos, super_index, w=node_size)
            if self.use_pos:
                x = self.feature_fusion(normalized_pos, x)
        if self.use_diameter:
            diam = diameter if diameter is not None else \
                torch.zeros((N, 1), dtype=dtype, device=device)
            x = self.feature_fusion(diam, x)
        if self.use_diameter_parent:
            if diameter_parent is None:
                diam = torch.zeros((N, 1), dtype=dtype, device=device)
            elif super_index is None:
                diam = diameter_parent.repeat(N, 1)
            else:
                diam = diameter_parent[super_index]
            x = self.feature_fusion(diam, x)

        if self.in_mlp is not None:
            x = self.in_mlp(x, batch=norm_index)
        if self.transformer_blocks is not None:
            for block in self.transformer_blocks:
                x, norm_index, edge_index = block(
                    x, norm_index, edge_index=edge_index, edge_attr=edge_attr)
        if self.out_mlp is not None:
            x = self.out_mlp(x, batch=norm_index)
        return x, diameter_parent


class DownNFuseStage(Stage):
    """x_child --pool--> fuse with x_parent --> Stage (reference :316-444). The
    attribute name `down_pool_block` selects the low-LR parameter group upstream."""

    def __init__(self, *args, pool='max', fusion='cat', **kwargs):
        super().__init__(*args, **kwargs)
        self.down_pool_block = pool_factory(pool)
        self.fusion = fusion_factory(fusion)

    def forward(self, x_parent, x_child, norm_index, pool_index, pos=None, diameter=None,
                node_size=None, super_index=None, edge_index=None, edge_attr=None,
                v_edge_attr=None, num_super=None):
        x_pooled = self.down_pool_block(x_child, x_parent, pool_index,
                                        edge_attr=v_edge_attr, num_pool=num_super)
        x_fused = self.fusion(x_parent, x_pooled)
        return super().forward(x_fused, norm_index, pos=pos, node_size=node_size,
                               super_index=super_index, edge_index=edge_index,
                               edge_attr=edge_attr)


class UpNFuseStage(Stage):
    """x_parent --unpool--> fuse with x_child --> Stage (reference :447-571)."""

    def __init__(self, *args, unpool='index', fusion='cat', **kwargs):
        super().__init__(*args, **kwargs)
        if unpool != 'index':
            raise NotImplementedError(f'Unknown unpool={unpool} mode')
        self.unpool = IndexUnpool()
        self.fusion = fusion_factory(fusion)

    def forward(self, x_child, x_parent, norm_index, unpool_index, pos=None, diameter=None,
                node_size=None, super_index=None, edge_index=None, edge_attr=None):
        x_unpool = self.unpool(x_parent, unpool_index)
        x_fused = self.fusion(x_child, x_unpool)
        return super().forward(x_fused, norm_index, pos=pos, node_size=node_size,
                               super_index=super_index, edge_index=edge_index,
                               edge_attr=edge_attr)


class PointStage(Stage):
    """Level-0 stage: MLP on [pos, parent diameter, point features], no attention
    (reference :574-806 with cnn_blocks=False; the torchsparse CNN branch is out of
    scope)."""

    def __init__(self, in_mlp, mlp_activation=nn.LeakyReLU(), mlp_norm=BatchNorm,
                 mlp_drop=None, use_pos=True, use_diameter_parent=False,
                 cnn_blocks=False, version_holder=None, **kwargs):
        if cnn_blocks:
            raise NotImplementedError("sparse-CNN PointStage is out of scope")
        assert len(in_mlp) > 1, 'in_mlp should be a list of 2 or more integers'
        super().__init__(in_mlp[-1], num_blocks=0, in_mlp=in_mlp, out_mlp=None,
                         mlp_activation=mlp_activation, mlp_norm=mlp_norm, mlp_drop=mlp_drop,
                         use_pos=use_pos, use_diameter=False,
                         use_diameter_parent=use_diameter_parent,
                         version_holder=version_holder)
