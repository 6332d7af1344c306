"""SPT: the UNet-like network walking the NAG levels — the CALLER of the hot path.

Same constructor keywords, sub-module names (state-dict keys: `first_stage`,
`down_stages.{i}`, `up_stages.{i}`, `node_mlps`, `h_edge_mlps`, `v_edge_mlps`,
`...transformer_blocks.{j}.sa.{qkv,k_rpe,q_rpe,v_rpe,out_proj}`, `...sa_norm`,
`...down_pool_block`) and forward contract as reference
src/models/components/spt.py:288-944, so a reference Lightning module can hold
this class as `self.net`.  The torchsparse point-CNN options are out of scope
 and raise.
"""
import torch
from torch import nn

from . import ops
from .nn import (Stage, PointStage, DownNFuseStage, UpNFuseStage, BatchNorm, CatFusion,
                 MLP, LayerNorm)
from .nn.pool import BaseAttentivePool, pool_factory
from .nn.linear import Linear
from .utils.nn import VersionHolder, listify_with_reference

__all__ = ['SPT']

__version__ = '3.0.0'


def _stage_rpe_specs(rpe, num_stages, in_dim, out_dim, stages_share):
    """reference spt.py:947-968: one Linear shared by every stage, or the bool."""
    if not isinstance(rpe, bool):
        assert stages_share, \
            "A prebuilt RPE encoder is passed to all stages: set stages_share_rpe=True"
        return [rpe] * num_stages
    if stages_share and rpe:
        return [Linear(in_dim, out_dim)] * num_stages
    return [rpe] * num_stages


def _hf_mlps(layers, num_stage, activation, norm, shared):
    """reference spt.py:971-981"""
    if layers is None:
        return [None] * num_stage
    if shared:
        return nn.ModuleList([MLP(layers, activation=activation, norm=norm)] * num_stage)
    return nn.ModuleList(
        [MLP(layers, activation=activation, norm=norm) for _ in range(num_stage)])


class SPT(nn.Module):
    def __init__(
            self, point_hf=[], post_cnn_point_hf=[], segment_hf=[], point_mlp=None,
            point_drop=None, point_cnn_blocks=False, point_cnn=None,
            point_cnn_kernel_size=None, point_cnn_dilation=None, point_cnn_norm=None,
            point_cnn_activation=None, point_cnn_residual=False,
            point_cnn_global_residual=False, point_mlp_on_cnn_feats=False, nano=False,
            down_dim=None, down_pool_dim=None, down_in_mlp=None, down_out_mlp=None,
            down_mlp_drop=None, down_num_heads=1, down_num_blocks=0, down_ffn_ratio=4,
            down_residual_drop=None, down_attn_drop=None, down_drop_path=None,
            up_dim=None, up_in_mlp=None, up_out_mlp=None, up_mlp_drop=None,
            up_num_heads=1, up_num_blocks=0, up_ffn_ratio=4, up_residual_drop=None,
            up_attn_drop=None, up_drop_path=None, node_mlp=None, h_edge_mlp=None,
            v_edge_mlp=None, mlp_activation=nn.LeakyReLU(), mlp_norm=BatchNorm, qk_dim=8,
            qkv_bias=True, qk_scale=None, in_rpe_dim=18, activation=nn.LeakyReLU(),
            norm=LayerNorm, pre_norm=True, no_sa=False, no_ffn=False, k_rpe=False,
            q_rpe=False, v_rpe=False, k_delta_rpe=False, q_delta_rpe=False,
            qk_share_rpe=False, q_on_minus_rpe=False, share_hf_mlps=False,
            stages_share_rpe=False, blocks_share_rpe=False, heads_share_rpe=False,
            use_pos=True, use_node_hf=True, use_diameter=False, use_diameter_parent=False,
            pool='max', unpool='index', fusion='cat', norm_mode='graph',
            output_stage_wise=False, store_features=False):
        super().__init__()
        if point_cnn_blocks:
            raise NotImplementedError("sparse point-CNN (EZ-SP) is out of scope")
        self.nano = nano
        self.use_pos = use_pos
        self.use_node_hf = use_node_hf
        self.use_diameter = use_diameter
        self.use_diameter_parent = use_diameter_parent
        self.norm_mode = norm_mode
        self.stages_share_rpe = stages_share_rpe
        self.blocks_share_rpe = blocks_share_rpe
        self.heads_share_rpe = heads_share_rpe
        self.output_stage_wise = output_stage_wise
        self.store_features = store_features
        self.share_hf_mlps = share_hf_mlps
        self.point_mlp_on_cnn_feats = point_mlp_on_cnn_feats
        self.point_hf = point_hf
        self.segment_hf = segment_hf
        self.post_cnn_point_hf = post_cnn_point_hf
        self.version_holder = VersionHolder(__version__)

        (down_dim, down_pool_dim, down_in_mlp, down_out_mlp, down_mlp_drop, down_num_heads,
         down_num_blocks, down_ffn_ratio, down_residual_drop, down_attn_drop,
         down_drop_path, pool) = listify_with_reference(
            down_dim, down_pool_dim, down_in_mlp, down_out_mlp, down_mlp_drop,
            down_num_heads, down_num_blocks, down_ffn_ratio, down_residual_drop,
            down_attn_drop, down_drop_path, pool)
        (up_dim, up_in_mlp, up_out_mlp, up_mlp_drop, up_num_heads, up_num_blocks,
         up_ffn_ratio, up_residual_drop, up_attn_drop, up_drop_path) = \
            listify_with_reference(
                up_dim, up_in_mlp, up_out_mlp, up_mlp_drop, up_num_heads, up_num_blocks,
                up_ffn_ratio, up_residual_drop, up_attn_drop, up_drop_path)

        nano_i = int(self.nano)
        num_down = len(down_dim) - nano_i
        num_up = len(up_dim)
        needs_h_edge_hf = any(x > 0 for x in down_num_blocks + up_num_blocks)
        needs_v_edge_hf = num_down > 0 and isinstance(
            pool_factory(pool[0], down_pool_dim[0]), BaseAttentivePool)

        node_mlp = node_mlp if use_node_hf else None
        self.node_mlps = _hf_mlps(node_mlp, num_down + nano_i, mlp_activation, mlp_norm,
                                  share_hf_mlps)
        h_edge_mlp = h_edge_mlp if needs_h_edge_hf else None
        self.h_edge_mlps = _hf_mlps(h_edge_mlp, num_down + nano_i, mlp_activation, mlp_norm,
                                    share_hf_mlps)
        v_edge_mlp = v_edge_mlp if needs_v_edge_hf else None
        self.v_edge_mlps = _hf_mlps(v_edge_mlp, num_down, mlp_activation, mlp_norm,
                                    share_hf_mlps)

        common = dict(
            mlp_activation=mlp_activation, mlp_norm=mlp_norm, qk_dim=qk_dim,
            qkv_bias=qkv_bias, qk_scale=qk_scale, in_rpe_dim=in_rpe_dim,
            activation=activation, norm=norm, pre_norm=pre_norm, no_sa=no_sa, no_ffn=no_ffn,
            v_rpe=v_rpe, k_delta_rpe=k_delta_rpe, q_delta_rpe=q_delta_rpe,
            qk_share_rpe=qk_share_rpe, q_on_minus_rpe=q_on_minus_rpe, use_pos=use_pos,
            use_diameter=use_diameter, use_diameter_parent=use_diameter_parent,
            blocks_share_rpe=blocks_share_rpe, heads_share_rpe=heads_share_rpe,
            version_holder=self.version_holder)

        if self.nano:
            self.first_stage = Stage(
                down_dim[0], num_blocks=down_num_blocks[0], in_mlp=down_in_mlp[0],
                out_mlp=down_out_mlp[0], mlp_drop=down_mlp_drop[0],
                num_heads=down_num_heads[0], ffn_ratio=down_ffn_ratio[0],
                residual_drop=down_residual_drop[0], attn_drop=down_attn_drop[0],
                drop_path=down_drop_path[0], k_rpe=k_rpe, q_rpe=q_rpe, **common)
        else:
            self.first_stage = PointStage(
                point_mlp, mlp_activation=mlp_activation, mlp_norm=mlp_norm,
                mlp_drop=point_drop, use_pos=use_pos,
                use_diameter_parent=use_diameter_parent,
                version_holder=self.version_holder)

        self.feature_fusion = CatFusion()

        self.down_stages = None
        if num_down > 0:
            k_specs = _stage_rpe_specs(k_rpe, num_down, 18, qk_dim, stages_share_rpe)
            k_on = (not isinstance(k_rpe, bool)) or k_rpe
            q_arg = q_rpe if not isinstance(q_rpe, bool) else \
                (q_rpe and not (k_on and qk_share_rpe))
            q_specs = _stage_rpe_specs(q_arg, num_down, 18, qk_dim, stages_share_rpe)
            stages = []
            for i in range(num_down):
                j = i + nano_i
                stages.append(DownNFuseStage(
                    down_dim[j], num_blocks=down_num_blocks[j], in_mlp=down_in_mlp[j],
                    out_mlp=down_out_mlp[j], mlp_drop=down_mlp_drop[j],
                    num_heads=down_num_heads[j], ffn_ratio=down_ffn_ratio[j],
                    residual_drop=down_residual_drop[j], attn_drop=down_attn_drop[j],
                    drop_path=down_drop_path[j], k_rpe=k_specs[i], q_rpe=q_specs[i],
                    pool=pool_factory(pool[j], down_pool_dim[j]), fusion=fusion, **common))
            self.down_stages = nn.ModuleList(stages)

        self.up_stages = None
        if num_up > 0:
            k_specs = _stage_rpe_specs(k_rpe, num_up, 18, qk_dim, stages_share_rpe)
            k_on = (not isinstance(k_rpe, bool)) or k_rpe
            q_arg = q_rpe if not isinstance(q_rpe, bool) else \
                (q_rpe and not (k_on and qk_share_rpe))
            q_specs = _stage_rpe_specs(q_arg, num_up, 18, qk_dim, stages_share_rpe)
            self.up_stages = nn.ModuleList([
                UpNFuseStage(
                    up_dim[i], num_blocks=up_num_blocks[i], in_mlp=up_in_mlp[i],
                    out_mlp=up_out_mlp[i], mlp_drop=up_mlp_drop[i],
                    num_heads=up_num_heads[i], ffn_ratio=up_ffn_ratio[i],
                    residual_drop=up_residual_drop[i], attn_drop=up_attn_drop[i],
                    drop_path=up_drop_path[i], k_rpe=k_specs[i], q_rpe=q_specs[i],
                    unpool=unpool, fusion=fusion, **common)
                for i in range(num_up)])

        assert self.num_up_stages > 0 or not self.output_stage_wise, \
            "At least one up stage is needed for output_stage_wise=True"
        assert bool(self.down_stages) != bool(self.up_stages) \
            or self.num_down_stages >= self.num_up_stages, \
            "The number of Up stages should be <= the number of Down stages."
        assert self.nano or self.num_down_stages > self.num_up_stages \
            or self.num_down_stages == 0, \
            "The number of Up stages should be < the number of Down stages."

    # ------------------------------------------------------------------ props
    @property
    def num_down_stages(self):
        return len(self.down_stages) if self.down_stages is not None else 0

    @property
    def num_up_stages(self):
        return len(self.up_stages) if self.up_stages is not None else 0

    @property
    def out_dim(self):
        if self.output_stage_wise:
            out = [s.out_dim for s in self.up_stages][::-1]
            return out + [self.down_stages[-1].out_dim]
        if self.up_stages is not None:
            return self.up_stages[-1].out_dim
        if self.down_stages is not None:
            return self.down_stages[-1].out_dim
        return self.first_stage.out_dim

    @property
    def version(self):
        return self.version_holder.value

    @version.setter
    def version(self, v):
        self.version_holder.value = v

    # ---------------------------------------------------------------- forward
    def _edge_norm_index(self, data):
        """norm index of every edge = norm index of its source node (reference
        spt.py:829-831); the number of graphs is known from the node-level index,
        so the GraphNorm segment count needs no extra host sync."""
        node_idx = data.norm_index(mode=self.norm_mode)
        edge_idx = node_idx[data.edge_index[0]]
        ops.register_num_segments(edge_idx, ops.num_segments(node_idx))
        return edge_idx

    def _encode_handcrafted(self, nag, i_level, node_mlp, h_edge_mlp, v_edge_mlp):
        d = nag[i_level]
        if node_mlp is not None:
            d.x = node_mlp(d.x, batch=d.norm_index(mode=self.norm_mode))
        if h_edge_mlp is not None and d.edge_attr is not None:
            d.edge_attr = h_edge_mlp(d.edge_attr, batch=self._edge_norm_index(d))
        if v_edge_mlp is not None:
            child = nag[i_level - 1]
            v = d.v_edge_attr  # (sic) the reference reads it on the parent level
            if v is not None:
                child.v_edge_attr = v_edge_mlp(
                    v, batch=child.norm_index(mode=self.norm_mode))

    def forward(self, nag):
        assert int(self.nano) == nag.start_i_level, \
            "`nano` mode should be consistent between the model and the data"
        if not self.nano:
            nag.add_keys_to(level=0, keys=self.point_hf, to='x', delete_after=False)
            nag.add_keys_to(level=0, keys=self.post_cnn_point_hf, to='x_mlp',
                            delete_after=not self.store_features)
        nag.add_keys_to(level='1+', keys=self.segment_hf, to='x',
                        delete_after=not self.store_features)

        if self.nano:
            self._encode_handcrafted(
                nag, 1, self.node_mlps[0] if self.node_mlps is not None else None,
                self.h_edge_mlps[0] if self.h_edge_mlps is not None else None, None)

        start = nag.start_i_level
        d0 = nag[start]
        x, diameter = self.first_stage(
            d0.x if self.use_node_hf else None, d0.norm_index(mode=self.norm_mode),
            pos=d0.pos, diameter=None, node_size=d0.node_size, super_index=d0.super_index,
            edge_index=d0.edge_index, edge_attr=d0.edge_attr)
        nag[start + 1].diameter = diameter

        down_outputs = [x] if self.nano else []
        nano_i = int(self.nano)
        for i_stage in range(self.num_down_stages):
            stage = self.down_stages[i_stage]
            i_level = i_stage + 1 + nano_i
            self._encode_handcrafted(
                nag, i_level, self.node_mlps[i_stage + nano_i],
                self.h_edge_mlps[i_stage + nano_i], self.v_edge_mlps[i_stage])
            d = nag[i_level]
            is_last = i_level == nag.end_i_level
            x, diameter = stage(
                d.x if self.use_node_hf else None, x, d.norm_index(mode=self.norm_mode),
                nag[i_level - 1].super_index, pos=d.pos, diameter=d.diameter,
                node_size=d.node_size, super_index=None if is_last else d.super_index,
                edge_index=d.edge_index, edge_attr=d.edge_attr,
                v_edge_attr=nag[i_level - 1].v_edge_attr, num_super=d.num_nodes)
            down_outputs.append(x)
            if i_level < nag.absolute_num_levels - 1:
                nag[i_level + 1].diameter = diameter

        up_outputs = []
        for i_stage in range(self.num_up_stages):
            stage = self.up_stages[i_stage]
            i_level = self.num_down_stages - i_stage - 1 + nano_i
            d = nag[i_level]
            x_skip = down_outputs[-(2 + i_stage)]
            x_hf = d.x if self.use_node_hf else None
            x, _ = stage(
                self.feature_fusion(x_skip, x_hf), x, d.norm_index(mode=self.norm_mode),
                d.super_index, pos=d.pos, diameter=None, node_size=d.node_size,
                super_index=d.super_index, edge_index=d.edge_index, edge_attr=d.edge_attr)
            up_outputs.append(x)

        if self.output_stage_wise:
            return [x] + up_outputs[::-1][1:] + [down_outputs[-1]]
        return x
