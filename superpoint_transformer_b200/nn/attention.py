"""SelfAttentionBlock — sparse multi-head attention over the superpoint graph
with edge-feature relative positional encodings.

Same constructor / forward signature / parameter names as the reference
(src/nn/attention.py:83-325) so it drops in under TransformerBlock/Stage/SPT and
loads reference checkpoints; the arithmetic between the `qkv` and `out_proj`
Linears is one fused CUDA pass per direction (csrc/attention.cu) instead of the
~20 library launches listed in SURVEY.md §3.3.
"""
import torch
from torch import nn

from .linear import Linear

from .. import ops
from ..utils.nn import build_qk_scale

__all__ = ['SelfAttentionBlock']


def _encoder(spec, in_dim, out_dim, build):
    """RPE encoder argument: a bool (build a Linear or not) or a prebuilt, possibly
    shared, module (reference src/nn/attention.py:127-157)."""
    if not isinstance(spec, bool):
        return spec
    return Linear(in_dim, out_dim) if (spec and build) else None


class SelfAttentionBlock(nn.Module):
    def __init__(self, dim, num_heads=1, in_dim=None, out_dim=None, qkv_bias=True,
                 qk_dim=8, qk_scale=None, attn_drop=None, drop=None, in_rpe_dim=18,
                 k_rpe=False, q_rpe=False, v_rpe=False, k_delta_rpe=False,
                 q_delta_rpe=False, qk_share_rpe=False, q_on_minus_rpe=False,
                 heads_share_rpe=False):
        super().__init__()
        assert dim % num_heads == 0, "dim must be a multiple of num_heads"
        self.dim = dim
        self.num_heads = num_heads
        self.qk_dim = qk_dim
        self.qk_scale = build_qk_scale(dim, num_heads, qk_scale)
        self.heads_share_rpe = heads_share_rpe
        self.qk_share_rpe = qk_share_rpe
        self.q_on_minus_rpe = q_on_minus_rpe

        self.qkv = Linear(dim, qk_dim * 2 * num_heads + dim, bias=qkv_bias)

        qk_out = qk_dim if heads_share_rpe else qk_dim * num_heads
        v_out = dim // num_heads if heads_share_rpe else dim
        k_on = (not isinstance(k_rpe, bool)) or k_rpe
        kd_on = (not isinstance(k_delta_rpe, bool)) or k_delta_rpe
        self.k_rpe = _encoder(k_rpe, in_rpe_dim, qk_out, True)
        self.q_rpe = _encoder(q_rpe, in_rpe_dim, qk_out, not (k_on and qk_share_rpe))
        self.k_delta_rpe = _encoder(k_delta_rpe, dim, qk_out, True)
        self.q_delta_rpe = _encoder(q_delta_rpe, dim, qk_out,
                                    not (kd_on and qk_share_rpe))
        self.v_rpe = _encoder(v_rpe, in_rpe_dim, v_out, True)

        self.in_proj = Linear(in_dim, dim) if in_dim is not None else None
        self.out_proj = Linear(dim, out_dim) if out_dim is not None else None

        # attention-weight dropout (reference :162-163, 310-311): the mask is drawn by torch and
        # applied inside the kernels (spt_attn_extras.drop_mask), forward and backward
        self.attn_drop = nn.Dropout(attn_drop) \
            if attn_drop is not None and attn_drop > 0 else None
        self.out_drop = nn.Dropout(drop) if drop is not None and drop > 0 else None

    # -- helpers -----------------------------------------------------------
    def _linear_wb(self, lin, negate_input=False):
        """Linear encoder -> (W [H*d, F], b [H*d]) as the kernel expects them;
        heads_share_rpe tiles the single-head encoder over heads
        (rpe.repeat(1, H), reference src/nn/attention.py:229-230)."""
        if lin is None:
            return None, None
        if not isinstance(lin, nn.Linear):
            raise NotImplementedError(
                "RPE encoders must be nn.Linear modules for the fused kernel")
        W, b = lin.weight, lin.bias
        if negate_input:  # enc(-a) = -W a + b
            W = -W
        if self.heads_share_rpe:
            W = W.repeat(self.num_heads, 1)
            b = b.repeat(self.num_heads) if b is not None else None
        return W, b

    def _delta_terms(self, x, qkv, has_edge_attr):
        """Node-difference encodings (reference :259-291).  The encoders are Linear, so
        enc(x_t - x_s) = W x_t - W x_s + b: one dense product per node, then per-row /
        per-target addends of the fused kernel.  Returns (q_row_add, q_tgt_add, k_row_add,
        kv) with kv = [k + W_k x | v] when the key encoder is on, else None."""
        H, HD = self.num_heads, self.num_heads * self.qk_dim

        def proj(lin):
            if not isinstance(lin, nn.Linear):
                raise NotImplementedError("delta RPE encoders must be nn.Linear")
            u = ops.linear(x, lin.weight)                 # [N, D or HD], no bias
            b = lin.bias
            if self.heads_share_rpe:
                u = u.repeat(1, H)
                b = b.repeat(H) if b is not None else None
            return u, b

        q_row = q_tgt = k_row = kv = None
        if self.k_delta_rpe is not None:
            uk, bk = proj(self.k_delta_rpe)
            k_row = -uk if bk is None else bk - uk        # the x_s part (+ bias)
            kv = torch.cat((qkv[:, HD:2 * HD] + uk, qkv[:, 2 * HD:]), dim=1)   # x_t part folded in k
        q_enc = self.q_delta_rpe
        if q_enc is None and self.k_delta_rpe is not None and self.qk_share_rpe and has_edge_attr:
            q_enc = self.k_delta_rpe                      # reference :281-291 (incl. its condition)
        if q_enc is not None:
            uq, bq = proj(q_enc)
            sign = -1.0 if self.q_on_minus_rpe else 1.0   # enc(x_s - x_t) when on_minus
            q_tgt = sign * uq
            q_row = -sign * uq if bq is None else bq - sign * uq
        return q_row, q_tgt, k_row, kv

    def forward(self, x, edge_index, edge_attr=None, attn_drop_mask=None):
        """x [N, Cx]; edge_index [2, E] (row 0 = querying node, row 1 = key node;
        any order); edge_attr [E, F] or None.  Returns [N, out_dim or dim].
        `attn_drop_mask` [E, H] (extension, tests): the dropout multipliers to use instead of
        drawing them (original edge order, already scaled by 1 / (1 - p))."""
        N = x.shape[0]
        H, D = self.num_heads, self.qk_dim
        if self.in_proj is not None:
            x = self.in_proj(x)
        qkv = self.qkv(x)

        g = ops.graph_index(edge_index, N)
        q_row = q_tgt = k_row = kv = None
        if self.k_delta_rpe is not None or self.q_delta_rpe is not None:
            q_row, q_tgt, k_row, kv = self._delta_terms(x, qkv, edge_attr is not None)
        mask = attn_drop_mask
        if mask is None and self.attn_drop is not None and self.training:
            E = edge_index.shape[1]
            mask = self.attn_drop(torch.ones((E, H), dtype=qkv.dtype, device=qkv.device))
        if mask is not None and g.perm is not None:
            mask = ops._gather_rows(mask.detach().float().contiguous(), g.perm)

        # which encoders act on edge_attr (reference attention.py:225-256, 294-301)
        Wq = bq = Wk = bk = None
        use_v = False
        if edge_attr is not None:
            if self.k_rpe is not None:
                Wk, bk = self._linear_wb(self.k_rpe)
            if self.q_rpe is not None:
                Wq, bq = self._linear_wb(self.q_rpe, self.q_on_minus_rpe)
            elif self.k_rpe is not None and self.qk_share_rpe:
                Wq, bq = self._linear_wb(self.k_rpe, self.q_on_minus_rpe)
            use_v = self.v_rpe is not None
        a = None
        if edge_attr is not None and (Wq is not None or Wk is not None or use_v):
            a = ops.permute_rows_cached(edge_attr, g.perm)

        mode, value = self.qk_scale
        qsrc = qkv if kv is None else qkv[:, :H * D].contiguous()
        agg, abar, sump = ops.attention_core(qsrc, kv, a, Wq, bq, Wk, bk, g, H, D, mode,
                                             value, want_abar=use_v, q_row_add=q_row,
                                             q_tgt_add=q_tgt, k_row_add=k_row, drop_mask=mask)
        y = agg
        if use_v:
            # sum_e p_e (Wv a_e + bv) = Wv abar + bv * sum_e p_e
            if not isinstance(self.v_rpe, nn.Linear):
                raise NotImplementedError("v_rpe must be nn.Linear for the fused kernel")
            Dv = self.dim // H
            Wv, bv = self.v_rpe.weight, self.v_rpe.bias
            F = abar.shape[2]
            HF = H * F
            if (self.dim % 4 == 0 and self.dim <= 256 and HF % 4 == 0 and N > 0
                    and Wv.dtype == torch.float32):
                # agg + Wbd abar + bv (x) sump: one tensor-core GEMM + the glue kernels of
                # csrc/vrpe.cu (block-diagonal weight built on the device every step)
                y = ops.value_rpe(agg, abar, sump, Wv, bv, H, self.heads_share_rpe)
            else:
                blocks = [Wv] * H if self.heads_share_rpe else list(Wv.view(H, Dv, F))
                Wbd = torch.block_diag(*blocks)              # [C, H*F]
                rv = ops.linear(abar.reshape(N, HF), Wbd)    # [N, C]
                if bv is not None:
                    b_full = bv.repeat(H) if self.heads_share_rpe else bv
                    rv = rv + sump.repeat_interleave(Dv, dim=1) * b_full.view(1, -1)
                y = y + rv
        if self.out_proj is not None:
            y = self.out_proj(y)
        if self.out_drop is not None:
            y = self.out_drop(y)
        return y

    def extra_repr(self) -> str:
        return f'dim={self.dim}, num_heads={self.num_heads}'
