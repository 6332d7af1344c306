"""Device-side A/B of the row-tile attention kernels (csrc/attention_tile.cuh) and of the split
kernels (csrc/attention_split.cuh, `rpw=split`) against the
round-1 per-edge kernels (attention_fast.cuh, 1 row per warp — the configuration the golden
vectors pin): prints the max abs difference of every output and gradient.  Diagnosis tool;
the parity tests proper are tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superpoint_transformer_b200 as S  # noqa: E402
from superpoint_transformer_b200 import ops  # noqa: E402
from superpoint_transformer_b200.synthetic import _trimmed_graph  # noqa: E402

DEV = 'cuda'


def graph(n, seed, hub=False):
    rng = np.random.default_rng(seed)
    se = torch.from_numpy(_trimmed_graph(rng, n, 16))
    ei = torch.cat([se, se.flip(0), torch.arange(n).repeat(2, 1)], dim=1)
    if hub:
        g = torch.Generator().manual_seed(seed)
        ei = torch.cat([ei, torch.stack((torch.zeros(300, dtype=torch.long),
                                         torch.randint(1, n, (300,), generator=g))),
                        torch.stack((torch.randint(0, n - 50, (40,), generator=g),
                                     torch.full((40,), n - 1)))], dim=1)
        ei = ei[:, ei[0] < n - 20]          # last 20 rows: no outgoing edges
    g = torch.Generator().manual_seed(seed + 1)
    return ei[:, torch.randperm(ei.shape[1], generator=g)]


def run(N, seed, hub, want_abar, use_q, use_k, env, split=False, shape=(4, 4, 128, 32)):
    for k in ('SPT_ATTN_NO_TILE', 'SPT_ATTN_ROWS_PER_WARP'):
        os.environ.pop(k, None)
    os.environ.update(env)
    ops.set_attention_split(split)
    H, D, C, F = shape
    g = torch.Generator().manual_seed(seed)
    ei = graph(N, seed, hub).to(DEV)
    E = ei.shape[1]
    gi = ops.build_graph_index(ei, N)
    qkv = torch.randn(N, 2 * H * D + C, generator=g).to(DEV).requires_grad_(True)
    a = ops.permute_rows(torch.randn(E, F, generator=g).to(DEV), gi.perm).detach().requires_grad_(True)
    mk = lambda *s: (torch.randn(*s, generator=g) * 0.3).to(DEV).requires_grad_(True)  # noqa: E731
    Wq, bq = (mk(H * D, F), mk(H * D)) if use_q else (None, None)
    Wk, bk = (mk(H * D, F), mk(H * D)) if use_k else (None, None)
    agg, abar, sump = ops.attention_core(qkv, None, a, Wq, bq, Wk, bk, gi, H, D,
                                         ops.SCALE_D_TIMES_G, (C // H) ** -0.5,
                                         want_abar=want_abar)
    pr1 = torch.randn(agg.shape, generator=g).to(DEV)
    loss = (agg * pr1).sum()
    if abar is not None:
        pr2 = torch.randn(abar.shape, generator=g).to(DEV)
        loss = loss + (abar * pr2).sum()
    loss.backward()
    torch.cuda.synchronize()
    out = dict(agg=agg, sump=sump, dqkv=qkv.grad, da=a.grad)
    if abar is not None:
        out['abar'] = abar
    for nm, p in (('dWq', Wq), ('dbq', bq), ('dWk', Wk), ('dbk', bk)):
        if p is not None:
            out[nm] = p.grad
    out['dq'] = qkv.grad[:, :H * D]
    out['dk'] = qkv.grad[:, H * D:2 * H * D]
    out['dv'] = qkv.grad[:, 2 * H * D:]
    return {k: v.detach().clone() for k, v in out.items()}


def main():
    cases = [(403, 3, True, True, True, True), (3001, 4, False, True, True, True),
             (3001, 5, True, False, False, True), (20000, 6, False, True, True, True)]
    if os.environ.get('SMALL'):      # compute-sanitizer runs
        cases = cases[:2]
    worst = 0.0
    for N, seed, hub, want_abar, use_q, use_k in cases:
        ref = run(N, seed, hub, want_abar, use_q, use_k,
                  dict(SPT_ATTN_NO_TILE='1', SPT_ATTN_ROWS_PER_WARP='1'))
        for rpw in ('1', '3', '8', 'split'):
            if rpw == 'split':     # edge pass + row pass (csrc/attention_split.cuh)
                got = run(N, seed, hub, want_abar, use_q, use_k, {}, split=True)
            else:
                got = run(N, seed, hub, want_abar, use_q, use_k, dict(SPT_ATTN_ROWS_PER_WARP=rpw))
            line = []
            for k in ref:
                sc = max(float(ref[k].abs().max()), 1e-6)
                err = float((got[k] - ref[k]).abs().max()) / sc
                nan = bool(torch.isnan(got[k]).any())
                worst = max(worst, err if not nan else 1e9)
                line.append(f'{k}={err:.1e}' + ('(NaN)' if nan else ''))
            print(f'N={N} hub={hub} abar={want_abar} q={use_q} rpw={rpw}: ' + ' '.join(line),
                  flush=True)
    # the shipped head layout (H = 16; C = 64 and 128): split16 kernels against the generic ones
    for shape in ((16, 4, 64, 32), (16, 4, 128, 32)):
        for N, seed, hub, want_abar, use_q, use_k in cases[:3]:
            ref = run(N, seed, hub, want_abar, use_q, use_k, {}, split=False, shape=shape)
            got = run(N, seed, hub, want_abar, use_q, use_k, {}, split=True, shape=shape)
            line = []
            for k in ref:
                sc = max(float(ref[k].abs().max()), 1e-6)
                err = float((got[k] - ref[k]).abs().max()) / sc
                nan = bool(torch.isnan(got[k]).any())
                if not (k == 'dbk' and not use_q):
                    worst16 = max(globals().get('worst16', 0.0), err if not nan else 1e9)
                    globals()['worst16'] = worst16
                line.append(f'{k}={err:.1e}' + ('(NaN)' if nan else ''))
            print(f'H16 C={shape[2]} N={N} hub={hub} abar={want_abar} q={use_q} rpw=split16: '
                  + ' '.join(line), flush=True)
    print('WORST relative-to-scale difference:', worst)
    print('WORST H=16 difference:', globals().get('worst16'))


if __name__ == '__main__':
    main()
