"""Runs the attention core (fwd + bwd) at a cfg-2 level shape a few times: the command ncu
captures (tools/README).  N=<rows> ITERS=<n> python tools/run_attn.py; prints CUDA-event times."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superpoint_transformer_b200 import ops  # noqa: E402
from superpoint_transformer_b200.synthetic import _trimmed_graph  # noqa: E402

N = int(os.environ.get('N', 100_000))
ITERS = int(os.environ.get('ITERS', 3))
MORTON = os.environ.get('SORTED', '0') == '1'
dev = 'cuda'
rng = np.random.default_rng(1)
se = torch.from_numpy(_trimmed_graph(rng, N, 16))
if MORTON:   # neighbours close in id space (stands for Morton-sorted node ids)
    off = torch.from_numpy(rng.integers(1, 64, size=se.shape[1]))
    se = torch.stack((se[0], (se[0] + off) % N))
ei = torch.cat([se, se.flip(0), torch.arange(N).repeat(2, 1)], dim=1).to(dev)
order = torch.argsort(ei[0], stable=True)
ei = ei[:, order].contiguous()
ops.mark_csr_ordered(ei)
gi = ops.build_graph_index(ei, N)
E = ei.shape[1]
H, D, C, F = (int(os.environ.get(k, v)) for k, v in (('H', 4), ('D', 4), ('C', 128), ('F', 32)))
g = torch.Generator().manual_seed(0)
qkv = torch.randn(N, 2 * H * D + C, generator=g).to(dev).requires_grad_(True)
a = torch.randn(E, F, generator=g).to(dev).requires_grad_(True)
mk = lambda *s: (torch.randn(*s, generator=g) * 0.2).to(dev).requires_grad_(True)  # noqa: E731
Wq, bq, Wk, bk = mk(H * D, F), mk(H * D), mk(H * D, F), mk(H * D)
ops.set_attention_storage(os.environ.get('SPT_ATTN_STORAGE', 'fp32'))
ops.enable_event_timing(True)


def step():
    agg, abar, sump = ops.attention_core(qkv, None, a, Wq, bq, Wk, bk, gi, H, D,
                                         ops.SCALE_D_TIMES_G, (C // H) ** -0.5)
    (agg.sum() + abar.sum()).backward()


for it in range(ITERS):
    step()
torch.cuda.synchronize()
if os.environ.get('PROFILE', '0') == '1':   # per-kernel device times (torch.profiler / CUPTI)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for it in range(3):
            step()
        torch.cuda.synchronize()
    for ev in sorted(prof.key_averages(), key=lambda x: -x.device_time_total):
        if ev.device_time_total > 0 and ('spt' in ev.key or 'umma' in ev.key):
            print(f'  [kernel] {ev.device_time_total / ev.count / 1e3:.4f} ms x{ev.count // 3}/step  '
                  f'{ev.key[:90]}')
# bare gather ceiling: out[row] = sum_e v[col[e]] (the 512-byte rows the forward gathers, nothing
# else) with the CSR segment-sum kernel — 64 resident warps/SM, 2 rows in flight per warp
seg = ops.SegmentIndex(gi.rowptr, gi.col, None, E, N, None)
vmat = qkv.detach()[:, 2 * H * D:].contiguous()
for it in range(ITERS):
    ops._segment_pool_fwd(vmat, seg, 'sum')
torch.cuda.synchronize()
acc = {}
for tag, meta, s, e in ops.timing_records():
    acc.setdefault(tag, []).append(s.elapsed_time(e))
print(f'N={N} E={E} C={C} H={H} D={D} F={F} sorted={MORTON} storage={ops.ATTN_STORAGE} split={ops.ATTN_SPLIT}')
for k, v in acc.items():
    print(f'  {k}: min {min(v):.4f} ms  last {v[-1]:.4f} ms')
if 'segment_pool_fwd' in acc:
    t = min(acc['segment_pool_fwd'])
    gb = (E * C * 4 + N * C * 4 + E * 4) / 1e9
    print(f'  bare gather of the v rows (segment-sum by col): {t:.4f} ms = {gb / t * 1e3:.0f} GB/s '
          f'of gathered + written bytes')
