"""(Timing script, not a test: pytest does not collect it; it lives here because only
tests/ may import oracle/.)
Time NAG.select (csrc/select.cu) on a BASELINE-config partition against the oracle's CPU
restatement of the reference algorithm (sort-based relabel), level by level:
    python tests/bench_select.py [cfg2|cfg3|cfg5] [fraction]
Wall clock around the call with a device synchronize on both sides (the call reads a few
8-byte counters back, so it is not a pure device region).  Prints one JSON line per level."""
import json
import sys
import time

import torch

sys.path.insert(0, '.')
from oracle import select as O                                   # noqa: E402
from superpoint_transformer_b200 import ops                       # noqa: E402
from superpoint_transformer_b200.data import Cluster              # noqa: E402
from superpoint_transformer_b200.synthetic import make_nag, CONFIGS   # noqa: E402


def level_of(data):
    out = {}
    for k in data.keys:
        if k.startswith('_'):
            continue
        v = data[k]
        out[k] = {'pointers': v.pointers, 'points': v.points} if isinstance(v, Cluster) else v
    return out


def level_bytes(level):
    n = 0
    for v in level.values():
        n += sum(t.numel() * t.element_size() for t in v.values()) if isinstance(v, dict) \
            else v.numel() * v.element_size()
    return n


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
    nag = make_nag(**CONFIGS[cfg])
    levels = [level_of(nag[i]) for i in nag.level_range]
    dev = nag.cuda()
    g = torch.Generator().manual_seed(0)
    for fused, i_level in [(f, l) for f in (True, False) for l in nag.level_range]:
        ops.set_select_fused(fused)
        n = nag[i_level].num_nodes
        idx = torch.randperm(n, generator=g)[:int(frac * n)]
        idx_d = idx.cuda()
        for _ in range(3):
            res = dev.select(i_level, idx_d)
        torch.cuda.synchronize()
        before = ops.launch_count() if hasattr(ops, 'launch_count') else 0
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            res = dev.select(i_level, idx_d)
        torch.cuda.synchronize()
        gpu_ms = (time.perf_counter() - t0) / reps * 1e3
        launches = ((ops.launch_count() - before) // reps) if hasattr(ops, 'launch_count') else None
        t0 = time.perf_counter()
        want = O.nag_select(levels, nag.start_i_level, i_level, idx)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        out_bytes = sum(level_bytes(level_of(res[i])) for i in res.level_range)
        in_bytes = sum(level_bytes(l) for l in levels)
        print(json.dumps({
            'config': cfg, 'path': 'fused' if fused else 'primitives', 'i_level': i_level, 'selected': int(idx.numel()), 'of': n,
            'gpu_ms': round(gpu_ms, 3), 'cpu_oracle_ms': round(cpu_ms, 1),
            'cpu_threads': torch.get_num_threads(), 'speedup': round(cpu_ms / gpu_ms, 1),
            'launches': launches, 'in_MB': round(in_bytes / 1e6, 1),
            'out_MB': round(out_bytes / 1e6, 1),
            'moved_GBps': round((in_bytes + out_bytes) / gpu_ms / 1e6, 1)}))


if __name__ == '__main__':
    main()
