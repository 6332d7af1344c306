"""(Timing script, not a test: pytest does not collect it; it lives here because only
tests/ may import oracle/.)
Time `sparse_sample` (csrc/sample.cu) beside the oracle's CPU restatement of the reference
(randperm + sort): N elements in G segments, the shape of SampleSubNodes on level 0.
    python tests/bench_sample.py [N] [G] [n_max] [n_min]
Device time: CUDA events around the call (it reads one scalar back); prints one JSON line."""
import json
import sys
import time

import torch

sys.path.insert(0, '.')
from oracle import sampling as OS                       # noqa: E402
from superpoint_transformer_b200 import ops             # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    n_max = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    n_min = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    g = torch.Generator().manual_seed(0)
    idx = torch.cat((torch.arange(G), torch.randint(0, G, (N - G,), generator=g)))
    idx = idx[torch.randperm(N, generator=g)]
    idx_d = idx.cuda()
    for _ in range(3):
        out = ops.sparse_sample(idx_d, n_max, n_min, num_segments=G, seed=1)
    torch.cuda.synchronize()
    reps = 20
    t0 = time.perf_counter()
    for i in range(reps):
        out = ops.sparse_sample(idx_d, n_max, n_min, num_segments=G, seed=i)
    torch.cuda.synchronize()
    cached_ms = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    for i in range(5):
        out = ops.sparse_sample(idx_d.clone(), n_max, n_min, num_segments=G, seed=i)
    torch.cuda.synchronize()
    cold_ms = (time.perf_counter() - t0) / 5 * 1e3
    t0 = time.perf_counter()
    ref, _ = OS.sparse_sample(idx, n_max, n_min)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({'N': N, 'G': G, 'n_max': n_max, 'n_min': n_min,
                      'sampled': int(out.numel()), 'oracle_sampled': int(ref.numel()),
                      'gpu_ms_csr_cached': round(cached_ms, 3),
                      'gpu_ms_with_csr_build': round(cold_ms, 3),
                      'cpu_oracle_ms': round(cpu_ms, 1), 'cpu_threads': torch.get_num_threads(),
                      'alg_MB': round((N * 4 + G * 24 + out.numel() * 8) / 1e6, 1)}))


if __name__ == '__main__':
    main()
