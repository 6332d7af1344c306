"""Determinism / race stress for the tcgen05 GEMMs: repeated launches must be bit-identical."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from superpoint_transformer_b200 import ops
dev = 'cuda'
for (M, N, K) in [(100003, 160, 128), (100003, 128, 160), (100000, 128, 128), (20000, 160, 128),
                  (50000, 256, 128), (100003, 32, 20)]:
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 2 + 0.5).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.3).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    ref = (a.double() @ w.double().t() + b.double())
    first = None
    bad = 0
    worst = 0.0
    for it in range(40):
        out = ops._gemm_nt(a, w, b)
        # interleave another shape to vary the machine state between launches
        if it % 3 == 0:
            ops._gemm_nt(a[:4096], w, None)
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        worst = max(worst, err)
        if first is None:
            first = out.clone()
        elif not torch.equal(out, first):
            bad += 1
            d = (out - first).abs()
            idx = d.flatten().argmax().item()
            print(f"  it {it}: mismatch rows/cols around {idx // N},{idx % N} max {float(d.max()):.3e} "
                  f"n_bad_rows {int((d.amax(1) > 0).sum())}", flush=True)
    print(f"M={M} N={N} K={K}: worst relerr {worst:.2e}, nondeterministic launches {bad}/39", flush=True)
