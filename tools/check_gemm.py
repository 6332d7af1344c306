"""GPU check of spt_gemm_nt (tcgen05 path) against an fp64 product: accuracy + timing."""
import sys, os, faulthandler
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from superpoint_transformer_b200 import ops
faulthandler.dump_traceback_later(120, exit=True)
dev = 'cuda'
shapes = [(50000, 256, 128), (128, 128, 32), (128, 128, 128), (1000, 128, 128), (4096, 16, 12), (5000, 160, 128),
          (20000, 256, 128), (7777, 300, 64), (3001, 128, 256), (100000, 128, 128),
          (100000, 160, 128), (100000, 128, 256), (100000, 256, 128)]
if os.environ.get("SHAPES"):
    shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ["SHAPES"].split(",")]
worst = 0.0
for (M, N, K) in shapes:
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.2).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    ref = (a.double() @ w.double().t() + b.double())
    out = ops._gemm_nt(a, w, b)
    torch.cuda.synchronize()
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    out2 = ops._gemm_nt(a, w, None)
    err2 = float((out2.double() - (ref - b.double())).abs().max() / ref.abs().max())
    # timing: inputs > L2 for the big shapes, 20 launches
    for _ in range(3):
        ops._gemm_nt(a, w, b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops._gemm_nt(a, w, b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gb = (M * K + M * N + N * K) * 4 / 1e9
    print(f"M={M} N={N} K={K} relerr={err:.2e} nobias={err2:.2e} {ms*1e3:.1f} us "
          f"{gb/ms*1e3:.0f} GB/s", flush=True)
    worst = max(worst, err, err2)
print("worst", worst, "OK" if worst < 4e-6 else "FAIL")
