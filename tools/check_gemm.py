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

# ---- gemm_tn_acc: dW[N,K] += g[M,N]^T x[M,K], db[N] += colsum(g) ----
from superpoint_transformer_b200 import _lib
lib = _lib.load()
_p = lambda t: None if t is None else t.data_ptr()
tn_shapes = [(5000, 128, 128), (4096, 32, 20), (3000, 160, 128), (10007, 128, 292), (2048, 16, 12),
             (100000, 128, 128), (100000, 160, 128), (1696398, 32, 32), (1696398, 32, 20),
             (20000, 256, 128)]
if os.environ.get("TN_SHAPES"):
    tn_shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ["TN_SHAPES"].split(",")]
worst = 0.0
for (M, N, K) in tn_shapes:
    gen = torch.Generator().manual_seed(M + 3 * N + K)
    g = torch.randn(M, N, generator=gen).to(dev)
    x = torch.randn(M, K, generator=gen).to(dev)
    def run(dW, db):
        _lib.check(lib.spt_gemm_tn_acc(_p(g), M, N, g.stride(0), _p(x), K, x.stride(0), _p(dW), K,
                                       _p(db), torch.cuda.current_stream().cuda_stream), "tn")
    dW = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
    run(dW, db); torch.cuda.synchronize()
    ref = g.double().t() @ x.double(); refb = g.double().sum(0)
    err = float((dW.double() - ref).abs().max() / ref.abs().max())
    errb = float((db.double() - refb).abs().max() / refb.abs().max())
    for _ in range(3):
        run(dW, db)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run(dW, db)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    gb = (M * K + M * N) * 4 / 1e9
    print(f"TN M={M} N={N} K={K} relerr={err:.2e} colsum={errb:.2e} {ms*1e3:.1f} us {gb/ms*1e3:.0f} GB/s", flush=True)
    worst = max(worst, err, errb)
print("tn worst", worst, "OK" if worst < 3e-5 else "FAIL")  # fp32 accumulation over up to 1.7M rows
