import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from superpoint_transformer_b200 import _lib
lib = _lib.load()
dev = 'cuda'
_p = lambda t: None if t is None else t.data_ptr()
def run(g, x):
    M, N = g.shape; K = x.shape[1]
    dW = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
    _lib.check(lib.spt_gemm_tn_acc(_p(g), M, N, g.stride(0), _p(x), K, x.stride(0), _p(dW), K,
                                   _p(db), torch.cuda.current_stream().cuda_stream), "tn")
    torch.cuda.synchronize()
    return dW
M, N, K = 2048, 128, 128
g = torch.ones(M, N, device=dev); x = torch.ones(M, K, device=dev)
dW = run(g, x)
print("ones: min/max/mean", float(dW.min()), float(dW.max()), float(dW.mean()), "expect", M)
g = torch.zeros(M, N, device=dev); x = torch.zeros(M, K, device=dev)
g[:, 3] = 1; x[:, 5] = 1
dW = run(g, x)
nz = dW.nonzero()
print("delta: nonzeros", nz[:10].tolist(), dW[nz[:, 0], nz[:, 1]][:10].tolist(), "expect [[3,5]]", M)
g = torch.zeros(M, N, device=dev); x = torch.zeros(M, K, device=dev)
g[7, 40] = 1; x[7, 77] = 2
dW = run(g, x)
nz = dW.nonzero()
print("single row: nonzeros", nz[:10].tolist(), dW[nz[:, 0], nz[:, 1]][:10].tolist(), "expect [[40,77]] 2")
g = torch.randn(M, N, device=dev); x = torch.randn(M, K, device=dev)
dW = run(g, x); ref = g.double().t() @ x.double()
print("rand: |dW| mean", float(dW.abs().mean()), "|ref| mean", float(ref.abs().mean()),
      "corr", float((dW.double() * ref).sum() / (ref * ref).sum()))
