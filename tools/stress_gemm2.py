import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from superpoint_transformer_b200 import ops
dev = 'cuda'
M, N, K = 20000, 160, 128
g = torch.Generator().manual_seed(1)
a = (torch.randn(M, K, generator=g) * 2 + 0.5).to(dev)
w = (torch.randn(N, K, generator=g) * 0.3).to(dev)
b = torch.randn(N, generator=g).to(dev)
ref = (a.double() @ w.double().t() + b.double()).float()
for it in range(300):
    out = ops._gemm_nt(a, w, b)
    d = (out - ref).abs()
    bad = (d > 1e-3)
    rows = bad.any(1).nonzero().flatten()
    cols = bad.any(0).nonzero().flatten()
    if rows.numel() or it == 299: print(f"it {it}: bad rows {rows.numel()} tiles {sorted(set((rows // 128).tolist()))[:20]} "
          f"row%128 {sorted(set((rows % 128).tolist()))} cols {cols.tolist()[:40]}")
    if rows.numel():
        r = int(rows[0]); 
        print("   out", out[r, 124:136].tolist()); print("   ref", ref[r, 124:136].tolist())
        # is the bad slab equal to some other slab of the same row / same slab of other row?
        for s in range(4):
            if torch.allclose(out[r, 128:160], ref[r, 32*s:32*s+32], atol=1e-3): print("   == slab", s, "of same row")
        hits = (ref[:, 128:160] - out[r, 128:160]).abs().amax(1) < 1e-3
        print("   matches slab 4 of rows", hits.nonzero().flatten().tolist()[:5])
