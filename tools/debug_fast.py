import sys, os, time, faulthandler
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import superpoint_transformer_b200 as S
from superpoint_transformer_b200 import ops
faulthandler.dump_traceback_later(100, exit=True)
dev = 'cuda'
def log(*a):
    print(*a, flush=True)
for N in [int(x) for x in os.environ.get("NS", "403,5000,20000,100000").split(",")]:
    g = torch.Generator().manual_seed(0)
    E = N * 17
    ei = torch.randint(0, N, (2, E), generator=g).to(dev)
    gi = ops.build_graph_index(ei, N)
    torch.cuda.synchronize(); log(N, 'graph ok')
    H, D, C, F = 4, 4, 128, 32
    qkv = torch.randn(N, 2*H*D + C, device=dev, requires_grad=True)
    a = torch.randn(E, F, device=dev, requires_grad=True)
    Wq = (torch.randn(H*D, F, device=dev) * 0.1).requires_grad_(True)
    Wk = (torch.randn(H*D, F, device=dev) * 0.1).requires_grad_(True)
    bq = torch.zeros(H*D, device=dev, requires_grad=True); bk = torch.zeros(H*D, device=dev, requires_grad=True)
    t = time.time()
    agg, abar, sump = ops.attention_core(qkv, None, a, Wq, bq, Wk, bk, gi, H, D, ops.SCALE_D_TIMES_G, 32 ** -0.5)
    torch.cuda.synchronize(); log(N, 'fwd ok', time.time() - t, float(agg.abs().mean()))
    t = time.time()
    if os.environ.get('SYNC_EACH'):
        import superpoint_transformer_b200._lib as L
        lib = L.load()
        for name in ('spt_attn_bwd_rows', 'spt_attn_bwd_targets'):
            orig = getattr(lib, name)
            def wrap(*a, _o=orig, _n=name):
                r = _o(*a); torch.cuda.synchronize(); log('   done', _n); return r
            try:
                setattr(lib, name, wrap)
            except Exception as ex:
                log('cannot wrap', ex)
    (agg.sum() + abar.sum()).backward()
    torch.cuda.synchronize(); log(N, 'bwd ok', time.time() - t, float(a.grad.abs().mean()), float(Wq.grad.abs().mean()))
log('done')
