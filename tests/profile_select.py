"""(Profiling script, not a test.)  One NAG.select of level 1 of the cfg-2 partition after one
warm-up call, for an ncu launch list of the selection kernels:
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv \
        --log-file launches.csv python tests/profile_select.py"""
import sys

import torch

sys.path.insert(0, '.')
from superpoint_transformer_b200.synthetic import make_nag, CONFIGS   # noqa: E402

nag = make_nag(**CONFIGS['cfg2']).cuda()
g = torch.Generator().manual_seed(0)
n = nag[1].num_nodes
idx = torch.randperm(n, generator=g)[:int(0.6 * n)].cuda()
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = nag.select(1, idx)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(out)
