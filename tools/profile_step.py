"""Where does a resident-input step go?  torch.profiler kernel table (CUDA time) +
wall-clock per step.  Diagnostic only (numbers under a profiler are never bench values)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import superpoint_transformer_b200 as S  # noqa: E402
from superpoint_transformer_b200.distributed import FlatGradients  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    net = S.SPT(mlp_norm=S.nn.GraphNorm, norm=S.nn.GraphNorm, **bench.model_kwargs(S))
    net.apply(S.init_weights)
    head = S.nn.Classifier(bench.DIM, bench.NUM_CLASSES)
    model = torch.nn.ModuleDict(dict(net=net, head=head)).to(dev)
    params = list(model.parameters())
    flat = FlatGradients(params)
    opt = torch.optim.AdamW(params, lr=1e-3, fused=True)
    host_nag, host_labels, _ = bench.host_scene(bench.LEVELS, seed=1)
    nag = bench.device_transforms(S, host_nag.to(dev))
    labels = host_labels.to(dev)
    base = {l: (nag[l].edge_attr, nag[l]['hf']) for l in nag.level_range}

    def step():
        for l, (ea, hf) in base.items():
            d = nag[l]
            d.x, d.edge_attr, d['hf'], d.diameter = None, ea, hf, None
        flat.release()
        out = net(nag)
        loss = torch.nn.functional.cross_entropy(head(out), labels)
        loss.backward()
        flat.collect()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    print(f"wall per step: {(time.time() - t) / 5 * 1e3:.2f} ms")
    # host-only cost: how long does it take to ENQUEUE a step?
    t = time.time()
    step()
    t_enq = time.time() - t
    torch.cuda.synchronize()
    print(f"enqueue time of one step (host): {t_enq * 1e3:.2f} ms")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        step()
        torch.cuda.synchronize()
    # which aten ops launch the glue kernels (fills / copies / adds / reductions), by input shape
    glue = []
    for e in prof.key_averages(group_by_input_shape=True):
        if e.self_device_time_total > 0 and e.key.startswith('aten::'):
            glue.append((e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:90]))
    glue.sort(reverse=True)
    print("aten ops with their own kernels (self device time):")
    for t_us, n, k, shp in glue[:70]:
        print(f"{t_us / 1e3:9.3f} ms  x{n:<4d} {k:28s} {shp}")
    rows = []
    for e in prof.key_averages():
        if e.device_time_total > 0:
            rows.append((e.device_time_total, e.count, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"total device time {tot / 1e3:.2f} ms over {sum(r[1] for r in rows)} launches")
    for t_us, n, k in rows[:45]:
        print(f"{t_us / 1e3:9.3f} ms {100 * t_us / tot:5.1f}%  x{n:<4d} {k[:110]}")


if __name__ == '__main__':
    main()
