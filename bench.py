#!/usr/bin/env python
"""bench.py — superpoints/s (fwd+bwd) of the hierarchical superpoint-graph
attention + pooling stack on the BASELINE.json cfg-2 workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm (default): one process per GPU (torchrun for N>1), every rank owns one
cfg-2 scene (scene-shard data parallelism, weak scaling), one flat NCCL gradient
all-reduce per step.  Prints ONE JSON line (contract in the task statement) with
`value` (inputs resident in HBM), `e2e` (host buffers -> H2D -> on-device
transforms -> CSR build -> fwd+bwd+step -> D2H loss), `roofline` (dominant
kernel, timed live with CUDA events), `cpu_baseline` (oracle port on host cores,
bounded sample).

Reference arm (--impl reference): the reference's CPU path for the same metric —
the oracle restatement of its source (oracle/path.py; the reference itself is
Python and cannot be installed here: torch_scatter / torch_geometric /
lightning / hydra are absent) on all host threads, each step a
bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "superpoints/sec (fwd+bwd) on 100k-SP 3-level NAG"
UNIT = "superpoints/s"
LEVELS = [100_000, 20_000, 4_000]
MEAN_DEGREE = 16
DIM, HEADS, QK_DIM, RPE_DIM, HF_DIM, NUM_CLASSES = 128, 4, 4, 32, 12, 13
WORKLOAD = ("cfg2: S3DIS-shaped 3-level NAG 100k/20k/4k superpoints, sym. degree "
            "~clamp(Poisson16,1,30)+self-loop, random node ids; nano-3 SPT C=128 H=4 "
            "qk_dim=4 F=32 (h_edge_mlp 18->32->32, node_mlp 12->32->32), 3 blocks/down "
            "level + 1 block/up level, k/q/v RPE, max-pool, GraphNorm, 13-class CE head, "
            "AdamW step; fp32 (ieee matmul)")


def model_kwargs(S):
    inj = 3 + 1 + 32
    return dict(
        nano=True, segment_hf=['hf'], down_dim=[DIM] * 3,
        down_in_mlp=[[inj, DIM, DIM], [inj + DIM, DIM, DIM], [inj + DIM, DIM, DIM]],
        down_num_heads=HEADS, down_num_blocks=3, down_ffn_ratio=1, up_dim=[DIM] * 2,
        up_in_mlp=[[inj + 2 * DIM, DIM, DIM], [inj + 2 * DIM, DIM, DIM]], up_num_heads=HEADS,
        up_num_blocks=1, node_mlp=[HF_DIM, 32, 32], h_edge_mlp=[18, RPE_DIM, RPE_DIM],
        qk_dim=QK_DIM, in_rpe_dim=RPE_DIM, k_rpe=True, q_rpe=True, v_rpe=True, no_ffn=True,
        use_diameter_parent=True, pool='max')


# --------------------------------------------------------------------------- #
#  clocks sampler (nvidia-smi while the timed region runs)
# --------------------------------------------------------------------------- #
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                  'sw_power_cap'), parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- #
#  workload
# --------------------------------------------------------------------------- #
def host_scene(levels, seed):
    """compact host-side NAG as the reference ships it to the device (trimmed graph,
    fp16 raw edge attributes: configs/datamodule/semantic/default.yaml:40-42), pinned."""
    from superpoint_transformer_b200.synthetic import make_nag
    nag = make_nag(levels, mean_degree=MEAN_DEGREE, seed=seed)
    nbytes = 0
    for d in nag:
        d.edge_attr = d.edge_attr.half()
        d.sub = None  # rebuilt on device from super_index (bit-exact, tests)
        for k in d.keys:
            v = d[k]
            if torch.is_tensor(v):
                d[k] = v.pin_memory()
                nbytes += v.numel() * v.element_size()
    labels = torch.randint(0, NUM_CLASSES, (levels[0],),
                           generator=torch.Generator().manual_seed(seed + 7)).pin_memory()
    nbytes += labels.numel() * 8
    return nag, labels, nbytes


def device_transforms(S, nag):
    nag = S.transforms.NodeSize()(nag)
    # csr_order: edges emitted grouped by source (the model is invariant to edge order), so the
    # attention blocks read edge_attr in place
    return S.transforms.OnTheFlyHorizontalEdgeFeatures(add_self_loops=True, csr_order=True)(nag)


def attn_bytes(tag, m, elt=4, idx=4):
    """ALGORITHMIC (compulsory) bytes of one launch: every input / output tensor once
    (DESIGN.md §Kernels).  HD2 = 2*H*D, C = H*Dv."""
    R, E, H, D, Dv, F = m['R'], m['E'], m['H'], m['D'], m['Dv'], m['F']
    T = m.get('T', R)
    C, HD2, HF = H * Dv, 2 * H * D, H * F
    abar = HF if m.get('abar') else 0
    if tag == 'attn_fwd':       # qkv in; a in; rowptr+col; agg+abar+sump+m+z out
        return (R * (HD2 + C) + E * F + R * (C + abar + 3 * H)) * elt + (R + 1 + E) * idx
    if tag == 'attn_bwd_rows':  # qkv, a, stats, agg/abar, dY/dabar in; dq, da, P, G out
        da = E * F if m.get('da') else 0
        return (R * (HD2 + C) + E * F + 2 * R * H + 2 * R * (C + abar) + R * HD2 // 2 + da +
                E * H + E * HD2) * elt + (R + 1 + E) * idx
    if tag == 'attn_bwd_targets':  # P, dk_e half of G, dY in; dk, dv out
        return (E * H + E * HD2 // 2 + R * C + T * (HD2 // 2 + C)) * elt + (T + 1 + 2 * E) * idx
    if tag == 'attn_bwd_weights':  # G, a in
        return (E * HD2 + E * F) * elt
    return 0


def kernel_bytes(tag, m):
    if tag.startswith('gemm'):   # A [M,K] + W [N,K] + C [M,N] once each (dW: A, B in, C out)
        return (m['M'] * m['K'] + m['N'] * m['K'] + m['M'] * m['N']) * 4
    return attn_bytes(tag, m)


def run_own(args):
    import superpoint_transformer_b200 as S
    from superpoint_transformer_b200 import ops
    from superpoint_transformer_b200.distributed import (FlatGradients,
                                                         init_process_group_from_env)
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (own arm) needs a CUDA device; there is no CPU fallback")
    rank, world, local = init_process_group_from_env()
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False   # fp32 parity setting ('highest')
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)

    net = S.SPT(mlp_norm=S.nn.GraphNorm, norm=S.nn.GraphNorm, **model_kwargs(S))
    net.apply(S.init_weights)
    head = S.nn.Classifier(DIM, NUM_CLASSES)
    model = torch.nn.ModuleDict(dict(net=net, head=head)).to(dev)
    params = list(model.parameters())
    flat = FlatGradients(params)
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-4, fused=True,
                            capturable=not args.no_graph)

    host_nag, host_labels, h2d_bytes = host_scene(LEVELS, seed=1 + rank)
    n1 = LEVELS[0]

    def fwd_bwd(nag, labels):
        flat.release()
        out = net(nag)
        loss = torch.nn.functional.cross_entropy(head(out), labels)
        loss.backward()
        flat.collect()
        return loss

    def step(nag, labels):
        loss = fwd_bwd(nag, labels)
        flat.all_reduce()
        opt.step()
        return loss

    class GraphedStep:
        """fwd+bwd and the optimizer step captured as two CUDA graphs (the NCCL
        gradient all-reduce stays between them, eager).  The launch-bound inner loop
        (~800 kernels / step) replays without Python or driver launch overhead."""

        def __init__(self, body):
            self.loss = None
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    body()
                    flat.all_reduce()
                    opt.step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g1):
                self.loss = body()
            self.g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g2):
                opt.step()

        def __call__(self):
            self.g1.replay()
            flat.all_reduce()
            self.g2.replay()
            return self.loss

    def fresh_device_nag():
        nag = host_nag.to(dev, non_blocking=True)
        return device_transforms(S, nag), host_labels.to(dev, non_blocking=True)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), t0, t1

    # ---- resident-input phase ------------------------------------------------
    res_nag, res_labels = fresh_device_nag()
    base = {l: (res_nag[l].x, res_nag[l].edge_attr, res_nag[l]['hf']) for l in res_nag.level_range}

    def resident_step():
        # SPT.forward rewrites x / edge_attr / hf on the NAG: restore the inputs (no copy)
        for l, (x, ea, hf) in base.items():
            d = res_nag[l]
            d.x, d.edge_attr, d['hf'] = x, ea, hf
            d.diameter = None
        return step(res_nag, res_labels)

    # CUDA graphs are captured first (their own warm-up runs on the capture stream, before
    # any eager step creates autograd nodes on the default stream)
    graph_mode, run_resident = False, resident_step
    if not args.no_graph:
        try:
            def body():
                for l, (x, ea, hf) in base.items():
                    d = res_nag[l]
                    d.x, d.edge_attr, d['hf'] = x, ea, hf
                    d.diameter = None
                return fwd_bwd(res_nag, res_labels)
            run_resident = GraphedStep(body)
            graph_mode = True
        except Exception as ex:  # noqa: BLE001
            sys.stderr.write(f"[bench] CUDA graph capture failed, running eager: {ex!r}\n")
            torch.cuda.synchronize()
            run_resident = resident_step
    for _ in range(args.warmup):
        run_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ms, t0, t1 = timed(run_resident, args.steps)
    clocks = sampler.stop(t0, t1) if sampler else None
    # count my launches per step and time my kernels with CUDA events (eager steps: the
    # same kernels on the same inputs; events cannot be read back from a graph replay)
    for _ in range(2):
        resident_step()
    ops.enable_event_timing(True)
    l0 = ops.launch_count()
    n_evt_steps = 3
    ms_eager, _, _ = timed(resident_step, n_evt_steps)
    launches_per_step = (ops.launch_count() - l0) // n_evt_steps
    records = ops.timing_records()
    ops.enable_event_timing(False)
    ms_eager_step = ms_eager / n_evt_steps
    launches = launches_per_step * args.steps
    ms_per_step = ms / args.steps
    value = world * n1 / (ms_per_step * 1e-3)

    # ---- per-kernel roofline from the live events ------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # noqa: BLE001
        pass
    peak_gbs, peak_src = (peaks['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)') \
        if 'hbm_gbs' in peaks else (6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)')
    agg = {}
    for tag, meta, s, e in records:
        size = meta.get('E', meta.get('M'))
        shape = f"E={meta['E']},rows={meta['R']}" if 'E' in meta else \
            f"M={meta['M']},N={meta['N']},K={meta['K']}"
        key = (tag, shape)
        a = agg.setdefault(key, dict(tag=tag, E=size, R=shape, ms=0.0, n=0,
                                     bytes=kernel_bytes(tag, meta)))
        a['ms'] += s.elapsed_time(e)
        a['n'] += 1
    kernels = []
    for a in agg.values():
        avg = a['ms'] / a['n']
        gbs = a['bytes'] / (avg * 1e-3) / 1e9
        kernels.append(dict(kernel=a['tag'], E=a['E'], shape=a['R'], launches=a['n'],
                            avg_ms=round(avg, 4), share_of_step=round(a['ms'] / ms_eager, 4),
                            algorithmic_MB=round(a['bytes'] / 1e6, 2),
                            achieved_GBs=round(gbs, 1), frac=round(gbs / peak_gbs, 4)))
    kernels.sort(key=lambda k: -k['share_of_step'])
    if args.kernels_out and rank == 0:
        with open(args.kernels_out, 'w') as fh:
            json.dump(dict(eager_ms_per_step=ms_eager_step, steps=n_evt_steps, kernels=kernels),
                      fh, indent=1)
    roofline = None
    if kernels:
        top = kernels[0]
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, 'profiles', 'dram_traffic.json')))
            traffic = prof.get(f"{top['kernel']}:{top['E']}")
        except Exception:  # noqa: BLE001
            pass
        roofline = dict(bound='hbm', kernel=f"{top['kernel']} ({top['shape']})",
                        achieved=top['achieved_GBs'], peak=peak_gbs, unit='GB/s',
                        frac=top['frac'], traffic=traffic, peak_source=peak_src,
                        algorithmic_bytes=int(top['algorithmic_MB'] * 1e6),
                        avg_launch_ms=top['avg_ms'], share_of_step=top['share_of_step'])

    # ---- end-to-end phase: host buffers every step -------------------------------
    def e2e_step():
        nag, labels = fresh_device_nag()
        loss = step(nag, labels)
        return float(loss.item())  # D2H read of the step's result

    e2e_graph = False
    e2e_h2d = "eager: one pinned->device copy per tensor on the compute stream"
    if graph_mode:
        def make_static_set():
            # static device input buffers + the graphs (transforms + CSR build + fwd + bwd | step)
            snag = host_nag.to(dev)
            slab = host_labels.to(dev)
            pairs = []
            for l in host_nag.level_range:
                hd, sd_ = host_nag[l], snag[l]
                for k in hd.keys:
                    if torch.is_tensor(hd[k]):
                        pairs.append((sd_[k], hd[k]))
            pairs.append((slab, host_labels))

            def body():
                nag = snag.clone()
                nag = device_transforms(S, nag)
                return fwd_bwd(nag, slab)
            return pairs, GraphedStep(body)

        try:
            # double-buffered inputs: while step i replays on the compute stream, the pinned-memory
            # H2D copy of step i+1's inputs runs on a copy stream into the other buffer set.
            # Every step still uploads its 43 MB and reads its loss back.
            sets = [make_static_set(), make_static_set()]
            copy_stream = torch.cuda.Stream()
            h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
            state = {'i': 0}

            def enqueue_h2d(b):
                # the graph that last read buffer set b is already enqueued on the compute stream
                copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(copy_stream):
                    for dst, src in sets[b][0]:
                        dst.copy_(src, non_blocking=True)
                    h2d_done[b].record(copy_stream)

            enqueue_h2d(0)

            def e2e_step():  # noqa: F811
                b = state['i'] & 1
                enqueue_h2d(1 - b)
                torch.cuda.current_stream().wait_event(h2d_done[b])
                loss = sets[b][1]()
                state['i'] += 1
                return float(loss.item())
            e2e_step()
            e2e_graph = True
            e2e_h2d = ("double-buffered: the pinned->device copy of step i+1 overlaps the compute "
                       "of step i (copy stream), every step uploads all inputs")
        except Exception as ex:  # noqa: BLE001
            sys.stderr.write(f"[bench] pipelined e2e failed ({ex!r}); single-buffer graphs\n")
            torch.cuda.synchronize()
            try:
                pairs, e2e_graphed = make_static_set()

                def e2e_step():  # noqa: F811
                    for dst, src in pairs:
                        dst.copy_(src, non_blocking=True)
                    loss = e2e_graphed()
                    return float(loss.item())
                e2e_graph = True
                e2e_h2d = "single buffer: H2D copies on the compute stream before each replay"
            except Exception as ex2:  # noqa: BLE001
                sys.stderr.write(f"[bench] e2e CUDA graph capture failed, running eager: {ex2!r}\n")
                torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    e2e_steps = args.steps
    ms_e2e, _, _ = timed(e2e_step, e2e_steps)
    e2e_value = world * n1 / (ms_e2e / e2e_steps * 1e-3)

    # ---- CPU baseline (oracle port, rank 0, N=1 only) ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_sample(target_seconds=20.0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "superpoints_per_rank": n1,
                       "edges_level1": int(res_nag[1].edge_index.shape[1]),
                       "parallelism": f"scene-shard dp{world} (one scene per GPU, flat NCCL "
                                      f"grad all-reduce)",
                       "l2": "inputs_exceed_l2 (per-step working set > 126 MB; no flush needed)",
                       "cuda_graph": {"value": graph_mode, "e2e": e2e_graph,
                                      "eager_ms_per_step": round(ms_eager_step, 4)},
                       "matmul": "fp32-accurate 3xTF32 on tcgen05 tensor cores, TMEM accumulators, TMA (csrc/gemm_umma.cu)",
                       "csr": "graph CSR cached across steps in `value` (amortised, SURVEY §8d); "
                              "rebuilt every step in `e2e`; on-the-fly edges emitted in CSR order "
                              "(OnTheFlyHorizontalEdgeFeatures(csr_order=True))"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT,
                    "ms_per_step": round(ms_e2e / e2e_steps, 4),
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 4,
                    "h2d": e2e_h2d},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "kernels": kernels[:8],
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------- #
#  CPU reference (oracle port)
# --------------------------------------------------------------------------- #
def _cpu_scene(levels, seed):
    """same generator + the oracle's CPU transforms -> NAG ready for spt_forward"""
    from oracle import path as P
    from superpoint_transformer_b200.synthetic import make_nag
    nag = make_nag(levels, mean_degree=MEAN_DEGREE, seed=seed)
    size = nag[1].node_size
    for l in nag.level_range:
        d = nag[l]
        ei, ea = P.horizontal_edge_features(d.edge_index, d.edge_attr, d.pos, d.normal,
                                            d['log_length'], d['log_surface'],
                                            d['log_volume'], d['log_size'])
        d.edge_index, d.edge_attr = P.add_self_loops(ei, ea, d.num_nodes)
        if l > 1:
            size = torch.from_numpy(P.node_size_np(nag[l - 1].super_index.numpy(), d.num_nodes,
                                                   size.numpy()))
            d.node_size = size
    return nag


def _cpu_state_dict():
    """random-init parameters with the reference key names (built from the product's
    module tree on CPU — construction only, no product compute)"""
    import superpoint_transformer_b200 as S
    torch.manual_seed(0)
    net = S.SPT(mlp_norm=S.nn.GraphNorm, norm=S.nn.GraphNorm, **model_kwargs(S))
    net.apply(S.init_weights)
    head = S.nn.Classifier(DIM, NUM_CLASSES)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point())
          for k, v in net.state_dict().items()}
    hw = head.classifier.weight.detach().clone().requires_grad_(True)
    hb = head.classifier.bias.detach().clone().requires_grad_(True)
    return sd, hw, hb


def cpu_step(sd, hw, hb, nag, labels):
    from oracle import path as P
    for v in list(sd.values()) + [hw, hb]:
        if v.grad is not None:
            v.grad = None
    out = P.spt_forward(sd, nag, num_heads=HEADS, qk_dim=QK_DIM, nano=True, num_down=2, num_up=2,
                        use_diameter_parent=True, pool_reduce='max')
    loss = torch.nn.functional.cross_entropy(torch.nn.functional.linear(out, hw, hb), labels)
    loss.backward()
    return float(loss.detach())


def _scaled_levels(n1):
    return [n1, max(n1 // 5, 2), max(n1 // 25, 1)]


def _available_cores():
    """host cores this process may use: affinity mask capped by the cgroup CPU quota"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()[:2]
        if quota != 'max':
            n = min(n, max(1, int(float(quota) / float(period) + 0.999)))
    except Exception:  # noqa: BLE001
        pass
    return max(1, n)


def _pick_threads(sd, hw, hb):
    """The reference's torch CPU path does not scale to every core of a many-core
    host on graphs this small; give it its best case: probe 8,16,32,... threads on
    a 1k-superpoint NAG and keep the fastest.  Returns (threads, seconds per
    superpoint at that setting)."""
    avail = _available_cores()
    nag = _cpu_scene(_scaled_levels(1000), seed=1)
    labels = torch.randint(0, NUM_CLASSES, (1000,))
    cands = sorted({c for c in (8, 16, 32, 64, avail) if c <= avail} or {avail})
    best = None
    for c in cands:
        torch.set_num_threads(c)
        cpu_step(sd, hw, hb, nag, labels)
        t = time.time()
        cpu_step(sd, hw, hb, nag, labels)
        dt = time.time() - t
        if best is None or dt < best[1]:
            best = (c, dt)
        elif dt > 1.5 * best[1]:
            break
    torch.set_num_threads(best[0])
    return best[0], best[1] / 1000, avail


def cpu_reference_sample(target_seconds=20.0):
    sd, hw, hb = _cpu_state_dict()
    threads, per_sp, avail = _pick_threads(sd, hw, hb)
    n1 = int(min(max(target_seconds / max(per_sp, 1e-9), 1000), LEVELS[0]))
    n1 = max(1000, (n1 // 1000) * 1000)
    nag = _cpu_scene(_scaled_levels(n1), seed=1)
    labels = torch.randint(0, NUM_CLASSES, (n1,))
    t = time.time()
    cpu_step(sd, hw, hb, nag, labels)
    dt = time.time() - t
    return {"value": round(n1 / dt, 1), "unit": UNIT, "cores": threads, "kind": "port",
            "host_cores_available": avail,
            "sample": f"one fwd+bwd of the same model on a {n1}/{n1 // 5}/{n1 // 25}-superpoint "
                      f"NAG of the same law ({dt:.1f} s); oracle/path.py (reference glue "
                      f"restated, torch CPU leaves), fp32, {threads} threads (fastest of the "
                      f"probed thread counts on this {avail}-core host)"}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    sd, hw, hb = _cpu_state_dict()
    threads, per_sp, avail = _pick_threads(sd, hw, hb)
    budget = 180.0
    total_steps = args.steps + args.warmup
    n1 = int(budget / total_steps / max(per_sp, 1e-9))
    n1 = max(1000, min((n1 // 1000) * 1000, LEVELS[0]))
    nag = _cpu_scene(_scaled_levels(n1), seed=1)
    labels = torch.randint(0, NUM_CLASSES, (n1,))
    for _ in range(args.warmup):
        cpu_step(sd, hw, hb, nag, labels)
    t = time.time()
    for _ in range(args.steps):
        cpu_step(sd, hw, hb, nag, labels)
    dt = time.time() - t
    ms = dt / args.steps * 1e3
    value = n1 / (ms * 1e-3)
    sample = (f"each step = fwd+bwd on a {n1}/{n1 // 5}/{n1 // 25}-superpoint NAG of the cfg-2 "
              f"law (bounded sample of the 100k workload), oracle/path.py on {threads} threads "
              f"(fastest probed; host has {avail} cores)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": threads,
                         "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='eager launches, no CUDA graphs')
    ap.add_argument('--kernels-out', default=None, help='write the full per-kernel timing table (JSON)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'own' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_own(args)


if __name__ == '__main__':
    main()
