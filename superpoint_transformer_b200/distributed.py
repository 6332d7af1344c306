"""Scene-shard data parallelism: one process per GPU, scenes (independent NAGs)
sharded across ranks, ONE flat NCCL all-reduce of the gradients per step.

The reference only ever runs DDP through Lightning (configs/trainer/ddp.yaml:8-13);
no tensor crosses GPUs inside the hot path because batch items are disjoint
graphs.  The model is tiny (0.2-0.8 M parameters), so the
all-reduce is latency-bound: all gradients live in one pre-allocated flat buffer
(parameters' .grad are views into it) and a single collective is issued.
"""
import torch
import torch.distributed as dist

__all__ = ['shard_indices', 'FlatGradients', 'init_process_group_from_env']


def shard_indices(num_items, rank, world_size, weights=None):
    """Items -> ranks.  Without weights: round-robin (item i -> rank i % W), the
    reference DistributedSampler layout.  With weights (e.g. edge counts): greedy
    longest-processing-time balancing, deterministic on every rank."""
    if weights is None:
        return list(range(rank, num_items, world_size))
    order = sorted(range(num_items), key=lambda i: (-float(weights[i]), i))
    load = [0.0] * world_size
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += float(weights[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


class FlatGradients:
    """Owns one contiguous gradient buffer for `params`; `all_reduce()` averages it
    across the process group with a single collective."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        p0 = self.params[0]
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero_(self):
        self.flat.zero_()

    def release(self):
        """Detach the parameters from the flat buffer before a backward pass: autograd
        then *assigns* fresh gradients instead of launching one accumulate-add kernel per
        parameter (233 for the cfg-2 model); `collect()` packs them afterwards."""
        for p in self.params:
            p.grad = None
        if self.flat.is_cuda:
            from . import ops
            ops.zero_pool.begin_step(self.flat.device)

    def collect(self, accumulate=False):
        """Pack the freshly produced gradients into the flat buffer with one multi-tensor
        copy and re-point `.grad` at the flat views (missing gradients count as zero).
        `accumulate=True` adds them instead (gradient accumulation over the micro-batches of
        one optimizer step)."""
        if self.flat.is_cuda:
            from . import ops
            ops.zero_pool.end_step()
        if not hasattr(self, '_views'):
            self._views, off = [], 0
            for p in self.params:
                n = p.numel()
                self._views.append(self.flat[off:off + n].view_as(p))
                off += n
        have = [(v, p.grad) for v, p in zip(self._views, self.params) if p.grad is not None]
        if accumulate:
            if have:
                torch._foreach_add_([v for v, _ in have], [g for _, g in have])
        else:
            if len(have) != len(self.params):
                self.flat.zero_()
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, p in zip(self._views, self.params):
            p.grad = v

    def rebind(self):
        """re-point .grad at the flat buffer (after optimizer.zero_grad(set_to_none=True))"""
        off = 0
        for p in self.params:
            n = p.numel()
            view = self.flat[off:off + n].view_as(p)
            if p.grad is not view:
                if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                    view.copy_(p.grad)
                p.grad = view
            off += n

    def all_reduce(self, group=None, async_op=False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        self.flat.div_(dist.get_world_size(group))
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def init_process_group_from_env(backend=None):
    """RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT / LOCAL_RANK as set by torchrun."""
    import os
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local
