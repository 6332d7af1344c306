"""CPU, gloo, world_size 2: scene sharding + the single flat gradient all-reduce
give every rank the gradient of the full-batch mean loss (the N>1 path of bench.py,
minus the CUDA kernels which cannot run here)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from superpoint_transformer_b200.distributed import (FlatGradients, shard_indices,
                                                         init_process_group_from_env)
    r, w, _ = init_process_group_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.LeakyReLU(), torch.nn.Linear(5, 3))
    flat = FlatGradients(model.parameters())
    scenes = [torch.randn(10 + i, 6, generator=torch.Generator().manual_seed(i)) for i in range(6)]
    mine = shard_indices(len(scenes), rank, world)
    flat.release()
    for i in mine:  # local mean over this rank's scenes
        (model(scenes[i]).pow(2).mean() / len(mine)).backward()
    flat.collect()
    flat.all_reduce()
    out[rank] = flat.flat.clone()
    dist.barrier()
    dist.destroy_process_group()


def test_flat_all_reduce_matches_full_batch_gradient():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.LeakyReLU(), torch.nn.Linear(5, 3))
    scenes = [torch.randn(10 + i, 6, generator=torch.Generator().manual_seed(i)) for i in range(6)]
    loss = sum(model(s).pow(2).mean() for s in scenes) / len(scenes)
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    torch.testing.assert_close(out[0], ref, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(out[1], ref, atol=1e-6, rtol=1e-5)
