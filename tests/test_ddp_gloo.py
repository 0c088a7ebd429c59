"""CPU, world_size 2 over gloo: the multi-rank semantics the GPU path must reproduce --
real torch DistributedDataParallel around the oracle's Net (rank-0 init wins, gradients
averaged, replicas identical) against the oracle's single-process world=2 restatement."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import ref_step as O


def _worker(rank, world, port, steps, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = O.make_model("net", 0 if rank == 0 else 99 + rank)       # ranks seeded differently
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    opt = torch.optim.SGD(ddp.parameters(), lr=O.LR, momentum=O.MOMENTUM)
    x, y = O.synthetic_batch(O.data_generator(0), 4 * world)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]   # shard rows [r*B,(r+1)*B)
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(ddp(xs), ys)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    q.put((rank, losses, [t.numpy().copy() for _, t in O.named_params(model)]))
    dist.barrier()
    dist.destroy_process_group()


def test_real_ddp_world2_equals_restatement():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, 29631, 6, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    ref = O.train("net", seed=0, batch=4, steps=6, world=2)
    for a, b in zip(res[0][2], res[1][2]):
        assert np.array_equal(a, b)                       # replicas identical
    np.testing.assert_allclose(res[0][1], ref["loss"], atol=2e-6)
    for a, (n, t) in zip(res[0][2], ref["final"]):
        np.testing.assert_allclose(a, t.numpy(), rtol=1e-5, atol=2e-7, err_msg=n)
