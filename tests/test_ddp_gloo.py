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


# ---- host plumbing of the package's own DDP (ddp-tutorial_b200/parallel.py), 2 processes, no GPU -------------
def _plumbing_worker(rank, world, port, q):
    import tempfile
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ddp_tutorial_b200 import parallel as P
    import ddp_tutorial_b200 as b2
    P._state.update(rank=rank, world=world, distributed=True, pg=True)
    out = {"rank": rank}
    # (1) handle exchange (cudaIpc handles / multicast descriptors travel as opaque bytes through gloo)
    got = P._exchange((bytes([rank]) * 64, f"buf{rank}"))
    out["handles"] = [(g[0][:1], g[1]) for g in got]
    # (2) POSIX file descriptors (cuMemExportToShareableHandle output) over Unix sockets, SCM_RIGHTS: all-to-all ...
    files = []
    for k in range(2):
        f = tempfile.TemporaryFile()
        f.write(f"rank{rank}-fd{k}".encode()); f.flush()
        files.append(f)
    fds = P._exchange_fds([f.fileno() for f in files])
    seen = []
    for r in range(world):
        row = []
        for fd in fds[r]:
            os.lseek(fd, 0, os.SEEK_SET)
            row.append(os.read(fd, 64).decode())
        seen.append(row)
    out["all_to_all"] = seen
    # ... and root-only (the multicast object's descriptor comes from rank 0)
    fds0 = P._exchange_fds([files[0].fileno()] if rank == 0 else [], root=0)
    os.lseek(fds0[0][0], 0, os.SEEK_SET)
    out["from_root"] = os.read(fds0[0][0], 64).decode()
    out["non_root_empty"] = all(len(fds0[r]) == 0 for r in range(1, world))
    # (3) the sampler shards of the two ranks partition the (padded) epoch
    s = b2.DistributedSampler(103, num_replicas=world, rank=rank, shuffle=True, seed=5)
    s.set_epoch(2)
    out["indices"] = list(s)
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_b2_ddp_host_plumbing_world2():
    """SURVEY 8e host side: handle exchange, SCM_RIGHTS descriptor passing (the NVLS set-up path) and the sampler's
    rank shards, with two real processes over gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_plumbing_worker, args=(r, 2, 29641, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda d: d["rank"])
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for d in res:
        assert d["handles"] == [(b"\x00", "buf0"), (b"\x01", "buf1")]
        assert d["all_to_all"] == [["rank0-fd0", "rank0-fd1"], ["rank1-fd0", "rank1-fd1"]]
        assert d["from_root"] == "rank0-fd0" and d["non_root_empty"]
    a, b = res[0]["indices"], res[1]["indices"]
    assert len(a) == len(b) == 52
    both = sorted(a + b)
    assert set(both) == set(range(103)) and len(both) == 104          # padded by one repeated index, as torch does
