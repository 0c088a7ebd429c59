"""Generate the committed golden vectors from the REAL reference module.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

It imports /root/reference/cifar_example_ddp.py and cifar_example.py with the
two absent, hot-path-irrelevant imports (matplotlib, torchmetrics) stubbed,
then drives the reference's own objects -- ``Net`` (cifar_example_ddp.py:23-40),
``nn.CrossEntropyLoss()`` (:85), ``optim.SGD(lr=0.001, momentum=0.9)`` (:86),
``DistributedDataParallel`` (:83, gloo backend, CPU processes) -- through the
reference step (:101-110) on seeded synthetic input and writes small fixtures:

  net_b4_w1.npz     Net, B=4 (as written), fixed batch, 50 steps: x, y, init
                    params, loss curve, step-0 grads, final params
  net_b16_fresh.npz Net, B=16, fresh batch every step, 50 steps: losses, final params
  net_ddp_w2.npz    Net under real DDP, W=2 x B=4, 10 steps: rank losses, final params
  resnet18_b32.json torchvision resnet18(num_classes=10) in the cifar_example.py
                    loop, B=32 fixed batch, 20 steps: loss curve, per-tensor
                    fingerprints of init / step-0 grads / final params
  resnet18_b512.json same at BASELINE's B=512, 50 steps (loss curve + fingerprints)
  resnet50_b32.json  resnet50, B=32, 5 steps
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def load_reference():
    for n in ("matplotlib", "matplotlib.pyplot", "torchmetrics"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, "/root/reference")
    import cifar_example
    import cifar_example_ddp
    return cifar_example, cifar_example_ddp


def synth(gen, b):
    x = torch.randn(b, 3, 32, 32, generator=gen)
    y = torch.randint(0, 10, (b,), generator=gen)
    return x, y


def run_reference_loop(model, batch, steps, seed, fixed):
    """cifar_example.py:63-83 with synthetic data."""
    criterion = nn.CrossEntropyLoss()
    optimizer = optim.SGD(model.parameters(), lr=0.001, momentum=0.9)
    gen = torch.Generator().manual_seed(seed + 1)
    x0, y0 = synth(gen, batch)
    losses, grads0 = [], None
    for step in range(steps):
        inputs, labels = (x0, y0) if (fixed or step == 0) else synth(gen, batch)
        optimizer.zero_grad()
        outputs = model(inputs)
        loss = criterion(outputs, labels)
        loss.backward()
        if step == 0:
            grads0 = [p.grad.clone() for p in model.parameters()]
        optimizer.step()
        losses.append(loss.item())
    return x0, y0, losses, grads0


def fingerprint(t):
    t = t.double()
    return [float(t.sum()), float(t.abs().sum()), float((t * t).sum())]


def ddp_worker(rank, world, seed, batch, steps, port, q):
    import torch.distributed as dist
    _, R = load_reference()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(seed + 100 * rank)      # ranks are never seeded alike; rank 0 wins (distributed.py:883-890)
    if rank == 0:
        torch.manual_seed(seed)
    net = torch.nn.parallel.DistributedDataParallel(R.Net())
    criterion = nn.CrossEntropyLoss()
    optimizer = optim.SGD(net.parameters(), lr=0.001, momentum=0.9)
    gen = torch.Generator().manual_seed(seed + 1)
    x, y = synth(gen, batch * world)
    xs, ys = x[rank * batch:(rank + 1) * batch], y[rank * batch:(rank + 1) * batch]
    losses = []
    for _ in range(steps):
        optimizer.zero_grad()
        loss = criterion(net(xs), ys)
        loss.backward()
        optimizer.step()
        losses.append(loss.item())
    q.put((rank, losses, [p.detach().numpy().copy() for p in net.module.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def main():
    C, R = load_reference()
    torch.set_num_threads(8)
    # ---- Net, B=4 as written, fixed batch
    torch.manual_seed(0)
    net = R.Net()
    names = [n for n, _ in net.named_parameters()]
    init = [p.detach().clone() for p in net.parameters()]
    x, y, losses, g0 = run_reference_loop(net, 4, 50, 0, fixed=True)
    d = {"x": x.numpy(), "y": y.numpy(), "loss": np.array(losses, dtype=np.float64)}
    for n, a, b, c in zip(names, init, g0, net.parameters()):
        d["init." + n], d["grad0." + n], d["final." + n] = a.numpy(), b.numpy(), c.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "net_b4_w1.npz"), **d)
    # ---- Net, B=16, fresh batches (cifar_example.Net is the same class)
    torch.manual_seed(0)
    net = C.Net()
    _, _, losses, _ = run_reference_loop(net, 16, 50, 0, fixed=False)
    d = {"loss": np.array(losses, dtype=np.float64)}
    for n, c in zip(names, net.parameters()):
        d["final." + n] = c.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "net_b16_fresh.npz"), **d)
    # ---- real DDP over gloo, W=2
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=ddp_worker, args=(r, 2, 0, 4, 10, 29611, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get() for _ in procs], key=lambda t: t[0])
    [p.join() for p in procs]
    d = {"loss_rank0": np.array(res[0][1]), "loss_rank1": np.array(res[1][1])}
    for n, a, b in zip(names, res[0][2], res[1][2]):
        assert np.array_equal(a, b), "DDP replicas diverged"
        d["final." + n] = a
    np.savez_compressed(os.path.join(HERE, "net_ddp_w2.npz"), **d)
    # ---- ResNets in the cifar_example.py loop
    import torchvision
    for name, b, steps in (("resnet18", 32, 20), ("resnet18", 512, 50), ("resnet50", 32, 5)):
        torch.manual_seed(0)
        m = getattr(torchvision.models, name)(num_classes=10)
        pn = [n for n, _ in m.named_parameters()]
        init_fp = {n: fingerprint(p.detach()) for n, p in m.named_parameters()}
        _, _, losses, g0 = run_reference_loop(m, b, steps, 0, fixed=True)
        out = {"model": name, "batch": b, "steps": steps, "seed": 0, "fixed_batch": True,
               "torch": torch.__version__, "torchvision": torchvision.__version__,
               "threads": torch.get_num_threads(), "loss": losses, "init": init_fp,
               "grad0": {n: fingerprint(g) for n, g in zip(pn, g0)},
               "final": {n: fingerprint(p.detach()) for n, p in m.named_parameters()}}
        with open(os.path.join(HERE, f"{name}_b{b}.json"), "w") as f:
            json.dump(out, f)
        print(name, b, "loss[0]=%.6f loss[-1]=%.6f" % (losses[0], losses[-1]))


if __name__ == "__main__":
    main()
