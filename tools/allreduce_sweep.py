"""BASELINE configs[4]: all-reduce bus-bandwidth sweep, 1 KB .. 256 MB fp32, this repo's peer-memory
kernel (b2_allreduce, in place, sum) vs torch.distributed NCCL on the same GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 \
        --master-port 29544 tools/allreduce_sweep.py [--out gpurun_out/allreduce_W.json]

Timing: CUDA events around `iters` back-to-back calls after `warm` warm-up calls, max over ranks;
algbw = S/t, busbw = algbw * 2(W-1)/W (nccl-tests convention).  Correctness of every size is
checked against the NCCL result first.
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import ddp_tutorial_b200 as b2
from ddp_tutorial_b200 import parallel as P
from ddp_tutorial_b200._lib import call, lib

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--max-mb", type=float, default=256.0)
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--warm", type=int, default=20)
args = ap.parse_args()

ns = argparse.Namespace()
b2.init_distributed(ns)                      # gloo rendezvous + b2_comm_init
W, rank, dev = ns.world_size, int(os.environ["RANK"]), ns.gpu
nccl = dist.new_group(backend="nccl")
device = torch.device("cuda", dev)

n_max = int(args.max_mb * (1 << 20) / 4)
flags = P._CommBuffer(lib().b2_comm_flag_bytes() // 4, device)
nvls, gran = P._nvls_available(device, None)
mc = None
if nvls:            # cuMemCreate memory bound to an NVSwitch multicast object: in-switch reduction available
    grads = P._VmmBuffer(n_max, device, gran)
    ptrs, mc = P._share_vmm([grads], device)
    g = ptrs[0]
    fh = P._exchange(flags.handle())
    f = []
    for q in range(W):
        if q == rank:
            f.append(flags.ptr)
        else:
            out = C.c_void_p(); call("b2_comm_import", fh[q], C.byref(out)); f.append(out.value)
else:
    grads = P._CommBuffer(n_max, device)
    handles = P._exchange((grads.handle(), flags.handle()))
    g, f = [], []
    for q in range(W):
        if q == rank:
            g.append(grads.ptr); f.append(flags.ptr)
        else:
            ptrs = []
            for h in handles[q]:
                out = C.c_void_p(); call("b2_comm_import", h, C.byref(out)); ptrs.append(out.value)
            g.append(ptrs[0]); f.append(ptrs[1])
arr = lambda xs: (C.c_void_p * W)(*xs)
call("b2_comm_register", arr(g), None, arr(f), n_max)
if mc is not None:
    call("b2_comm_register_multicast", C.c_void_p(mc[0]), None)
torch.cuda.synchronize(); P._exchange("ready")

sizes = []
s = 1024
while s <= args.max_mb * (1 << 20):
    sizes.append(int(s)); s *= 4
if sizes[-1] != int(args.max_mb * (1 << 20)):
    sizes.append(int(args.max_mb * (1 << 20)))
stream = None                                 # legacy default stream, same as torch's current stream


def timed(fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier(group=nccl); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=nccl)
    return float(t) * 1e-3


rows = []
for S in sizes:
    n = S // 4
    iters = args.iters if S <= (16 << 20) else max(20, args.iters // 4)
    x = torch.randn(n, device=device, generator=torch.Generator(device=device).manual_seed(rank + 1))
    ref = x.clone(); dist.all_reduce(ref, group=nccl)
    grads.tensor[:n].copy_(x); torch.cuda.synchronize(); dist.barrier(group=nccl)
    errs, times = {}, {}
    for mode in (["nvls", "p2p"] if nvls else ["p2p"]):
        lib().b2_comm_set_nvls(1 if mode == "nvls" else 0)
        grads.tensor[:n].copy_(x); torch.cuda.synchronize(); dist.barrier(group=nccl)
        call("b2_allreduce", n, 1.0, 0, stream); torch.cuda.synchronize()
        errs[mode] = float((grads.tensor[:n] - ref).abs().max() / (ref.abs().max() + 1e-30))
        times[mode] = timed(lambda: call("b2_allreduce", n, 1.0, 0, stream), iters, args.warm)
    y = x.clone()
    t_nccl = timed(lambda: dist.all_reduce(y, group=nccl), iters, args.warm)
    P.check()
    k = 2 * (W - 1) / W
    t_b2 = min(times.values())
    row = {"bytes": S, "b2_us": t_b2 * 1e6, "nccl_us": t_nccl * 1e6, "b2_busbw_GBs": S / t_b2 * k / 1e9,
           "nccl_busbw_GBs": S / t_nccl * k / 1e9, "max_rel_err_vs_nccl": max(errs.values())}
    for mode in times:
        row[f"b2_{mode}_us"] = times[mode] * 1e6
        row[f"b2_{mode}_busbw_GBs"] = S / times[mode] * k / 1e9
    rows.append(row)
    if rank == 0:
        r = rows[-1]
        per = " ".join(f"{m} {times[m] * 1e6:8.1f} us {S / times[m] * k / 1e9:6.1f} GB/s" for m in times)
        print(f"{S:>11d} B  {per} | nccl {r['nccl_us']:9.1f} us {r['nccl_busbw_GBs']:7.1f} GB/s | err {r['max_rel_err_vs_nccl']:.1e}", flush=True)
if rank == 0 and args.out:
    with open(args.out, "w") as fh:
        json.dump({"world": W, "dtype": "f32", "nvls_available": bool(nvls), "rows": rows, "nccl": ".".join(map(str, torch.cuda.nccl.version()))}, fh, indent=1)
dist.barrier(group=nccl)
b2.destroy()
