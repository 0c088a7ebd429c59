#!/usr/bin/env python
"""Benchmark of the DDP-Tutorial hot path on B200 (BASELINE.json metric):
images/sec, device-timed, max over ranks, ResNet-18 (torchvision topology, num_classes=10),
synthetic CIFAR-shaped input, per-GPU batch 512, bf16 storage / fp32 accumulate + fp32 master.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference ...                           # the reference CPU loop

One "step" = the reference's six statements (cifar_example_ddp.py:101-110): zero_grad,
forward, criterion, backward, optimizer.step (fused with the gradient all-reduce when N>1).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "resnet18"
BATCH = 512                      # per GPU (BASELINE configs[1], configs[2])
METRIC = "images/sec (device-timed, max over ranks) ResNet-18 CIFAR-shape"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------
# reference arm: the reference's CPU loop (cifar_example.py:66-87 semantics) via the oracle
# ----------------------------------------------------------------------------------------
def usable_cores() -> int:
    """Host threads this process can really use: affinity mask capped by the cgroup CPU quota
    (os.cpu_count() reports the whole machine and oversubscribes a container)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


def cpu_reference_loop(batch: int, steps: int, warmup: int):
    import torch
    from oracle import ref_step as O       # the CPU baseline IS the oracle (kind="port")
    torch.set_num_threads(min(usable_cores(), 64))
    sys.stderr.write(f"[bench] cpu reference: os.cpu_count()={os.cpu_count()} usable={usable_cores()} "
                     f"torch threads={torch.get_num_threads()}\n")
    model = O.make_model(MODEL, 0)
    crit = torch.nn.CrossEntropyLoss()
    opt = torch.optim.SGD(model.parameters(), lr=O.LR, momentum=O.MOMENTUM)
    x, y = O.synthetic_batch(O.data_generator(0), batch)
    t0 = None
    for s in range(warmup + steps):
        if s == warmup:
            t0 = time.perf_counter()
        opt.zero_grad()
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        loss.item()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, ms, threads = cpu_reference_loop(BATCH, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload naming as this repo's arm (run_b2); the CPU times a bounded sample of it: one batch of BATCH per step
        "config": {"workload": f"{MODEL}(num_classes=10) training step, per-GPU batch {BATCH}, 3x32x32 synthetic "
                               f"(BASELINE configs[{(1 if args.gpus == 1 else 2) if MODEL == 'resnet18' else 3}])",
                   "global_batch": args.gpus * BATCH, "parallelism": f"dp{args.gpus}", "precision": "fp32",
                   "where": "host CPU cores, torch CPU kernels (the reference's own code path: cifar_example.py:66-87)"},
        "cpu_baseline": {"value": ips, "unit": "images/sec", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of one batch of {BATCH} (rank 0's shard of the global batch) after {args.warmup} warm-up"},
        "e2e": {"value": ips, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------
TRAIN_FLOP = {"resnet18": 217.28e6, "resnet50": 495.77e6}      # SURVEY.md 8(d), dense count per image


def ddp_parity(b2, O, ns, model_name, batch, precision, dev, rank, world, overlap, bucket_mb):
    """Multi-GPU parity on the benchmarked model, run before timing (the driver's SCALE runs see it):
    rank-0 init wins (ranks are seeded differently, cifar_example_ddp.py:83 -> distributed.py:883-890),
    replicas stay bit-identical, rank 0's loss and the parameter update follow the oracle's world=N run."""
    import torch
    import torch.distributed as dist
    steps = 2
    torch.manual_seed(0 if rank == 0 else 1234 + rank)
    net = getattr(b2, model_name)(num_classes=10, precision=precision).to(dev)
    model = b2.DistributedDataParallel(net, device_ids=[dev], overlap=overlap, bucket_cap_mb=bucket_mb)
    crit = b2.CrossEntropyLoss()
    opt = b2.SGD(model.parameters(), lr=0.001, momentum=0.9)
    e = net.engine

    def replicas_identical():
        flat = e.flat_params[:e.n_params].cpu()
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([1 if torch.equal(flat, ref) else 0])
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        return bool(same.item())

    init_same = replicas_identical()
    p0 = e.flat_params[:e.n_params].cpu().clone()
    r = None
    if rank == 0:                                      # the oracle's world=N run on the host (torchrun pins OMP threads to 1)
        nt = torch.get_num_threads()
        torch.set_num_threads(usable_cores())
        r = O.train(model_name, seed=0, batch=batch, steps=steps, world=world)
        torch.set_num_threads(nt)
    dist.barrier()                                     # (gloo) no rank enters the exchange kernels while rank 0 is still on the host
    xg, yg = O.synthetic_batch(O.data_generator(0), batch * world)
    x, y = xg[rank * batch:(rank + 1) * batch].to(dev), yg[rank * batch:(rank + 1) * batch].to(dev)
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    torch.cuda.synchronize()
    b2.check(net)
    same = replicas_identical()
    out = {"replicas_identical_after_init": init_same, "replicas_identical_after_steps": same, "steps": steps}
    ok = init_same and same
    if rank == 0:
        names = {s.name: s for s in e.specs}
        upd, ref_upd = [], []
        init = dict(r["init"])
        for n, t in r["final"]:
            s_ = names[n]
            mine = torch.as_strided(e.flat_params.cpu(), s_.shape, s_.stride, s_.offset) - torch.as_strided(p0, s_.shape, s_.stride, s_.offset)
            upd.append(mine.flatten().double()); ref_upd.append((t - init[n]).flatten().double())
        upd, ref_upd = torch.cat(upd), torch.cat(ref_upd)
        out["loss"] = losses
        out["loss_err"] = [abs(a - b) for a, b in zip(losses, r["loss"])]
        out["update_rel_err"] = float((upd - ref_upd).norm() / ref_upd.norm())
        out["update_cos"] = float((upd @ ref_upd) / (upd.norm() * ref_upd.norm()))
        tol = 1e-3 if precision == "fp32" else (2e-2 if model_name == "resnet18" else 1e-1)
        out["loss_tol"] = tol
        ok = ok and max(out["loss_err"]) < tol and out["update_cos"] > (0.999 if precision == "fp32" else 0.85)
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    out["ok"] = bool(flag.item())
    del opt, model, net
    return out


def measure(b2, args, ns, model_name, batch, dev, rank, full):
    """Build `model_name` at per-GPU `batch`, time K steps resident + K steps end-to-end; `full` adds the
    roofline leg and the clock sampler."""
    import torch
    torch.manual_seed(0 if rank == 0 else 1234 + rank)        # rank 0's init wins under DDP
    net = getattr(b2, model_name)(num_classes=10, precision=args.precision).to(dev)
    model = (b2.DistributedDataParallel(net, device_ids=[dev], overlap=not args.no_overlap, bucket_cap_mb=args.bucket_mb)
             if ns.distributed else net)
    criterion = b2.CrossEntropyLoss()
    optimizer = b2.SGD(model.parameters(), lr=0.001, momentum=0.9)

    gen = torch.Generator().manual_seed(1 + rank)
    x_host = torch.randn(batch, 3, 32, 32, generator=gen).pin_memory()
    y_host = torch.randint(0, 10, (batch,), generator=gen).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step_resident():
        optimizer.zero_grad()
        loss = criterion(model(x_dev), y_dev)
        loss.backward()
        optimizer.step()
        return loss

    # e2e: every step copies its batch from pinned host memory (through the repo's DevicePrefetcher:
    # the copy of batch i+1 runs on a copy stream under step i) and reads the loss back to the host
    def host_batches():
        while True:
            yield x_host, y_host
    e2e_iter = iter(b2.DevicePrefetcher(host_batches(), dev, depth=2))

    def step_e2e():
        xi, yi = next(e2e_iter)
        optimizer.zero_grad()
        loss = criterion(model(xi), yi)
        loss.backward()
        optimizer.step()
        return loss.item()                     # device->host read of the step's result

    def sync_all():
        torch.cuda.synchronize()
        if ns.distributed:
            b2.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if ns.distributed:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)       # gloo, host side, outside the timed region
            ms = float(t[0])
        sync_all()
        return ms, last

    sampler = ClockSampler(dev) if (full and rank == 0 and not args.no_clocks) else None   # nvidia-smi needs ~1 s to start
    untimed = max(args.warmup, 3)
    for _ in range(untimed):
        step_resident()
    if full and not args.no_clocks:
        for _ in range(800):                                 # ~1 s under load on EVERY rank (same collective count)
            step_resident()                                  # so that nvidia-smi samples flow before the timed region
        untimed += 800
    launches0 = b2.lib().b2_launch_count()
    ms, last = timed(step_resident, args.steps)
    launches = b2.lib().b2_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    final_loss = last.item()
    b2.check(net)
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    res = {"ms": ms, "ms_e2e": ms_e2e, "launches": int(launches), "clocks": clocks, "final_loss": final_loss,
           "untimed_steps": untimed, "h2d": x_host.numel() * 4 + y_host.numel() * 8, "roofline": None}

    # ---- roofline leg: events around every op of the dominant kernel class (separate pass)
    if full:                                                 # every rank runs the same steps (collectives must match)
        e = net.engine
        m = e._model(batch)
        b2._lib.call("b2_model_profile", m.h, 1)
        nprof = 3
        for _ in range(nprof):
            step_resident()
        arr_d, arr_i = (C.c_double * 8), (C.c_int * 8)
        pms, pfl, pby, pln = arr_d(), arr_d(), arr_d(), arr_i()
        b2._lib.call("b2_model_profile_collect", m.h, pms, pfl, pby, pln)
        b2._lib.call("b2_model_profile", m.h, 0)
        cats = ["conv_fprop", "conv_dgrad", "conv_wgrad", "bn_fwd", "bn_bwd", "pool", "other"]
        per = {c: {"ms_per_step": pms[i] / nprof, "gflop_per_step": pfl[i] / nprof / 1e9,
                   "mb_per_step": pby[i] / nprof / 1e6, "ops_per_step": pln[i] // nprof} for i, c in enumerate(cats)}
        hbm, tf_burst, tf_sus, src = _peaks()
        conv_ms_eager = sum(per[c]["ms_per_step"] for c in cats[:3])
        total_ms = sum(v["ms_per_step"] for v in per.values())
        # dominant kernel class in isolation: every conv launch of the step, 20 back-to-back launches each
        kms, kfl, kln = (C.c_double * 3)(), (C.c_double * 3)(), (C.c_int * 3)()
        b2._lib.call("b2_model_time_conv_kernels", m.h, 20, kms, kfl, kln, None)
        conv_ms = sum(kms); conv_fl = sum(kfl) / 1e9
        achieved = conv_fl / conv_ms if conv_ms > 0 else 0.0          # GFLOP/ms == TFLOP/s
        kernel_avg = {n: {"ms_per_step": kms[i], "gflop_per_step": kfl[i] / 1e9, "launches": kln[i],
                          "tflops": (kfl[i] / 1e9 / kms[i]) if kms[i] > 0 else 0.0} for i, n in enumerate(("fprop", "dgrad", "wgrad"))}
        traffic, traffic_note = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                tj = json.load(f)
            traffic = tj["conv_dram_bytes_per_launch"]
            traffic_note = tj.get("note")
        except Exception:
            pass
        dense_gflop = TRAIN_FLOP[model_name] * batch / 1e9
        res["roofline"] = {
            "bound": "tensor", "kernel": "conv implicit-GEMM kernels (fprop+dgrad incl. the fused BatchNorm work, wgrad) of one step",
            "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst,
            "peak_source": f"{src} MEASURED_PEAKS bf16_tflops (burst: the conv launches are timed alone)",
            "frac_of_sustained_peak": achieved / tf_sus, "traffic": traffic, "traffic_note": traffic_note,
            "flops_counted": f"executed (padding-only taps skipped); SURVEY 8d dense count would be {dense_gflop:.2f} GFLOP/step",
            "whole_step_tflops_dense_count": dense_gflop / (ms / args.steps),
            "how": "each conv launch of the step timed alone: 20 back-to-back launches between two CUDA events "
                   "(b2_model_time_conv_kernels; BatchNorm sum buffers and barrier counters are reset before every launch, so the "
                   "grid barrier of the fused conv+BN kernels is paid in full), summed over the step's launches; per_class = eager "
                   "per-op event pass (includes launch gaps)",
            "kernel_avg": kernel_avg,
            "share_of_step": conv_ms_eager / total_ms if total_ms else None, "per_class": per}
    del optimizer, model
    return res, net


def run_b2(args):
    import torch
    import ddp_tutorial_b200 as b2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    ns = argparse.Namespace()
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # keep stdout to the one JSON line
        b2.init_distributed(ns)
    dev = ns.gpu
    torch.cuda.set_device(dev)

    parity = None
    if ns.distributed and not args.no_parity:
        from oracle import ref_step as O               # checker only (the timed region never touches it)
        parity = ddp_parity(b2, O, ns, MODEL, BATCH, args.precision, dev, rank, world, not args.no_overlap, args.bucket_mb)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "ddp parity check failed", "ddp_parity": parity}), flush=True)
            b2.destroy()
            raise SystemExit(3)

    main_res, net = measure(b2, args, ns, MODEL, BATCH, dev, rank, full=True)
    del net
    secondary = None
    if MODEL == "resnet18" and not args.no_secondary:
        # BASELINE configs[3]: ResNet-50, per-GPU batch 256, same N, shorter run (no roofline leg)
        sec, net2 = measure(b2, args, ns, "resnet50", 256, dev, rank, full=False)
        del net2
        secondary = {"metric": METRIC.replace("ResNet-18", "ResNet-50"), "value": args.gpus * 256 * args.steps / (sec["ms"] / 1e3),
                     "unit": "images/sec", "ms_per_step": sec["ms"] / args.steps,
                     "e2e": {"value": args.gpus * 256 * args.steps / (sec["ms_e2e"] / 1e3), "unit": "images/sec",
                             "ms_per_step": sec["ms_e2e"] / args.steps, "h2d_bytes_per_step": sec["h2d"], "d2h_bytes_per_step": 4},
                     "gpu_launches": sec["launches"], "final_loss": sec["final_loss"],
                     "config": {"workload": "resnet50(num_classes=10) training step, per-GPU batch 256, 3x32x32 synthetic (BASELINE configs[3])",
                                "global_batch": args.gpus * 256, "parallelism": f"dp{args.gpus}"},
                     "whole_step_tflops_dense_count": TRAIN_FLOP["resnet50"] * 256 / 1e9 / (sec["ms"] / args.steps)}

    if rank == 0:
        n = args.gpus
        ms, ms_e2e = main_res["ms"], main_res["ms_e2e"]
        ips = n * BATCH * args.steps / (ms / 1e3)
        ips_e2e = n * BATCH * args.steps / (ms_e2e / 1e3)
        cpu = None
        if n == 1 and not args.no_cpu_baseline:
            cips, cms, threads = cpu_reference_loop(128, 12, 3)      # BASELINE configs[0]: batch 128
            cpu = {"value": cips, "unit": "images/sec", "cores": threads, "kind": "port",
                   "sample": "cifar_example.py loop via oracle/ref_step.py, resnet18 batch 128, 12 steps after 3 warm-up"}
        line = {
            "metric": METRIC, "value": ips, "unit": "images/sec", "n_gpus": n, "steps": args.steps,
            "warmup": main_res["untimed_steps"], "warmup_requested": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"{MODEL}(num_classes=10) training step, per-GPU batch {BATCH}, 3x32x32 synthetic "
                                   f"(BASELINE configs[{(1 if n == 1 else 2) if MODEL == 'resnet18' else 3}])",
                       "global_batch": n * BATCH, "parallelism": f"dp{n}",
                       "precision": "bf16 activations+weight copies, fp32 accumulate, fp32 master+momentum"
                       if args.precision == "bf16" else "fp32",
                       "l2": "no explicit flush: one step streams >600 MB of activations/gradients/params, "
                             "several times the 126 MB L2"},
            "final_loss": main_res["final_loss"], "final_loss_after_steps": main_res["untimed_steps"] + args.steps,
            "e2e": {"value": ips_e2e, "unit": "images/sec", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": main_res["h2d"], "d2h_bytes_per_step": 4},
            "gpu_launches": main_res["launches"],
            "roofline": main_res["roofline"], "cpu_baseline": cpu, "clocks": main_res["clocks"],
            "ddp_parity": parity, "secondary": secondary,
            "env_toggles": {k: v for k, v in sorted(os.environ.items()) if k.startswith("B2_")},
        }
        print(json.dumps(line), flush=True)
    if ns.distributed:
        b2.destroy()


def main():
    global MODEL, BATCH, METRIC
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b2", choices=["b2", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--model", default="resnet18", choices=["resnet18", "resnet50"],
                    help="resnet18 = the headline config; resnet50 (per-GPU batch 256) = BASELINE configs[3]")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 512 for resnet18, 256 for resnet50)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="DDP: one fused exchange at optimizer.step() instead of bucketed overlap")
    ap.add_argument("--bucket-mb", type=float, default=32.0)
    ap.add_argument("--no-clocks", action="store_true", help="skip the nvidia-smi sampler (profiler runs)")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the DDP parity phase that runs before timing")
    ap.add_argument("--no-secondary", action="store_true", help="skip the ResNet-50 (BASELINE configs[3]) secondary measurement")
    args = ap.parse_args()
    MODEL = args.model
    BATCH = args.batch or (512 if MODEL == "resnet18" else 256)
    if MODEL != "resnet18":
        METRIC = METRIC.replace("ResNet-18", "ResNet-50")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b2(args)


if __name__ == "__main__":
    main()
