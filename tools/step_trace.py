"""In-graph timeline of ONE training step (b2_trace_start / b2_trace_stop, csrc/common.cuh:trace_mark).

  python tools/step_trace.py [--model resnet18] [--batch 512] [--bucket-mb 32] [--out gpurun_out/trace.txt]
  python -m torch.distributed.run --nproc-per-node 2 ... tools/step_trace.py      (every rank writes <out>.rank<r>)

Each record is block 0 / thread 0 of a kernel passing griddepcontrol.wait ("begin") or finishing ("end", only the
kernels that carry an end mark: k_conv, k_wgrad, k_allreduce*, k_sgd).  Times are %globaltimer ns relative to the
first record of the traced step."""
import argparse, ctypes as C, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TU = {1: "comm.cu", 2: "conv_simt.cu", 3: "conv_umma.cu", 4: "ops_elementwise.cu"}


def kernel_of(tu, line, cache={}):
    """name of the __global__ function enclosing csrc/<tu>:<line>"""
    f = TU.get(tu)
    if f is None:
        return f"tu{tu}:{line}"
    if f not in cache:
        cache[f] = open(os.path.join(ROOT, "ddp-tutorial_b200", "csrc", f)).read().split("\n")
    src = cache[f]
    for i in range(min(line, len(src)) - 1, -1, -1):
        if "__global__" in src[i]:
            m = re.search(r"\b(k_\w+)\s*\(", " ".join(src[i:i + 3]))
            return m.group(1) if m else f"{f}:{line}"
    return f"{f}:{line}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet18")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--bucket-mb", type=float, default=32.0)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--out", default="gpurun_out/step_trace.txt")
    a = ap.parse_args()
    import torch
    import ddp_tutorial_b200 as b2
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    ns = argparse.Namespace()
    b2.init_distributed(ns)
    dev = ns.gpu
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    net = getattr(b2, a.model)(num_classes=10, precision="bf16").to(dev)
    model = b2.DistributedDataParallel(net, device_ids=[dev], overlap=not a.no_overlap, bucket_cap_mb=a.bucket_mb) if world > 1 else net
    crit = b2.CrossEntropyLoss()
    opt = b2.SGD(model.parameters(), lr=0.001, momentum=0.9)
    g = torch.Generator().manual_seed(1 + rank)
    x = torch.randn(a.batch, 3, 32, 32, generator=g).to(dev)
    y = torch.randint(0, 10, (a.batch,), generator=g).to(dev)

    def step():
        opt.zero_grad()
        loss = crit(model(x), y)
        loss.backward()
        opt.step()

    for _ in range(30):
        step()
    torch.cuda.synchronize()
    if world > 1:
        b2.barrier(); torch.cuda.synchronize()
    cap = 4096
    b2._lib.call("b2_trace_start", cap)
    for _ in range(a.steps + 1):          # the first traced step still overlaps the tail of nothing: keep the LAST one
        step()
    torch.cuda.synchronize()
    rec = (C.c_ulonglong * (2 * cap))()
    n = C.c_int(0)
    b2._lib.call("b2_trace_stop", rec, cap, C.byref(n))
    ev = sorted(((rec[2 * i + 1], rec[2 * i]) for i in range(n.value)))
    # last step = from the last k_stem_pack_weights / first forward kernel begin
    names = [(t, kernel_of(tag >> 24, (tag >> 4) & 0xFFFFF), tag & 15) for t, tag in ev]
    starts = [i for i, (t, k, kind) in enumerate(names) if kind == 0 and k in ("k_stem_pack_weights", "k_nchw_to_nhwc", "k_stem_im2col")]
    i0 = starts[-1] if starts else 0
    t0 = names[i0][0]
    out = a.out + (f".rank{rank}" if world > 1 else "")
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    open_begin = {}
    lines = []
    for t, k, kind in names[i0:]:
        if kind == 0:
            open_begin[k] = t
            lines.append(f"{(t - t0) / 1e3:9.2f} us  begin {k}")
        else:
            d = (t - open_begin.get(k, t)) / 1e3
            lines.append(f"{(t - t0) / 1e3:9.2f} us  end   {k}  (+{d:.2f} us)")
    with open(out, "w") as f:
        f.write(f"# {a.model} B={a.batch} world={world} bucket_mb={a.bucket_mb} overlap={not a.no_overlap} records={n.value}\n")
        f.write("\n".join(lines) + "\n")
    if rank == 0:
        print(f"wrote {out}: {len(lines)} records, step span {(names[-1][0] - t0) / 1e3:.1f} us")


if __name__ == "__main__":
    main()
