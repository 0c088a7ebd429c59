# usage: bash tools/run_n8.sh W     (inside gpurun --gpus W): the driver-style bench line + a few exchange-shape variants
W=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
show() { python - <<PY
import json
try:
    d=json.load(open("$1")); s=d.get("secondary") or {}; p=d.get("ddp_parity") or {}
    print("$2", round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],4), {k:p.get(k) for k in ("ok","replicas_identical_after_steps","loss_err")}, "R50", s.get("ms_per_step"), s.get("value"))
except Exception as e: print("$2 parse", e)
PY
}
timeout 600 $TR --master-port 29555 bench.py --gpus $W --steps 20 --warmup 5 > gpurun_out/r02_bench_n$W.json 2> gpurun_out/bench_n$W.err; tail -2 gpurun_out/bench_n$W.err; show gpurun_out/r02_bench_n$W.json "N=$W default"
for th in 256 128; do
B2_COMM_THREADS=$th timeout 300 $TR --master-port 29577 bench.py --gpus $W --steps 20 --warmup 5 --no-secondary --no-parity --no-cpu-baseline > gpurun_out/r02_bench_n${W}_t$th.json 2> gpurun_out/sw.err; show gpurun_out/r02_bench_n${W}_t$th.json "N=$W nvls t$th"
done
B2_COMM_BLOCKS=48 timeout 300 $TR --master-port 29578 bench.py --gpus $W --steps 20 --warmup 5 --no-secondary --no-parity --no-cpu-baseline > gpurun_out/r02_bench_n${W}_cb48.json 2> gpurun_out/sw.err; show gpurun_out/r02_bench_n${W}_cb48.json "N=$W nvls cb48"
B2_NVLS=0 timeout 300 $TR --master-port 29579 bench.py --gpus $W --steps 20 --warmup 5 --no-secondary --no-parity --no-cpu-baseline > gpurun_out/r02_bench_n${W}_p2p.json 2> gpurun_out/sw.err; show gpurun_out/r02_bench_n${W}_p2p.json "N=$W p2p"
