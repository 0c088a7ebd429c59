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
for cfg in "48 512" "32 512" "24 512" "48 256" "64 512"; do
set -- $cfg
B2_COMM_BLOCKS=$1 B2_COMM_THREADS=$2 timeout 300 $TR --master-port 29577 bench.py --gpus $W --steps 20 --warmup 5 --no-secondary --no-parity --no-cpu-baseline > gpurun_out/r02_bench_n${W}_cb$1_t$2.json 2> gpurun_out/sw.err; show gpurun_out/r02_bench_n${W}_cb$1_t$2.json "N=$W nvls cb$1 t$2"
done
B2_NVLS=0 B2_COMM_BLOCKS=48 timeout 300 $TR --master-port 29579 bench.py --gpus $W --steps 20 --warmup 5 --no-secondary --no-parity --no-cpu-baseline > gpurun_out/r02_bench_n${W}_p2p_cb48.json 2> gpurun_out/sw.err; show gpurun_out/r02_bench_n${W}_p2p_cb48.json "N=$W p2p cb48"
