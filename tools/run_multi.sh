# usage: bash tools/run_multi.sh W      (inside gpurun --gpus W)
W=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1"
if [ "$W" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --timeout=400 --timeout-method=thread > gpurun_out/t_multi_w$W.log 2>&1; tail -5 gpurun_out/t_multi_w$W.log
fi
timeout 300 $TR --master-port 29533 tools/ddp_check.py resnet18 4 1 bf16 > gpurun_out/ddp_check_r18_w$W.log 2>&1; tail -4 gpurun_out/ddp_check_r18_w$W.log
timeout 400 $TR --master-port 29544 tools/allreduce_sweep.py --max-mb 256 --out gpurun_out/r02_allreduce_w$W.json > gpurun_out/sweep_w$W.log 2>&1; tail -11 gpurun_out/sweep_w$W.log
show() { python - <<PY
import json
try:
    d=json.load(open("$1")); s=d.get("secondary") or {}
    print("$2", round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],4), d.get("ddp_parity",{}) and {k:d["ddp_parity"][k] for k in ("ok","replicas_identical_after_steps","loss_err")}, "R50", s.get("ms_per_step"), s.get("value"))
except Exception as e: print("$2 parse", e)
PY
}
B2_NVLS=1 timeout 600 $TR --master-port 29555 bench.py --gpus $W --steps 20 --warmup 5 > gpurun_out/r02_bench_n$W.json 2> gpurun_out/bench_n$W.err; tail -3 gpurun_out/bench_n$W.err; show gpurun_out/r02_bench_n$W.json "N=$W nvls b32"
B2_NVLS=0 timeout 600 $TR --master-port 29566 bench.py --gpus $W --steps 20 --warmup 5 > gpurun_out/r02_bench_n${W}_p2p.json 2> gpurun_out/bench_n${W}_p2p.err; show gpurun_out/r02_bench_n${W}_p2p.json "N=$W p2p b32"
for nv in 1 0; do
B2_NVLS=$nv timeout 300 $TR --master-port 29577 bench.py --gpus $W --steps 20 --warmup 5 --no-secondary --no-parity --no-cpu-baseline --bucket-mb 16 > gpurun_out/r02_bench_n${W}_nvls${nv}_b16.json 2> gpurun_out/sw.err; show gpurun_out/r02_bench_n${W}_nvls${nv}_b16.json "N=$W nvls$nv b16"
B2_NVLS=$nv timeout 300 $TR --master-port 29588 tools/step_trace.py --bucket-mb 32 --out gpurun_out/r02_trace_n${W}_nvls$nv.txt 2>&1 | grep -v Warn | tail -1
done
