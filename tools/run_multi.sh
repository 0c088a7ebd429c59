# usage: bash tools/run_multi.sh W      (inside gpurun --gpus W)
W=$1
mkdir -p gpurun_out
if [ "$W" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --timeout=400 --timeout-method=thread > gpurun_out/t_multi_w$W.log 2>&1; tail -25 gpurun_out/t_multi_w$W.log
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py resnet18 4 1 bf16 > gpurun_out/ddp_check_r18_w$W.log 2>&1; tail -12 gpurun_out/ddp_check_r18_w$W.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29544 tools/allreduce_sweep.py --max-mb 256 --out gpurun_out/r02_allreduce_w$W.json > gpurun_out/sweep_w$W.log 2>&1; tail -14 gpurun_out/sweep_w$W.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $W --steps 20 --warmup 5 > gpurun_out/r02_bench_n$W.json 2> gpurun_out/bench_n$W.err; tail -5 gpurun_out/bench_n$W.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_n$W.json")); print("N=$W", d.get("ms_per_step"), d.get("value"), d.get("ddp_parity"), (d.get("secondary") or {}).get("ms_per_step"), (d.get("secondary") or {}).get("value"))
except Exception as e: print("bench parse", e)
PY
B2_NVLS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus $W --steps 20 --warmup 5 --no-secondary > gpurun_out/r02_bench_n${W}_p2p.json 2> gpurun_out/bench_n${W}_p2p.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_n${W}_p2p.json")); print("N=$W p2p", d.get("ms_per_step"), d.get("value"))
except Exception as e: print("bench parse", e)
PY
