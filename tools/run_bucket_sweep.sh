# usage: bash tools/run_bucket_sweep.sh W   (inside gpurun --gpus W): step time vs bucket size / comm launch shape / NVLS
W=$1
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29577 \
      bench.py --gpus $W --steps 20 --warmup 5 --no-secondary --no-parity --no-cpu-baseline $EXTRA > gpurun_out/sw_$label.json 2> gpurun_out/sw_$label.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sw_$label.json")); print("$label", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4))
except Exception as e: print("$label ERR", e)
PY
}
for th in 128 256 512; do for cb in 96 148; do for mb in 4 16; do
  EXTRA="--bucket-mb $mb" run w${W}_p2p_b${mb}_t${th}_cb$cb B2_NVLS=0 B2_COMM_THREADS=$th B2_COMM_BLOCKS=$cb
  EXTRA="--bucket-mb $mb" run w${W}_nvls_b${mb}_t${th}_cb$cb B2_NVLS=1 B2_COMM_THREADS=$th B2_COMM_BLOCKS=$cb
done; done; done
for mb in 1 2 8 1000; do
  EXTRA="--bucket-mb $mb" run w${W}_p2p_b${mb} B2_NVLS=0
  EXTRA="--bucket-mb $mb" run w${W}_nvls_b${mb} B2_NVLS=1
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29544 tools/allreduce_sweep.py --max-mb 256 --out gpurun_out/r02_allreduce_w$W.json > gpurun_out/sweep_w$W.log 2>&1; tail -12 gpurun_out/sweep_w$W.log
B2_COMM_THREADS=512 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29545 tools/allreduce_sweep.py --max-mb 256 --out gpurun_out/r02_allreduce_w${W}_t512.json > gpurun_out/sweep_w${W}_t512.log 2>&1; tail -12 gpurun_out/sweep_w${W}_t512.log
