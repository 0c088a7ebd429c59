# usage: bash tools/run_trace.sh W   (inside gpurun --gpus W)
W=$1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bench_config.py -x -q -k "resnet18_bf16_b512_per_op or fused" --timeout=300 > gpurun_out/t_pf.log 2>&1; tail -3 gpurun_out/t_pf.log
timeout 200 python tools/step_trace.py --out gpurun_out/r02_trace_n1_pf.txt 2>&1 | tail -1
for sb in 0 32 64 128 256; do
B2_SGD_BLOCKS=$sb timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_sb$sb.json 2> gpurun_out/bench_h.err; python -c "
import json; d=json.load(open('gpurun_out/bench_sb$sb.json')); print('N=1 sgd_blocks=$sb', d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])"
done
B2_SGD_BLOCKS=64 timeout 200 python tools/step_trace.py --out gpurun_out/r02_trace_n1_sb64.txt 2>&1 | tail -1
B2_FUSE_POOL_BWD=0 timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_hnp.json 2> gpurun_out/bench_h.err; python -c "
import json; d=json.load(open('gpurun_out/bench_hnp.json')); print('N=1 nopoolfuse', d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])"
if [ "$W" != "1" ]; then
for cb in 96 48 32; do for mb in 32; do for th in 512 256; do
  export B2_NVLS=0 B2_COMM_BLOCKS=$cb B2_COMM_THREADS=$th
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29588 \
     tools/step_trace.py --bucket-mb $mb --out gpurun_out/r02_trace_n${W}_b${mb}_cb${cb}_t$th.txt 2>&1 | grep -v Warn | tail -1
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29577 \
      bench.py --gpus $W --steps 20 --warmup 5 --no-secondary --no-parity --no-cpu-baseline --bucket-mb $mb > gpurun_out/sw4_w${W}_b${mb}_cb${cb}_t$th.json 2> gpurun_out/sw4.err
  python -c "
import json; d=json.load(open('gpurun_out/sw4_w${W}_b${mb}_cb${cb}_t$th.json')); print('N=$W b$mb cb$cb t$th', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"
done; done; done
fi
