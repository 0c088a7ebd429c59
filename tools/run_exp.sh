mkdir -p gpurun_out
for pf in 1 0; do for mult in 2 4 8; do
B2_FUSE_POOL_BWD=$pf B2_BN_GRID_MULT=$mult timeout 300 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_x.json 2> gpurun_out/bench_h.err; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print('N=1 poolfuse=$pf mult=$mult', d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])"
done; done
B2_FUSE_POOL_BWD=1 B2_BN_GRID_MULT=8 timeout 200 python tools/step_trace.py --out gpurun_out/r02_trace_n1_pf8.txt 2>&1 | tail -1
B2_FUSE_POOL_BWD=0 B2_BN_GRID_MULT=8 timeout 200 python tools/step_trace.py --out gpurun_out/r02_trace_n1_nopf8.txt 2>&1 | tail -1
