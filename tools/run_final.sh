# usage: bash tools/run_final.sh [notests]    (inside gpurun, one GPU): the round's closing validation + evidence
mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 --timeout-method=thread > gpurun_out/t_gpu_final.log 2>&1; tail -4 gpurun_out/t_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
fi
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02_bench_n1.json")); s=d.get("secondary") or {}
    print("N=1", d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], "R50", s.get("ms_per_step"), s.get("value"), "launches", d["gpu_launches"], d["clocks"])
except Exception as e: print("bench parse", e)
PY
timeout 200 python tools/step_trace.py --out gpurun_out/r02_trace_n1_final.txt 2>&1 | tail -1
B="python bench.py --steps 2 --warmup 3 --no-clocks --no-cpu-baseline --no-secondary --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -s 500 -c 300 --csv --log-file gpurun_out/r02_launches_dram.csv $B > gpurun_out/ncu_dram.log 2>&1; tail -1 gpurun_out/ncu_dram.log | cut -c1-200
# full-section capture: 16 launches only (the report must stay far below gpurun's 64 MiB return limit)
timeout 400 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"k_conv|k_wgrad" -s 312 -c 16 \
   -f -o gpurun_out/r02_prof_conv_wgrad $B > gpurun_out/ncu_conv.log 2>&1; tail -1 gpurun_out/ncu_conv.log
ls -la gpurun_out/*.ncu-rep | tail -2; du -sh gpurun_out
