mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout=300 --timeout-method=thread > gpurun_out/t_parity.log 2>&1; tail -3 gpurun_out/t_parity.log
timeout 900 python -m pytest tests/test_gpu_bench_config.py -q --timeout=600 --timeout-method=thread > gpurun_out/t_bench_config.log 2>&1; tail -6 gpurun_out/t_bench_config.log
timeout 400 python bench.py --no-cpu-baseline --no-secondary > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; tail -3 gpurun_out/bench_g.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 500 -c 400 --csv --log-file gpurun_out/r02_launches_g.csv python bench.py --steps 2 --warmup 3 --no-clocks --no-cpu-baseline --no-secondary > gpurun_out/ncu_g.log 2>&1
python - <<'PY'
import json
for f in ["bench_g"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["final_loss"], d["roofline"]["frac"], {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernel_avg"].items()})
    except Exception as e: print(f, "ERR", e)
PY
