mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -x -q --timeout=180 --timeout-method=thread > gpurun_out/t_umma.log 2>&1; tail -5 gpurun_out/t_umma.log
timeout 300 python tools/conv_timeline.py > gpurun_out/r02_conv_timeline_b.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout=300 --timeout-method=thread > gpurun_out/t_parity.log 2>&1; tail -5 gpurun_out/t_parity.log
timeout 900 python -m pytest tests/test_gpu_bench_config.py -q --timeout=600 --timeout-method=thread > gpurun_out/t_bench_config.log 2>&1; tail -8 gpurun_out/t_bench_config.log
timeout 400 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -3 gpurun_out/bench_c.err
B2_FUSE_BN_APPLY=0 timeout 400 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/bench_c_noapply.json 2> gpurun_out/bench_c_noapply.err
B2_SPLITK=0 timeout 400 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/bench_c_nosplit.json 2> gpurun_out/bench_c_nosplit.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 400 --csv --log-file gpurun_out/r02_launches_c.csv python bench.py --steps 2 --warmup 3 --no-clocks --no-cpu-baseline --no-secondary > gpurun_out/ncu_c2.log 2>&1
python - <<'PY'
import json
for f in ["bench_c","bench_c_noapply","bench_c_nosplit"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["final_loss"])
    except Exception as e: print(f, "ERR", e)
PY
