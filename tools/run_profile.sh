# usage: bash tools/run_profile.sh      (inside gpurun, one GPU) -- the ncu evidence of a round
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-clocks --no-cpu-baseline --no-secondary --no-parity"
# (1) warm per-launch durations + DRAM bytes of every kernel of the timed steps (graph replays; L2 left as the step leaves it)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
   -s 500 -c 300 --csv --log-file gpurun_out/r02_launches_dram.csv $B > gpurun_out/ncu_dram.log 2>&1
# (2) full-section captures of the three dominant kernels (one training step's worth of each)
timeout 1200 ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_conv -s 195 -c 39 \
   -f -o gpurun_out/r02_prof_conv $B > gpurun_out/ncu_conv.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_wgrad -s 105 -c 21 \
   -f -o gpurun_out/r02_prof_wgrad $B > gpurun_out/ncu_wgrad.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none -k regex:"k_sgd|k_bn_apply|k_ce_loss|k_head|k_maxpool" -s 200 -c 30 \
   -f -o gpurun_out/r02_prof_elementwise $B > gpurun_out/ncu_elem.log 2>&1
ls -la gpurun_out/*.ncu-rep
