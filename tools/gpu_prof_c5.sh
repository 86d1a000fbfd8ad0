# ncu --set full of the staged rollout (config 5, uint8 schedules) -> gpurun_out/prof_rollout_staged.ncu-rep
python tools/bench_configs.py --only 5 --scale 0.25 --no-check > gpurun_out/c5_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_staged -s 1 -c 1 -o gpurun_out/prof_rollout_staged -f \
  python tools/bench_configs.py --only 5 --scale 0.25 --no-check > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log
