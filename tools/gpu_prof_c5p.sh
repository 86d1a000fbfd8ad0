# ncu --set full of the in-kernel Philox rollout (config 5) -> gpurun_out/prof_rollout_philox.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/prof_rollout_philox -f \
  python tools/bench_configs.py --only 5 --scale 0.25 --no-check > gpurun_out/ncu_c5p.log 2>&1
tail -2 gpurun_out/ncu_c5p.log
