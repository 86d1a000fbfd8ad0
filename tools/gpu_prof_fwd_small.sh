# forward pass of the 8-GPU shard: the piped kernel (3-warp SM-exclusive CTAs) and the spread one-warp kernel
B="python bench.py --regions 30 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary"
EPI_PIPE=2 EPI_PIPE_CHUNKS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ekf_forward_piped -s 2 -c 1 -o gpurun_out/prof_fwd_piped_small -f $B > gpurun_out/ncu_fps.log 2>&1
tail -1 gpurun_out/ncu_fps.log
EPI_PIPE=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ekf_forward -s 2 -c 1 -o gpurun_out/prof_fwd_spread_small -f $B > gpurun_out/ncu_fss.log 2>&1
tail -1 gpurun_out/ncu_fss.log
