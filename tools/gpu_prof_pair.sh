B="python bench.py --regions 30 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary"
export EPI_PAIR=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forward_pair -s 2 -c 1 -o gpurun_out/prof_pair -f $B > gpurun_out/ncu_pair.log 2>&1
tail -2 gpurun_out/ncu_pair.log
