B="python bench.py --regions 30 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary"
export EPI_ROWS=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${1:-eks_backward} -s 2 -c 1 -o gpurun_out/prof_small -f $B > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log
