B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary"
$B > gpurun_out/plain.log 2> gpurun_out/plain.err || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${1:-eks_gain} -s 2 -c 1 -o gpurun_out/prof_${1:-eks_gain} -f $B > gpurun_out/ncu_${1:-eks_gain}.log 2>&1
tail -2 gpurun_out/ncu_${1:-eks_gain}.log
