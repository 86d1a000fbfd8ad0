bash tools/gpu_r2c.sh "$@"
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eks_gain -s 2 -c 1 -o gpurun_out/prof_eks_gain -f $B > gpurun_out/ncu_eks_gain.log 2>&1
tail -1 gpurun_out/ncu_eks_gain.log
