# ncu --set full of one kernel of the 8-GPU shard (30 regions): bash tools/gpu_prof_small.sh [kernel regex] [env...]
K=${1:-eks_backward}; shift
B="python bench.py --regions 30 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary"
env "$@" timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_small -f $B > gpurun_out/ncu_small.log 2>&1
tail -2 gpurun_out/ncu_small.log
