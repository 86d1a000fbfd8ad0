run() { # name, env...
  name=$1; shift
  env "$@" timeout 60 python bench.py --regions ${REG:-30} --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/pipe_$name.log 2> gpurun_out/pipe_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/pipe_$name.log").read().strip().splitlines()[-1])
    print("$name ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$name FAILED", e)
PY
}
run seq_daysync1 EPI_PIPE=2 EPI_PIPE_CHUNKS=1 EPI_PIPE_DAYSYNC=1
run piped_daysync1 EPI_PIPE=1 EPI_PIPE_DAYSYNC=1
