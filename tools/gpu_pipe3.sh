for v in pw2 pw3 pw6; do
  export EPI_B200_LIB=$PWD/epidemicmodeling_b200/variants/$v/libepi_b200.so
  for R in 30 59; do
  EPI_PIPE_DAYSYNC=0 timeout 60 python bench.py --regions $R --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary > gpurun_out/pipe_$v.log 2> gpurun_out/pipe_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/pipe_$v.log").read().strip().splitlines()[-1])
    print("$v regions $R ms/step", round(d["ms_per_step"],3), {k:round(x["ms"],3) for k,x in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$v FAILED", e)
PY
  done
done
