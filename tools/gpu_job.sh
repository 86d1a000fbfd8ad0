timeout 600 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
for v in main g3 g5; do
  if [ $v = main ]; then L=""; else L=$PWD/epidemicmodeling_b200/variants/libepi_$v.so; fi
  EPI_B200_LIB=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err
done
timeout 600 python tools/bench_configs.py > gpurun_out/configs.log 2>&1
python - <<PY
import json
for n in ("main","g3","g5"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.log").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], {k:round(v["ms"],3) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
for l in open("gpurun_out/configs.log"):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    if d["config"]==5: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k not in ("hbm_peak_gbs","fp64_peak_tflops_measured","kernel_ms")})
PY
