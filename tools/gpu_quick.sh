B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-lean"
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
$B > gpurun_out/plain.log 2> gpurun_out/plain.err
tail -3 gpurun_out/pytest.log
python - <<PY
import json
d=json.loads(open("gpurun_out/plain.log").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], {k:round(v["ms"],3) for k,v in d["roofline"]["kernels"].items()}, d["clocks"])
PY
