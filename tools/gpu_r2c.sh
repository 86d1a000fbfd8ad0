# quick: sweep/ekf6 parity subset + bench of the main lib and variants
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=600 -k "sweep or ekf6 or optctrl or smoke or eks" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -2 gpurun_out/pytest.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary"
$B > gpurun_out/plain.log 2> gpurun_out/plain.err
python - <<PY
import json
d=json.loads(open("gpurun_out/plain.log").read().strip().splitlines()[-1])
print("main ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], {k:round(v["ms"],3) for k,v in d["roofline"]["kernels"].items()})
PY
if [ -n "$1" ]; then bash tools/gpu_variants.sh "$@"; fi
