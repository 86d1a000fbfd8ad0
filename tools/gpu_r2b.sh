# round 2: parity of the pair-at-a-time Jacobi + timing of its occupancy variants
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-lean --no-secondary"
$B > gpurun_out/plain.log 2> gpurun_out/plain.err
python - <<PY
import json
d=json.loads(open("gpurun_out/plain.log").read().strip().splitlines()[-1])
print("main ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], {k:round(v["ms"],3) for k,v in d["roofline"]["kernels"].items()}, d["clocks"])
PY
bash tools/gpu_variants.sh "$@"
