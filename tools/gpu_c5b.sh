# usage: bash tools/gpu_c5b.sh [variant ...]  -- config 5 (random-NPI scoring) with the default library and the named variants
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rollout or random or schedul or config5 or monte" 2>&1 | tail -2
for v in default "$@"; do
  if [ "$v" = default ]; then unset EPI_B200_LIB; else export EPI_B200_LIB=$PWD/epidemicmodeling_b200/variants/$v/libepi_b200.so; fi
  python tools/bench_configs.py --only 5 --scale ${C5_SCALE:-0.5} > gpurun_out/c5_$v.log 2> gpurun_out/c5_$v.err
  python - <<PY
import json
try:
    for ln in open("gpurun_out/c5_$v.log"):
        ln=ln.strip()
        if ln.startswith("{"):
            d=json.loads(ln)
            if d.get("kernel","").startswith("rollout_cost"): print("$v", d["kernel"], "ms", round(d["ms"],3), "td/s %.3e"%d["trajectory_days_per_s"], d.get("oracle_spot_check"))
except Exception as e:
    print("$v FAILED", e)
PY
done
