for v in "$@"; do
  export EPI_B200_LIB=$PWD/epidemicmodeling_b200/variants/$v/libepi_b200.so
  python tools/bench_configs.py --only 5 --scale 0.5 > gpurun_out/c5_$v.log 2> gpurun_out/c5_$v.err
  python - <<PY
import json
try:
    for ln in open("gpurun_out/c5_$v.log"):
        ln=ln.strip()
        if ln.startswith("{"):
            d=json.loads(ln)
            if d.get("kernel","").startswith("rollout_cost[u8]"): print("$v", d["kernel"], "ms", round(d["ms"],3), "td/s %.3e"%d["trajectory_days_per_s"], d["oracle_spot_check"])
except Exception as e:
    print("$v FAILED", e)
PY
done
