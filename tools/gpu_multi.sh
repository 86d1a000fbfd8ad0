# usage: bash tools/gpu_multi.sh N [extra bench args]   (run under gpurun --gpus N)
N=$1; shift
timeout ${MULTI_TIMEOUT:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 "$@" > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N.log").read().strip().splitlines()[-1])
    s=d.get("strong_scaling") or {}
    print("N=$N value %.4e ms/step %.3f e2e %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step")))
    print("serial", s.get("ms_per_step_serial"), "floor", s.get("collective_and_pareto_floor_ms"), "schedule", (s.get("schedule") or "")[:40])
    print("rank0 kernels", (s.get("kernel_ms_per_rank") or [None])[0])
    print("host_multi", json.dumps(d.get("host_multi_call"))[:300])
    print("lean", json.dumps(d.get("lean_mode"))[:200])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_n$N.err").read()[-1500:])
PY
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N.log").read().strip().splitlines()[-1])
    print("config5_sharded", json.dumps(d.get("config5_sharded"))[:900])
except Exception as e:
    print("FAILED", e)
PY
