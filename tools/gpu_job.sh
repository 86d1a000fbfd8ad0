# Round-end style GPU job: tests, smoke, bench, ncu launch list + full captures, secondary configs.
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-lean"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
$B > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
for k in ekf_forward eks_gain eks_backward; do timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_$k -f $B > gpurun_out/ncu_$k.log 2>&1; done
C="python tools/bench_configs.py --only 2 --scale 0.5"
$C > gpurun_out/c2_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:seirp_staged -s 1 -c 1 -o gpurun_out/prof_seirp_staged -f $C > gpurun_out/ncu_seirp.log 2>&1
C5="python tools/bench_configs.py --only 5 --scale 0.25"
$C5 > gpurun_out/c5_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_staged -s 1 -c 1 -o gpurun_out/prof_rollout_staged -f $C5 > gpurun_out/ncu_rollout.log 2>&1
timeout 800 python tools/bench_configs.py > gpurun_out/configs.log 2>&1
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/pytest.log; tail -2 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_reference.log | tail -1
python - <<PY
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_step_median"], {k:round(v["ms"],3) for k,v in d["roofline"]["kernels"].items()}, d["clocks"], "cpu %.3e"%d["cpu_baseline"]["value"], d["cpu_baseline"]["interpreted_proxy"]["value"], "lean", d["lean_mode"]["ms_per_step"])
PY
