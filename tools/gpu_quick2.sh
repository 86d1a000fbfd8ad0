timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 600 python tools/bench_configs.py --only 5 > gpurun_out/c5.log 2>&1; tail -4 gpurun_out/c5.log | cut -c1-400
