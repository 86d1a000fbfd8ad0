timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
timeout 300 python tools/prescribe_from_csv.py --regions 16 --t-hist 150 --t-fore 30 --eps 50 2>&1 | tail -4
head -3 gpurun_out/prescriptions.csv
