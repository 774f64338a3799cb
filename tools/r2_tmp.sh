timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c_pytest.log
PYTHONPATH=. python tools/curve_time.py 2>&1 | tail -4
SR_CURVE_INTERP=0 PYTHONPATH=. python tools/curve_time.py 2>&1 | tail -4 | head -2
