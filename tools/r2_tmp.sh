SR_DEBUG_SYNC=1 timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SR_DEBUG_SYNC=1 SR_PIPELINE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_random.py -m gpu -x -q 2>&1 | tail -2
