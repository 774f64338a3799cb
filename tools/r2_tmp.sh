SR_DEBUG_SYNC=1 SR_PIPELINE=0 timeout 900 python -m pytest tests/test_bunny.py -m gpu -x -q 2>&1 | grep "SrError\|passed\|failed" | head
