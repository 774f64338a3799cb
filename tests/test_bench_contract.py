"""The driver's bench contract on the legs that run without a GPU: `bench.py --impl reference` (the
reference's own MultiViewStereo from oracle/_ref when it is there, else the oracle port) prints ONE JSON line
with the contract's keys.  (The product arm's line is produced on the GPU box by the driver itself;
profiles/r1_bench_default.json is a committed copy.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--workload", "small", "--steps", "1", "--warmup", "0", "--cpu-rows", "8"])
    assert d["impl"] == "reference" and d["unit"] == "Mpix*disp/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    from oracle import oracle_api as O
    if O.ref_lib() is not None:
        assert cb["kind"] == "reference" and "of the reference itself" in cb["sample"]
