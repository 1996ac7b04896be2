"""CPU-side checks of the bench.py contract the driver depends on: the reference arm (the reference's own CPU
binary, oracle/_ref, timed on the host cores) prints one JSON line with the agreed keys; the B200 arm refuses to
run without a CUDA device (no CPU fallback); the roofline helpers read the committed profile."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "phantom_env_cpu_debug")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref is built where /root/reference exists")
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-histories-per-proc", "400"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                   # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "histories/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_b200_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_roofline_helpers_read_the_committed_profile():
    """The committed ncu capture is quoted only for the launch size it was taken at and only while the library still
    contains the kernel it profiled (SASS hash); anything else is reported as stale, never as a number."""
    sys.path.insert(0, ROOT)
    import bench
    from moquimc_b200 import build as B
    peak, src = bench.measured_peaks()
    assert 3000.0 < peak < 9000.0 and ("measured" in src or "fallback" in src)
    prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    assert len(prof["sass_sha256"]) == 64 and prof["histories_per_launch"] == 10_000_000
    cap, note = bench.committed_capture(12345)
    assert cap is None and "launch size" in note            # only at the launch size of the capture
    ident = B.kernel_identity()
    if ident is None:
        pytest.skip("cuobjdump or the library is missing")
    cap, note = bench.committed_capture(10_000_000)
    if ident["sass_sha256"] == prof["sass_sha256"]:
        t = float(cap["dram_bytes_per_launch"])
        assert 1e10 < t < bench.STEPS_PER_HISTORY * bench.BYTES_PER_STEP * 1e7   # below the algorithmic bytes: L2 hits
        i = bench.ncu_issue(cap, 76.0, 1965.0)
        assert i is not None and 0.5 < i["frac"] < 1.0
        assert bench.ncu_issue(cap, 76.0, None) is None
    else:
        assert cap is None and note.startswith("stale")
    assert bench.ncu_issue(None, 76.0, 1965.0) is None
