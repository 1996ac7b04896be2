"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/mqi_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import pytest

from moquimc_b200 import capi


def test_library_exports_every_declared_symbol():
    L = capi.load()
    names = capi.header_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(L, n), n
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported
    # nothing but the C ABI leaks out of the library
    assert all(e.startswith("mqi_") for e in exported), exported - set(names)


def test_version_and_error_strings():
    L = capi.load()
    assert b"sm_100a" in L.mqi_version()
    assert isinstance(L.mqi_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.MqiError) as e:
        capi.Engine(0)
    assert e.value.code == capi.ENODEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "moquimc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "mqi_oracle" not in text and "oracle_lib" not in text, os.path.join(dp, f)
                assert "libmqi_oracle" not in text


def test_kernel_and_oracle_share_the_philox_round_count():
    """The RNG protocol is one constant in two places: the kernel's MQI_K_PHILOX_ROUNDS and the
    oracle's MQO_PHILOX_ROUNDS.  Identical-stream parity tests only mean something while they agree."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    k = re.search(r"#define\s+MQI_K_PHILOX_ROUNDS\s+(\d+)", open(os.path.join(root, "moquimc_b200", "csrc", "mqi_kernels.h")).read())
    o = re.search(r"#define\s+MQO_PHILOX_ROUNDS\s+(\d+)", open(os.path.join(root, "oracle", "mqi_oracle.h")).read())
    assert k and o and int(k.group(1)) == int(o.group(1)) == 7
