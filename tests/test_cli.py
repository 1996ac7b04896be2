"""phantom_env CLI (C++ host over the C ABI): flag handling and error behaviour on CPU, full run on GPU."""
import os
import subprocess

import numpy as np
import pytest

from moquimc_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "moquimc_b200", "bin", "phantom_env")
C1 = ["--lxyz", "100", "100", "350", "--pxyz", "0.0", "0.0", "-175", "--nxyz", "200", "200", "350",
      "--spot_energy", "200.0", "0.0", "--spot_position", "0", "0", "0.5", "--spot_size", "30.0", "30.0"]


def run_cli(args):
    return subprocess.run([EXE] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def test_cli_is_built():
    assert os.access(EXE, os.X_OK), "run python -m moquimc_b200.build"


def test_cli_requires_output_prefix_and_phantom_path(tmp_path):
    r = run_cli(C1 + ["--histories", "10"])
    assert r.returncode != 0 and "output_path is required." in r.stderr
    r = run_cli(C1 + ["--histories", "10", "--output_prefix", str(tmp_path)])
    assert r.returncode != 0 and "phantom_path is required." in r.stderr


def test_cli_echoes_flags_like_the_reference(tmp_path):
    r = run_cli(["--histories", "10", "--bogus_flag", "1", "--output_prefix", str(tmp_path)])
    assert "# of arguments: 7" in r.stdout
    assert "--histories : 10 " in r.stdout
    assert "--bogus_flag" not in r.stdout   # unknown options are ignored (mqi_cli.hpp:79-81)


def test_cli_fails_loudly_without_a_gpu(tmp_path):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    ph = tmp_path / "ph.raw"
    np.zeros(8, dtype=np.int16).tofile(ph)
    r = run_cli(["--lxyz", "2", "2", "2", "--pxyz", "0", "0", "-1", "--nxyz", "2", "2", "2", "--histories", "10",
                 "--output_prefix", str(tmp_path), "--phantom_path", str(ph), "--random_seed", "1"])
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr
    assert not (tmp_path / "0_water_dE_total.raw").exists()


@pytest.mark.gpu
def test_cli_c1_output_matches_the_library_and_the_reference_layout(tmp_path, golden_dir):
    import dose_metrics as M
    ph = tmp_path / "water_phantom.raw"
    np.zeros((350, 200, 200), dtype=np.int16).tofile(ph)
    n = 200000
    out = {}
    for fmt in ("raw", "mhd", "mha"):
        od = tmp_path / fmt
        od.mkdir()
        r = run_cli(C1 + ["--histories", str(n), "--phantom_path", str(ph), "--output_prefix", str(od),
                          "--random_seed", "12345", "--gpu_id", "0", "--output_format", fmt])
        assert r.returncode == 0, r.stderr
        assert "Number of particles tracked %d" % n in r.stdout
        assert "Time taken by MC engine" in r.stdout
        out[fmt] = od
    d = np.fromfile(out["raw"] / "0_water_dE_total.raw", dtype=np.float64)
    assert d.size == 200 * 200 * 350      # float64 [nz][ny][nx], unscaled sum over histories
    # same seed, same history range through the Python binding: same dose up to fp64 summation order
    e = capi.Engine(0, physics=capi.PHYSICS_DEBUG)
    e.set_grid_hu(capi.uniform_edges(-50, 50, 200), capi.uniform_edges(-50, 50, 200), capi.uniform_edges(-350, 0, 350),
                  np.zeros((350, 200, 200), dtype=np.int16))
    e.add_scorer(capi.SCORER_DOSE, "water_dE_total")
    e.set_beamlets([capi.make_beamlet(200.0, [0, 0, 0.5, 0, 0, -1], [30, 30, 0, 0, 0, 0], uniform=True)], [n])
    e.run(12345, 0, n)
    np.testing.assert_allclose(d, e.get_dense(0).ravel(), rtol=1e-9, atol=1e-22)
    # against the reference's own output for this configuration (golden: per history)
    gold = np.load(os.path.join(golden_dir, "c1_water200_debug.npz"))
    idd = d.reshape(350, 200, 200).sum(axis=(1, 2)) / n
    assert abs(M.r80_mm(idd) - M.r80_mm(gold["water_dE_total_idd"])) < 0.1
    assert abs(idd.sum() / float(gold["water_dE_total_total"]) - 1.0) < 5e-3
    # mhd = header + the same raw; mha = header + inline data (mqi_io.hpp:493-591)
    hdr = (out["mhd"] / "0_water_dE_total.mhd").read_text()
    assert "DimSize = 200 200 350" in hdr and "ElementType = MET_DOUBLE" in hdr
    assert "ElementSpacing = 0.5 0.5 1" in hdr and "Offset -49.75 -49.75 -349.5" in hdr
    assert "ElementDataFile = 0_water_dE_total.raw" in hdr
    # (separate runs: identical histories, fp64 atomics in a different order)
    np.testing.assert_allclose(np.fromfile(out["mhd"] / "0_water_dE_total.raw", dtype=np.float64), d, rtol=1e-9, atol=1e-22)
    blob = (out["mha"] / "0_water_dE_total.mha").read_bytes()
    cut = blob.index(b"ElementDataFile = LOCAL\n") + len(b"ElementDataFile = LOCAL\n")
    assert b"Origin = -49.75 -49.75 -349.5\n" in blob[:cut] and b"HeaderSize = -1\n" in blob[:cut]
    np.testing.assert_allclose(np.frombuffer(blob[cut:], dtype=np.float64), d, rtol=1e-9, atol=1e-22)
