#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the reference itself.

Runs only in the build container (needs oracle/_ref built from /root/reference by build_ref.sh).
  kat_<variant>.npz        deterministic vectors dumped by oracle/ref_kat.cpp (reference headers)
  c1_*.npz, c2_*.npz ...   reduced dose files of the reference CPU phantom_env (oracle/ref_run.py);
                           the exact commands are listed in tests/golden/README.md
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "..", "tests", "golden")
DT = {"f32": np.float32, "i32": np.int32, "u32": np.uint32, "u64": np.uint64, "i16": np.int16}


def kat(variant):
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call([os.path.join(HERE, "_ref", "ref_kat_" + variant), d], stdout=subprocess.DEVNULL)
        out = {}
        for f in sorted(os.listdir(d)):
            name, ext = f.rsplit(".", 1)
            out[name] = np.fromfile(os.path.join(d, f), dtype=DT[ext])
        # the physics tables live in moquimc_b200/data/mqi_tables_v1.bin; keep only a checksum here
        for k in ("tables", "density_correction"):
            out[k + "_sum"] = np.array(out.pop(k).astype(np.float64).sum())
        np.savez_compressed(os.path.join(GOLD, "kat_%s.npz" % variant), **out)
        print("kat", variant, {k: v.shape for k, v in out.items()})


def c2(energy, procs=8, per_proc=125000):
    """Config C2: water / bone (HU +1000, 50-70 mm) / lung (HU -741, 70-100 mm) / water slabs on the C1
    grid, 10 mm uniform square spot, RELEASE physics (LETd is NaN-poisoned in the debug build, B18),
    scorers Dose + LETd_numer + LETd_denom through oracle/ref_harness.cpp -- three scorers, so the
    reference's Dose is scored twice per step (quirk B2)."""
    import argparse
    sys.path.insert(0, HERE)
    import ref_run
    a = argparse.Namespace(variant="release", procs=procs, histories_per_proc=per_proc, energy=float(energy), spot_size=10.0,
                           nxyz=[200, 200, 350], lxyz=[100.0, 100.0, 350.0], slab=[[50.0, 70.0, 1000.0], [70.0, 100.0, -741.0]],
                           seed=4242, rebin=8, harness=True, scorers="dose+letd", gauss=None, out=None)
    res, meta = ref_run.run(a)
    keep = {}
    for k, v in res.items():
        if k == "meta" or k.endswith(("_idd", "_idd_se", "_total", "_total_se")) or k in ("Dose_xz", "Dose_xz_se", "LETd_numer_xz", "LETd_denom_xz"):
            keep[k] = v.astype(np.float32) if k.endswith("_se") and v.ndim > 1 else v
    np.savez_compressed(os.path.join(GOLD, "c2_slabs%d_release.npz" % int(energy)), **keep)
    print("c2", energy, {k: getattr(v, "shape", None) for k, v in keep.items()})


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "c2":
        for e in (70, 150, 230):
            c2(e)
        sys.exit(0)
    for v in ("debug", "release"):
        kat(v)
