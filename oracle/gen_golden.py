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


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    for v in ("debug", "release"):
        kat(v)
