#!/usr/bin/env python
"""Generator of moquimc_b200/data/mqi_tables_v1.bin, the physics-table blob compiled into libmqi_b200.so.

    python oracle/gen_tables.py            # rewrite the blob from the reference's own headers
    python oracle/gen_tables.py --check    # byte-for-byte comparison with the committed blob (exit 1 on a difference)

Source of the numbers: oracle/_ref/ref_kat_release (oracle/ref_kat.cpp section 7, built by oracle/build_ref.sh),
which includes the reference's headers from /root/reference and dumps, in their own float precision,
  tables.f32               [6][600]: cs_p_ion_table, restricted_stopping_power_table, range_steps
                           (physics/mqi_physics_data.hpp), cs_pp_e_g4_table, cs_po_e_g4_table, cs_po_i_g4_table
  density_correction.f32   [3996]  : materials/mqi_patient_materials.hpp:11-412
Layout: "MQITBL1\\0" | uint32 600 | uint32 3996 | 6 x 600 float32 | 3996 float32  (little endian, 30 400 bytes).
Runs only where oracle/_ref exists (the build container); tests/test_oracle_kat.py::test_tables_blob_is_what_the_
reference_headers_hold repeats the check in the CPU suite.  Test / build infrastructure only.
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BLOB = os.path.join(HERE, "..", "moquimc_b200", "data", "mqi_tables_v1.bin")
KAT = os.path.join(HERE, "_ref", "ref_kat_release")


def build_blob():
    with tempfile.TemporaryDirectory() as d:
        # the reference's own code dies with a signal now and then (undefined behaviour upstream; seen once
        # in this dump too): such a run is repeated, any other failure is one
        for attempt in range(3):
            rc = subprocess.call([KAT, d], stdout=subprocess.DEVNULL)
            if rc >= 0:
                break
        if rc != 0:
            raise subprocess.CalledProcessError(rc, [KAT, d])
        tabs = np.fromfile(os.path.join(d, "tables.f32"), dtype="<f4")
        corr = np.fromfile(os.path.join(d, "density_correction.f32"), dtype="<f4")
    assert tabs.size == 6 * 600 and corr.size == 3996
    return b"MQITBL1\0" + struct.pack("<II", 600, 3996) + tabs.tobytes() + corr.tobytes()


def main():
    blob = build_blob()
    if "--check" in sys.argv:
        same = open(BLOB, "rb").read() == blob
        print("mqi_tables_v1.bin %s the reference headers" % ("matches" if same else "DIFFERS from"))
        return 0 if same else 1
    open(BLOB, "wb").write(blob)
    print("wrote", BLOB, len(blob))
    return 0


if __name__ == "__main__":
    sys.exit(main())
