#!/usr/bin/env python
"""gpurun_out/golden_gpu/*.npz (written on the GPU box by oracle/gen_golden_gpu.py) -> tests/golden/, with the large
arrays quantised so that the fixtures stay small: doses as uint16 levels of the maximum (the step, maximum / levels,
is a small fraction of the statistical error and of the 1 % gamma criterion), standard errors as float16.
Test infrastructure only."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "gpurun_out", "golden_gpu")
DST = os.path.join(HERE, "..", "tests", "golden")


def quant(a, levels):
    m = float(a.max())
    return np.round(a.astype(np.float64) / m * levels).astype(np.uint16), m


def main():
    for f in sorted(os.listdir(SRC)):
        src = os.path.join(SRC, f)
        if f.startswith("c1_water200_") and f.endswith("_3d.npz"):
            d = np.load(src)
            meta = json.loads(str(d["meta"]))
            lv = 2047
            q = np.round(d["q"].astype(np.float64) * (lv / meta["levels"])).astype(np.uint16)
            meta["levels"] = lv
            np.savez_compressed(os.path.join(DST, f), q=q, se_block=d["se_block"].astype(np.float16), idd=d["idd"], idd_se=d["idd_se"],
                                total=d["total"], meta=np.array(json.dumps(meta)))
        elif f == "c3like_head_release.npz":
            d = np.load(src)
            meta = json.loads(str(d["meta"]))
            dq, dmax = quant(d["dose"], 4095)
            rq, rmax = quant(d["dij_rows_full"], 4095)
            meta.update(dose_levels=4095, dose_max=dmax, rows_levels=4095, rows_max=rmax, rows_full=[0, 9, 19])
            np.savez_compressed(os.path.join(DST, f), dose_q=dq, dose_se=d["dose_se"].astype(np.float16), dij_row_total=d["dij_row_total"],
                                dij_row_idd=d["dij_row_idd"], dij_row_xy=d["dij_row_xy"].astype(np.float16),
                                dij_nnz_per_row=d["dij_nnz_per_row"], dij_rows_full_q=rq, meta=np.array(json.dumps(meta)))
        elif f == "a15_stat_release.npz":
            d = np.load(src)
            np.savez_compressed(os.path.join(DST, f), **{k: d[k] for k in d.files if not k.endswith("_dose")})
        elif f.endswith(".npz") or f.endswith(".json"):
            with open(src, "rb") as a, open(os.path.join(DST, f), "wb") as b:
                b.write(a.read())
        else:
            continue
        print(f, os.path.getsize(os.path.join(DST, f)))


if __name__ == "__main__":
    sys.exit(main())
