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
            q0 = d["q"].astype(np.float64) / meta["levels"]          # dose / dmax on the generator's box
            # the gamma test looks at voxels above 10 % of the maximum and 1 mm around them: crop to the box of the
            # voxels above 8 % plus three voxels (lateral limits on multiples of 4: the block grid of the errors)
            zz, yy, xx = np.nonzero(q0 > 0.08)
            z0, z1 = max(zz.min() - 3, 0), min(zz.max() + 4, q0.shape[0])
            y0, y1 = max(yy.min() - 3, 0) // 4 * 4, min((yy.max() + 7) // 4 * 4, q0.shape[1])
            x0, x1 = max(xx.min() - 3, 0) // 4 * 4, min((xx.max() + 7) // 4 * 4, q0.shape[2])
            lv = 1023
            q = np.round(q0[z0:z1, y0:y1, x0:x1] * lv).astype(np.uint16)
            B = meta["box"]
            meta["box"] = [int(B[0] + z0), int(B[0] + z1), int(B[2] + y0), int(B[2] + y1), int(B[4] + x0), int(B[4] + x1)]
            meta["levels"] = lv
            meta["units"] = ("dose per primary history = q / levels * dmax; se_block_rel * dmax = standard error of the mean dose, root of "
                             "the run-to-run variance averaged over 4x4 lateral voxel blocks")
            se_rel = (d["se_block"].astype(np.float64) / meta["dmax"])[z0:z1, y0 // 4:y1 // 4, x0 // 4:x1 // 4]
            np.savez_compressed(os.path.join(DST, f), q=q, se_block_rel=se_rel.astype(np.float16), idd=d["idd"], idd_se=d["idd_se"],
                                total=d["total"], meta=np.array(json.dumps(meta)))
        elif f == "c3like_head_release.npz":
            d = np.load(src)
            meta = json.loads(str(d["meta"]))
            dq, dmax = quant(d["dose"], 4095)
            rq, rmax = quant(d["dij_rows_full"], 4095)
            meta.update(dose_levels=4095, dose_max=dmax, rows_levels=4095, rows_max=rmax, rows_full=[0, 9, 19])
            xy_max = float(d["dij_row_xy"].max())
            meta.update(row_xy_max=xy_max)
            np.savez_compressed(os.path.join(DST, f), dose_q=dq, dose_se_rel=(d["dose_se"].astype(np.float64) / dmax).astype(np.float16),
                                dij_row_total=d["dij_row_total"], dij_row_idd=d["dij_row_idd"],
                                dij_row_xy_rel=(d["dij_row_xy"].astype(np.float64) / xy_max).astype(np.float16),
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
