#!/usr/bin/env python
"""Run the reference's own CPU phantom_env (oracle/_ref, built by build_ref.sh) as P independent
single-threaded processes with distinct --random_seed, and reduce the dose files to a small golden
fixture (depth dose, projections, a laterally rebinned 3-D grid, and batch standard errors).

Test infrastructure only.  Used (a) here, in the build container, to generate tests/golden/*.npz and
(b) by bench.py's cpu_baseline / --impl reference leg on the GPU box's host cores.

The reference CPU path is single-threaded (mqi_phantom_env.hpp:420), so whole-host throughput is
P processes x N/P histories.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def make_phantom(path, nxyz, slabs):
    """HU volume int16 [nz][ny][nx]; slabs = list of (depth0_mm, depth1_mm, HU) measured from the
    beam-entry face (z = top of the box); index rule k = nz - 1 - floor(depth / dz)."""
    nx, ny, nz = nxyz
    hu = np.zeros((nz, ny, nx), dtype=np.int16)
    for d0, d1, h, dz in slabs:
        k_hi = nz - 1 - int(np.floor(d0 / dz))
        k_lo = nz - int(np.floor(d1 / dz))
        hu[k_lo:k_hi + 1] = h
    hu.tofile(path)
    return hu


def reduce_dose(d, nxyz, rebin):
    nx, ny, nz = nxyz
    d = d.reshape(nz, ny, nx)
    out = {
        "idd": d.sum(axis=(1, 2)),
        "xz": d.sum(axis=1),
        "yz": d.sum(axis=2),
        "xy": d.sum(axis=0),
        "reb": d.reshape(nz, ny // rebin, rebin, nx // rebin, rebin).sum(axis=(2, 4)),
        "total": np.array(d.sum()),
    }
    return out


def run(args):
    nx, ny, nz = args.nxyz
    lx, ly, lz = args.lxyz
    work = tempfile.mkdtemp(prefix="mqi_ref_")
    try:
        slabs = [(a, b, h, lz / nz) for a, b, h in args.slab]
        ph = os.path.join(work, "phantom.raw")
        make_phantom(ph, (nx, ny, nz), slabs)
        exe = os.path.join(REF, ("ref_harness_" if args.harness else "phantom_env_cpu_") + args.variant)
        procs = []
        t0 = time.time()
        for p in range(args.procs):
            od = os.path.join(work, "out%d" % p)
            os.makedirs(od)
            cmd = [exe, "--lxyz", str(lx), str(ly), str(lz), "--pxyz", "0.0", "0.0", str(-0.5 * lz),
                   "--nxyz", str(nx), str(ny), str(nz),
                   "--spot_energy", str(args.energy), "0.0", "--spot_position", "0", "0", str(getattr(args, "spot_z", 0.5)),
                   "--spot_size", str(args.spot_size), str(args.spot_size),
                   "--histories", str(args.histories_per_proc), "--phantom_path", ph,
                   "--output_prefix", od, "--random_seed", str(args.seed + 7919 * p), "--gpu_id", "0"]
            if args.harness:
                cmd += ["--scorers", args.scorers]
                if args.gauss:
                    cmd += ["--gauss"] + [str(g) for g in args.gauss]
                cmd += [str(x) for x in (getattr(args, "extra", None) or [])]
            procs.append((subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT), od))
        for pr, _ in procs:
            rc = pr.wait()
            if rc != 0:
                raise RuntimeError("reference process failed rc=%d" % rc)
        wall = time.time() - t0
        import re
        # <child index>_<scorer>.raw: the phantom is child 0, or the last child behind beamline nodes
        names = sorted(f for f in os.listdir(procs[0][1]) if re.match(r"^\d+_.*\.raw$", f))
        result = {}
        for name in names:
            key = name.split("_", 1)[1][:-4]
            batches = []
            for _, od in procs:
                d = np.fromfile(os.path.join(od, name), dtype=np.float64)
                batches.append(reduce_dose(d, (nx, ny, nz), args.rebin))
            for k in batches[0]:
                stack = np.stack([b[k] for b in batches])
                n_total = args.procs * args.histories_per_proc
                # per-history mean and its standard error from the P batches
                mean = stack.sum(axis=0) / n_total
                result["%s_%s" % (key, k)] = mean
                if args.procs > 1:
                    per_batch = stack / args.histories_per_proc
                    result["%s_%s_se" % (key, k)] = per_batch.std(axis=0, ddof=1) / np.sqrt(args.procs)
        meta = dict(variant=args.variant, energy=args.energy, histories=args.procs * args.histories_per_proc,
                    procs=args.procs, nxyz=[nx, ny, nz], lxyz=[lx, ly, lz], spot_size=args.spot_size,
                    slab=args.slab, seed=args.seed, rebin=args.rebin, wall_s=wall, scorers=args.scorers,
                    harness=bool(args.harness), gauss=args.gauss,
                    units="per primary history (reference output / histories)")
        result["meta"] = np.array(json.dumps(meta))
        if args.out:
            small = ("reb", "reb_se", "xz_se", "yz_se")
            np.savez_compressed(args.out, **{k: (v.astype(np.float32) if k.endswith(small) else v)
                                             for k, v in result.items()})
        print(json.dumps(dict(histories=meta["histories"], wall_s=wall,
                              hist_per_s=meta["histories"] / wall, procs=args.procs)))
        return result, meta
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="debug", choices=["debug", "release"])
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--histories-per-proc", type=int, default=10000)
    ap.add_argument("--energy", type=float, default=200.0)
    ap.add_argument("--spot-size", type=float, default=30.0)
    ap.add_argument("--nxyz", type=int, nargs=3, default=[200, 200, 350])
    ap.add_argument("--lxyz", type=float, nargs=3, default=[100.0, 100.0, 350.0])
    ap.add_argument("--slab", type=float, nargs=3, action="append", default=[],
                    metavar=("D0", "D1", "HU"), help="depth range [mm] from the entry face and its HU")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--rebin", type=int, default=8)
    ap.add_argument("--harness", action="store_true")
    ap.add_argument("--scorers", default="dose")
    ap.add_argument("--gauss", type=float, nargs=5, default=None)
    ap.add_argument("--spot-z", type=float, default=0.5)
    ap.add_argument("--extra", nargs=argparse.REMAINDER, default=None, help="further ref_harness flags")
    ap.add_argument("--out", default=None)
    run(ap.parse_args())


if __name__ == "__main__":
    sys.exit(main())
