#!/usr/bin/env python
"""High-statistics golden fixtures from the reference's own CUDA path, generated ON THE GPU BOX.

    gpurun -- 'python oracle/gen_golden_gpu.py probe c1 c2sweep stat dose2 c3like [--budget-s S]'

Runs oracle/_ref/ref_harness_gpu_<variant> (oracle/ref_harness.cpp compiled with nvcc -x cu by oracle/build_ref.sh:
the reference's transport_particles_patient / _stat kernels, its scorers, its host beam sampler and, for the
stopping criterion, its calculate_standard_deviation kernel -- reference headers included from /root/reference at
build time, never copied) and reduces the outputs to small arrays under gpurun_out/golden_gpu/, which are then
committed under tests/golden/ (README there).  /root/reference is not needed at run time: the binaries travel with
the repo snapshot.  Test infrastructure only.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
OUT = os.path.join(ROOT, "gpurun_out", "golden_gpu")
NCPU = os.cpu_count() or 8
C1 = dict(nxyz=(200, 200, 350), lxyz=(100.0, 100.0, 350.0))


def harness_start(variant, out_dir, nxyz, lxyz, phantom, energy, histories, seed, scorers="dose", spot_size=30.0, spot_z=0.5,
                  pxyz=None, batches=1, threads=None, extra=()):
    os.makedirs(out_dir, exist_ok=True)
    pxyz = pxyz or (0.0, 0.0, -0.5 * lxyz[2])
    cmd = [os.path.join(REF, os.environ.get("MQI_GOLD_HARNESS", "ref_harness_gpu_") + variant),
           "--lxyz"] + [str(v) for v in lxyz] + ["--pxyz"] + [str(v) for v in pxyz] + ["--nxyz"] + [str(v) for v in nxyz] + [
           "--spot_energy", str(energy), "0.0", "--spot_position", "0", "0", str(spot_z), "--spot_size", str(spot_size), str(spot_size),
           "--histories", str(int(histories)), "--phantom_path", phantom, "--output_prefix", out_dir, "--random_seed", str(seed),
           "--gpu_id", "0", "--scorers", scorers, "--batches", str(batches), "--sample_threads", str(min(NCPU, 32))]
    if threads:
        cmd += ["--threads"] + [str(t) for t in threads]
    cmd += [str(x) for x in extra]
    log = open(os.path.join(out_dir, "stdout.log"), "w")
    return dict(proc=subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), log=log, out_dir=out_dir, t0=time.time())


def harness_finish(h):
    rc = h["proc"].wait()
    h["log"].close()
    wall = time.time() - h["t0"]
    if rc != 0:
        raise RuntimeError("reference harness failed rc=%d: %s" % (rc, open(os.path.join(h["out_dir"], "stdout.log")).read()[-2000:]))
    st = {}
    for ln in open(os.path.join(h["out_dir"], "harness_stats.txt")):
        t = ln.split()
        st[t[0]] = float(t[1]) if len(t) == 2 else [float(x) for x in t[1:]]
    st["wall_s"] = wall
    return st


def harness(*a, **kw):
    """one run; a process of the reference that dies (its own undefined behaviour) is run again with
    another seed, twice at most"""
    for attempt in range(3):
        try:
            return harness_finish(harness_start(*a, **kw))
        except RuntimeError as ex:
            if attempt == 2:
                raise
            print("harness failed (%s...), another seed" % str(ex)[:120], flush=True)
            if "seed" in kw:
                kw["seed"] = kw["seed"] + 1000003
            else:
                a = list(a)
                a[7] = a[7] + 1000003
    return None


def water(path, nxyz, slabs=()):
    nx, ny, nz = nxyz
    hu = np.zeros((nz, ny, nx), dtype=np.int16)
    for d0, d1, h in slabs:   # depth [mm] from the entry face, 1 mm slabs on the C1 grid: k = nz - 1 - floor(depth)
        hu[nz - int(d1):nz - int(d0)] = h
    hu.tofile(path)
    return hu


def probe(work, args):
    """Throughput of the reference's own CUDA kernel on C1 (200 MeV, 30 mm spot, water 200x200x350) for the launch
    shapes its CLI offers: default (512 threads x ceil(N/512) blocks, one history per thread) and --threads T B."""
    ph = os.path.join(work, "water.raw")
    water(ph, C1["nxyz"])
    res = []
    n = int(args.probe_histories) if not args.tiny else 2000
    for variant in ("debug", "release"):
        shapes = [None, (256, 1184), (128, 2368), (512, 592), (256, 4736)] if variant == "debug" else [None, (256, 1184)]
        for th in shapes:
            try:
                st = harness(variant, os.path.join(work, "probe"), C1["nxyz"], C1["lxyz"], ph, 200.0, n, 12345, threads=th)
                r = dict(variant=variant, threads=th, histories=st["histories"], kernel_s=st["transport_seconds"],
                         init_threads_s=st["init_threads_seconds"], run_s=st["run_seconds"], wall_s=st["wall_s"],
                         blocks=st["blocks"], threads_per_block=st["threads"],
                         hist_per_s_kernel=st["histories"] / st["transport_seconds"],
                         hist_per_s_run=st["histories"] / st["run_seconds"])
            except Exception as ex:   # a launch shape that does not fit is part of the answer
                r = dict(variant=variant, threads=th, error=str(ex)[-300:])
            print(json.dumps(r), flush=True)
            res.append(r)
    # the unmodified reference executable (tests/mc/phantom/phantom_env.cpp compiled as CUDA): its own "Run done" line
    # prints milliseconds * 1e-4 (mqi_phantom_env.hpp:427), i.e. a tenth of the seconds run() took
    for variant in ("debug", "release"):
        od = os.path.join(work, "pe_" + variant)
        os.makedirs(od, exist_ok=True)
        cmd = [os.path.join(REF, "phantom_env_gpu_" + variant), "--lxyz", "100", "100", "350", "--pxyz", "0", "0", "-175",
               "--nxyz", "200", "200", "350", "--spot_energy", "200", "0", "--spot_position", "0", "0", "0.5", "--spot_size", "30", "30",
               "--histories", str(n), "--phantom_path", ph, "--output_prefix", od, "--random_seed", "12345", "--gpu_id", "0"]
        t0 = time.time()
        try:
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
            wall = time.time() - t0
            run_s = [float(ln.split()[2]) * 10.0 for ln in r.stdout.splitlines() if ln.startswith("Run done")]
            d = np.fromfile(os.path.join(od, "0_water_dE_total.raw"), dtype=np.float64)
            rec = dict(variant=variant, exe="phantom_env_gpu", rc=r.returncode, histories=n, wall_s=wall, run_s=run_s[0] if run_s else None,
                       hist_per_s_run=n / run_s[0] if run_s else None, dose_sum=float(d.sum()), tail=r.stdout[-400:])
        except Exception as ex:
            rec = dict(variant=variant, exe="phantom_env_gpu", error=str(ex)[-300:])
        print(json.dumps(rec), flush=True)
        res.append(rec)
    json.dump(res, open(os.path.join(OUT, "probe_reference_cuda.json"), "w"), indent=1)
    return res


def best_rate(variant):
    p = os.path.join(OUT, "probe_reference_cuda.json")
    if not os.path.exists(p):
        return 2.0e6, None
    ok = [r for r in json.load(open(p)) if r.get("variant") == variant and "error" not in r and "hist_per_s_kernel" in r]
    if not ok:
        return 2.0e6, None
    b = max(ok, key=lambda r: r["hist_per_s_kernel"])
    return b["hist_per_s_kernel"], b["threads"]


def c1(work, args):
    """C1 at high statistics, full 3-D: K independent runs (seeds) of B batches each; mean dose per history on the
    box of voxels above 5 % of the maximum (+ margin), 12-bit quantised, and its standard error from the K runs
    pooled over 4x4 lateral blocks."""
    ph = os.path.join(work, "water.raw")
    water(ph, C1["nxyz"])
    for variant in args.c1_variants.split(","):
        rate, th = best_rate(variant)
        K = 8
        per_batch = 12_500_000 if not args.tiny else 2000
        total = min(args.c1_histories_release if variant == "release" and args.c1_histories_release else args.c1_histories,
                    rate * args.budget_s)
        B = max(1, int(round(total / (K * per_batch))))
        nx, ny, nz = C1["nxyz"]
        s1 = np.zeros(nx * ny * nz)
        s2 = np.zeros(nx * ny * nz)
        kernel_s = 0.0
        t0 = time.time()
        for k in range(K):
            od = os.path.join(work, "c1_%s" % variant)
            st = harness(variant, od, C1["nxyz"], C1["lxyz"], ph, 200.0, per_batch, 1000 + 7919 * k, batches=B, threads=th)
            kernel_s += st["transport_seconds"]
            d = np.fromfile(os.path.join(od, "0_water_dE_total.raw"), dtype=np.float64) / st["histories"]
            s1 += d
            s2 += d * d
        n_hist = K * B * per_batch
        mean = (s1 / K).reshape(nz, ny, nx)
        var = np.maximum(s2 / K - (s1 / K) ** 2, 0.0).reshape(nz, ny, nx) * K / (K - 1) / K   # variance of the mean
        dmax = mean.max()
        zz, yy, xx = np.nonzero(mean > 0.05 * dmax)
        m = 4
        z0, z1 = max(zz.min() - m, 0), min(zz.max() + m + 1, nz)
        y0, y1 = max(yy.min() - m, 0) // 4 * 4, min((yy.max() + m + 4) // 4 * 4, ny)
        x0, x1 = max(xx.min() - m, 0) // 4 * 4, min((xx.max() + m + 4) // 4 * 4, nx)
        crop = mean[z0:z1, y0:y1, x0:x1]
        levels = 4095
        q = np.round(crop / dmax * levels).astype(np.uint16)
        np.save(os.path.join(OUT, "c1_water200_%s_sums.npy" % variant), np.stack([s1, s2]).astype(np.float32)) if args.keep_sums else None
        vc = var[z0:z1, y0:y1, x0:x1]
        vb = vc.reshape(z1 - z0, (y1 - y0) // 4, 4, (x1 - x0) // 4, 4).mean(axis=(2, 4))
        meta = dict(variant=variant, histories=n_hist, runs=K, batches=B, per_batch=per_batch, energy=200.0, spot_size=30.0,
                    nxyz=list(C1["nxyz"]), lxyz=list(C1["lxyz"]), box=[int(z0), int(z1), int(y0), int(y1), int(x0), int(x1)],
                    levels=levels, dmax=float(dmax), kernel_s=kernel_s, wall_s=time.time() - t0, threads=th,
                    hist_per_s_kernel=n_hist / kernel_s, generator="oracle/gen_golden_gpu.py c1",
                    binary="oracle/_ref/ref_harness_gpu_%s (reference CUDA kernel, sm_100a, -maxrregcount=128)" % variant,
                    units="dose per primary history = q / levels * dmax; se_block = standard error of the mean dose, "
                          "root of the run-to-run variance averaged over 4x4 lateral voxel blocks")
        np.savez_compressed(os.path.join(OUT, "c1_water200_%s_3d.npz" % variant), q=q, se_block=np.sqrt(vb).astype(np.float32),
                            idd=mean.sum(axis=(1, 2)), idd_se=np.sqrt(var.sum(axis=(1, 2))), total=np.array(mean.sum()),
                            meta=np.array(json.dumps(meta)))
        print(json.dumps(meta), flush=True)


C2_SLABS = [(50.0, 70.0, 1000), (70.0, 100.0, -741)]


def reduce_runs(work, tag, variant, nxyz, lxyz, ph, energy, per_run, runs, seed, scorers, keep_xz=(), extra=(), spot_size=10.0,
                threads=None):
    nx, ny, nz = nxyz
    acc = {}
    hist = 0
    # the runs are independent processes: started together, they share the GPU and overlap their host phases
    # (table set-up, download, reshape and file output dominate a run of this size)
    hs = [harness_start(variant, os.path.join(work, "%s_%d" % (tag, k)), nxyz, lxyz, ph, energy, per_run, seed + 7919 * k,
                        scorers=scorers, spot_size=spot_size, threads=threads, extra=extra) for k in range(runs)]
    for h in hs:
        st = harness_finish(h)
        od = h["out_dir"]
        hist += st["histories"]
        for f in sorted(os.listdir(od)):
            if not (f[0].isdigit() and f.endswith(".raw")):
                continue
            name = f.split("_", 1)[1][:-4]
            d = np.fromfile(os.path.join(od, f), dtype=np.float64).reshape(nz, ny, nx) / st["histories"]
            a = acc.setdefault(name, dict(idd=[], xz=0.0, total=[]))
            a["idd"].append(d.sum(axis=(1, 2)))
            a["total"].append(d.sum())
            if name in keep_xz:
                a["xz"] = a["xz"] + d.sum(axis=1) / runs
        shutil.rmtree(od, ignore_errors=True)
    out = {}
    for name, a in acc.items():
        idd = np.stack(a["idd"])
        out[name + "_idd"] = idd.mean(axis=0)
        out[name + "_idd_se"] = idd.std(axis=0, ddof=1) / np.sqrt(runs)
        out[name + "_total"] = np.array(np.mean(a["total"]))
        out[name + "_total_se"] = np.array(np.std(a["total"], ddof=1) / np.sqrt(runs))
        if name in keep_xz:
            out[name + "_xz"] = a["xz"].astype(np.float32)
    return out, hist


def c2sweep(work, args):
    """Config C2 in full: bone / lung slabs on the C1 grid, 70 ... 230 MeV in steps of 10, Dose + LETd numerator /
    denominator (three scorers: the reference's Dose is scored twice, quirk B2), release physics."""
    ph = os.path.join(work, "slabs.raw")
    water(ph, C1["nxyz"], C2_SLABS)
    rate, th = best_rate("release")
    runs = 8
    per_run = int(args.c2_histories // runs) if not args.tiny else 500
    out = {}
    t0 = time.time()
    energies = list(range(70, 231, 10)) if not args.tiny else [70, 150]
    for e in energies:
        r, hist = reduce_runs(work, "c2", "release", C1["nxyz"], C1["lxyz"], ph, float(e), per_run, runs, 4242 + e, "dose+letd",
                              keep_xz=("Dose",) if e in (70, 150, 230) else (), threads=th)
        for k, v in r.items():
            out["E%d_%s" % (e, k)] = v
        print("c2sweep E=%d done, %.0f s elapsed" % (e, time.time() - t0), flush=True)
    meta = dict(variant="release", energies=energies, histories=runs * per_run, runs=runs, slab=C2_SLABS, spot_size=10.0,
                nxyz=list(C1["nxyz"]), lxyz=list(C1["lxyz"]), scorers="dose+letd", wall_s=time.time() - t0, threads=th,
                generator="oracle/gen_golden_gpu.py c2sweep", binary="oracle/_ref/ref_harness_gpu_release",
                units="per primary history")
    np.savez_compressed(os.path.join(OUT, "c2_sweep_release.npz"), meta=np.array(json.dumps(meta)), **out)
    print(json.dumps(meta), flush=True)


def dose2(work, args):
    """dose_to_water_square (scorers/mqi_scorer_energy_deposit.hpp:64-77) next to dose_to_water on the slab phantom."""
    ph = os.path.join(work, "slabs.raw")
    water(ph, C1["nxyz"], C2_SLABS)
    out = {}
    for variant in ("release", "debug"):
        r, hist = reduce_runs(work, "d2", variant, C1["nxyz"], C1["lxyz"], ph, 150.0, 500_000 if not args.tiny else 500, 8, 99, "dose+dose2",
                              threads=best_rate(variant)[1])
        for k, v in r.items():
            out["%s_%s" % (variant, k)] = v
    meta = dict(energy=150.0, histories=4_000_000, runs=8, slab=C2_SLABS, spot_size=10.0, nxyz=list(C1["nxyz"]),
                lxyz=list(C1["lxyz"]), scorers="dose+dose2", generator="oracle/gen_golden_gpu.py dose2",
                units="per primary history; Dose2 = sum over steps of (dose of the step)^2")
    np.savez_compressed(os.path.join(OUT, "c2_slabs150_dose2.npz"), meta=np.array(json.dumps(meta)), **out)
    print(json.dumps(meta), flush=True)


def stat(work, args):
    """The stopping criterion on the reference's own CUDA kernels: transport_particles_patient_stat fills Dose and the
    stat pair (sum d, sum d^2), calculate_standard_deviation (kernel_functions/mqi_variables.hpp:20-48) turns them into
    float sigma / mean arrays, the host part of calculate_stat (mqi_tps_env.hpp:1409-1425) into one number."""
    nxyz, lxyz = (40, 40, 100), (100.0, 100.0, 150.0)
    ph = os.path.join(work, "w40.raw")
    water(ph, nxyz)
    out = {}
    for tag, thr, n in (("a", 0.5, 400_000), ("b", 0.1, 1_500_000)) if not args.tiny else (("a", 0.5, 3000),):
        od = os.path.join(work, "stat_" + tag)
        st = harness("release", od, nxyz, lxyz, ph, 100.0, n, 31 + len(tag), scorers="stat", spot_size=20.0,
                     extra=["--stat_threshold", thr])
        out[tag + "_sum"] = np.fromfile(os.path.join(od, "0_StatDose.raw"), dtype=np.float64)
        out[tag + "_sumsq"] = np.fromfile(os.path.join(od, "0_StatDose2.raw"), dtype=np.float64)
        out[tag + "_dose"] = np.fromfile(os.path.join(od, "0_Dose.raw"), dtype=np.float64)
        if os.path.exists(os.path.join(od, "stat_sd.raw")):   # CUDA build only
            out[tag + "_sd"] = np.fromfile(os.path.join(od, "stat_sd.raw"), dtype=np.float32)
            out[tag + "_mean"] = np.fromfile(os.path.join(od, "stat_mean.raw"), dtype=np.float32)
            out[tag + "_value"] = np.array(st["stat_value"])
            out[tag + "_count"] = np.array(st["stat_count"])
            out[tag + "_dose_max"] = np.array(st["stat_dose_max"])
        out[tag + "_n"] = np.array(st["tracked"])
        out[tag + "_threshold"] = np.array(thr)
        print("stat", tag, st, flush=True)
    meta = dict(nxyz=list(nxyz), lxyz=list(lxyz), energy=100.0, spot_size=20.0, variant="release",
                generator="oracle/gen_golden_gpu.py stat", binary="oracle/_ref/ref_harness_gpu_release --scorers stat")
    np.savez_compressed(os.path.join(OUT, "a15_stat_release.npz"), meta=np.array(json.dumps(meta)), **out)


C3L = dict(nxyz=(128, 128, 80), spacing=(1.0, 1.0, 2.5), angles=(10.0, 30.0, 0.0, 0.0), grid=(5, 4, 6.0), e0=90.0, de=2.0,
           gauss=(3.0, 3.0, 0.003, 0.003, 0.8), spot_z=140.0, per_spot=400_000, seed=2024)


def c3like(work, args):
    """A C3/C4-like case the reference can run: heterogeneous synthetic head CT (skull shell, air cavities) of
    128 x 128 x 80 voxels, 20 gaussian pbs beamlets (5 x 4 grid, 90 ... 128 MeV) from gantry 30 / collimator 10 degrees,
    release physics: dense Dose (full 3-D) and the rows of the Dij matrix."""
    sys.path.insert(0, ROOT)
    from moquimc_b200 import synthetic as S
    c = dict(C3L)
    if args.tiny:
        c["per_spot"] = 400
    hu, origin = S.head_ct(c["nxyz"], c["spacing"], seed=7)
    ph = os.path.join(work, "head.raw")
    hu.astype(np.int16).tofile(ph)
    nx, ny, nz = c["nxyz"]
    lxyz = (nx * c["spacing"][0], ny * c["spacing"][1], nz * c["spacing"][2])
    ns = c["grid"][0] * c["grid"][1]
    extra = ["--spot_angles"] + list(c["angles"]) + ["--spot_grid", c["grid"][0], c["grid"][1], c["grid"][2],
                                                     "--energy_step", c["de"], "--gauss"] + list(c["gauss"])
    common = dict(variant="release", nxyz=c["nxyz"], lxyz=lxyz, phantom=ph, energy=c["e0"], spot_size=0.0, spot_z=c["spot_z"],
                  pxyz=(0.0, 0.0, 0.0))
    runs = 8
    doses = []
    for k in range(runs):
        od = os.path.join(work, "c3d")
        st = harness(out_dir=od, histories=ns * c["per_spot"] // runs, seed=c["seed"] + 7919 * k, scorers="dose", extra=extra, **common)
        doses.append(np.fromfile(os.path.join(od, "0_water_dE_total.raw"), dtype=np.float64).reshape(nz, ny, nx) / st["histories"])
    doses = np.stack(doses)
    mean = doses.mean(axis=0)
    se = doses.std(axis=0, ddof=1) / np.sqrt(runs)
    od = os.path.join(work, "c3j")
    st = harness(out_dir=od, histories=ns * c["per_spot"] // 4, seed=c["seed"] + 1, scorers="dij", extra=extra, **common)
    k1 = np.fromfile(os.path.join(od, "dij_key1.raw"), dtype=np.uint32)
    k2 = np.fromfile(os.path.join(od, "dij_key2.raw"), dtype=np.uint32)
    v = np.fromfile(os.path.join(od, "dij_value.raw"), dtype=np.float64) / (st["histories"] / ns)
    rows = np.zeros((ns, nz, ny, nx), dtype=np.float32)
    np.add.at(rows.reshape(ns, -1), (k2, k1), v)
    meta = dict(c, hu_seed=7, rot=st.get("rot"), lxyz=list(lxyz), histories_dose=ns * c["per_spot"] // runs * runs, runs=runs,
                histories_dij=int(st["histories"]), dij_nnz=int(k1.size), generator="oracle/gen_golden_gpu.py c3like",
                binary="oracle/_ref/ref_harness_gpu_release", units="dose per primary history; dij rows per history of the spot")
    np.savez_compressed(os.path.join(OUT, "c3like_head_release.npz"), dose=mean.astype(np.float32), dose_se=se.astype(np.float32),
                        dij_row_total=rows.reshape(ns, -1).sum(axis=1), dij_row_idd=rows.sum(axis=(2, 3)),
                        dij_row_xy=rows.sum(axis=1), dij_nnz_per_row=np.bincount(k2, minlength=ns),
                        dij_rows_full=rows[[0, 9, 19]], hu=hu.astype(np.int16), meta=np.array(json.dumps(meta)))
    print(json.dumps(meta), flush=True)


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="+", choices=["probe", "c1", "c2sweep", "dose2", "stat", "c3like"])
    ap.add_argument("--budget-s", type=float, default=120.0, help="kernel seconds the C1 golden may take per variant")
    ap.add_argument("--c1-histories", type=float, default=1.0e9)
    ap.add_argument("--c1-histories-release", type=float, default=0.0)
    ap.add_argument("--c1-variants", default="debug,release")
    ap.add_argument("--c2-histories", type=float, default=4.0e6)
    ap.add_argument("--probe-histories", type=float, default=1.0e7)
    ap.add_argument("--keep-sums", action="store_true")
    ap.add_argument("--tiny", action="store_true", help="dry run of the script itself with a few thousand histories")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    work = tempfile.mkdtemp(prefix="mqi_gold_")
    try:
        for w in args.what:
            t0 = time.time()
            try:
                globals()[w](work, args)
            except Exception as ex:   # keep going: every target writes its own file
                print("TARGET %s FAILED: %s" % (w, str(ex)[-1500:]), flush=True)
            print("== %s: %.0f s" % (w, time.time() - t0), flush=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
