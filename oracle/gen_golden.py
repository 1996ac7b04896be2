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
        fmt = {}
        for f in sorted(os.listdir(d)):
            name, ext = f.rsplit(".", 1)
            if name.startswith("fmt_"):   # files written by the reference's own output writers: kept as bytes
                fmt[f.replace(".", "_")] = np.fromfile(os.path.join(d, f), dtype=np.uint8)
                continue
            out[name] = np.fromfile(os.path.join(d, f), dtype=DT[ext])
        if variant == "release":
            np.savez_compressed(os.path.join(GOLD, "fmt_writers.npz"), **fmt)
            print("fmt", {k: v.shape for k, v in fmt.items()})
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


def c2_scorers(energy, scorers, tag, procs=8, per_proc=50000, seed=777):
    import argparse
    sys.path.insert(0, HERE)
    import ref_run
    a = argparse.Namespace(variant="release", procs=procs, histories_per_proc=per_proc, energy=float(energy), spot_size=10.0,
                           nxyz=[200, 200, 350], lxyz=[100.0, 100.0, 350.0], slab=[[50.0, 70.0, 1000.0], [70.0, 100.0, -741.0]],
                           seed=seed, rebin=8, harness=True, scorers=scorers, gauss=None, out=None)
    res, meta = ref_run.run(a)
    keep = {k: v for k, v in res.items() if k == "meta" or k.endswith(("_idd", "_idd_se", "_total", "_total_se"))}
    np.savez_compressed(os.path.join(GOLD, "c2_slabs%d_%s_release.npz" % (int(energy), tag)), **keep)
    print("c2", scorers, energy, {k: getattr(v, "shape", None) for k, v in keep.items()})


G1_GAUSS = [4.0, 3.0, 0.004, 0.003, 1.5]     # sigma x, y [mm], x', y' [rad], energy [MeV]: the pbs beamlet of configs 3-5


def g1_gauss(procs=8, per_proc=50000):
    """A gaussian pencil beam (phsp_6d + norm_1d, the beamlet treatment_machine_pbs builds per spot) of 150 MeV into
    the C1 water phantom through oracle/ref_harness.cpp --gauss: pins the source sampling at transport level."""
    import argparse
    sys.path.insert(0, HERE)
    import ref_run
    a = argparse.Namespace(variant="release", procs=procs, histories_per_proc=per_proc, energy=150.0, spot_size=0.0,
                           nxyz=[200, 200, 350], lxyz=[100.0, 100.0, 350.0], slab=[], seed=4711, rebin=8, harness=True,
                           scorers="dose", gauss=G1_GAUSS, out=None)
    res, meta = ref_run.run(a)
    keep = {k: v for k, v in res.items() if k == "meta" or k.endswith(("_idd", "_idd_se", "_total", "_total_se", "_xz", "_yz"))}
    np.savez_compressed(os.path.join(GOLD, "g1_gauss150_release.npz"), **keep)
    print("g1", {k: getattr(v, "shape", None) for k, v in keep.items()})


B1_ARGS = dict(pph=25000.0, snout=250.0, collimator=15.0, gantry=45.0, couch=10.0, iso=(1.5, -2.5, 40.0),
               n_layers=4, pitch=10.0, half_width=20.0, seed=1)


B1_BLOCK = [(-10.0, -8.0), (10.0, -8.0), (10.0, 8.0), (-10.0, 8.0)]
B1_BEAMLINES = {
    "rs_id_block": ["--rs-id", "RS1", "--block", "20", "60"] + ["%g" % v for p in B1_BLOCK for v in p],
    "rs_two_ids": ["--rs-id", "RS1", "RS1"],
    "rs_wet": ["--rs-wet", "46", "120"],
}


def b1_beam_model():
    """The beam model at work: the reference's own mqi::pbs (beam data file, spot -> beamlet, histories per spot),
    treatment_machine_ion::create_beamsource / create_coordinate_transform and beam_module_ion, unmodified, on the
    synthetic machine file and a 52-spot plan built in memory (oracle/ref_tps_kat.cpp over oracle/ref_dataset_stub.hpp)."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from moquimc_b200 import synthetic as S
    a = B1_ARGS
    with tempfile.TemporaryDirectory() as d:
        S.write_beam_model(os.path.join(d, "machine.txt"))
        spots = S.spot_list(n_layers=a["n_layers"], pitch=a["pitch"], half_width=a["half_width"], seed=a["seed"])
        with open(os.path.join(d, "spots.txt"), "w") as f:
            f.write("".join("%.6g %.6g %.6g %.8g\n" % s for s in spots))   # as synthetic.write_plan prints them
        out = subprocess.check_output([os.path.join(HERE, "_ref", "ref_tps_kat"), os.path.join(d, "machine.txt"),
                                       os.path.join(d, "spots.txt"), "%g" % a["pph"], "%g" % a["snout"], "%g" % a["collimator"],
                                       "%g" % a["gantry"], "%g" % a["couch"]] + ["%g" % v for v in a["iso"]]).decode()
        lines = [ln for ln in out.splitlines() if ln.startswith(("angles", "trans", "spot "))]
        assert len(lines) == len(spots) + 2
        # beamline devices through create_beamline (characterize_rangeshifter / characterize_aperture): range shifter by
        # ID + one block; two range shifter IDs; range shifter by water-equivalent thickness (a machine file without
        # the [rangeshifter_thickness] table selects that branch)
        base = [os.path.join(HERE, "_ref", "ref_tps_kat"), os.path.join(d, "machine.txt"), os.path.join(d, "spots.txt"), "%g" % a["pph"],
                "%g" % a["snout"], "%g" % a["collimator"], "%g" % a["gantry"], "%g" % a["couch"]] + ["%g" % v for v in a["iso"]]
        geo = {}
        for key, extra in B1_BEAMLINES.items():
            cmd = list(base)
            if key == "rs_wet":
                text = open(os.path.join(d, "machine.txt")).read()
                i, j = text.index("[rangeshifter_thickness]"), text.index("[spot]")
                with open(os.path.join(d, "machine_nors.txt"), "w") as f:
                    f.write(text[:i] + text[j:])
                cmd[1] = os.path.join(d, "machine_nors.txt")
            o = subprocess.check_output(cmd + extra, stderr=subprocess.DEVNULL).decode()
            geo[key] = "\n".join(ln for ln in o.splitlines() if ln.startswith("geo "))
            print("b1", key, geo[key].replace("\n", " | "))
    np.savez_compressed(os.path.join(GOLD, "b1_beam_model.npz"), text=np.frombuffer("\n".join(lines).encode(), dtype=np.uint8),
                        meta=np.array(str(a)), **{"geo_" + k: np.frombuffer(v.encode(), dtype=np.uint8) for k, v in geo.items()})
    print("b1", len(lines), lines[0], lines[1])


D1_SPOTS, D1_PITCH, D1_ENERGY, D1_SPOT = 3, 15.0, 120.0, 3.0


def d1_dij(procs=8, per_proc=30000):
    """Dij at transport level: three 120 MeV spots 15 mm apart in the C1 water phantom through oracle/ref_harness.cpp
    --scorers dij (the reference's transport_particles_patient with its scorer_offset_vector = spot of each history,
    insert_hashtable keyed by (voxel, spot)); the occupied slots are reduced to one depth profile and one lateral
    centroid per spot."""
    import time
    nx, ny, nz = 200, 200, 350
    exe = os.path.join(HERE, "_ref", "ref_harness_release")
    idd = np.zeros((D1_SPOTS, nz))
    cx = np.zeros(D1_SPOTS)
    tot = np.zeros((procs, D1_SPOTS))
    with tempfile.TemporaryDirectory() as work:
        ph = os.path.join(work, "phantom.raw")
        np.zeros(nx * ny * nz, dtype=np.int16).tofile(ph)
        runs = []
        for p in range(procs):
            od = os.path.join(work, "o%d" % p)
            os.makedirs(od)
            cmd = [exe, "--lxyz", "100", "100", "350", "--pxyz", "0.0", "0.0", "-175.0", "--nxyz", str(nx), str(ny), str(nz),
                   "--spot_energy", str(D1_ENERGY), "0.0", "--spot_position", "0", "0", "0.5", "--spot_size", str(D1_SPOT), str(D1_SPOT),
                   "--histories", str(per_proc), "--phantom_path", ph, "--output_prefix", od, "--random_seed", str(5150 + 7919 * p),
                   "--gpu_id", "0", "--scorers", "dij", "--nspots", str(D1_SPOTS), "--spot_pitch", str(D1_PITCH)]
            runs.append((subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT), od))
        for p, (pr, od) in enumerate(runs):
            assert pr.wait() == 0
            k1 = np.fromfile(os.path.join(od, "dij_key1.raw"), dtype=np.uint32).astype(np.int64)
            k2 = np.fromfile(os.path.join(od, "dij_key2.raw"), dtype=np.uint32).astype(np.int64)
            v = np.fromfile(os.path.join(od, "dij_value.raw"), dtype=np.float64)
            assert k2.max() == D1_SPOTS - 1
            z, x = k1 // (nx * ny), k1 % nx
            np.add.at(idd, (k2, z), v)
            np.add.at(cx, k2, v * ((x + 0.5) * 0.5 - 50.0))
            np.add.at(tot[p], k2, v)
    per_spot = procs * (per_proc // D1_SPOTS)
    total = tot.sum(axis=0)
    out = {"dij_idd": idd / per_spot, "dij_centroid_x": cx / total, "dij_total": total / per_spot,
           "dij_total_se": tot.std(axis=0, ddof=1) * np.sqrt(procs) / per_spot,
           "meta": np.array(str({"spots": D1_SPOTS, "pitch": D1_PITCH, "energy": D1_ENERGY, "spot_size": D1_SPOT,
                                 "histories_per_spot": per_spot, "variant": "release"}))}
    np.savez_compressed(os.path.join(GOLD, "d1_dij3_release.npz"), **out)
    print("d1", {k: getattr(v, "shape", None) for k, v in out.items()}, out["dij_centroid_x"], out["dij_total"])


def c2_debug(energy=150, procs=8, per_proc=50000):
    """The C2 slab phantom, Dose only, with the DEBUG physics variant (-D__PHYSICS_DEBUG__: the water shortcut of
    spr_default, zero-energy delta daughters, recoil daughters) through the reference's own phantom_env."""
    import argparse
    sys.path.insert(0, HERE)
    import ref_run
    a = argparse.Namespace(variant="debug", procs=procs, histories_per_proc=per_proc, energy=float(energy), spot_size=10.0,
                           nxyz=[200, 200, 350], lxyz=[100.0, 100.0, 350.0], slab=[[50.0, 70.0, 1000.0], [70.0, 100.0, -741.0]],
                           seed=9001, rebin=8, harness=False, scorers="dose", gauss=None, out=None)
    res, meta = ref_run.run(a)
    keep = {k: v for k, v in res.items() if k == "meta" or k.endswith(("_idd", "_idd_se", "_total", "_total_se"))}
    np.savez_compressed(os.path.join(GOLD, "c2_slabs%d_debug.npz" % int(energy)), **keep)
    print("c2_debug", energy, {k: getattr(v, "shape", None) for k, v in keep.items()})


def c2_lett(energy=150):
    """The C2 slab phantom with the track-averaged LET scorers (scorers/mqi_scorer_energy_deposit.hpp:141-177:
    numerator = step length x LET, denominator = step length) through oracle/ref_harness.cpp --scorers lett."""
    c2_scorers(energy, "lett", "lett", seed=778)


def c2_edep(energy=150, procs=8, per_proc=50000):
    """The C2 slab phantom with the EnergyDeposition scorer alone (scorers/mqi_scorer_energy_deposit.hpp:14-22: dE +
    local dE in MeV per step, no division by the voxel mass) through oracle/ref_harness.cpp --scorers edep."""
    import argparse
    sys.path.insert(0, HERE)
    import ref_run
    a = argparse.Namespace(variant="release", procs=procs, histories_per_proc=per_proc, energy=float(energy), spot_size=10.0,
                           nxyz=[200, 200, 350], lxyz=[100.0, 100.0, 350.0], slab=[[50.0, 70.0, 1000.0], [70.0, 100.0, -741.0]],
                           seed=777, rebin=8, harness=True, scorers="edep", gauss=None, out=None)
    res, meta = ref_run.run(a)
    keep = {k: v for k, v in res.items() if k == "meta" or k.endswith(("_idd", "_idd_se", "_total", "_total_se"))}
    np.savez_compressed(os.path.join(GOLD, "c2_slabs%d_edep_release.npz" % int(energy)), **keep)
    print("c2_edep", energy, {k: getattr(v, "shape", None) for k, v in keep.items()})


F3_GRID = dict(nxyz=[100, 100, 200], lxyz=[100.0, 100.0, 200.0])
F3_RS = [100.0, 140.0, 150.0, 1.19]          # zlo zhi half-width density[g/cm3]
F3_AP = [40.0, 60.0, 40.0, 12.0, 8.0]        # zlo zhi half-width open_hx open_hy
F3_ANGLE_Z = 30.0                            # the beamline frame is rotated about z and shifted
F3_TRANS = [2.0, -3.0, 0.0]


def f3_frame():
    a = np.deg2rad(F3_ANGLE_Z)
    c, s = float(np.float32(np.cos(a))), float(np.float32(np.sin(a)))
    return [c, -s, 0.0, s, c, 0.0, 0.0, 0.0, 1.0] + F3_TRANS


def f4_mask_total(nxyz=(100, 100, 200)):
    """Sum of two overlapping 0/1 masks (tests/test_gpu_beamline_roi.py::make_mask_total)."""
    nx, ny, nz = nxyz
    m1 = np.zeros((nz, ny, nx), dtype=np.uint8)
    m1[60:190, 30:70, 35:65] = 1
    m2 = np.zeros((nz, ny, nx), dtype=np.uint8)
    m2[100:120, 40:60, 35:50] = 1
    return m1 + m2


def f3_beamline(procs=8, per_proc=150000):
    """range shifter slab + voxelised aperture as beamline children in front of a water
    phantom (100 x 100 x 200 mm, 1 mm voxels), 150 MeV, 20 mm uniform square spot starting at z = 180,
    beamline frame rotated 30 deg about z and shifted by (2, -3, 0); release physics, Dose."""
    import argparse
    sys.path.insert(0, HERE)
    import ref_run
    extra = ["--rangeshifter"] + F3_RS + ["--aperture"] + F3_AP + ["--frame"] + f3_frame()
    a = argparse.Namespace(variant="release", procs=procs, histories_per_proc=per_proc, energy=150.0, spot_size=20.0,
                           slab=[], seed=777, rebin=4, harness=True, scorers="dose", gauss=None, out=None,
                           spot_z=180.0, extra=extra, **F3_GRID)
    res, meta = ref_run.run(a)
    keep = {k: v for k, v in res.items() if k == "meta" or k.endswith(("_idd", "_idd_se", "_total", "_total_se", "_xy", "_xy_se", "_xz"))}
    keep = {k: (v.astype(np.float32) if getattr(v, "ndim", 0) > 1 else v) for k, v in keep.items()}
    np.savez_compressed(os.path.join(GOLD, "f3_beamline_release.npz"), **keep)
    print("f3", {k: getattr(v, "shape", None) for k, v in keep.items()})


def f4_roi(procs=8, per_proc=100000):
    """CONTOUR roi (mask_reader::mask_to_roi) on the Dose scorer; two overlapping masks."""
    import argparse
    sys.path.insert(0, HERE)
    import ref_run
    with tempfile.TemporaryDirectory() as d:
        mpath = os.path.join(d, "mask_total.raw")
        f4_mask_total().tofile(mpath)
        a = argparse.Namespace(variant="release", procs=procs, histories_per_proc=per_proc, energy=150.0, spot_size=25.0,
                               slab=[], seed=991, rebin=4, harness=True, scorers="dose", gauss=None, out=None,
                               spot_z=0.5, extra=["--roi_mask", mpath], **F3_GRID)
        res, meta = ref_run.run(a)
    keep = {k: v for k, v in res.items() if k == "meta" or k.endswith(("_idd", "_idd_se", "_total", "_total_se", "_xy", "_xz", "_yz"))}
    keep = {k: (v.astype(np.float32) if getattr(v, "ndim", 0) > 1 else v) for k, v in keep.items()}
    np.savez_compressed(os.path.join(GOLD, "f4_roi_release.npz"), **keep)
    print("f4", {k: getattr(v, "shape", None) for k, v in keep.items()})


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "f3":
        f3_beamline()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "f4":
        f4_roi()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "b1":
        b1_beam_model()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "d1":
        d1_dij()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c2_debug":
        c2_debug()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "g1":
        g1_gauss()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c2_lett":
        c2_lett()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c2_edep":
        c2_edep()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c2":
        for e in (70, 150, 230):
            c2(e)
        sys.exit(0)
    for v in ("debug", "release"):
        kat(v)
