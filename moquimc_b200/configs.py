"""The BASELINE.json configurations beside the headline one, as short fixed-size workloads through the C ABI
(bench.py's `configs` block, scripts/config_bench.py).  Synthetic inputs, seeds fixed:

  C2  water / bone / lung slabs on the C1 grid, 150 MeV, 10 mm spot, Dose + LETd (3 dense scorers), release physics
  C3  synthetic head-and-neck CT 512 x 512 x 200, ~2 000-spot PBS plan through tps_env (the reference-facing CLI),
      Dose + the two stat scorers, one pass of the stopping loop
  C4  Dij: 5 000 gaussian spots x 1e4 histories on a 256 x 256 x 150 CT, sparse hash scoring, at the reference's
      table size (393 216 001 slots, mqi_tps_env.hpp:922) and at a table sized from the free HBM
  RS  range shifter + voxelised aperture in front of a water phantom (multi-node world), Dose

Every leg returns {"value": histories/s from the kernel's CUDA-event time, "kernel_ms", "histories", ...} plus the
algorithmic bytes per scored step of its scorer set, from which bench.py forms a per-config roofline.
"""
import os
import re
import subprocess
import tempfile

import numpy as np

from . import capi, synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TPS_ENV = os.path.join(ROOT, "moquimc_b200", "bin", "tps_env")


def slab_hu():
    hu = np.zeros((350, 200, 200), dtype=np.int16)
    hu[350 - 70:350 - 50] = 1000     # bone 50-70 mm
    hu[350 - 100:350 - 70] = -741    # lung 70-100 mm
    return hu


def steps_per_history(make_engine, n=200_000):
    """scored voxel steps per primary history of a workload: the kernel's own counter (option count_steps, which runs
    the general kernel; tests/test_gpu_parity.py holds it to the oracle's count within 0.5 %)"""
    e, run = make_engine()
    e.set_option("count_steps", 1)
    st = run(e, n)
    e.close()
    return st.steps / max(1, st.histories)


def _timed(e, run, n, reps):
    run(e, min(n, 200_000))            # warm-up: module load, table upload
    ms, hist = [], 0
    for _ in range(reps):
        e.clear_scorers()
        st = run(e, n)
        ms.append(st.kernel_ms)
        hist = st.histories
    return hist, sum(ms) / len(ms), st


def c2(device=0, histories=10_000_000, reps=2, energy=150.0):
    xe, ye, ze = capi.uniform_edges(-50, 50, 200), capi.uniform_edges(-50, 50, 200), capi.uniform_edges(-350, 0, 350)

    def make():
        e = capi.Engine(device, physics=capi.PHYSICS_RELEASE)
        e.set_grid_hu(xe, ye, ze, slab_hu())
        for k, nm in ((capi.SCORER_DOSE, "Dose"), (capi.SCORER_LETD_NUMER, "LETd_numer"), (capi.SCORER_LETD_DENOM, "LETd_denom")):
            e.add_scorer(k, nm)
        e.set_beamlets([capi.make_beamlet(energy, [0, 0, 0.5, 0, 0, -1], [10, 10, 0, 0, 0, 0], uniform=True)], [4 * histories])
        return e, (lambda eng, n: eng.run(7, 0, n))
    sph = steps_per_history(make)
    e, run = make()
    hist, ms, st = _timed(e, run, histories, reps)
    e.close()
    return {"workload": "C2: bone/lung slabs on the C1 grid, %g MeV, Dose + LETd (3 fp64 scorers), release physics" % energy,
            "histories": hist, "kernel_ms": ms, "value": hist / (ms * 1e-3), "steps_per_history": sph, "bytes_per_step": 4 + 3 * 16,
            "scorers": 3}


def c4_setup(device, capacity, n_spots=5000, per=10_000):
    n, sp = (256, 256, 150), (1.5, 1.5, 2.0)
    hu, origin = S.head_ct(n, sp, seed=4)
    edges = [(np.float32(origin[a] - sp[a] / 2) + np.arange(n[a] + 1, dtype=np.float32) * np.float32(sp[a])).astype(np.float32) for a in range(3)]
    rng = np.random.default_rng(5)
    g = np.arange(-30.0, 30.0 + 1e-6, 4.0)
    pos = [(x, z) for x in g for z in g if x * x + z * z <= 30.0 ** 2 + 1e-6]
    R = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float32)   # beam along +y
    bl = []
    for e_mev in np.linspace(80.0, 160.0, 20):
        for i in rng.choice(len(pos), size=n_spots // 20, replace=True):
            x, z = pos[i]
            bl.append(capi.make_beamlet(float(e_mev), [x, z, 250.0, 0, 0, -1], [3.0, 3.0, 0.0, 0.003, 0.003, 0.0], uniform=False,
                                        sigma_energy=0.6, rot=R))
    e = capi.Engine(device, physics=capi.PHYSICS_RELEASE)
    e.set_grid_hu(edges[0], edges[1], edges[2], hu)
    s = e.add_scorer(capi.SCORER_DIJ, "Dij", capacity=int(capacity) | 1)
    e.set_beamlets(bl, [per] * len(bl))
    return e, s, len(bl) * per


def c4(device=0, capacity=393_216_001, reps=1, n_spots=5000, per=10_000):
    # steps per history from a small run of the same source with the step counter (general kernel) switched on
    e, s, total = c4_setup(device, 4_000_001, n_spots, per)
    e.set_option("count_steps", 1)
    st = e.run(77, 0, 400_000, per_spot=True)
    sph = st.steps / max(1, st.histories)
    e.close()
    e, s, total = c4_setup(device, capacity, n_spots, per)
    run = lambda eng, n: eng.run(77, 0, n, per_spot=True)   # noqa: E731
    hist, ms, st = _timed(e, run, total, reps)
    nnz = e.get_sparse_count(s)
    full = st.dij_table_full
    e.close()
    return {"workload": "C4: Dij, %d gaussian spots x %d histories on a 256x256x150 CT, table of %d slots" % (n_spots, per, int(capacity) | 1),
            "histories": hist, "kernel_ms": ms, "value": hist / (ms * 1e-3), "nnz": int(nnz), "load_factor": nnz / float(int(capacity) | 1),
            "table_full": int(full), "steps_per_history": sph, "bytes_per_step": 4 + 16, "scorers": 1}


def rs_aperture(device=0, histories=8_000_000, reps=2, nodes=2):
    NX, NY, NZ = 100, 100, 200

    def make():
        e = capi.Engine(device, physics=capi.PHYSICS_RELEASE)
        e.set_grid_hu(capi.uniform_edges(-50, 50, NX), capi.uniform_edges(-50, 50, NY), capi.uniform_edges(-200, 0, NZ),
                      np.zeros((NZ, NY, NX), np.int16))
        if nodes >= 1:   # 40 mm range shifter slab (1.19 g/cm3), create_rangeshifter style: one voxel
            e.add_beamline_node(np.float32([-150, 150]), np.float32([-150, 150]), np.float32([100, 140]), np.float32([1.19e-3]))
        if nodes >= 2:   # 20 mm block voxelised at 1 mm, 36 x 36 mm opening
            axe, aze = capi.uniform_edges(-40, 40, 80), capi.uniform_edges(40, 60, 20)
            xc = 0.5 * (axe[1:] + axe[:-1])
            op = (np.abs(xc)[:, None] < 18.0) & (np.abs(xc)[None, :] < 18.0)
            e.add_beamline_node(axe, axe, aze, np.broadcast_to(np.where(op, np.float32(1e-8), np.float32(100.0)), (20, 80, 80)).astype(np.float32).copy())
        e.add_scorer(capi.SCORER_DOSE, "Dose")
        e.set_beamlets([capi.make_beamlet(180.0, [0, 0, 180.0, 0, 0, -1], [15, 15, 0, 0, 0, 0], uniform=True)], [4 * histories])
        return e, (lambda eng, n: eng.run(1, 0, n))
    sph = steps_per_history(make)
    e, run = make()
    hist, ms, st = _timed(e, run, histories, reps)
    e.close()
    return {"workload": "range shifter (40 mm) + voxelised aperture in front of a 100x100x200 water phantom, 180 MeV, Dose, release physics",
            "histories": hist, "kernel_ms": ms, "value": hist / (ms * 1e-3), "steps_per_history": sph, "bytes_per_step": 20, "scorers": 1,
            "nodes": nodes + 1}


def c3_case(root, n=(512, 512, 200), spacing=(1.0, 1.0, 2.5)):
    """the synthetic C3 inputs (CT .mha, beam model, ~2 000-spot plan) under `root`"""
    if not os.path.exists(os.path.join(root, "ct.mha")):
        S.make_case(root, n=n, spacing=spacing, n_layers=25, pitch=5.0, half_width=25.0, ParticlesPerHistory=400.0)
    return root


def run_tps(inp, timeout=1500):
    r = subprocess.run([TPS_ENV, inp], capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("tps_env failed: " + r.stdout[-1500:] + r.stderr[-1500:])
    return r.stdout


def parse_tps(out):
    m = re.findall(r"Transport kernels ([0-9.eE+-]+) ms on (\d+) GPU\(s\): ([0-9.eE+-]+) histories/s", out)
    runs = [(int(a), float(b)) for a, b in re.findall(r"Run (\d+): current uncertainty ([0-9.eE+-]+) %", out)]
    tracked = [int(x) for x in re.findall(r"Number of particles tracked (\d+)", out)]
    st = re.findall(r"Stopping criterion: (\d+) evaluations ([0-9.eE+-]+) ms .* final gather of the dense scorers ([0-9.eE+-]+) ms", out)
    return {"kernel_ms": float(m[-1][0]) if m else None, "gpus": int(m[-1][1]) if m else None, "value": float(m[-1][2]) if m else None,
            "passes": runs, "histories": tracked[-1] if tracked else 0,
            "stat_ms": float(st[-1][1]) if st else None, "gather_ms": float(st[-1][2]) if st else None}


def c3(device_ids=(0,), root=None, passes=1, criteria=1.0, pph=400.0):
    """C3 through tps_env: `passes` = 1 times one pass of the stopping loop (the bench leg); passes = None runs the
    loop to the criterion (strong-scaling record)."""
    own = root is None
    root = root or tempfile.mkdtemp(prefix="mqi_c3_")
    c3_case(root)
    od = os.path.join(root, "o_c3_%d" % len(device_ids))
    inp = os.path.join(root, "c3_%d.in" % len(device_ids))
    S.write_input(inp, root, od, ParticlesPerHistory=pph, StoppingStatistics="true", StoppingCriteria=criteria, StatThreshold=0.5,
                  MaxStatPasses=passes if passes else 400, GPUID=",".join(str(d) for d in device_ids), OutputFormat="raw")
    import time
    t0 = time.time()
    out = run_tps(inp)
    wall = time.time() - t0
    r = parse_tps(out)
    r.update({"workload": "C3: synthetic head-and-neck CT 512x512x200, ~2 000-spot PBS plan through tps_env, Dose + the two stat "
                          "scorers (3 fp64 grids), release physics, %s" % ("%d pass(es) of the stopping loop" % passes if passes
                                                                           else "stopping loop run to %g %%" % criteria),
              "wall_s": wall, "bytes_per_step": 4 + 3 * 16, "scorers": 3})
    if own:
        import shutil
        shutil.rmtree(root, ignore_errors=True)
    return r
