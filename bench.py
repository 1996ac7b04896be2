#!/usr/bin/env python
"""bench.py -- primary proton histories/s of the per-history transport hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json north_star target case): config C1 -- 200 MeV protons, 30 mm uniform square
spot, water phantom 100x100x350 mm on a 200x200x350 grid (HU 0), phantom_env physics
(__PHYSICS_DEBUG__), one Dose (dose-to-water) scorer accumulated in fp64.  One step = one pass of
the hot path over one batch of --histories primaries per GPU (default 1e7): device beam source ->
transport -> scoring, plus (N > 1) one NCCL reduce of the per-GPU dose grids to rank 0.  Histories
are sharded by index range, no other collective: scaling is weak (per-GPU work fixed).

  value  whole-job histories/s with the HU volume, beam model and dose grid resident in HBM,
         timed with CUDA events on the launching stream, max over ranks;
  e2e    the same through the reference-facing C ABI with HOST buffers: every step uploads the int16
         HU volume (pinned) and the beam model, transports, and downloads the fp64 dose grid;
  roofline  HBM bound of the transport kernel: algorithmic bytes = 20 B per scored step (4 B density
         read + 8 B + 8 B fp64 dose read-modify-write) x 446.3 oracle-measured steps per 200 MeV
         history = 8 926 B per history, against MEASURED_PEAKS.json;
  cpu_baseline  the reference's own CPU phantom_env (oracle/_ref, built from /root/reference with
         the two documented patches) on all host cores, bounded sample, rank 0, N = 1.

  configs   (N = 1) short fixed-size legs of the other BASELINE configs after the headline: C2 (slabs, Dose + LETd),
         C3 (head-and-neck CT through tps_env, Dose + the stat pair, one pass), C4 (Dij at the reference's table
         size and at a table sized from the free HBM), range shifter + aperture; each with its kernel time and a
         per-config roofline from its own steps/history (moquimc_b200/configs.py);
  gpu_reference_baseline  (N = 1) the reference's own CUDA kernel compiled for sm_100a (oracle/_ref/ref_harness_gpu_debug,
         -maxrregcount=128, its default launch shape) on the headline workload, CUDA-event time of its kernel;
  strong    C3 with the 1 % statistical stopping criterion run to the criterion on N GPUs (strong scaling: the work is
         fixed): time to criterion, passes, and the share of the collectives (moquimc_b200/parallel.py StoppingLoop).

--impl reference times that CPU reference alone (all host cores, bounded sample per step).
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "primary proton histories/s (C1: 200 MeV, water phantom 200x200x350, Dose scorer)"
UNIT = "histories/s"
NXYZ = (200, 200, 350)
LXYZ = (100.0, 100.0, 350.0)
ENERGY = 200.0
SPOT = 30.0
STEPS_PER_HISTORY = 446.3      # SURVEY.md section 8d, oracle-measured scored steps per 200 MeV history
BYTES_PER_STEP = 20.0          # 4 B density + 8 B + 8 B fp64 dose RMW
WORKLOAD = ("C1 phantom_env case at throughput size: 200 MeV pencil beam, 30 mm uniform square spot, "
            "water phantom 100x100x350 mm on a 200x200x350 grid, debug physics, 1 Dose scorer (fp64)")


# --------------------------------------------------------------------------------------------------
# clocks sampled during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        self.proc = subprocess.Popen([exe, "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                      "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (oracle/_ref), all host cores
# --------------------------------------------------------------------------------------------------
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "phantom_env_cpu_debug")


_JSON_FD = None


def emit(line):
    """the one JSON line of this run, on the real stdout"""
    text = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(text.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, text)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_sample(procs, histories_per_proc, seed):
    """P independent single-threaded reference processes (the reference CPU path is single-threaded,
    mqi_phantom_env.hpp:420) on the bench workload.  Returns (total histories, seconds) where seconds
    is the slowest process's transport phase: the reference's own 'Run done <x> s' print, rescaled
    by 10 because it multiplies milliseconds by 1e-4 (mqi_phantom_env.hpp:427)."""
    import numpy as np
    if not os.path.exists(REF_EXE):
        raise RuntimeError("oracle/_ref/phantom_env_cpu_debug is missing (run `make -C oracle ref` in the build container)")
    work = tempfile.mkdtemp(prefix="mqi_bench_ref_")
    try:
        ph = os.path.join(work, "phantom.raw")
        np.zeros((NXYZ[2], NXYZ[1], NXYZ[0]), dtype=np.int16).tofile(ph)
        def launch(p, attempt):
            od = os.path.join(work, "o%d_%d" % (p, attempt))
            os.makedirs(od)
            cmd = [REF_EXE, "--lxyz", "100", "100", "350", "--pxyz", "0.0", "0.0", "-175", "--nxyz", "200", "200", "350",
                   "--spot_energy", str(ENERGY), "0.0", "--spot_position", "0", "0", "0.5",
                   "--spot_size", str(SPOT), str(SPOT), "--histories", str(histories_per_proc),
                   "--phantom_path", ph, "--output_prefix", od, "--random_seed", str(seed + 7919 * p + 104729 * attempt), "--gpu_id", "0"]
            return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)

        ps = [launch(p, 0) for p in range(procs)]
        slowest = 0.0
        for p, pr in enumerate(ps):
            # the reference's CPU build reads a few uninitialised values; a process that dies or
            # prints no timing is run again with another seed (alone: its time still counts as the slowest) before
            # the sample is given up
            for attempt in range(3):
                out, _ = pr.communicate()
                m = re.search(r"Run done ([0-9.eE+-]+) s", out) if pr.returncode == 0 else None
                if m and float(m.group(1)) > 0.0:
                    break
                sys.stderr.write("bench.py: reference process %d failed (rc=%s), attempt %d\n" % (p, pr.returncode, attempt))
                if attempt == 2:
                    raise RuntimeError("reference process failed three times: rc=%s\n%s" % (pr.returncode, out[-500:]))
                pr = launch(p, attempt + 1)
            slowest = max(slowest, 10.0 * float(m.group(1)))
        return procs * histories_per_proc, slowest
    finally:
        shutil.rmtree(work, ignore_errors=True)


def reference_procs():
    # each process holds ~0.35 GB (224 MB scorer table + 56 MB density + 28 MB HU); stay within RAM
    cores = host_cores()
    try:
        avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2**30
        cores = max(1, min(cores, int(avail_gb / 0.5)))
    except (ValueError, OSError):
        pass
    return cores


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    procs = reference_procs()
    n = args.ref_histories_per_proc
    for _ in range(args.warmup):
        run_reference_sample(procs, max(200, n // 10), 1)
    total_h, total_s = 0, 0.0
    for s in range(args.steps):
        h, sec = run_reference_sample(procs, n, 1000 + s)
        total_h += h
        total_s += sec
    value = total_h / total_s
    sample = "%d processes x %d histories per step of the bench workload, transport phase only" % (procs, n)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 transport, f64 dose accumulation", "data": "synthetic",
        "config": {"workload": WORKLOAD, "histories_per_step": procs * n, "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------------
# the B200 path
# --------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_capture(histories_per_launch):
    """profiles/roofline_traffic.json (dram bytes and warp instructions per launch of the headline kernel from an
    ncu --set full capture) if it was taken at this launch size AND from the kernel this library contains: the file
    carries the SASS hash of the kernel it profiled (moquimc_b200.build.kernel_identity), a library built from other
    source no longer matches and the capture is refused as stale instead of being quoted."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None, "no committed capture"
    if int(d.get("histories_per_launch", -1)) != int(histories_per_launch):
        return None, "capture taken at another launch size"
    from moquimc_b200 import build as B
    ident = B.kernel_identity()
    if ident is None:
        return None, "cuobjdump unavailable: kernel identity unchecked, capture not used"
    if d.get("sass_sha256") != ident["sass_sha256"]:
        return None, "stale: the capture's kernel (%s...) is not the library's (%s...)" % (str(d.get("sass_sha256"))[:12], ident["sass_sha256"][:12])
    return d, "profiles/roofline_traffic.json, kernel identity %s... verified" % ident["sass_sha256"][:12]


def ncu_issue(capture, kernel_ms, sm_mhz, sm_count=148):
    """Issue-slot utilisation of the transport kernel, the bound that actually holds: warp
    instructions per launch from the committed ncu capture (smsp__inst_executed.sum at this launch size) over
    the live kernel time, against SMs x 4 schedulers x the SM clock sampled during the timed region."""
    if not capture or not sm_mhz:
        return None
    achieved = float(capture["warp_instructions_per_launch"]) / (kernel_ms * 1e-3)
    peak = int(sm_count) * 4 * float(sm_mhz) * 1e6
    return {"achieved_warp_inst_per_s": achieved, "peak_warp_inst_per_s": peak, "frac": achieved / peak,
            "source": "smsp__inst_executed.sum of profiles/roofline_traffic.json over the live kernel time"}


# --------------------------------------------------------------------------------------------------
# the reference's own CUDA path on this GPU (oracle/_ref/ref_harness_gpu_debug): baseline, N = 1
# --------------------------------------------------------------------------------------------------
def gpu_reference_baseline(histories=10_000_000):
    import numpy as np
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_harness_gpu_debug")
    if not os.path.exists(exe):
        return {"value": None, "kind": "reference-cuda", "note": "oracle/_ref/ref_harness_gpu_debug is missing (make -C oracle ref in the build container)"}
    work = tempfile.mkdtemp(prefix="mqi_bench_refgpu_")
    try:
        ph = os.path.join(work, "phantom.raw")
        np.zeros((NXYZ[2], NXYZ[1], NXYZ[0]), dtype=np.int16).tofile(ph)
        cmd = [exe, "--lxyz", "100", "100", "350", "--pxyz", "0.0", "0.0", "-175", "--nxyz", "200", "200", "350",
               "--spot_energy", str(ENERGY), "0.0", "--spot_position", "0", "0", "0.5", "--spot_size", str(SPOT), str(SPOT),
               "--histories", str(int(histories)), "--phantom_path", ph, "--output_prefix", work, "--random_seed", "12345",
               "--gpu_id", os.environ.get("LOCAL_RANK", "0"), "--scorers", "dose", "--sample_threads", str(min(host_cores(), 32))]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        if r.returncode != 0:
            return {"value": None, "kind": "reference-cuda", "note": "failed rc=%d: %s" % (r.returncode, r.stdout[-300:])}
        st = {}
        for ln in open(os.path.join(work, "harness_stats.txt")):
            t = ln.split()
            if len(t) == 2:
                st[t[0]] = float(t[1])
        return {"value": st["histories"] / st["transport_seconds"], "unit": UNIT, "kind": "reference-cuda",
                "threads": [int(st["threads"]), int(st["blocks"])], "regcap": 128, "histories": int(st["histories"]),
                "kernel_s": st["transport_seconds"], "run_s": st["run_seconds"],
                "value_incl_sampling_and_uploads": st["histories"] / st["run_seconds"],
                "what": "the reference's transport_particles_patient<float> (nvcc -x cu --use_fast_math -maxrregcount=128, sm_100a, "
                        "__PHYSICS_DEBUG__) on the bench workload, its default launch shape, CUDA events around its kernel "
                        "(oracle/ref_harness.cpp); initialize_threads and the host vertex sampling are outside kernel_s"}
    except Exception as ex:
        return {"value": None, "kind": "reference-cuda", "note": "failed: %s" % ex}
    finally:
        shutil.rmtree(work, ignore_errors=True)


# --------------------------------------------------------------------------------------------------
# the other BASELINE configs, N = 1 (moquimc_b200/configs.py)
# --------------------------------------------------------------------------------------------------
def config_legs(device, peak):
    from moquimc_b200 import configs as K
    legs = (("C2", lambda: K.c2(device)), ("C3", lambda: K.c3((device,))), ("C4_reference_table", lambda: K.c4(device, 393_216_001)),
            ("C4_auto_table", lambda: K.c4(device, 1_600_000_001)), ("rangeshifter_aperture", lambda: K.rs_aperture(device)))
    out = {}
    for name, fn in legs:
        try:
            r = fn()
            r.pop("passes", None)
            if r.get("steps_per_history") and r.get("kernel_ms"):
                alg = r["histories"] * r["steps_per_history"] * r["bytes_per_step"]
                ach = alg / (r["kernel_ms"] * 1e-3) / 1e9
                r["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                 "algorithmic_bytes_per_launch": alg,
                                 "bytes_per_step": "4 B density + 16 B fp64 read-modify-write per dense scorer (SURVEY 8d)"}
            out[name] = r
        except Exception as ex:
            out[name] = {"value": None, "note": "failed: %s" % str(ex)[-300:]}
    return out


# --------------------------------------------------------------------------------------------------
# strong scaling: C3 with the 1 % stopping criterion over the ranks
# --------------------------------------------------------------------------------------------------
def strong_c3(rank, world, local, dev, stream, seed, criteria=1.0):
    import numpy as np
    import torch
    import torch.distributed as dist
    from moquimc_b200 import capi, configs as K, parallel as P, synthetic as S
    n3, sp3 = (512, 512, 200), (1.0, 1.0, 2.5)
    hu, origin = S.head_ct(n3, sp3, seed=1)
    edges = [(np.float32(origin[a] - sp3[a] / 2) + np.arange(n3[a] + 1, dtype=np.float32) * np.float32(sp3[a])).astype(np.float32) for a in range(3)]
    spots = [None]
    if rank == 0:   # the plan's beamlets as tps_env builds them (beam model, histories per spot): --dry-run, no GPU
        root = tempfile.mkdtemp(prefix="mqi_strong_")
        try:
            K.c3_case(root)
            inp = os.path.join(root, "c3.in")
            S.write_input(inp, root, os.path.join(root, "out"), ParticlesPerHistory=400.0)
            r = subprocess.run([K.TPS_ENV, "--dry-run", inp], capture_output=True, text=True, timeout=300)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("DRYRUN ")][-1]
            spots[0] = json.loads(line[len("DRYRUN "):])["beams"][0]["spots"]
        finally:
            shutil.rmtree(root, ignore_errors=True)
    if world > 1:
        dist.broadcast_object_list(spots, src=0)
    spots = spots[0]
    bl = [capi.make_beamlet(s["energy"], s["mean"], s["sigma"], uniform=False, sigma_energy=s["sigma_energy"], rot=s["rot"],
                            trans=s["trans"]) for s in spots]
    hist = [s["histories"] for s in spots]
    total = int(sum(hist))
    nvox = n3[0] * n3[1] * n3[2]
    n_pad = P.StoppingLoop.padded_len(nvox)
    eng = capi.Engine(local, physics=capi.PHYSICS_RELEASE)
    eng.set_stream(stream.cuda_stream)
    d_hu = torch.from_numpy(hu.reshape(-1)).to(dev)
    eng.set_grid_hu_device(edges[0], edges[1], edges[2], d_hu.data_ptr())
    bufs = [torch.zeros(n_pad, dtype=torch.float64, device=dev) for _ in range(3)]
    for kind, name, b in ((capi.SCORER_DOSE, "Dose", bufs[0]), (capi.SCORER_DOSE, "Dose_stat", bufs[1]), (capi.SCORER_DOSE_SQ, "DoseSquare_stat", bufs[2])):
        eng.bind_scorer_buffer(eng.add_scorer(kind, name), b.data_ptr())
    eng.set_beamlets(bl, hist)
    kernel_ms = []

    def transport_pass(k):
        # interleaved sharding (chunks of 32 histories dealt round-robin over the ranks): the plan is sorted by energy
        # layer, contiguous history ranges would leave the rank with the highest layers working longest
        st = eng.run_sharded(seed + k, 0, total, world, rank)
        kernel_ms.append(st.kernel_ms)
        return st.histories

    def evaluate(s, q, n, mx):
        return eng.stat_partial_buffers(s.data_ptr(), q.data_ptr(), s.numel(), n, 0.5, mx)
    # warm-up, untimed: one short pass through the same loop (kernel module, torch's reduction kernels, the NCCL
    # channels of the reduce-scatter and of the final reduce), then the buffers are cleared
    def warm_pass(k):
        return eng.run_sharded(seed + 1000, 0, min(total, 100_000), world, rank).histories
    P.StoppingLoop(0.0, warm_pass, evaluate, threshold=0.5, max_passes=1).run(bufs[1], bufs[2])
    if world > 1:   # NCCL sets up the channels of a collective on its first large call: not part of a pass
        w = torch.zeros(4 * 1024 * 1024, dtype=torch.float64, device=dev)
        for _ in range(2):
            P.reduce_scatter_sum(torch.empty(w.numel() // world, dtype=torch.float64, device=dev), w)
        dist.reduce(bufs[0], dst=0)
    def timed_loop():
        for b in bufs:
            b.zero_()
        del kernel_ms[:]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        lp = P.StoppingLoop(criteria, transport_pass, evaluate, threshold=0.5, max_passes=400, histories_per_pass=total)
        res = lp.run(bufs[1], bufs[2])
        t1 = time.perf_counter()
        if world > 1:
            dist.reduce(bufs[0], dst=0, op=dist.ReduceOp.SUM)    # the dose travels once
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        tt = torch.tensor([t2 - t0, sum(kernel_ms) * 1e-3, lp.stat_seconds, t2 - t1], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return lp, res, [float(x) for x in tt.tolist()]

    # Rehearsal, untimed: the whole loop once with the real pass size.  The short warm pass above does not reach the sizes of
    # the real exchange, and whatever is set up on their first use (allocator blocks of the packed buffers, NCCL for messages
    # of that size) showed up as 14 ms per pass in the first run of the loop and 1 ms in the second (measured on two GPUs).
    timed_loop()
    loop, (tracked, current, passes), (total_s, transport_s, stat_s, reduce_s) = timed_loop()
    checksum = float(bufs[0][:nvox].sum().item()) if rank == 0 else 0.0
    reversed_fetch = None
    if world > 1:   # the same loop with every launch handing out its histories last to first (option fetch_order): the plan is
        eng.set_option("fetch_order", 1)   # listed by ascending energy, so the tail of each small launch is made of short histories
        _, (tr2, cur2, p2), (tot2, trs2, _, _) = timed_loop()
        eng.set_option("fetch_order", 0)
        reversed_fetch = {"time_to_criterion_s": tot2, "transport_s": trs2, "passes": p2, "uncertainty_percent": cur2, "histories": tr2}
    eng.set_stream(None)
    eng.close()
    return {"workload": "C3: synthetic head-and-neck CT 512x512x200, %d-spot PBS plan, Dose + the two stat scorers, release physics, "
                        "statistical stopping at %g %% (StatThreshold 0.5), %d histories per pass" % (len(bl), criteria, total),
            "scaling": "strong", "n_gpus": world, "criteria_percent": criteria, "uncertainty_percent": current, "passes": passes,
            "histories": tracked, "time_to_criterion_s": total_s, "histories_per_s": tracked / total_s,
            "transport_s": transport_s, "stat_s": stat_s, "stat_phases_s_rank0": loop.phase_seconds, "final_reduce_s": reduce_s,
            "overhead_share": (total_s - transport_s) / total_s,   # everything but the slowest rank's kernels: selection, exchange,
                                                                   # evaluation, the final reduce, launch and host latencies
            "collective": ("per pass %d collectives: max-all-reduce of %d chunk flags, one ncclReduceScatter each of the packed chunks of sum d "
                           "and sum d^2 that can hold a voxel above the dose threshold (%d of %d values each), max- and sum-all-reduce "
                           "of three doubles; once: ncclReduce of the dose grid" % (loop.collectives_per_pass, n_pad // P.StoppingLoop.CHUNK,
                                                                                   loop.exchanged_values, nvox))
                          if world > 1 else "none (1 GPU): the criterion is evaluated on the device where the grids are",
            "timer": "host wall clock between device synchronisations and barriers, max over ranks; the phases by CUDA events on rank 0", "dose_checksum": checksum,
            "reversed_fetch_order": reversed_fetch}


def bench_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from moquimc_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with --nproc-per-node %d" % (args.gpus, args.gpus))
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the transport path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    H = int(args.histories)
    K, W = args.steps, args.warmup
    nx, ny, nz = NXYZ
    nvox = nx * ny * nz
    xe, ye, ze = capi.uniform_edges(-50, 50, nx), capi.uniform_edges(-50, 50, ny), capi.uniform_edges(-350, 0, nz)
    beamlet = capi.make_beamlet(ENERGY, [0, 0, 0.5, 0, 0, -1], [SPOT, SPOT, 0, 0, 0, 0], uniform=True)
    total = (W + K) * world * H

    # ---------------- HBM-resident leg ----------------
    eng = capi.Engine(local, physics=capi.PHYSICS_DEBUG)
    # the transport kernel is launched on torch's current stream so that torch events bracket it; a
    # side stream, because the legacy default stream has handle 0 (= "use the handle's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    d_hu = torch.zeros(nvox, dtype=torch.int16, device=dev)
    eng.set_grid_hu_device(xe, ye, ze, d_hu.data_ptr())
    s_dose = eng.add_scorer(capi.SCORER_DOSE, "Dose")
    dose = torch.zeros(nvox, dtype=torch.float64, device=dev)       # this rank's grid of the current step
    eng.bind_scorer_buffer(s_dose, dose.data_ptr())
    total_dose = torch.zeros(nvox, dtype=torch.float64, device=dev) if world > 1 and rank == 0 else None
    eng.set_beamlets([beamlet], [total])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(s):
        """one batch on this rank: histories [(s*world + rank)*H, +H), then the dose reduce"""
        if world > 1:
            dose.zero_()
        eng.run_async(args.seed, (s * world + rank) * H, H)
        if world > 1:
            dist.reduce(dose, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                total_dose.add_(dose)

    for s in range(W):
        step(s)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kernel_ms, launches, overflows = [], 0, 0
    t_wall = time.perf_counter()
    for s in range(K):
        flush.fill_(s & 0xff)            # evict the dose / material volumes from L2 between steps (untimed)
        ev[s][0].record(stream)
        step(W + s)
        ev[s][1].record(stream)
        st = eng.run_stats()             # waits for the kernel; device time of this launch
        kernel_ms.append(st.kernel_ms)
        launches += st.launches
        overflows += st.stack_overflows     # secondaries the per-lane stack had no room for (the reference drops them too, B10)
        assert st.histories == H, (st.histories, H)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * H * K / (ms * 1e-3)
    checksum = float((total_dose if total_dose is not None else dose).sum().item()) if rank == 0 else 0.0
    eng.set_stream(None)
    eng.close()
    del dose, total_dose, d_hu, flush
    torch.cuda.empty_cache()

    # ---------------- end-to-end leg: host buffers through the C ABI ----------------
    hu_host = torch.zeros(nvox, dtype=torch.int16).pin_memory()
    out_host = torch.empty(nvox, dtype=torch.float64).pin_memory()
    red = torch.zeros(nvox, dtype=torch.float64, device=dev) if world > 1 else None
    e2 = capi.Engine(local, physics=capi.PHYSICS_DEBUG)
    e2_s = e2.add_scorer(capi.SCORER_DOSE, "Dose")
    if world > 1:
        e2.set_stream(stream.cuda_stream)
        e2.bind_scorer_buffer(e2_s, red.data_ptr())

    e2e_kernel_ms = []

    def e2e_step(s):
        e2.set_grid_hu(xe, ye, ze, hu_host.numpy().reshape(nz, ny, nx))          # H2D: HU volume
        e2.set_beamlets([beamlet], [total])                                       # H2D: beam model
        if world > 1:
            red.zero_()
        else:
            e2.clear_scorers()
        st = e2.run(args.seed, (s * world + rank) * H, H)
        e2e_kernel_ms.append(st.kernel_ms)
        if world > 1:
            dist.reduce(red, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                out_host.copy_(red, non_blocking=False)                          # D2H: reduced dose
            torch.cuda.synchronize()
        else:
            e2.get_dense(e2_s, out=out_host.numpy())                              # D2H: dose grid

    ke = max(1, min(K, args.e2e_steps))
    e2e_step(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(ke):
        e2e_step(W + s)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_serial = world * H * ke / e2e_s
    h2d = world * (nvox * 2 + 4 * (nx + ny + nz + 3) + 128 + 8)
    d2h = nvox * 8
    e2.close()
    del red

    # The same steps the way a caller with more than one batch runs them (beams of a plan, batches of the stopping
    # loop): two handles, double-buffered.  Step s uploads its HU volume and beam model and launches on one handle
    # while the other handle's fp64 dose grid of step s-1 travels to the host; every step still moves all of its
    # bytes inside the timed region, only the copies no longer wait for each other's kernels.
    pipe = [capi.Engine(local, physics=capi.PHYSICS_DEBUG) for _ in range(2)]
    pipe_s = [e.add_scorer(capi.SCORER_DOSE, "Dose") for e in pipe]
    pipe_out = [out_host, torch.empty(nvox, dtype=torch.float64).pin_memory()]
    pipe_red = [torch.zeros(nvox, dtype=torch.float64, device=dev) for _ in range(2)] if world > 1 else None
    copy_stream = torch.cuda.Stream(device=dev)
    reduced = [torch.cuda.Event() for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    if world > 1:   # the reduces of one communicator stay on one stream; only the download leaves it
        for e, sc, r in zip(pipe, pipe_s, pipe_red):
            e.set_stream(stream.cuda_stream)
            e.bind_scorer_buffer(sc, r.data_ptr())

    def pipe_issue(s):
        i = s & 1
        e = pipe[i]
        e.set_grid_hu(xe, ye, ze, hu_host.numpy().reshape(nz, ny, nx))           # H2D: HU volume
        e.set_beamlets([beamlet], [total])                                        # H2D: beam model
        if world > 1:
            stream.wait_event(copied[i])          # the download of step s-2 has left this buffer
            pipe_red[i].zero_()
        else:
            e.clear_scorers()
        e.run_async(args.seed, (s * world + rank) * H, H)
        if world > 1:
            dist.reduce(pipe_red[i], dst=0, op=dist.ReduceOp.SUM)
            reduced[i].record(stream)

    def pipe_collect(s):
        i = s & 1
        if world > 1:
            if rank == 0:
                copy_stream.wait_event(reduced[i])
                with torch.cuda.stream(copy_stream):
                    pipe_out[i].copy_(pipe_red[i], non_blocking=True)             # D2H: reduced dose
                    copied[i].record(copy_stream)
                copied[i].synchronize()
            else:
                reduced[i].synchronize()
        else:
            pipe[i].get_dense(pipe_s[i], out=pipe_out[i].numpy())                 # D2H: dose grid (waits for the kernel)
        assert pipe[i].run_stats().histories == H

    def pipe_run(first, n):
        pipe_issue(first)
        for s in range(first + 1, first + n):
            pipe_issue(s)
            pipe_collect(s - 1)
        pipe_collect(first + n - 1)

    pipe_run(0, 2)                                # both handles once, untimed
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pipe_run(W, ke)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * H * ke / e2e_s
    e2e_checksum = float(pipe_out[(W + ke - 1) & 1].sum().item()) if rank == 0 else 0.0
    for e in pipe:
        if world > 1:
            e.set_stream(None)
        e.close()
    del pipe_red

    strong = None
    if not args.no_strong:
        try:
            strong = strong_c3(rank, world, local, dev, stream, args.seed)
        except Exception as ex:
            strong = {"scaling": "strong", "n_gpus": world, "time_to_criterion_s": None, "note": "failed: %s" % str(ex)[-300:]}
    if rank == 0:
        peak, peak_src = measured_peaks()
        capture, capture_note = committed_capture(H)
        k_ms = sum(kernel_ms) / len(kernel_ms)
        alg_bytes = H * STEPS_PER_HISTORY * BYTES_PER_STEP
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 transport, f64 dose accumulation", "data": "synthetic",
            "config": {"workload": WORKLOAD, "histories_per_gpu_per_step": H, "global_histories_per_step": world * H,
                       "parallelism": "histories sharded over %d GPU(s), one NCCL reduce of the dose grid per step" % world
                       if world > 1 else "1 GPU",
                       "l2": "256 MB flush written between timed steps (untimed); working set 140 MB > 126 MB L2",
                       "timer": "CUDA events on the launching stream per step, max over ranks",
                       "dose_checksum": checksum, "wall_s_timed_region": t_wall,
                       "parity": "gamma 1 %/1 mm >= 99 %, R80 within 0.1 mm against the reference's CPU dose on this workload: "
                                 "tests/test_gpu_parity.py::test_c1_dose_against_reference_golden"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": ke, "kernel_ms_per_step": sum(e2e_kernel_ms[1:]) / max(1, len(e2e_kernel_ms) - 1),
                    "how": "two handles, double-buffered: every step uploads its HU volume and beam model from pinned host memory, "
                           "transports, and downloads its fp64 dose grid; step s's download overlaps step s+1's kernel",
                    "serial_value": e2e_serial, "serial_how": "one handle, every copy and the kernel one after the other",
                    "last_step_dose_checksum": e2e_checksum,
                    "timer": "host wall clock around the C-ABI calls of all steps, synchronised both sides, max over ranks"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": float(capture["dram_bytes_per_launch"]) if capture else None, "traffic_source": capture_note,
                         "peak_source": peak_src, "kernel": "transport_kernel<debug, SET_DOSE>",
                         "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "latency/issue bound, not HBM bound: the dose grid hot set stays in L2 (see DESIGN.md)",
                         "issue_slots": ncu_issue(capture, k_ms, (clocks or {}).get("sm_mhz"),
                                                  torch.cuda.get_device_properties(dev).multi_processor_count)},
            "stack_overflows": overflows,
        }
        if strong is not None:
            line["strong"] = strong
        if world == 1 and not args.no_configs:
            line["configs"] = config_legs(local, peak)
        if world == 1 and not args.no_gpu_baseline:
            g = gpu_reference_baseline(args.gpu_ref_histories)
            if g.get("value"):
                g["this_over_reference_cuda"] = value / g["value"]
            line["gpu_reference_baseline"] = g
        if world == 1 and not args.no_cpu_baseline:
            try:
                procs = reference_procs()
                n = args.ref_histories_per_proc
                h, sec = run_reference_sample(procs, n, 4242)
                line["cpu_baseline"] = {"value": h / sec, "unit": UNIT, "cores": procs, "kind": "reference",
                                        "sample": "%d reference processes x %d histories of the bench workload, "
                                                  "transport phase only" % (procs, n)}
            except Exception as ex:   # the checker is missing: report it, never fake a number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %s" % ex}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--histories", type=float, default=1e7, help="primaries per GPU per step")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-histories-per-proc", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2 / C3 / C4 / beamline legs (N = 1)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the reference's own CUDA kernel (N = 1)")
    ap.add_argument("--no-strong", action="store_true", help="skip the C3 strong-scaling record")
    ap.add_argument("--gpu-ref-histories", type=float, default=1e7)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    # stdout carries exactly ONE line, the JSON: anything libraries write to file descriptor 1 on the way
    # (NCCL prints its version there) goes to stderr instead
    sys.stdout.flush()
    global _JSON_FD
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return bench_reference(args)
    return bench_b200(args)


if __name__ == "__main__":
    sys.exit(main())
