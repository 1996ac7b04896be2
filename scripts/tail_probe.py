"""Kernel time of one C3 pass on one GPU as a whole launch and as the share of one of 8 ranks (interleaved shard 0 of 8), for
several values of the option tail_percent (and fetch orders): python scripts/tail_probe.py [tail_percent ...]"""
import json, os, shutil, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from moquimc_b200 import capi, configs as K, synthetic as S

n3, sp3 = (512, 512, 200), (1.0, 1.0, 2.5)
hu, origin = S.head_ct(n3, sp3, seed=1)
edges = [(np.float32(origin[a] - sp3[a] / 2) + np.arange(n3[a] + 1, dtype=np.float32) * np.float32(sp3[a])).astype(np.float32) for a in range(3)]
root = tempfile.mkdtemp(prefix="mqi_tail_")
try:
    K.c3_case(root)
    inp = os.path.join(root, "c3.in")
    S.write_input(inp, root, os.path.join(root, "out"), ParticlesPerHistory=400.0)
    r = subprocess.run([K.TPS_ENV, "--dry-run", inp], capture_output=True, text=True, timeout=300)
    spots = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("DRYRUN ")][-1][len("DRYRUN "):])["beams"][0]["spots"]
finally:
    shutil.rmtree(root, ignore_errors=True)
bl = [capi.make_beamlet(s["energy"], s["mean"], s["sigma"], uniform=False, sigma_energy=s["sigma_energy"], rot=s["rot"], trans=s["trans"]) for s in spots]
hist = [s["histories"] for s in spots]
total = int(sum(hist))
e = capi.Engine(0, physics=capi.PHYSICS_RELEASE)
e.set_grid_hu(edges[0], edges[1], edges[2], hu)
for kind, name in ((capi.SCORER_DOSE, "Dose"), (capi.SCORER_DOSE, "Dose_stat"), (capi.SCORER_DOSE_SQ, "DoseSquare_stat")):
    e.add_scorer(kind, name)
e.set_beamlets(bl, hist)
e.run_sharded(1, 0, min(total, 200000), 1, 0)
for tail in [int(x) for x in sys.argv[1:]] or [0, 100]:
    e.set_option("tail_percent", tail)
    for order in (0, 1):
        e.set_option("fetch_order", order)
        out = []
        for shards in (1, 8):
            ms = [e.run_sharded(10 + k, 0, total, shards, 0).kernel_ms for k in range(3)]
            out.append("1/%d of the pass: %.3f ms" % (shards, sorted(ms)[1]))
        print("tail_percent %3d fetch_order %d: %s" % (tail, order, ", ".join(out)), flush=True)
