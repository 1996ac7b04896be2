"""Where does an end-to-end C-ABI step spend its time? (development aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from moquimc_b200 import capi
nx, ny, nz = 200, 200, 350
xe, ye, ze = capi.uniform_edges(-50, 50, nx), capi.uniform_edges(-50, 50, ny), capi.uniform_edges(-350, 0, nz)
hu = torch.zeros(nx * ny * nz, dtype=torch.int16).pin_memory()
out = torch.empty(nx * ny * nz, dtype=torch.float64).pin_memory()
e = capi.Engine(0, physics=capi.PHYSICS_DEBUG)
s = e.add_scorer(capi.SCORER_DOSE, "Dose")
b = capi.make_beamlet(200.0, [0, 0, 0.5, 0, 0, -1], [30, 30, 0, 0, 0, 0], uniform=True)
H = 2_000_000
for it in range(3):
    t = [time.perf_counter()]
    e.set_grid_hu(xe, ye, ze, hu.numpy().reshape(nz, ny, nx)); t.append(time.perf_counter())
    e.set_beamlets([b], [H * 10]); t.append(time.perf_counter())
    e.clear_scorers(); t.append(time.perf_counter())
    st = e.run(1, it * H, H); t.append(time.perf_counter())
    e.get_dense(s, out=out.numpy()); t.append(time.perf_counter())
    print("set_grid %.1f ms, set_beamlets %.1f, clear %.1f, run %.1f (kernel %.1f), get_dense %.1f" %
          tuple([1e3 * (t[i + 1] - t[i]) for i in range(4)] + [st.kernel_ms, 1e3 * (t[5] - t[4])]), flush=True)
