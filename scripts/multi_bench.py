"""Throughput of the multi-node kernel: range shifter + aperture in front of a water phantom (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from moquimc_b200 import capi
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
NX, NY, NZ = 100, 100, 200
e = capi.Engine(0, physics=capi.PHYSICS_RELEASE)
e.set_grid_hu(capi.uniform_edges(-50, 50, NX), capi.uniform_edges(-50, 50, NY), capi.uniform_edges(-200, 0, NZ), np.zeros((NZ, NY, NX), np.int16))
if nodes >= 1:
    e.add_beamline_node(np.float32([-150, 150]), np.float32([-150, 150]), np.float32([100, 140]), np.float32([1.19e-3]))
if nodes >= 2:
    axe = capi.uniform_edges(-40, 40, 80); aze = capi.uniform_edges(40, 60, 20)
    xc = 0.5 * (axe[1:] + axe[:-1])
    op = (np.abs(xc)[:, None] < 18.0) & (np.abs(xc)[None, :] < 18.0)
    e.add_beamline_node(axe, axe, aze, np.broadcast_to(np.where(op, np.float32(1e-8), np.float32(100.0)), (20, 80, 80)).astype(np.float32).copy())
e.add_scorer(capi.SCORER_DOSE, "Dose")
e.set_beamlets([capi.make_beamlet(180.0, [0, 0, 180.0, 0, 0, -1], [15, 15, 0, 0, 0, 0], uniform=True)], [n * 4])
e.set_option("count_steps", int(os.environ.get("MQI_COUNT_STEPS", "0")))   # counting runs the general kernel
for i in range(3):
    st = e.run(1, i * n, n)
    print("nodes %d: %d histories kernel %.2f ms -> %.3e hist/s, steps/hist %.1f" % (nodes, st.histories, st.kernel_ms, st.histories / (st.kernel_ms * 1e-3), st.steps / st.histories), flush=True)
