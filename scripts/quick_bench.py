"""Quick throughput probe on one GPU (development aid; bench.py is the contract)."""
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from moquimc_b200 import capi

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
accum = int(sys.argv[3]) if len(sys.argv) > 3 else 0
energy = float(sys.argv[4]) if len(sys.argv) > 4 else 200.0
spot = float(sys.argv[5]) if len(sys.argv) > 5 else 30.0
e = capi.Engine(0, physics=variant)
xe, ye, ze = capi.uniform_edges(-50, 50, 200), capi.uniform_edges(-50, 50, 200), capi.uniform_edges(-350, 0, 350)
e.set_grid_hu(xe, ye, ze, np.zeros((350, 200, 200), dtype=np.int16))
e.add_scorer(capi.SCORER_DOSE, "dose")
e.set_accumulation(accum)
e.set_beamlets([capi.make_beamlet(energy, [0, 0, 0.5, 0, 0, -1], [spot, spot, 0, 0, 0, 0], uniform=True)], [n * 8])
e.set_option("count_steps", int(os.environ.get("MQI_COUNT_STEPS", "0")))   # the counter costs ~2 %
if os.environ.get("MQI_L2_PERSIST") is not None:
    e.set_option("l2_persist", int(os.environ["MQI_L2_PERSIST"]))
st = e.run(1, 0, min(n, 200000))
for i in range(3):
    t = time.time()
    st = e.run(1, i * n, n)
    dt = time.time() - t
    print("variant %d accum %d E %.0f: %d histories kernel %.2f ms wall %.2f ms -> %.3e hist/s, steps/hist %.1f, sec/hist %.3f"
          % (variant, accum, energy, st.histories, st.kernel_ms, dt * 1e3, st.histories / (st.kernel_ms * 1e-3),
             st.steps / st.histories, st.secondaries / st.histories), flush=True)
