"""Throughput of the BASELINE configs beside C1 on one GPU (development aid; bench.py carries the same legs):
python scripts/config_bench.py [c2 c3 c4 c4big rs ...]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moquimc_b200 import configs as K

legs = {"c2": lambda: K.c2(), "c3": lambda: K.c3((0,)), "c4": lambda: K.c4(capacity=393_216_001), "c4big": lambda: K.c4(capacity=1_600_000_001),
        "rs": lambda: K.rs_aperture(nodes=2), "rs1": lambda: K.rs_aperture(nodes=1), "rs0": lambda: K.rs_aperture(nodes=0),
        "c2_70": lambda: K.c2(energy=70.0), "c2_230": lambda: K.c2(energy=230.0)}
for name in sys.argv[1:] or ["c2", "c3", "c4", "c4big", "rs"]:
    r = legs[name]()
    r.pop("passes", None)
    print(name, json.dumps(r), flush=True)
