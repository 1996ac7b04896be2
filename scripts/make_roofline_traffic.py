"""profiles/roofline_traffic.json from the raw CSV page of an ncu --set full capture of the headline kernel
(scripts/gpu_ncu.sh <tag> -> gpurun_out/<tag>.raw.csv), tied to the SASS hash of that kernel in the library the capture
ran (moquimc_b200.build.kernel_identity): python scripts/make_roofline_traffic.py gpurun_out/<tag>.raw.csv <histories> "<source note>" """
import csv, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moquimc_b200 import build as B

raw, hist, note = sys.argv[1], int(float(sys.argv[2])), sys.argv[3]
r = list(csv.reader(open(raw)))
h, u, v = r[0], r[1], r[2]
col = {k: i for i, k in enumerate(h)}


def val(name):
    x, unit = float(v[col[name]].replace(",", "")), u[col[name]]
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(unit, 1.0)


kernel = v[col["Kernel Name"]]
assert "transport_kernel<1, 1, 0, 0>" in kernel.replace("(int)", "").replace("(bool)", ""), kernel
ident = B.kernel_identity()
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
out = {"kernel": kernel, "histories_per_launch": hist, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
       "warp_instructions_per_launch": val("smsp__inst_executed.sum"), "sass_sha256": ident["sass_sha256"],
       "sass_instructions": ident["sass_instructions"], "source": note}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_traffic.json"), "w"), indent=1)
print(out)
