"""Registers / stack / spills of every transport_kernel instantiation: python scripts/ptxas_report.py
(rebuilds libmqi_b200.so with -Xptxas -v and parses the log)."""
import os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moquimc_b200 import build as B

B.gen_tables_inc()
cu = [os.path.join(B.CSRC, f) for f in ("mqi_transport.cu", "mqi_capi.cu")]
cmd = [B.nvcc(), "-Xptxas=-v"] + B.NVCC_FLAGS + sys.argv[1:] + ["-shared", "-o", B.LIB] + cu + ["-ldl"]
log = subprocess.run(cmd, capture_output=True, text=True)
if log.returncode:
    print(log.stderr[-3000:]); sys.exit(1)
cur = None
for ln in log.stderr.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        cur = m.group(1); first = True; continue
    if cur and "transport_kernel" in cur:
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m and first:
            stack, st, ld = m.groups(); first = False
        m = re.search(r"Used (\d+) registers", ln)
        if m:
            t = re.search(r"ILi(\d)ELb(\d)ELb(\d)ELb(\d)E", cur)
            names = re.findall(r"IL[ib](\d+)E|ELb(\d)|ELi(\d+)", cur)
            print("%-70s regs %s stack %s spill st %s ld %s" % (cur[12:70], m.group(1), stack, st, ld))
            cur = None
