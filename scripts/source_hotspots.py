"""Per-source-line instruction counts from an ncu capture (compile with -lineinfo, capture with --import-source on):
    ncu -i x.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python scripts/source_hotspots.py x.csv [N]
Prints the N source lines with the most executed warp instructions: share of all warp instructions, share of
the stall samples, average active threads per instruction."""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
f = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        f = os.path.basename(r[1]); continue
    if len(r) < 10 or r[0] in ("Line No", "Function Name"): continue
    if r[2] != "-": continue            # SASS rows carry an address; source rows carry "-"
    try:
        out.append((f, int(r[0]), r[1].strip(), int(r[6]), int(r[7]), int(r[8])))
    except ValueError:
        pass
tw = sum(o[4] for o in out); ts = sum(o[3] for o in out)
print("total warp inst", tw, "samples", ts)
for f, ln, src, smp, wi, ti in sorted(out, key=lambda o: -o[4])[:top]:
    print("%-18s %5d inst %5.2f%% samp %5.2f%% eff %4.1f | %s" % (f, ln, 100.0 * wi / tw, 100.0 * smp / max(ts, 1), ti / max(wi, 1), src[:95]))
