"""Top CUDA source lines of one kernel in an .ncu-rep by sampled stalls (needs -lineinfo and --import-source on).
usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fpath, hdr, data = "", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) >= len(hdr) - 2 and r[0].isdigit():
        data.append((fpath, r))
col = {k: hdr.index(k) for k in ("# Samples", "Instructions Executed", "stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait",
                                 "stall_math", "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_mio", "stall_lg",
                                 "Avg. Threads Executed", "L2 Theoretical Sectors Global")}
def num(r, k):
    try:
        return int(r[col[k]])
    except ValueError:
        return 0
tot = sum(num(r, "# Samples") for _, r in data)
inst = sum(num(r, "Instructions Executed") for _, r in data)
print("total samples %d, warp instructions %d" % (tot, inst))
print("%6s %5s | %6s %6s %6s %6s %6s %6s | %10s %4s %10s | line" % ("samp", "%", "barr", "longsb", "shrtsb", "wait", "math", "nsel", "inst", "thr", "L2sect"))
for f, r in sorted(data, key=lambda x: -num(x[1], "# Samples"))[:top]:
    s = num(r, "# Samples")
    print("%6d %5.1f | %6d %6d %6d %6d %6d %6d | %10d %4s %10d | %s:%s %s" % (
        s, 100.0 * s / max(tot, 1), num(r, "stall_barrier"), num(r, "stall_long_sb"), num(r, "stall_short_sb"), num(r, "stall_wait"),
        num(r, "stall_math"), num(r, "stall_not_selected"), num(r, "Instructions Executed"), r[col["Avg. Threads Executed"]],
        num(r, "L2 Theoretical Sectors Global"), f, r[0], r[1].strip()[:90]))
