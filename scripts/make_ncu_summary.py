"""profiles/ncu_summary.json from the two `ncu --set full` captures of the bench launch shape (walk launch and last push launch of one
48-query wave): the DRAM bytes per launch that bench.py reports as roofline.traffic, with the commit the captures were taken at.
usage: python scripts/make_ncu_summary.py <walk.ncu-rep> <push.ncu-rep> <slots> <commit> <txt name under profiles/>"""
import csv, json, subprocess, sys

def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]

def f(d, k):
    try:
        return float(d[k].replace(",", ""))
    except Exception:
        return None

def entry(d, launch, fname, slots):
    rd, wr = f(d, "dram__bytes_read.sum"), f(d, "dram__bytes_write.sum")
    return {"kernel_name": d.get("Kernel Name"), "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
            "duration_ms_under_ncu": f(d, "gpu__time_duration.sum"), "l2_hit_rate_pct": f(d, "lts__t_sector_hit_rate.pct"),
            "issue_active_pct": f(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "active_threads_per_inst": f(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "registers_per_thread": f(d, "launch__registers_per_thread"), "grid": f(d, "launch__grid_size"), "block": f(d, "launch__block_size"),
            "launch": launch, "file": fname, "slots": slots}

walk, push, slots, commit, txt = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
w, p = raw(walk)[0], raw(push)[0]
# ncu reports bytes with a unit column; normalise to bytes
def unit_scale(path, key):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    u = dict(zip(rows[0], rows[1])).get(key, "byte").lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)
out = {"shape": "lj", "commit": commit, "source": "ncu --set full --clock-control none, python scripts/profile_run.py %d %d (LJ-shape, FORA eps=0.5 --balanced --opt, one wave = the launch shape of bench.py's default)" % (slots, slots), "kernels": {}}
for name, d, path, launch in (("walk_kernel", w, walk, "the walk launch of a %d-query wave" % slots), ("push_kernel", p, push, "4th (last, largest) of the 4 balanced rounds of a %d-query wave" % slots)):
    e = entry(d, launch, "profiles/" + txt, slots)
    for k in ("dram_bytes_read", "dram_bytes_write"):
        e[k] *= unit_scale(path, "dram__bytes_read.sum" if k.endswith("read") else "dram__bytes_write.sum")
    e["dram_bytes_per_launch"] = e["dram_bytes_read"] + e["dram_bytes_write"]
    out["kernels"][name] = e
print(json.dumps(out, indent=1))
