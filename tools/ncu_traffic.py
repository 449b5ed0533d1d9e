"""profiles/ncu_traffic.json <- DRAM bytes per launch of a kernel from an `ncu --set full` report (what bench.py's roofline.traffic reads).
    python tools/ncu_traffic.py <report.ncu-rep> <kernel-name regex> <key> [launch index, default: median-duration launch]"""
import csv, json, os, re, subprocess, sys
rep, pat, key = sys.argv[1], re.compile(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def val(r, name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(units[i], 1.0)
sel = [r for r in rows[2:] if pat.search(r[hdr.index("Kernel Name")])]
assert sel, "no launch matches"
sel.sort(key=lambda r: val(r, "gpu__time_duration.sum"))
r = sel[int(sys.argv[4])] if len(sys.argv) > 4 else sel[len(sel) // 2]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
d = json.load(open(out)) if os.path.exists(out) else {}
d[key] = {"dram_bytes_per_launch": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"), "dram_read": val(r, "dram__bytes_read.sum"),
          "dram_write": val(r, "dram__bytes_write.sum"), "us": val(r, "gpu__time_duration.sum"), "kernel": r[hdr.index("Kernel Name")][:120],
          "source": f"ncu --set full capture {os.path.basename(rep)} (cold cache: ncu flushes L2 between replays), launches matched {len(sel)}"}
json.dump(d, open(out, "w"), indent=1, sort_keys=True)
print(key, d[key])
