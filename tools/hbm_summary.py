"""Per-kernel achieved HBM bandwidth from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --csv` launch
lists: median duration, DRAM bytes per launch, DRAM GB/s and L2 GB/s, against the measured copy bandwidth (MEASURED_PEAKS.json, 6549 GB/s)."""
import collections, csv, json, os, re, sys
peak = 6549.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
rows = collections.defaultdict(dict)
for path in sys.argv[1].split(","):
    lines = [l for l in open(path) if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if "Metric Value" not in r or r["Metric Value"] in ("", "n/a"):
            continue
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        else:
            v = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) * v
        key = (path, r["ID"])
        rows[key][r["Metric Name"]] = v
        rows[key]["name"] = re.sub(r"\(.*", "", r["Kernel Name"]) + " " + r.get("Grid Size", "").replace(" ", "")
agg = collections.defaultdict(list)
for d in rows.values():
    if pat is None or pat.search(d["name"]):
        agg[d["name"]].append(d)
print(f"{'kernel  (grid)':78s} {'n':>3s} {'us(med)':>8s} {'DRAM MB':>8s} {'DRAM GB/s':>9s} {'frac':>5s} {'L2 MB':>8s} {'L2 GB/s':>8s}")
for k, ds in sorted(agg.items()):
    ds = sorted(ds, key=lambda d: d.get("gpu__time_duration.sum", 0))
    d = ds[len(ds) // 2]
    t = d.get("gpu__time_duration.sum", 0.0)
    mb = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    l2 = d.get("lts__t_bytes.sum", 0.0)
    print(f"{k[:78]:78s} {len(ds):3d} {t:8.2f} {mb:8.2f} {mb / t * 1e3 if t else 0:9.0f} {mb / t * 1e3 / peak if t else 0:5.2f} {l2:8.2f} {l2 / t * 1e3 if t else 0:8.0f}")
