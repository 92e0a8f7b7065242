"""Turn an `ncu --csv` metric dump of tools/kernel_zoo.py into profiles/rNN_kernels_vN.csv:
per kernel: launches, average duration, DRAM bytes, achieved DRAM GB/s, share of the measured HBM peak,
tensor-pipe activity.  Usage: python tools/kernels_summary.py gpurun_out/zoo.csv profiles/r01_kernels_v3.csv"""
import collections, csv, json, os, sys

src, dst = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6553.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
except Exception:
    pass
rows = [r for r in csv.reader(open(src)) if len(r) > 8]
hdr = rows[0]
ki, mi, ui, vi, idi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
per = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "").strip()
    v = float(r[vi].replace(",", "")) if r[vi] not in ("", "n/a") else 0.0
    unit = r[ui]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
    per.setdefault(name, collections.defaultdict(dict))[r[idi]][r[mi]] = v * scale
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "avg_us", "avg_dram_read_MB", "avg_dram_write_MB", "dram_GBps",
                "pct_of_measured_hbm_%.1f" % peak, "tensor_pipe_pct_of_active"])
    for name, launches in per.items():
        n = len(launches)
        t = sum(l.get("gpu__time_duration.sum", 0.0) for l in launches.values()) / n
        rd = sum(l.get("dram__bytes_read.sum", 0.0) for l in launches.values()) / n
        wr = sum(l.get("dram__bytes_write.sum", 0.0) for l in launches.values()) / n
        tp = sum(l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) for l in launches.values()) / n
        gbps = (rd + wr) * 1e6 / (t * 1e-6) / 1e9 if t > 0 else 0.0
        w.writerow([name, n, round(t, 2), round(rd, 3), round(wr, 3), round(gbps, 1), round(100 * gbps / peak, 1), round(tp, 1)])
print(open(dst).read())
