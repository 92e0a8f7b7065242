"""profiles/r02_traffic.json from `ncu --set full` captures: DRAM bytes per launch of the kernels bench.py quotes a
roofline for.  Usage:
    python tools/ncu_traffic.py profiles/r02_traffic.json  KERNEL_REGEX:shape:report.ncu-rep [...]
e.g. assoc_tc_kernel:2000,2000,2048,1:gpurun_out/r02_assoc_full.ncu-rep
The shape is the tuple bench.py looks the record up with.  Reads the report with `ncu -i ... --page raw --csv`
(works in the build container: no GPU needed)."""
import csv, io, json, re, subprocess, sys


def read_report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return v * scale


def main():
    dst = sys.argv[1]
    kernels = []
    for spec in sys.argv[2:]:
        regex, shape, path = spec.split(":")
        hdr, units, rows = read_report(path)
        ik = hdr.index("Kernel Name")
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        sel = [r for r in rows if re.search(regex, r[ik])]
        if not sel:
            raise SystemExit(f"no kernel matching {regex} in {path}")
        rd = [to_bytes(r[ir], units[ir]) for r in sel]
        wr = [to_bytes(r[iw], units[iw]) for r in sel]
        name = re.sub(r"\(.*", "", sel[0][ik]).split("::")[-1].split("<")[0].strip()
        kernels.append({"kernel": name, "shape": [int(x) for x in shape.split(",")], "launches_in_capture": len(sel),
                        "dram_bytes_read_per_launch": sum(rd) / len(rd), "dram_bytes_write_per_launch": sum(wr) / len(wr),
                        "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd),
                        "ncu_duration_us": [float(r[it]) for r in sel], "capture": f"ncu --set full --clock-control none, {path.split('/')[-1]}",
                        "note": "cold-cache, serialised replay; writes that are still dirty in the 126 MB L2 when the kernel ends are not in dram__bytes_write"})
    json.dump({"kernels": kernels}, open(dst, "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
