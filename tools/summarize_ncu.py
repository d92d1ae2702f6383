"""ncu --page raw --csv  ->  per-kernel roofline summary (JSON lines + a text table).

    ncu -i prof.ncu-rep --page raw --csv > raw.csv ;  python tools/summarize_ncu.py raw.csv [MEASURED_PEAKS.json]

Per kernel name (launches of one capture averaged): duration, DRAM bytes read+written per launch (`roofline.traffic`),
achieved DRAM GB/s against the measured copy bandwidth, tensor-pipe active %, issue-slot %, registers."""
import csv, json, os, sys, collections

def num(x):
    try:
        return float(str(x).replace(",", ""))
    except Exception:
        return None

def main():
    path = sys.argv[1]
    peaks = json.load(open(sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    rows = list(csv.reader(open(path, newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {h: i for i, h in enumerate(hdr)}
    want = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
            "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "issue_pct": "smsp__issue_active.avg.pct", "regs": "launch__registers_per_thread",
            "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed"}
    scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.OrderedDict()
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        a = agg.setdefault(name, collections.defaultdict(list))
        for k, m in want.items():
            if m in col:
                v = num(r[col[m]])
                if v is None:
                    continue
                u = units[col[m]]
                if k in ("dur", "rd", "wr"):
                    v *= scale.get(u, 1.0)
                a[k].append(v)
    out = []
    for name, a in agg.items():
        mean = lambda k: (sum(a[k]) / len(a[k])) if a[k] else None
        dur, rd, wr = mean("dur"), mean("rd"), mean("wr")
        rec = {"kernel": name, "launches": len(a["dur"]), "duration_us": round(dur * 1e6, 1) if dur else None,
               "dram_read_MB": round(rd / 1e6, 2) if rd is not None else None, "dram_write_MB": round(wr / 1e6, 2) if wr is not None else None,
               "traffic_bytes_per_launch": (rd + wr) if rd is not None and wr is not None else None}
        if dur and rec["traffic_bytes_per_launch"] is not None:
            gbs = rec["traffic_bytes_per_launch"] / dur / 1e9
            rec["dram_GBps"] = round(gbs, 1)
            rec["frac_of_measured_hbm_peak"] = round(gbs / peaks["hbm_gbs"], 3)
        for k in ("dram_pct", "tensor_pct", "issue_pct", "l2_pct", "warps_pct", "regs"):
            v = mean(k)
            rec[k] = round(v, 2) if v is not None else None
        out.append(rec)
    out.sort(key=lambda r: -(r["duration_us"] or 0) * r["launches"])
    for r in out:
        print(json.dumps(r))
    print("# %-34s %5s %10s %9s %9s %7s %7s %7s" % ("kernel", "n", "us/launch", "GB/s", "frac_hbm", "dram%", "tensor%", "issue%"), file=sys.stderr)
    for r in out:
        print("# %-34s %5d %10.1f %9s %9s %7s %7s %7s" % (r["kernel"][:34], r["launches"], r["duration_us"] or 0, r.get("dram_GBps"),
              r.get("frac_of_measured_hbm_peak"), r["dram_pct"], r["tensor_pct"], r["issue_pct"]), file=sys.stderr)

if __name__ == "__main__":
    main()
