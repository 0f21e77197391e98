"""Developer tool (CPU): table of the bandwidth-bound kernels from `ncu --set full` reports of tools/eager_steps.py:
per kernel (aggregated over its launches in the report) duration, DRAM bytes read + written, achieved DRAM GB/s and its
fraction of the measured HBM copy bandwidth (MEASURED_PEAKS.json), L2 throughput, grid size.
usage: bw_summary.py report1.(ncu-rep|csv) [...]   (markdown on stdout)"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = {"t": "gpu__time_duration.sum", "r": "dram__bytes_read.sum", "w": "dram__bytes_write.sum",
        "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "grid": "launch__grid_size", "regs": "launch__registers_per_thread",
        "occ": "sm__warps_active.avg.pct_of_peak_sustained_active", "bps": "dram__bytes.sum.per_second"}
UNIT = {"nsecond": 1e-9, "ns": 1e-9, "usecond": 1e-6, "us": 1e-6, "msecond": 1e-3, "ms": 1e-3, "second": 1.0, "s": 1.0,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "": 1.0, "%": 1.0, "byte/second": 1.0,
        "Kbyte/second": 1e3, "Mbyte/second": 1e6, "Gbyte/second": 1e9, "Tbyte/second": 1e12, "byte/s": 1.0,
        "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}


def rows_of(rep):
    if rep.endswith(".csv"):            # already converted on the GPU box (`ncu -i X.ncu-rep --page raw --csv`)
        out = open(rep, errors="replace").read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = {"name": re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")}
        for k, col in COLS.items():
            if col in ix:
                try:
                    d[k] = float(r[ix[col]].replace(",", "")) * UNIT.get(units[ix[col]], 1.0)
                except ValueError:
                    d[k] = 0.0
        yield d


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    print("Peak = %.1f GB/s (MEASURED_PEAKS.json, STREAM-style copy).  `DRAM bytes` = dram__bytes_read.sum + "
          "dram__bytes_write.sum per launch; `GB/s` = DRAM bytes / gpu__time_duration (under ncu: clocks not locked, "
          "caches flushed between replays, so durations are cold-cache).\n" % peak)
    for rep in sys.argv[1:]:
        agg = collections.OrderedDict()
        for d in rows_of(rep):
            a = agg.setdefault((d["name"], int(d.get("grid", 0))), collections.Counter())
            a["n"] += 1
            for k in ("t", "r", "w", "l2", "dram", "occ"):
                a[k] += d.get(k, 0.0)
            if "r" not in d and "bps" in d:        # section-based capture: total DRAM bytes = rate x duration
                a["rw"] += d["bps"] * d.get("t", 0.0)
        print("### %s\n" % os.path.basename(rep))
        print("| kernel | grid | launches | us / launch | DRAM MB / launch (R + W) | GB/s | of HBM peak | DRAM % | L2 % | warps active % |")
        print("|---|---|---|---|---|---|---|---|---|---|")
        for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
            n = a["n"]
            t = a["t"] / n
            by = (a["r"] + a["w"] + a["rw"]) / n
            gbs = by / t / 1e9 if t > 0 else 0.0
            print("| `%s` | %d | %d | %.1f | %.2f%s | %.0f | %.2f | %.1f | %.1f | %.1f |" % (
                name[:48], grid, n, t * 1e6, by / 1e6,
                " (%.2f + %.2f)" % (a["r"] / n / 1e6, a["w"] / n / 1e6) if a["r"] + a["w"] > 0 else "", gbs, gbs / peak,
                a["dram"] / n, a["l2"] / n, a["occ"] / n))
        print()


if __name__ == "__main__":
    main()
