"""Developer tool (CPU): turn an ncu launch list (CSV of gpu__time_duration.sum) of `bench.py --no-graph` and a few
`ncu --set full` reports into the markdown summary committed under profiles/.
usage: profile_summary.py LAUNCHES.csv OUT_STEP.csv [name=report.ncu-rep:"description" ...]"""
import collections
import csv
import re
import subprocess
import sys


def load_launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ix = {k: i for i, k in enumerate(hdr)}
    out = []
    for r in rows[1:]:
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        u = r[ix["Metric Unit"]]
        v = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)
        out.append((re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", ""), v, r))
    return hdr, out


def one_step(launches):
    """The step starts with the stem im2col kernel; take the last complete [start, next start) window."""
    starts = [i for i, (n, _, _) in enumerate(launches) if "im2col_f32" in n]
    if len(starts) < 2:
        return launches
    return launches[starts[-2]:starts[-1]]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return {}
    return {k: (rows[2][i], rows[1][i]) for i, k in enumerate(rows[0])}


KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__cluster_size"]


def main():
    hdr, launches = load_launches(sys.argv[1])
    step = one_step(launches)
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(hdr)
        for _, _, r in step:
            w.writerow(r)
    agg = collections.OrderedDict()
    for n, v, _ in step:
        d = agg.setdefault(n, [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(v for _, v, _ in step)
    conv = sum(v for n, v, _ in step if "tc_gemm" in n)
    print("One training step (last complete im2col-to-im2col window of the launch list): %d launches, %.1f us "
          "serialised kernel time; tcgen05 conv engine %.1f %% (%.1f us).\n" % (len(step), tot, 100 * conv / tot, conv))
    print("| kernel | launches | us | share |\n|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v / tot < 0.0005:
            continue
        print("| `%s` | %d | %.1f | %.1f%% |" % (k[:70], n, v, 100 * v / tot))
    for spec in sys.argv[3:]:
        name, rest = spec.split("=", 1)
        rep, desc = rest.split(":", 1)
        m = raw_metrics(rep)
        print("\n### %s\n%s\n" % (name, desc))
        for k in KEYS:
            if k in m:
                print("- `%s` = %s %s" % (k, m[k][0], m[k][1]))


if __name__ == "__main__":
    main()
