"""Condense ncu CSV logs into the small tables kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.txt> "<command line>"      per-kernel totals / shares
  python tools/ncu_summary.py metrics  <raw.csv> <out.csv>                            --page raw dump -> metric x launch table
  python tools/ncu_summary.py traffic  <metrics.csv> <out.json> <kernel substring>    avg DRAM bytes per launch
"""
import csv
import json
import re
import sys


def rows_of(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return list(csv.reader(lines))


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name[:56]


def launches(path, out, cmd):
    rows = rows_of(path)
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    acc = {}
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        k = short(r[ik])
        t = float(r[iv].replace(",", ""))
        a = acc.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in acc.values())
    with open(out, "w") as f:
        f.write("# %s\n" % cmd)
        f.write("# cold-cache, serialised launch times: compare SHARES with bench.py's phases_ms_per_step, not absolutes.\n")
        f.write("# total captured: %.3f ms over %d launches\n" % (total / 1e6, sum(v[0] for v in acc.values())))
        for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
            f.write("%-56s n=%5d %10.3f ms %6.1f%%\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / total))


def metrics(path, out):
    rows = rows_of(path)
    hdr, units = rows[0], rows[1]
    first = hdr.index("Kernel Name")
    want = ("dram__bytes", "dram__throughput", "gpu__time_duration", "sm__throughput", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma",
            "sm__pipe_fma", "sm__pipe_alu", "smsp__issue_active", "sm__warps_active", "launch__registers", "launch__occupancy", "launch__grid_size",
            "launch__block_size", "data_bank_conflicts", "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate", "smsp__inst_executed.sum",
            "sm__inst_executed.sum", "smsp__warp_issue_stalled", "achieved_occupancy", "sm__cycles_elapsed.max")
    keep = [i for i, h in enumerate(hdr) if i > first and any(w in h for w in want)]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(rows) - 2)])
        w.writerow(["Kernel Name", ""] + [short(r[first]) for r in rows[2:]])
        for i in keep:
            w.writerow([hdr[i], units[i]] + [r[i] for r in rows[2:]])


def traffic(path, out, kernel):
    rows = rows_of(path)
    hdr = rows[0]
    ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}
    rd, wr, tm = [], [], []
    for r in rows[1:]:
        if kernel not in r[ik]:
            continue
        v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1)
        {"dram__bytes_read.sum": rd, "dram__bytes_write.sum": wr, "gpu__time_duration.sum": tm}.get(r[im], []).append(v)
    n = max(len(rd), 1)
    json.dump({"kernel": kernel, "launches": len(rd), "avg_dram_read_bytes": sum(rd) / n, "avg_dram_write_bytes": sum(wr) / n,
               "avg_traffic_bytes": (sum(rd) + sum(wr)) / n, "avg_time_ms_under_ncu": sum(tm) / max(len(tm), 1)}, open(out, "w"), indent=1)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    elif mode == "metrics":
        metrics(sys.argv[2], sys.argv[3])
    else:
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
