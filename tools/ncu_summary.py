#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report: one row per captured launch with the metrics the roofline
arithmetic uses (B200_PROFILING.md).  Runs where there is no GPU:  python tools/ncu_summary.py gpurun_out/x.ncu-rep
Optional: --sass <kernel regex> adds the opcode mix and stall-reason totals of the first matching kernel."""
import csv
import io
import subprocess
import sys
from collections import Counter

METRICS = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "DRAM rd GB"), ("dram__bytes_write.sum", "DRAM wr GB"),
           ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
           ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/LSU %"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
           ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp inst"), ("launch__grid_size", "grid")]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_float(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    scale = {"us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
    return x * scale.get(unit, 1.0)


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n in METRICS) + " |")
    print("|---|" + "---|" * len(METRICS))
    for r in rows:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "").strip()
        cells = []
        for m, _ in METRICS:
            if m not in ix:
                cells.append("-")
                continue
            v = to_float(r[ix[m]], units[ix[m]])
            cells.append("-" if v is None else ("%.3g" % v))
        print("| %s | %s |" % (name[:60], " | ".join(cells)))
    if "--sass" in sys.argv:
        pat = sys.argv[sys.argv.index("--sass") + 1]
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + pat],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[2:] if len(r) > ix["Instructions Executed"] and r[ix["Instructions Executed"]].isdigit()]
        tot = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
        ops = Counter()
        for r in data:
            toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
            ops[toks[0].split(".")[0]] += int(r[ix["Instructions Executed"]])
        print("\nSASS of `%s`: %d static instructions; executed mix: " % (pat, len(data)) +
              ", ".join("%s %.1f%%" % (o, 100.0 * n / tot) for o, n in ops.most_common(12)))
        stalls = {k: sum(int(r[ix[k]]) for r in data) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        st = sum(stalls.values()) or 1
        print("stall samples: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / st) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main()
