#!/usr/bin/env python
"""Summarises the ncu outputs of profiles/capture.sh into small tracked text files.

    python profiles/summarise.py <tag>     (reads gpurun_out/launches_<tag>.csv and gpurun_out/full_<tag>.ncu-rep)

Writes profiles/launches_<tag>.txt (per-kernel launch count, mean / total device time, share of the step) and
profiles/kernels_<tag>.txt (per kernel: duration, DRAM bytes read + written per launch, DRAM and SM throughput %,
achieved occupancy, registers, shared memory) — the numbers DESIGN.md and bench.py's roofline.traffic quote."""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def short(name):
    return name.split("(")[0].replace("plviwo::", "")


def launches():
    path = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", ""))
            if d.get("Metric Unit", "ns").startswith("us"):
                v *= 1e3
            agg.setdefault(short(d["Kernel Name"]), []).append(v)
    total = sum(sum(v) for v in agg.values())
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 40 --warmup 5 (tag %s)" % tag,
           "# cold-cache, serialised launches: compare SHARES, not absolutes", "%-22s %7s %12s %12s %7s" %
           ("kernel", "n", "mean_us", "total_ms", "share")]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append("%-22s %7d %12.2f %12.3f %6.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e6, 100 * sum(v) / total))
    open(os.path.join(ROOT, "profiles", "launches_%s.txt" % tag), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


WANT = collections.OrderedDict([
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_static", "smem_s"),
    ("launch__shared_mem_per_block_dynamic", "smem_d"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
])


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
    for k, m in mult.items():
        if u == k:
            return v * m
    return v


def full():
    rep = os.path.join(ROOT, "gpurun_out", "full_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        return
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[2:]:
        name = short(r[col["Kernel Name"]])
        rec = {}
        for m, key in WANT.items():
            if m not in col:
                continue
            val, unit = r[col[m]], units[col[m]]
            try:
                if key in ("dram_rd", "dram_wr"):
                    rec[key] = to_bytes(val, unit)
                elif key == "dur_us":
                    v = float(val.replace(",", ""))
                    rec[key] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
                else:
                    rec[key] = float(val.replace(",", ""))
            except ValueError:
                pass
        per.setdefault(name, []).append(rec)
    out = ["# ncu --set full --clock-control none, profiles/profile_driver.py (BASELINE.json configs[1], direct launches), tag %s" % tag,
           "# means over the captured launches of each kernel; dram_rd/dram_wr in bytes per launch",
           "%-18s %3s %9s %10s %10s %6s %6s %6s %6s %6s %5s %7s %7s %6s %6s" %
           ("kernel", "n", "dur_us", "dram_rd", "dram_wr", "dram%", "sm%", "l1%", "l2%", "occ%", "regs", "smem_s", "smem_d", "grid", "block")]
    for k, recs in per.items():
        def mean(key):
            v = [r[key] for r in recs if key in r]
            return sum(v) / len(v) if v else float("nan")
        out.append("%-18s %3d %9.2f %10.0f %10.0f %6.1f %6.1f %6.1f %6.1f %6.1f %5.0f %7.0f %7.0f %6.0f %6.0f" %
                   (k, len(recs), mean("dur_us"), mean("dram_rd"), mean("dram_wr"), mean("dram%"), mean("sm%"), mean("l1%"),
                    mean("l2%"), mean("occ%"), mean("regs"), mean("smem_s"), mean("smem_d"), mean("grid"), mean("block")))
    open(os.path.join(ROOT, "profiles", "kernels_%s.txt" % tag), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    import json
    traffic = {}
    for k, recs in per.items():
        v = [r.get("dram_rd", 0.0) + r.get("dram_wr", 0.0) for r in recs]
        traffic[k] = sum(v) / len(v)
    json.dump({"source": "gpurun_out/full_%s.ncu-rep (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)" % tag,
               "kernels": traffic}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)


launches()
full()
