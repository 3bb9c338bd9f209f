#!/usr/bin/env python
"""Summarises the ncu outputs of profiles/capture_group.sh into tracked text files.

    python profiles/summarise_group.py <tag>

Reads gpurun_out/launches_group_<tag>.csv and gpurun_out/full_group_<tag>.ncu-rep; writes profiles/launches_group_<tag>.txt
(per-kernel launch count, mean / max duration, share of the serialised kernel time), profiles/ncu_<tag>_metrics.txt (per kernel:
duration, DRAM bytes per launch, DRAM / SM / L1 / L2 throughput, issue-slot utilisation, achieved occupancy, registers, shared
memory, grid) and the "group_kernels" entry of profiles/traffic.json (DRAM bytes per FRAME per bench.py stage, which
bench.py's roofline.traffic quotes)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2z"
FRAMES_FRONT, FRAMES_LANE = 64, 16     # profiles/profile_group.py 64 streams, lookahead 2: one tick per front launch, 4 lanes


def short(name):
    name = name.split("(")[0].replace("plviwo::", "").replace("void ", "")
    return name.split("<")[0]


WANT = collections.OrderedDict([
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__inst_executed.sum", "winst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_static", "smem_s"),
    ("launch__shared_mem_per_block_dynamic", "smem_d"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
])
STAGE = {"k_hist_b": ("hist", FRAMES_FRONT), "k_eq_pyr1_b": ("eq_pyr1", FRAMES_FRONT), "k_pyr_down_b": ("pyr_rest", FRAMES_FRONT),
         "k_canny": ("canny", FRAMES_FRONT), "k_ccl_border": ("ccl", FRAMES_FRONT), "k_ccl_link": ("ccl", FRAMES_FRONT),
         "k_ccl_roots": ("ccl", FRAMES_FRONT), "k_fld_walk_cc": ("walk", FRAMES_FRONT), "k_fld_order": ("segments", FRAMES_FRONT),
         "k_fld_segments": ("segments", FRAMES_FRONT), "k_fld_compact": ("segments", FRAMES_FRONT), "k_fast_g": ("fast", FRAMES_LANE),
         "k_fast_select_g": ("select", FRAMES_LANE), "k_corner_subpix_g": ("subpix", FRAMES_LANE), "k_lk15w_g": ("lk", FRAMES_LANE),
         "k_group_detect": ("detect", FRAMES_LANE), "k_group_cands": ("cands", FRAMES_LANE), "k_group_accept": ("accept", FRAMES_LANE),
         "k_group_gate": ("gate", FRAMES_LANE), "k_group_lines": ("lines", FRAMES_LANE)}


def to_num(val, unit, key):
    v = float(val.replace(",", ""))
    u = unit.lower()
    if key in ("dram_rd", "dram_wr"):
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    if key == "dur_us":
        return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
    return v


def launches():
    path = os.path.join(ROOT, "gpurun_out", "launches_group_%s.csv" % tag)
    if not os.path.exists(path):
        return
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "summarise_launches.py"), path], capture_output=True, text=True).stdout
    hdr = ("# ncu --metrics gpu__time_duration.sum --clock-control none, profiles/profile_group.py 64 6 2 (64 streams, 6 ticks; a front launch\n"
           "# carries 64 frames, a lane launch 16 streams), tag %s.  Cold-cache, serialised launches: compare SHARES, not absolutes.\n" % tag)
    open(os.path.join(ROOT, "profiles", "launches_group_%s.txt" % tag), "w").write(hdr + txt)
    print(txt)


def full():
    rep = os.path.join(ROOT, "gpurun_out", "full_group_%s.ncu-rep" % tag)
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
            if m in col:
                try:
                    rec[key] = to_num(r[col[m]], units[col[m]], key)
                except ValueError:
                    pass
        per.setdefault(name, []).append(rec)
    out = ["# ncu --set full --clock-control none, profiles/profile_group.py 64 5 2 (one steady-state tick of 64 streams: a front launch",
           "# carries 64 frames, a lane launch 16 streams), tag %s.  Means over the captured launches of each kernel; dram_rd / dram_wr in" % tag,
           "# bytes per launch; winst = warp instructions executed per launch.",
           "%-18s %3s %9s %11s %11s %6s %6s %6s %6s %6s %6s %11s %5s %7s %7s %7s %6s" %
           ("kernel", "n", "dur_us", "dram_rd", "dram_wr", "dram%", "sm%", "l1%", "l2%", "issue%", "occ%", "winst", "regs", "smem_s", "smem_d", "grid", "block")]
    stage = collections.OrderedDict()
    for k, recs in per.items():
        def mean(key):
            v = [r[key] for r in recs if key in r]
            return sum(v) / len(v) if v else float("nan")
        out.append("%-18s %3d %9.2f %11.0f %11.0f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f %11.0f %5.0f %7.0f %7.0f %7.0f %6.0f" %
                   (k, len(recs), mean("dur_us"), mean("dram_rd"), mean("dram_wr"), mean("dram%"), mean("sm%"), mean("l1%"), mean("l2%"),
                    mean("issue%"), mean("occ%"), mean("winst"), mean("regs"), mean("smem_s"), mean("smem_d"), mean("grid"), mean("block")))
        if k in STAGE:
            st, frames = STAGE[k]
            n_per_tick = 3 if k == "k_pyr_down_b" else 1          # three pyramid levels per frame
            stage[st] = stage.get(st, 0.0) + (mean("dram_rd") + mean("dram_wr")) * n_per_tick / frames
    open(os.path.join(ROOT, "profiles", "ncu_%s_metrics.txt" % tag), "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        tr = json.load(open(tpath))
    except Exception:
        tr = {}
    tr["group_source"] = "gpurun_out/full_group_%s.ncu-rep (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch / frames per launch)" % tag
    tr["group_kernels"] = {st: {"dram_bytes_per_frame": v, "note": "ncu --set full, tag %s, cold-cache single launch" % tag} for st, v in stage.items()}
    json.dump(tr, open(tpath, "w"), indent=1)


launches()
full()
