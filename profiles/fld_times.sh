#!/bin/bash
# per-launch device times of the line-extraction kernels over 8 frames (scratch measurement, serialised launches)
PLVIWO_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"${1:-k_fld|k_ccl}" --csv --log-file gpurun_out/fld_times.csv python profiles/profile_driver.py 8 0 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/fld_times.csv")))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        agg.setdefault(d["Kernel Name"].split("(")[0],[]).append(float(d["Metric Value"].replace(",",""))/1e3)
for k,v in agg.items(): print(k, ["%.1f"%x for x in v])
PY
