#!/bin/bash
# ncu evidence for one round (run on the GPU box through gpurun; reports come back under gpurun_out/).
#   bash profiles/capture.sh <tag>
# 1. launch list of a short bench.py run (graph replays are profiled kernel by kernel)
# 2. --set full capture of every kernel of two steady-state frames (direct launches: PLVIWO_NO_GRAPHS=1)
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 800 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-stereo --multi-streams 0 > $OUT/launches_$TAG.bench.log 2>&1
PLVIWO_NO_GRAPHS=1 ncu --set full --clock-control none --import-source on -k regex:'^k_(?!signal)' -s 120 -c 44 -f \
    -o $OUT/full_$TAG python profiles/profile_driver.py 14 2 > $OUT/full_$TAG.log 2>&1
tail -3 $OUT/full_$TAG.log
ls -la $OUT
