#!/bin/bash
# ncu evidence for the stream group (run on the GPU box through gpurun; summaries come back under gpurun_out/):
#   bash profiles/capture_group.sh <tag>
# 1. launch list of a short group run (64 streams, 6 ticks, 64 frames per front launch, 16 streams per lane launch)
# 2. --set full capture of the kernels of one steady-state tick, summarised ON THE BOX (the report itself is ~100 MB with 60
#    kernels; gpurun brings back 64 MB at most): profiles/summarise_group.py writes the text files, the raw page goes to a csv
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_group_$TAG.csv \
    python profiles/profile_group.py 64 6 2 > $OUT/launches_group_$TAG.log 2>&1
tail -1 $OUT/launches_group_$TAG.log
ncu --set full --clock-control none -k regex:'^(void )?k_' -s 180 -c 56 -f \
    -o $OUT/full_group_$TAG python profiles/profile_group.py 64 5 2 > $OUT/full_group_$TAG.log 2>&1
tail -2 $OUT/full_group_$TAG.log
ncu -i $OUT/full_group_$TAG.ncu-rep --page raw --csv > $OUT/full_group_${TAG}_raw.csv 2>/dev/null
python profiles/summarise_group.py $TAG > $OUT/summarise_group_$TAG.log 2>&1
cp profiles/launches_group_$TAG.txt profiles/ncu_${TAG}_metrics.txt profiles/traffic.json $OUT/ 2>/dev/null
ls -la $OUT/full_group_$TAG.ncu-rep
rm -f $OUT/full_group_$TAG.ncu-rep
ls -la $OUT | grep -E "$TAG"
