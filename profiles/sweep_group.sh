# one-GPU sweeps of the stream group's pipeline knobs (experiments, not bench numbers): bash profiles/sweep_group.sh under gpurun
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/sw_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']))
except Exception as e: print('$tag', 'ERR', e)"; }
run s64_def PLVIWO_BENCH_STREAMS=64
run s32_def PLVIWO_BENCH_STREAMS=32
run s16_def PLVIWO_BENCH_STREAMS=16
run s8_def PLVIWO_BENCH_STREAMS=8
run s8_l8 PLVIWO_BENCH_STREAMS=8 PLVIWO_GROUP_LANES=8
