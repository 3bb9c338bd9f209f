# traced bench (stages of a front batch under load) + ncu evidence of the stream group
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']))
except Exception as e: print('$tag', 'ERR', e)"; grep "group trace" gpurun_out/q_$tag.err; }
run trace PLVIWO_BENCH_STREAMS=64 PLVIWO_GROUP_TRACE=1
timeout 900 bash profiles/capture_group.sh r2z
