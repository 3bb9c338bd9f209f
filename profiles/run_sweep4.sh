# one-GPU experiments (not bench numbers; rows are wrong with PLVIWO_EXP_SKIP): which part of the line path costs the throughput
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']))
except Exception as e: print('$tag', 'ERR', e)"; }
run nowalk PLVIWO_BENCH_STREAMS=64 PLVIWO_EXP_SKIP=1
run noassoc PLVIWO_BENCH_STREAMS=64 PLVIWO_EXP_SKIP=2
run nowalk_noassoc PLVIWO_BENCH_STREAMS=64 PLVIWO_EXP_SKIP=3
run nofront_lines PLVIWO_BENCH_STREAMS=64 PLVIWO_EXP_SKIP=4
run walk1 PLVIWO_BENCH_STREAMS=64 PLVIWO_WALK_CTAS=1
run walk4 PLVIWO_BENCH_STREAMS=64 PLVIWO_WALK_CTAS=4
