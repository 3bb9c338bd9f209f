# one-GPU experiments (not bench numbers; rows are wrong with PLVIWO_EXP_*): what the chain walk costs the rest of the pipeline
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']))
except Exception as e: print('$tag', 'ERR', e)"; }
export PLVIWO_BENCH_STREAMS=64 PLVIWO_BENCH_GROUP_LA=24 PLVIWO_WALK_CTAS=4
run noseg PLVIWO_EXP_SKIP=16
run pad30 PLVIWO_EXP_WALK_PAD_KB=30
run pad70 PLVIWO_EXP_WALK_PAD_KB=70
run w8 PLVIWO_WALK_CTAS=8
run w16 PLVIWO_WALK_CTAS=16
