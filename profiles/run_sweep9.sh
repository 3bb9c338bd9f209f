# one-GPU experiments (not bench numbers): LK with a 1.5 KB shared buffer per feature, 5 / 6 CTAs per SM
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group_gpu.py -x -q -m gpu > gpurun_out/q_pytest.txt 2>&1; tail -3 gpurun_out/q_pytest.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']), {k: round(v['avg_ms']*1000) for k,v in d['roofline'].get('per_kernel',{}).items() if k in ('walk','lk','cands','canny','ccl','segments')})
except Exception as e: print('$tag', 'ERR', e)"; }
export PLVIWO_BENCH_STREAMS=64 PLVIWO_BENCH_GROUP_LA=24
run occ5
run occ6 PLVIWO_LK_OCC=6
run occ6_nolines PLVIWO_LK_OCC=6 PLVIWO_BENCH_NO_LINES=1
run occ6_l8 PLVIWO_LK_OCC=6 PLVIWO_GROUP_LANES=8
