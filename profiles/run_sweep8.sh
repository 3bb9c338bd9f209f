# one-GPU experiments (not bench numbers): small components walked by a launch of their own (11 KB CTAs)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/q_pytest.txt 2>&1; tail -3 gpurun_out/q_pytest.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']), {k: round(v['avg_ms']*1000) for k,v in d['roofline'].get('per_kernel',{}).items() if k in ('walk','cands','canny','ccl','segments')})
except Exception as e: print('$tag', 'ERR', e)"; }
export PLVIWO_BENCH_STREAMS=64 PLVIWO_BENCH_GROUP_LA=24
run b2s2
run b1s2 PLVIWO_WALK_CTAS=1
run b1s4 PLVIWO_WALK_CTAS=1 PLVIWO_WALK_CTAS_SMALL=4
run b2s4 PLVIWO_WALK_CTAS_SMALL=4
run b4s4 PLVIWO_WALK_CTAS=4 PLVIWO_WALK_CTAS_SMALL=4
run b2s2_la12 PLVIWO_BENCH_GROUP_LA=12
