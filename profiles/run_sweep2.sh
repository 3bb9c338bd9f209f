# one-GPU experiment sweep (not bench numbers): ticks in flight, copy streams, front batch size after FAST moved on demand
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/q_pytest.txt 2>&1; tail -3 gpurun_out/q_pytest.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']), {k: round(v['avg_ms']*1000) for k,v in d['roofline'].get('per_kernel',{}).items()})
except Exception as e: print('$tag', 'ERR', e)"; }
run la12 PLVIWO_BENCH_STREAMS=64
run la12_c1 PLVIWO_BENCH_STREAMS=64 PLVIWO_GROUP_COPY_STREAMS=1
run la20 PLVIWO_BENCH_STREAMS=64 PLVIWO_BENCH_GROUP_LA=20
run la28 PLVIWO_BENCH_STREAMS=64 PLVIWO_BENCH_GROUP_LA=28
run la24_b2 PLVIWO_BENCH_STREAMS=64 PLVIWO_BENCH_GROUP_LA=24 PLVIWO_GROUP_FRONT_TICKS=2
run la24_w512 PLVIWO_BENCH_STREAMS=64 PLVIWO_BENCH_GROUP_LA=24 PLVIWO_WALK_CTAS=2
