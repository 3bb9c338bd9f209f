# one-GPU experiments (not bench numbers): after the line association was taken off the point chain (state-record ring)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/q_pytest.txt 2>&1; tail -3 gpurun_out/q_pytest.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']))
except Exception as e: print('$tag', 'ERR', e)"; }
run la12 PLVIWO_BENCH_STREAMS=64
run la24 PLVIWO_BENCH_STREAMS=64 PLVIWO_BENCH_GROUP_LA=24
run la24_w4 PLVIWO_BENCH_STREAMS=64 PLVIWO_BENCH_GROUP_LA=24 PLVIWO_WALK_CTAS=4
run la36_w4 PLVIWO_BENCH_STREAMS=64 PLVIWO_BENCH_GROUP_LA=36 PLVIWO_WALK_CTAS=4
run s8_la24 PLVIWO_BENCH_STREAMS=8
run s8_la48 PLVIWO_BENCH_STREAMS=8 PLVIWO_BENCH_GROUP_LA=48
