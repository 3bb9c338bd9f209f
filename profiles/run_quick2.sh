# quick check on one B200: kernel + group + front-end (line rows vs the oracle) tests, then short bench runs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_group_gpu.py tests/test_frontend_gpu.py -x -q -m gpu > gpurun_out/q_pytest.txt 2>&1; tail -3 gpurun_out/q_pytest.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']), {k: round(v['avg_ms']*1000) for k,v in d['roofline'].get('per_kernel',{}).items()})
except Exception as e: print('$tag', 'ERR', e)"; }
run s64 PLVIWO_BENCH_STREAMS=64
run s64_coop PLVIWO_BENCH_STREAMS=64 PLVIWO_WALK_COOP=1
