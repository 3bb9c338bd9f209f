# full GPU test suite + a traced bench run (front-batch latency vs interval), one B200
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_full.txt 2>&1; tail -3 gpurun_out/pytest_gpu_full.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']))
except Exception as e: print('$tag', 'ERR', e)"; grep "group trace" gpurun_out/q_$tag.err; }
export PLVIWO_BENCH_STREAMS=64 PLVIWO_GROUP_TRACE=1
run tr_la24
run tr_la12 PLVIWO_BENCH_GROUP_LA=12
run tr_la24_nolines PLVIWO_BENCH_NO_LINES=1
