# one-GPU experiments (not bench numbers): hardware work queues (18 CUDA streams share CUDA_DEVICE_MAX_CONNECTIONS = 8 queues by default)
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; python -c "
import json
try:
    d=json.loads(open('gpurun_out/q_$tag.json').read()); print('$tag', round(d['value']), round(d['e2e']['value']))
except Exception as e: print('$tag', 'ERR', e)"; }
run conn32 PLVIWO_BENCH_STREAMS=64 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_w4 PLVIWO_BENCH_STREAMS=64 CUDA_DEVICE_MAX_CONNECTIONS=32 PLVIWO_WALK_CTAS=4
run conn32_la24 PLVIWO_BENCH_STREAMS=64 CUDA_DEVICE_MAX_CONNECTIONS=32 PLVIWO_BENCH_GROUP_LA=24
run conn32_s8 PLVIWO_BENCH_STREAMS=8 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_s16 PLVIWO_BENCH_STREAMS=16 CUDA_DEVICE_MAX_CONNECTIONS=32
