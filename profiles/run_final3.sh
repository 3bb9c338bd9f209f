# final bench line of the round on one B200 (+ the tests that cover the line path of single handles and groups)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_frontend_gpu.py tests/test_group_gpu.py -x -q -m gpu > gpurun_out/pytest_gpu_final_b.txt 2>&1; tail -3 gpurun_out/pytest_gpu_final_b.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_1gpu_final.json 2> gpurun_out/bench_1gpu_final.err; tail -3 gpurun_out/bench_1gpu_final.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_1gpu_final.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline'].get('value'), 'launches', d['gpu_launches'], 'roofline', d['roofline'].get('kernel'), d['roofline'].get('frac'))
print('extras', {k: (v if not isinstance(v, dict) else v.get('value')) for k, v in d.get('extras', {}).items()})
"
