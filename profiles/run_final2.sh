# final evidence of the round on one B200: full GPU suite, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.txt 2>&1; tail -3 gpurun_out/pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.txt 2>&1; tail -1 gpurun_out/smoke_final.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_1gpu_final.json 2> gpurun_out/bench_1gpu_final.err; tail -3 gpurun_out/bench_1gpu_final.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_1gpu_final.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline'].get('value'), 'launches', d['gpu_launches'], 'roofline', d['roofline'].get('kernel'), d['roofline'].get('frac'), d['roofline'].get('whole_frame'))
print('extras', {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in list(v.items())[:4]}) for k, v in d.get('extras', {}).items()})
"
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; tail -3 gpurun_out/bench_ref_final.err; tail -c 600 gpurun_out/bench_ref_final.json
