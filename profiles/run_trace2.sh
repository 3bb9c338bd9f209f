# host-side trace of the stream group: where the submitting thread's time goes, resident vs end to end
mkdir -p gpurun_out
PLVIWO_BENCH_STREAMS=64 PLVIWO_GROUP_TRACE=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/q_trace2.json 2> gpurun_out/q_trace2.err
python -c "
import json
d=json.loads(open('gpurun_out/q_trace2.json').read()); print(round(d['value']), round(d['e2e']['value']))"
grep "group trace" gpurun_out/q_trace2.err
