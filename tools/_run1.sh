python -m pytest tests/test_stream_gpu.py tests/test_shard_gpu.py tests/test_golden.py -x -q 2>&1 | tail -3
python tools/kernel_times.py 100 text 9 1 > gpurun_out/kt5.json 2>gpurun_out/kt5.err
python tools/kernel_times.py 256 rep 9 1 > gpurun_out/kt5r.json 2>gpurun_out/kt5r.err
python - <<'PY'
import sys,json
for f in ('gpurun_out/kt5.json','gpurun_out/kt5r.json'):
    r=json.loads(open(f).read())
    print(r['corpus'], r['adler'], r['stage_ms'], r.get('libbz2_roundtrip'), r['bwt_stats'])
    print('   ', r['kernels'][:24])
PY
