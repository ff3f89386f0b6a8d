python -m pytest tests/test_stream_gpu.py tests/test_shard_gpu.py tests/test_golden.py tests/test_compress_block_gpu.py -x -q 2>&1 | tail -3
python tools/kernel_times.py 100 text 9 1 > gpurun_out/kt8.json 2>gpurun_out/kt8.err
python tools/kernel_times.py 256 rep 9 1 > gpurun_out/kt8r.json 2>gpurun_out/kt8r.err
python tools/kernel_times.py 256 mixed 5 1 > gpurun_out/kt8m.json 2>gpurun_out/kt8m.err
python - <<'PY'
import sys,json
for f in ('gpurun_out/kt8.json','gpurun_out/kt8r.json','gpurun_out/kt8m.json'):
    r=json.loads(open(f).read())
    print(r['corpus'], r['adler'], r['stage_ms'], r.get('libbz2_roundtrip'))
    print('   ', [k for k in r['kernels'] if 'rle' in k[0] or 'crc' in k[0]])
PY
