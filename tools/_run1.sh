python -m pytest tests/test_mtf_gpu.py tests/test_compress_block_gpu.py tests/test_golden.py -x -q 2>&1 | tail -2
for c in "100 text" "128 mixed"; do python tools/kernel_times.py $c 9 1 2>/dev/null | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print(r['corpus'], r['adler'], r['stage_ms']['mtf_rle2'], r.get('libbz2_roundtrip'), [k for k in r['kernels'] if 'mtf' in k[0]])"; done
