python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/kernel_times.py 1000 text 9 1 > gpurun_out/kt9_1g.json 2>gpurun_out/kt9.err
python - <<'PY'
import json
r=json.loads(open('gpurun_out/kt9_1g.json').read())
print(r['corpus'], r['mb'], r['wall_ms_untimed_mode'], r['stage_ms'], r.get('libbz2_roundtrip'), r['z'])
PY
python tools/run_configs.py decode 2>&1 | tail -1 | cut -c1-800
