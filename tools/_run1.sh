rm -f gpurun_out/kt4.jsonl
for g in 4 8 16 32 64 112; do
  BZ2B200_SWEEP_GROUP=$g python tools/kernel_times.py 100 text 9 >> gpurun_out/kt4.jsonl 2>gpurun_out/kt4.err
done
python - <<'PY'
import sys,json
for l in open('gpurun_out/kt4.jsonl'):
    r=json.loads(l)
    print(r['env'], r['adler'], r['stage_ms']['bwt'], [k for k in r['kernels'] if k[0].startswith('k_sweep')])
PY
python bench.py --steps 3 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; cat gpurun_out/bench4.json | cut -c1-1200
