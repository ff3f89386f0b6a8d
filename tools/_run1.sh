python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; cat gpurun_out/bench5.json | cut -c1-1500
python tools/run_configs.py > gpurun_out/configs5.jsonl 2> gpurun_out/configs5.err; cut -c1-400 gpurun_out/configs5.jsonl
