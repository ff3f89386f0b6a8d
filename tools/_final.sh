python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; cut -c1-400 gpurun_out/bench7.json
python tools/run_configs.py > gpurun_out/configs7.jsonl 2> gpurun_out/configs7.err; cut -c1-330 gpurun_out/configs7.jsonl
