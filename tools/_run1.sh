python -m pytest tests/test_decode_gpu.py -x -q 2>&1 | tail -15
python tools/run_configs.py decode 2>&1 | tail -2 | cut -c1-900
