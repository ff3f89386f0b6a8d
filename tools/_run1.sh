python bench.py --steps 3 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; cut -c1-700 gpurun_out/bench6.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_v3.csv python bench.py --steps 1 --warmup 3 --no-verify > gpurun_out/ncu_l.log 2>&1
wc -l gpurun_out/r01_launches_v3.csv
