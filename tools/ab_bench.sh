#!/bin/bash
# runs bench.py (no CPU baseline leg, few steps) for every gpurun_ab/*.so and the in-tree build; one JSON summary line each
cd "$(dirname "$0")/.."
for so in bzip2_rust_b200/libbz2b200.so gpurun_ab/*.so; do
  BZ2B200_LIB=$PWD/$so python bench.py --steps 5 --warmup 3 --no-verify --cpu-sample-mb 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k={x[0]:x[1] for x in d['kernels']}
print('$so $BZ2B200_REFINE_GROUP', 'value %.0f e2e %.0f dev_ms %.2f bwt %.2f' % (d['value'], d['e2e']['value'], d['device_ms_per_step'], d['stage_ms']['bwt']), {n:k.get(n) for n in ('k_sweep_gather','k_sweep_carry','k_byte_hist','k_key8_table','k_init_ranks','k_refine_local','k_mtf_emit','k_bwt_out')})
"
done
