#!/bin/bash
# quick look at the config-2 stage times (run on a B200 box):  bash tools/k1_quick.sh
python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('step', round(d['ms_per_step'],3), 'serial', round(d['serial_ms_per_step'],3), 'frac', round(r['frac'],3), r['kernel'], 'clk', d['clocks']['sm_mhz'])
print(' '.join('%s %.3f' % (k.split()[0], v['ms']) for k, v in d['kernels'].items()))
print(json.dumps(r.get('largest_kernel_by_time')))"
