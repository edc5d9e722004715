#!/bin/bash
# quick look at the single-scan latency (config 1) on a B200 box:  bash tools/c1_quick.sh
python bench.py --config 1 --steps 300 --warmup 30 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('device us', round(d['ms_per_step']*1e3,1), 'e2e us', round(d['e2e']['ms_per_step']*1e3,1), 'launches/step', d['gpu_launches']/d['steps'], 'parity', d.get('parity_vs_oracle_on_sample'), d.get('descriptor_parity_on_sample'))
print(' '.join('%s %.1f' % (k.split()[0], v['ms']*1e3) for k, v in d['kernels'].items()))"
