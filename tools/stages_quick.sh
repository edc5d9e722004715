#!/bin/bash
# stage times of configs 2, 3, 4 on a B200 box:  bash tools/stages_quick.sh [configs...]
for c in ${@:-2 3 4}; do
python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-subrecords 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('config $c step', round(d['ms_per_step'],3), 'frac', round(r['frac'],3), r['kernel'], '|', ' '.join('%s %.3f' % (k.split()[0], v['ms']) for k, v in d['kernels'].items()))"
done
