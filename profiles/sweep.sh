#!/bin/bash
# usage: sweep.sh "ENV=.. ENV2=.." ... ; one short bench line per variant (GPU box)
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --nx ${NX:-4096} --nz ${NZ:-4096} --steps 10 --no-cpu $ARGS 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']
print('GDOF/s %.2f step_ms %.3f kernel_ms %.3f frac %.3f k1_ms %.3f k1_frac %.3f e2e %.2f' % (j['value']/1e9, j['ms_per_step'], r['ms_per_launch'], r['frac'], r['k1_alone']['ms_per_launch'], r['k1_alone']['frac'], j['e2e']['value']/1e9))"
done
