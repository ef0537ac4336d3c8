#!/bin/bash
# usage: sweep_libs.sh "<libdir> [ENV=..]" ... ; one short bench line per library build / environment variant
# (GPU box).  <libdir> is a directory under sem2dpack_b200/ holding a differently built libsem2d_b200.so.
for v in "$@"; do
  set -- $v
  lib=$1; shift
  echo "== $lib $*"
  env S2D_LIB_PATH=$PWD/sem2dpack_b200/$lib/libsem2d_b200.so "$@" python bench.py --nx ${NX:-4096} --nz ${NZ:-4096} --steps ${STEPS:-10} --no-cpu --no-configs --generic-n 0 $ARGS 2>&1 | tail -1 | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); r=j['roofline']
    print('GDOF/s %.2f step_ms %.3f kernel_ms %.3f frac %.3f k1_ms %.3f k1_frac %.3f e2e %.2f clk %s %s' % (j['value']/1e9, j['ms_per_step'], r['ms_per_launch'], r['frac'], r['k1_alone']['ms_per_launch'], r['k1_alone']['frac'], j['e2e']['value']/1e9, j['clocks']['sm_mhz'], j['clocks']['reasons']))
except Exception as ex:
    print('FAILED', ex)"
done
