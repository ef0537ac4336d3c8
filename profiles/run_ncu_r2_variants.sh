#!/bin/bash
# Full ncu captures of the two kernel variants added late in round 2 (run under gpurun): the Coulomb-plasticity
# instantiation (profiles/bench_plastic.py, 2048^2) and the Kelvin-Voigt + Newmark fused instantiation (the TPV3 deck
# scaled x12 through sem2dsolve_b200).  Summaries: profiles/ncu_summary.py.
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_elem_strip -s 6 -c 1 -f -o /tmp/prof_r2_plastic \
    python profiles/bench_plastic.py 2048 1e-2 1.0 > gpurun_out/bench_under_ncu_plastic.log 2>&1
ncu -i /tmp/prof_r2_plastic.ncu-rep --page raw --csv > gpurun_out/prof_r2_plastic_raw.csv
ncu --set full --clock-control none --import-source on --target-processes all -k regex:k_elem_strip -s 6 -c 1 -f -o /tmp/prof_r2_kv \
    python bench.py --config tpv3 --config-scale 12 --steps 10 > gpurun_out/bench_under_ncu_kv.log 2>&1
ncu -i /tmp/prof_r2_kv.ncu-rep --page raw --csv > gpurun_out/prof_r2_kv_raw.csv
ls -la gpurun_out | tail -8
