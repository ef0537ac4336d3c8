#!/bin/bash
# Run on the GPU box (under gpurun): launch list of a short bench run + one full capture of the
# dominant kernel.  Outputs land in gpurun_out/ and are summarised into profiles/ by hand
# (profiles/ncu_summary.py, profiles/launch_summary.py).
set -x
NX=${NX:-2048}; NZ=${NZ:-2048}; TAG=${TAG:-r1}; ARGS=${ARGS:-}
mkdir -p gpurun_out
if [ -z "$SKIP_LIST" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --nx $NX --nz $NZ --steps 3 --warmup 3 --no-cpu --fint-reps 2 $ARGS > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
fi
ncu --set full --clock-control none --import-source on -k regex:k_elem_strip -s 4 -c 2 -f -o gpurun_out/prof_${TAG} \
    python bench.py --nx $NX --nz $NZ --steps 3 --warmup 3 --no-cpu --fint-reps 2 $ARGS > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv
ls -la gpurun_out
