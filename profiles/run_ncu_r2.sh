#!/bin/bash
# Run on the GPU box (under gpurun): (1) launch list of a short bench run at the benchmark size, (2) DRAM bytes of
# the dominant kernel at the benchmark size, (3) one full capture at 2048^2.  Outputs land in gpurun_out/ and are
# summarised into profiles/r2 (profiles/ncu_summary.py, profiles/launch_summary.py).  Numbers printed by bench.py
# under ncu are never bench values.
set -x
TAG=${TAG:-r2}; ARGS=${ARGS:-}; BIG=${BIG:-8192}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_${BIG}.csv \
    python bench.py --nx $BIG --nz $BIG --steps 3 --warmup 3 --no-cpu --no-configs --generic-n 0 --fint-reps 2 $ARGS > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_elem_strip -s 3 -c 3 --csv \
    --log-file gpurun_out/traffic_${TAG}_${BIG}.csv \
    python bench.py --nx $BIG --nz $BIG --steps 3 --warmup 3 --no-cpu --no-configs --generic-n 0 --fint-reps 2 $ARGS > gpurun_out/bench_under_ncu_traffic_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_elem_strip -s 4 -c 2 -f -o /tmp/prof_${TAG} \
    python bench.py --nx 2048 --nz 2048 --steps 3 --warmup 3 --no-cpu --no-configs --generic-n 0 --fint-reps 2 $ARGS > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
ncu -i /tmp/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv
ls -la gpurun_out
