#!/usr/bin/env bash
# Single-GPU evidence pack of one round: bench lines (default = C3 + c2, C1, C5, STRICT), ncu launch lists of the same
# commands' kernels, and ncu --set full captures of the dominant kernels. Run under gpurun; outputs in gpurun_out/<tag>_*.
# usage: scripts/collect_profiles.sh r02
set -uo pipefail
TAG="${1:-r02}"
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
python bench.py --workload c1 --steps 100 --warmup 10 --no-cpu-baseline > $O/${TAG}_bench_c1.json 2>> $O/${TAG}_bench_n1.err
python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c5.json 2>> $O/${TAG}_bench_n1.err
python bench.py --workload c2 --precision strict --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c2_strict.json 2>> $O/${TAG}_bench_n1.err
python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > $O/${TAG}_bench_c4_n1.json 2>> $O/${TAG}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench_n1.err
for w in c2 c3; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_$w.csv \
      python scripts/prof_step.py $w 3 > $O/${TAG}_ncu_$w.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches_dd8_emulated.csv \
    python scripts/prof_dd.py c3 8 2 > $O/${TAG}_ncu_dd8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_traverse2 -s 2 -c 1 -o $O/${TAG}_traverse2_c3 python scripts/prof_step.py c3 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_traverse2 -s 2 -c 1 -o $O/${TAG}_traverse2_c2 python scripts/prof_step.py c2 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_agg_level -s 12 -c 1 -o $O/${TAG}_agg_level_c3 python scripts/prof_step.py c3 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sort_onesweep -s 4 -c 1 -o $O/${TAG}_onesweep_c3 python scripts/prof_step.py c3 3 > /dev/null 2>&1
ls -la $O | grep ${TAG}_ | wc -l
