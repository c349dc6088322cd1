#!/bin/bash
# Launch lists of the bench command (ncu --metrics gpu__time_duration.sum, the recipe of B200_PROFILING.md), aggregated per
# kernel: profiles/r02_launches_<W>.txt holds name, launches, total and mean microseconds and the share of the kernel time.
# Per-launch times under ncu are cold-cache and serialised: the SHARES are what must agree with the bench, not the absolutes.
mkdir -p gpurun_out
for w in ${WORKLOADS:-C2 C3 C4 C5}; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_$w.csv \
      python bench.py --workload $w --steps 40 --warmup 3 --configs none --no-cpu-baseline --no-compact > gpurun_out/r02_launches_$w.json 2> gpurun_out/r02_launches_$w.err
  python scripts/aggregate_launches.py gpurun_out/r02_launches_$w.csv "$w: python bench.py --workload $w --steps 40 --warmup 3 --configs none --no-cpu-baseline --no-compact (first 600 launches)" > gpurun_out/r02_launches_$w.txt
  tail -8 gpurun_out/r02_launches_$w.txt
done
