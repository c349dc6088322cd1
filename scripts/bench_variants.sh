#!/bin/bash
# Bench each scratch_libs/lib_<name>.so in turn on workload $W (default C2); restores the built library afterwards.
cp wurm_b200/_C/libwurm_b200.so /tmp/lib_keep.so
for f in scratch_libs/lib_*.so; do
  n=$(basename $f .so); cp $f wurm_b200/_C/libwurm_b200.so
  python bench.py --workload ${W:-C2} --steps ${STEPS:-300} --warmup 20 --no-cpu-baseline | python scripts/bench_line.py $n
done
cp /tmp/lib_keep.so wurm_b200/_C/libwurm_b200.so
