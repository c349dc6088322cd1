#!/bin/bash
# Times a few MultiSnake geometries (default CTA sizes) with each scratch_libs/lib_<name>.so in turn.
cp wurm_b200/_C/libwurm_b200.so /tmp/lib_keep.so
for f in scratch_libs/lib_*.so; do
  n=$(basename $f .so); cp $f wurm_b200/_C/libwurm_b200.so
  for g in "65536 4 25" "16384 10 36" "32768 2 12" "8192 8 48" "32768 16 64" "16384 4 40"; do
    echo "$n E,K,S=$g: $(python scripts/sweep_multi.py $g)"
  done
done
cp /tmp/lib_keep.so wurm_b200/_C/libwurm_b200.so
