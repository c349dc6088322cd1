#!/bin/bash
# C2 (SingleSnake size 9, 2^20 envs, partial_2, dense state): one-shot tile CTAs (the shipped kernel) against the persistent
# multi-stage TMA ring (WURM_SINGLE_RING=stages, WURM_SINGLE_RING_CTAS=CTAs per SM); fused step+reset launches.
echo "shipped: $(python scripts/time_compact.py C2 400 dense | cut -c1-60)"
for st in 2 3 4; do for c in 6 8 12 16; do
  echo "ring stages $st, $c CTAs/SM: $(WURM_SINGLE_RING=$st WURM_SINGLE_RING_CTAS=$c python scripts/time_compact.py C2 400 dense | cut -c1-60)"
done; done
