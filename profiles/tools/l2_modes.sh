#!/bin/bash
# A/B of the L2 policy knob (experiment build): bit 0 spectra evict-first, bit 1 planes evict-last — kernel alone and whole step
export HSR_B200_EXPERIMENTAL_LIB=1
for round in 1 2; do
  for v in 1 2 3 0; do
    echo "--- HSR_L2_STREAM=$v (round $round)"
    HSR_L2_STREAM=$v timeout 100 python profiles/prof_invalid_knobs.py 2>&1 | tail -1 | cut -c1-110
    HSR_L2_STREAM=$v timeout 100 python profiles/prof_step_graph.py 300 2>&1 | grep "ms per step"
  done
done
