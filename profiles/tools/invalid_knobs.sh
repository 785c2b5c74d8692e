#!/bin/bash
# the fused gather + SRF kernel under the experiment build's knobs: granule, all-nodata grid, fixed cost
export HSR_B200_EXPERIMENTAL_LIB=1
run() { env "$@" timeout 120 python profiles/prof_invalid_knobs.py 2>&1 | tail -1; }
run A=1
run HSR_DRY_CONSUMER=1
run HSR_DRY_CONSUMER=2
run HSR_DRY_CONSUMER=4
run HSR_GLT_NO_RING=1
run HSR_STAGES=4
run HSR_STAGES=3
run HSR_L2_STREAM=0
