#!/bin/bash
# the fused gather + SRF kernel under the experiment build's knobs: granule, all-nodata grid, fixed cost
export HSR_B200_EXPERIMENTAL_LIB=1
run() { env "$@" timeout 120 python profiles/prof_invalid_knobs.py 2>&1 | tail -1; }
run A=1
run HSR_GLT_BULK=1
run HSR_DRY_CONSUMER=8
run HSR_DRY_CONSUMER=16
run HSR_DRY_CONSUMER=24
run HSR_DRY_CONSUMER=32
run HSR_DRY_CONSUMER=96
run HSR_DRY_CONSUMER=120
