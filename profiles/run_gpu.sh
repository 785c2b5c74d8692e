#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list of the bench command and one
# `ncu --set full` capture of every hot-path kernel.  Everything lands in gpurun_out/$TAG/.
#   gpurun --timeout 1500 -- 'bash profiles/run_gpu.sh r2a'            tests, smoke, benches, per-kernel timers (no profiler)
#   gpurun --timeout 900  -- 'bash profiles/run_gpu.sh r2a launches'   + the ncu launch list of the bench command
#   gpurun --timeout 900  -- 'bash profiles/run_gpu.sh r2a full'       + the ncu --set full capture of the step's kernels
# (one profiler invocation per call, and only after the same command has run clean without it)
# afterwards, here: python profiles/make_traffic_json.py gpurun_out/$TAG/step_full.ncu-rep  (roofline.traffic of bench.py)
TAG=${1:-r2}
MODE=${2:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$MODE" = launches ]; then
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu --no-variants > $OUT/bench_short.json 2> $OUT/bench_short.err || exit 1
  # launch list of the bench command (cold-cache, serialised per-launch times: shares only)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu --no-variants > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  exit 0
fi
if [ "$MODE" = full ]; then
  timeout 200 python profiles/prof_step.py 1 > $OUT/prof_step.log 2>&1 || exit 1
  # full capture of the hot-path kernels (one launch each)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'glt_stream|poly_moments|finalize|solve_apply' \
      -o $OUT/step_full -f python profiles/prof_step.py 1 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
  ls -la $OUT
  exit 0
fi
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/nvsmi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/nvsmi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
timeout 300 python profiles/prof_kernels.py all 20 > $OUT/prof_kernels.log 2>&1; cat $OUT/prof_kernels.log
timeout 300 python profiles/prof_next.py 10 > $OUT/prof_next.log 2>&1; cat $OUT/prof_next.log
timeout 300 python profiles/prof_ot.py > $OUT/prof_ot.log 2>&1; cat $OUT/prof_ot.log
timeout 300 python profiles/prof_warp.py 10 > $OUT/prof_warp.log 2>&1; cat $OUT/prof_warp.log
timeout 200 python profiles/prof_select.py 20 > $OUT/prof_select.log 2>&1; cat $OUT/prof_select.log
for c in ortho_srf tiles mosaic shards; do timeout 300 python bench.py --config $c --steps 20 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo "bench $c rc=$?"; done
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json
ls -la $OUT
