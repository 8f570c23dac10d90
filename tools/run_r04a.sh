#!/bin/bash
# round-2 session-4 GPU call A: gpu tests, PDL/event probe, rollout launch list, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r04a_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r04a_pytest.log
timeout 120 tools/bin/pdl_event_repro 3000 > gpurun_out/r04a_pdl_event_repro.txt 2>&1; cat gpurun_out/r04a_pdl_event_repro.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r04a_rollout_launches.csv \
  python tools/fwd_probe.py --net actor --batch 16384 --keep 0 --reps 2 > gpurun_out/r04a_ncu_rollout.log 2>&1; tail -1 gpurun_out/r04a_ncu_rollout.log
timeout 400 python bench.py > gpurun_out/r04a_bench.json 2> gpurun_out/r04a_bench.err; tail -c 600 gpurun_out/r04a_bench.json
