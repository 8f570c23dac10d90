#!/bin/bash
# compute-sanitizer passes over one SET forward + backward (actor through critic-1, twin critic) on the tcgen05 and the fp32
# SIMT paths, and over one captured-and-replayed Agent.update.  usage (GPU box): bash tools/sanitize.sh  -> gpurun_out/sanitize_*.log
# Summaries are committed under profiles/ (rNN_sanitize_summary.txt).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
for tool in memcheck racecheck synccheck initcheck; do
  for tc in 1 0; do
    log=gpurun_out/sanitize_${tool}_tc${tc}.log
    SGRL_SAN_TC=$tc timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitize_target.py > $log 2>&1
    echo "$tool tc=$tc exit=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
  done
done
