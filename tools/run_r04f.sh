#!/bin/bash
for shp in "2304 768 256 1" "2304 128 256" "2304 1024 256 1"; do
  for ts in 1 0; do SGRL_TC_TMA_STORE=$ts python tools/gemm_time.py $shp 2>&1 | tail -1; done
done
for ts in 1 0; do
echo "== update size TMA_STORE=$ts"; SGRL_TC_TMA_STORE=$ts python tools/gemm_trace.py 2304 768 256 1 2>&1 | tail -6
done
bash tools/ab_update.sh "" "SGRL_TC_TMA_STORE=0" "" "SGRL_TC_TMA_STORE=0"
