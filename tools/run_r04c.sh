#!/bin/bash
# rollout-size projections: two-CTAs-per-SM variant vs plain, and the pipeline timeline of a CTA in steady state
for shp in "147456 768 256" "147456 1024 256" "147456 128 256" "442368 252 128" "147456 512 256"; do
  for sm2 in 2 0; do SGRL_TC_SM2=$sm2 python tools/gemm_time.py $shp 2>&1 | tail -1; done
done
for sm2 in 2 0; do for cta in 0 3000; do
  echo "== SGRL_TC_SM2=$sm2 SGRL_TRACE_CTA=$cta"; SGRL_TC_SM2=$sm2 SGRL_TRACE_CTA=$cta python tools/gemm_trace.py 147456 768 256 1 2>&1 | tail -6
done; done
