#!/bin/bash
for shp in "147456 768 256 1" "147456 128 256" "442368 252 128" "147456 512 256"; do
  for sc in 0 9; do SGRL_TC_SM2=2 SGRL_TC_SCHED=$sc python tools/gemm_time.py $shp 2>&1 | tail -1; done
done
echo "== SM2"; SGRL_TC_SM2=2 SGRL_TRACE_CTA=3000 python tools/gemm_trace.py 147456 768 256 1 2>&1 | tail -6
bash tools/ab_update.sh "" "SGRL_TC_TMA_STORE=0"
timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-bf16 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['rollout']
print('rollout', r['value'], r['ms_per_forward'], r['share_ms'], r['gemm_frac_of_3xtf32_ceiling'])"
