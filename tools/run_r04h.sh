#!/bin/bash
# persistent tcgen05 projection kernel: correctness + timing vs the two-CTAs-per-SM variant
for shp in "147456 768 256 1" "147456 1024 256 1" "147456 128 256" "442368 252 128" "147456 512 256" "147456 128 128" "442368 128 256"; do
  SGRL_TC_PERSIST=2 timeout 60 python tools/gemm_time.py $shp 2>&1 | tail -1
  SGRL_TC_PERSIST=0 SGRL_TC_SM2=2 timeout 60 python tools/gemm_time.py $shp 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_rollout_gpu.py tests/test_forward_gpu.py tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-bf16 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['rollout']
print('update ms', d['ms_per_step'], 'rollout', r['value'], r['ms_per_forward'], r['share_ms'], r['gemm_frac_of_3xtf32_ceiling'])"
