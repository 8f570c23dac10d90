#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_fused_gpu.py tests/test_gemm_gpu.py tests/test_rollout_gpu.py tests/test_forward_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-bf16 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['rollout']
print('update ms', d['ms_per_step'], 'rollout', r['value'], r['ms_per_forward'], r['share_ms'], r['gemm_frac_of_3xtf32_ceiling'])"
