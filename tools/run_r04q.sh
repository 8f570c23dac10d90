#!/bin/bash
for v in 1 0; do echo "== SGRL_ATTN_V2=$v"; SGRL_ATTN_V2=$v python tools/k_bench.py 2>&1 | grep K2; SGRL_ATTN_V2=$v python tools/k_bench.py 147456 15 2>&1 | grep K2;  SGRL_ATTN_V2=$v python tools/k_bench.py 2304 9 2>&1 | grep K2; done
timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_rollout_gpu.py tests/test_backward_gpu.py tests/test_parity_wide_gpu.py tests/test_agent_gpu.py -x -q -m gpu 2>&1 | tail -3
