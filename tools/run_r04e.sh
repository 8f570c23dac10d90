#!/bin/bash
# TMA-store epilogue: correctness, then A/B against the st.global epilogue
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_fused_gpu.py tests/test_rollout_gpu.py tests/test_forward_gpu.py tests/test_backward_gpu.py tests/test_agent_gpu.py -x -q -m gpu 2>&1 | tail -3
for shp in "147456 768 256 1" "147456 1024 256 1" "147456 128 256" "442368 252 128" "147456 512 256" "2304 768 256 1" "2304 128 256"; do
  for ts in 1 0; do SGRL_TC_SM2=2 SGRL_TC_TMA_STORE=$ts python tools/gemm_time.py $shp 2>&1 | tail -1; done
done
for ts in 1 0; do
echo "== SM2 TMA_STORE=$ts"; SGRL_TC_TMA_STORE=$ts SGRL_TC_SM2=2 SGRL_TRACE_CTA=3000 python tools/gemm_trace.py 147456 768 256 1 2>&1 | tail -6
echo "== update size TMA_STORE=$ts"; SGRL_TC_TMA_STORE=$ts python tools/gemm_trace.py 2304 768 256 1 2>&1 | tail -6
done
bash tools/ab_update.sh "" "SGRL_TC_TMA_STORE=0" "" "SGRL_TC_TMA_STORE=0"
