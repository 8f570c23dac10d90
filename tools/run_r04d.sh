#!/bin/bash
# after the non-blocking barrier probe in the MMA issuer (+ reciprocal /F in the inference variant)
for shp in "147456 768 256 1" "147456 1024 256 1" "147456 128 256" "442368 252 128" "147456 512 256"; do
  for sm2 in 2 0; do SGRL_TC_SM2=$sm2 python tools/gemm_time.py $shp 2>&1 | tail -1; done
done
for sm2 in 2 0; do for cta in 3000; do
  echo "== SGRL_TC_SM2=$sm2 SGRL_TRACE_CTA=$cta"; SGRL_TC_SM2=$sm2 SGRL_TRACE_CTA=$cta python tools/gemm_trace.py 147456 768 256 1 2>&1 | tail -6
done; done
echo "== update size"; python tools/gemm_trace.py 2304 768 256 1 2>&1 | tail -6
python tools/gemm_bench.py 2>&1 | tail -40
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_fused_gpu.py tests/test_rollout_gpu.py tests/test_forward_gpu.py tests/test_backward_gpu.py -x -q -m gpu 2>&1 | tail -3
