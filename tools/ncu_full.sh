#!/bin/bash
# ncu --set full captures of the dominant kernel variants (1 GPU, eager single-stream so that launch indices are stable)
export SGRL_GRAPHS=0 SGRL_SIDE=0
CMD="python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-rollout"
cap() { ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$1" -s $2 -c 2 -f -o gpurun_out/$3 $CMD > gpurun_out/ncu_$3.log 2>&1; }
cap 'gemm_tc_kernel<\(int\)64, \(bool\)0, \(bool\)0, \(bool\)1>' 300 prof_gemm_fwd64
cap 'gemm_tc_kernel<\(int\)128, \(bool\)0, \(bool\)0, \(bool\)1>' 100 prof_gemm_fwd128
cap 'gemm_tc_kernel<\(int\)64, \(bool\)1, \(bool\)1, \(bool\)0>' 150 prof_gemm_wgrad64
cap 'gemm_tc_kernel<\(int\)128, \(bool\)1, \(bool\)1, \(bool\)0>' 20 prof_gemm_wgrad128
ls -la gpurun_out/*.ncu-rep
