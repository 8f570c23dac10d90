#!/bin/bash
# ncu --set full captures of the dominant kernels (1 GPU; never under torchrun).  Reports land in gpurun_out/; summarise them
# here with `python tools/ncu_summary.py gpurun_out/<name>.ncu-rep > profiles/<round>_<name>_summary.txt`.
# --kernel-name-base demangled makes -k match the full C++ name (namespace + template arguments).
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
BENCH="python bench.py --steps 2 --warmup 4 --no-cpu-baseline"
# tcgen05 GEMM launches of the B=256 update (gemm_tc_kernel<BN, AMN, BMN, BPRE, SM2>)
$NCU -k "regex:gemm_tc_kernel" -s 60 -c 12 -f -o gpurun_out/prof_gemm_update $BENCH --no-rollout > gpurun_out/ncu_gemm_update.log 2>&1
# the two-CTAs-per-SM variant only runs in the rollout leg
$NCU -k "regex:gemm_tc_kernel<\(int\)128, \(bool\)0, \(bool\)[01], \(bool\)1, \(bool\)1>" -s 10 -c 6 -f -o gpurun_out/prof_gemm_sm2 $BENCH > gpurun_out/ncu_gemm_sm2.log 2>&1
# K1 / K2 at rollout size through the C ABI
$NCU -k "regex:inv_feature_fwd_kernel<\(int\)0, \(int\)1>" -s 4 -c 1 -f -o gpurun_out/prof_k1 python tools/k_bench.py > gpurun_out/ncu_k1.log 2>&1
$NCU -k "regex:attention_fwd_kernel" -s 6 -c 1 -f -o gpurun_out/prof_k2 python tools/k_bench.py > gpurun_out/ncu_k2.log 2>&1
ls -la gpurun_out/*.ncu-rep
