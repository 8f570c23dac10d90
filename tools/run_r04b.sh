#!/bin/bash
# ncu --set full of one encoder layer's tcgen05 GEMMs at rollout size (second forward of tools/fwd_probe.py), exported as CSV
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tc_kernel" -s 47 -c 14 -f \
  -o /tmp/r04b python tools/fwd_probe.py --net actor --batch 16384 --keep 0 --reps 2 > gpurun_out/r04b_ncu.log 2>&1
tail -2 gpurun_out/r04b_ncu.log
ncu -i /tmp/r04b.ncu-rep --page raw --csv > gpurun_out/r04b_raw.csv
ncu -i /tmp/r04b.ncu-rep --page details --csv > gpurun_out/r04b_details.csv
# source page (SASS + source lines with per-instruction stall samples) of the QKV (5th) and L4 (14th) launches
ncu -i /tmp/r04b.ncu-rep --page source --csv --launch-skip 4 --launch-count 1 > gpurun_out/r04b_source_qkv.csv 2>/dev/null
ls -la gpurun_out/r04b_* /tmp/r04b.ncu-rep
