#!/bin/bash
# ncu --set full of one encoder layer's tcgen05 launches at rollout size with the persistent kernels, exported as CSV
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tc" -s 47 -c 14 -f \
  -o /tmp/r04p python tools/fwd_probe.py --net actor --batch 16384 --keep 0 --reps 2 > gpurun_out/r04p_ncu.log 2>&1
tail -2 gpurun_out/r04p_ncu.log
ncu -i /tmp/r04p.ncu-rep --page raw --csv > gpurun_out/r04p_raw.csv
ls -la gpurun_out/r04p_*
