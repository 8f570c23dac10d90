#!/bin/bash
# BASELINE.json config 5: SET actor inference at 1K-64K parallel humanoid-9 envs (limb-tokens/s, device timed, inputs resident)
for e in ${ENVS:-1024 4096 16384 65536}; do
  timeout 300 python bench.py --steps 2 --warmup 4 --no-cpu-baseline --rollout-envs $e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['rollout']
print(f\"envs {r['envs_per_gpu']:6d}  tokens {r['envs_per_gpu'] * 9:7d}  {r['ms_per_forward']:8.3f} ms/forward  {r['value'] / 1e6:7.3f} M limb-tokens/s  gemm {r['gemm_tflops']:6.1f} TF/s  K1 {r['feature_k1_gbs']:6.0f} GB/s  K2 {r['attention_k2_gbs']:6.0f} GB/s\")"
done
