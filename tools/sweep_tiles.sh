#!/bin/bash
# whole-step sweep of the tcgen05 tile cost-model knobs (csrc/gemm_tc.cuh gemm_tc()): ms per TD3 update, humanoid-9 B=256
run() { echo "== $*"; env "$@" timeout 120 python bench.py --steps 20 --warmup 6 --no-cpu-baseline --no-rollout 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3))"; }
for w in ${WAVES:-16 24 28 32 36}; do run SGRL_TC_WAVE=$w; done
for w in ${SPLIT_WAVES:-12 20 48 74}; do run SGRL_TC_WAVE_SPLIT=$w; done
