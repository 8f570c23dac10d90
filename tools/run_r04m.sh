#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r04m_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r04m_pytest.log
timeout 400 python bench.py > gpurun_out/r04m_bench.json 2> gpurun_out/r04m_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r04m_bench.json').read().strip().splitlines()[-1])
r=d['rollout']
print('update ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'rollout', r['value'], r['ms_per_forward'], r['share_ms'], r['gemm_frac_of_3xtf32_ceiling'], 'cpu', d['cpu_baseline']['value'])
PY
