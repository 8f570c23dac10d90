run() { echo "== $*"; timeout 300 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-rollout "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['config'].get('limb_tokens_per_step',''))"; }
run --set 3d_walkers --batch 100
run --set 3d_walkers --batch 100 --packed
run --set 3d_cwhh --batch 100
run --set 3d_cwhh --batch 100 --packed
run --set 3d_humanoids --batch 256
run --set 3d_humanoids --batch 256 --packed
