run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 20 --warmup 6 --no-cpu-baseline --no-rollout 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"; }
run SGRL_TC_WAVE=32
run SGRL_TC_WAVE=48
run SGRL_TC_WAVE=64
run SGRL_TC_WAVE=100
run SGRL_TC_WAVE=74 SGRL_TC_BLK128=700
run SGRL_TC_WAVE=74 SGRL_TC_BLK128=1400
run SGRL_TC_WAVE=74 SGRL_TC_FIX64=3000 SGRL_TC_FIX128=4000
run SGRL_TC_WAVE=74 SGRL_TC_FIX64=12000 SGRL_TC_FIX128=16000
run SGRL_TC_WAVE=74 SGRL_TC_SPLITFIX=0
run SGRL_TC_WAVE=48 SGRL_TC_FIX64=3000 SGRL_TC_FIX128=4000
