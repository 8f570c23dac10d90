run() { echo "== $*"; env "$@" timeout 120 python bench.py --steps 20 --warmup 6 --no-cpu-baseline --no-rollout 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3))"; }
run SGRL_TC_CSK=0
run SGRL_TC_CSK=1
run SGRL_TC_CSK=1 SGRL_TC_CSK_MINKB=24
run SGRL_TC_CSK=1 SGRL_TC_CSK_MAXT=40
run SGRL_TC_CSK=1 SGRL_TC_CSK_MINKB=24 SGRL_TC_CSK_MAXT=40
