run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 20 --warmup 6 --no-cpu-baseline --no-rollout $EXTRA 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"; }
for b in 32 64 128 512 1024; do EXTRA="--batch $b" run A=1; done
EXTRA="--batch 256 --morph 3d_cheetah_14_full" run A=1
