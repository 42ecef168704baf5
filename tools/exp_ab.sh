for i in 1 2; do
for st in 3 2; do MM_PLACE_STAGES=$st python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stages $st', round(d['ms_per_step']*1e3,1), d['phase_us'])"; done; done
