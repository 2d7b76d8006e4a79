one() { timeout 250 python bench.py --no-cpu-baseline --steps 40 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['phases_ms']['pairs'],3), round(d['phases_ms']['grid'],3))"; }
for rep in 1 2; do
one both
WVA_NO_CARVEOUT=1 one nocarve
WVA_NO_SPIN=1 one nospin
WVA_NO_CARVEOUT=1 WVA_NO_SPIN=1 one neither
done
