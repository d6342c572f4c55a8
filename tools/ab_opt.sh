# A/B on the cfg3 bench: TT_PREGATHER (stand-alone gather ahead of the fused tower launch for towers with bags)
run() { python bench.py --no-cpu-baseline --no-extras $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 $2', round(d['ms_per_step']*1000,2), 'us/step  e2e', round(d['e2e']['value']/1e6,2), d['kernels_us_in_graph'], d['kernels_us_per_step'])"; }
for g in 1 0 1 0; do TT_PREGATHER=$g run pregather=$g "--config cfg3 --steps 200"; done
