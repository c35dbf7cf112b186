B="python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['streaming']['ms_per_step'],3))"; }
$B 2>/dev/null | pick base
SCN_TC_SPLIT_MAX=40 $B 2>/dev/null | pick split40
SCN_TC_SPLIT_MAX=160 $B 2>/dev/null | pick split160
SCN_SKIP_TWIN=1 $B 2>/dev/null | pick skiptwin
SCN_PLAN_SORT=0 SCN_SKIP_TWIN=1 $B 2>/dev/null | pick skiptwin_nosort
