B="python bench.py --steps 60 --warmup 5 --no-extras --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['streaming']['ms_per_step'],3))"; }
$B 2>/dev/null | pick base
SCN_TC_SPLIT_MAX=0 $B 2>/dev/null | pick split0
SCN_TC_SPLIT_MAX=2 $B 2>/dev/null | pick split2
SCN_TC_SPLIT_MAX=8 $B 2>/dev/null | pick split8
SCN_SKIP_TWIN=1 $B 2>/dev/null | pick skiptwin
SCN_SKIP_TWIN=1 SCN_TC_SPLIT_MAX=0 $B 2>/dev/null | pick skiptwin_split0
$B 2>/dev/null | pick base
