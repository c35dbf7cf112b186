B="python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['streaming']['ms_per_step'],3), round(d['roofline']['ms_per_launch'],3))"; }
SCN_PDL=2 $B 2>/dev/null | pick pdl_small
SCN_PDL=0 $B 2>/dev/null | pick nopdl
SCN_PDL=2 $B 2>/dev/null | pick pdl_small
SCN_PDL=0 $B 2>/dev/null | pick nopdl
SCN_PDL=2 SCN_SKIP_TWIN=1 $B 2>/dev/null | pick pdl_small_skiptwin
SCN_PDL=1 SCN_SKIP_TWIN=1 $B 2>/dev/null | pick pdl_all_skiptwin
SCN_PDL=0 SCN_SKIP_TWIN=1 $B 2>/dev/null | pick nopdl_skiptwin
