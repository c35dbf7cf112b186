B="python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['streaming']['ms_per_step'],3))"; }
SCN_HI_STREAM=3 python -m pytest tests -x -q -m gpu -k "replay or program or prefetch" 2>&1 | tail -2
$B 2>/dev/null | pick base
SCN_HI_STREAM=1 $B 2>/dev/null | pick hi_only
SCN_HI_STREAM=3 $B 2>/dev/null | pick hi_lowtwin
SCN_HI_STREAM=2 $B 2>/dev/null | pick lowtwin_only
