B="python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['streaming']['ms_per_step'],3), round(d['roofline']['ms_per_launch'],3))"; }
for s in 148 144 140 136 128; do SCN_TC_SMS=$s $B 2>/dev/null | pick sms$s; done
SCN_SKIP_TWIN=1 $B 2>/dev/null | pick skiptwin
SCN_TC_SMS=140 SCN_SKIP_TWIN=1 $B 2>/dev/null | pick skiptwin140
