B="python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['streaming']['ms_per_step'],3), round(d['roofline']['ms_per_launch'],3))"; }
timeout 900 python -m pytest tests -x -q -m gpu -k "replay or program or prefetch or golden or fused or training or rotate_nms" 2>&1 | tail -3
$B 2>/dev/null | pick pdl
SCN_PDL=0 $B 2>/dev/null | pick nopdl
$B 2>/dev/null | pick pdl
SCN_PDL=0 $B 2>/dev/null | pick nopdl
