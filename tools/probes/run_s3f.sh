python -m pytest tests -x -q -m gpu 2>&1 | tail -2
B="python bench.py --steps 100 --warmup 5 --no-extras --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['streaming']['ms_per_step'],3), d['roofline']['ms_per_launch'], d['roofline']['frac'])"; }
$B 2>/dev/null | pick new
$B 2>/dev/null | pick new
