pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['batch64']['ms_passes'], d.get('rpn',{}).get('ms_per_step'))"; }
BENCH_SKIP=train python bench.py --no-cpu-baseline --steps 50 2>/dev/null | pick with_rpn
BENCH_SKIP=rpn,train python bench.py --no-cpu-baseline --steps 50 2>/dev/null | pick no_rpn
