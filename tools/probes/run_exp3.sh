B="python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['streaming']['ms_per_step'],3), round(d['roofline']['ms_per_launch'],3))"; }
timeout 300 python -m pytest tests -x -q -m gpu -k "tc_ or replay or program or fused or pins" 2>&1 | tail -2
SCN_TC_DYNAMIC=0 $B 2>/dev/null | pick static
$B 2>/dev/null | pick dynamic
SCN_TC_DYNAMIC=0 $B 2>/dev/null | pick static
$B 2>/dev/null | pick dynamic
for c in "32 32" "128 128"; do set -- $c; python tools/layer_kernel.py --cin $1 --cout $2 2>&1 | tail -1; done
