python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for v in 1 0; do
SCN_SIMT2=$v python bench.py --math fp32 --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fp32 simt2=$v', round(d['ms_per_step'],2), d['roofline']['ms_per_launch'])"
done
