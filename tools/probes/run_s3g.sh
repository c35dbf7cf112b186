python -m pytest tests -x -q -m gpu -k "rpn or nms or detector or postproc or anchor or proposals" 2>&1 | tail -2
python bench.py --config rpn --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rpn', d['ms_per_step'], d['value'], d['detail']['proposals_per_group'], d['detail']['finite'])"
python bench.py --config rpn --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rpn', d['ms_per_step'], d['value'], d['detail']['proposals_per_group'], d['detail']['finite'])"
python tools/probes/rpn_tail_profile.py 2>&1 | grep "backbone"
