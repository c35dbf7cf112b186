python -m pytest tests -x -q -m gpu -k "backward or training or grad or train" 2>&1 | tail -2
T="python bench.py --config train --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])['detail']; print('$1', round(d['ms_per_step'],2), round(d['forward_ms'],2), round(d['backward_ms'],2), d['loss'], d['params_with_grad'])"; }
$T 2>/dev/null | pick internal_plan
SCN_DW_PLAN=0 $T 2>/dev/null | pick internal_lists
$T 2>/dev/null | pick internal_plan
SCN_DW_PROF=1 python tools/probes/one_train_step.py 2 2>&1 | grep dwprof | tail -6
