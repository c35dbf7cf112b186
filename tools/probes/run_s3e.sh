python -m pytest tests -x -q -m gpu -k "backward or training or grad or train" 2>&1 | tail -2
python tools/strided_kernels.py dw --sorted
SCN_DW_PART=4096 python tools/strided_kernels.py dw --sorted
SCN_DW_PART=16384 python tools/strided_kernels.py dw --sorted
SCN_DW_PART=100000000 python tools/strided_kernels.py dw --sorted
T="python bench.py --config train --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])['detail']; print('$1', round(d['ms_per_step'],2), round(d['forward_ms'],2), round(d['backward_ms'],2), d['loss'], d['params_with_grad'])"; }
$T 2>/dev/null | pick part8k
SCN_DW_PART=4096 $T 2>/dev/null | pick part4k
SCN_DW_PART=100000000 $T 2>/dev/null | pick nopart
LIST=conv_dw python tools/trace_train.py 2>&1 | grep " us  \| ms " | tail -5
