python -m pytest tests -x -q -m gpu -k "training or train or batchnorm or bn or backward" 2>&1 | tail -2
T="python bench.py --config train --no-cpu-baseline"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1])['detail']; print('$1', round(d['ms_per_step'],2), round(d['forward_ms'],2), round(d['backward_ms'],2), d['loss'], d['params_with_grad'])"; }
$T 2>/dev/null | pick half_bn
SCN_TRAIN_HALF_BN=0 $T 2>/dev/null | pick fp32_bn
$T 2>/dev/null | pick half_bn
SCN_TRAIN_HALF_BN=0 $T 2>/dev/null | pick fp32_bn
