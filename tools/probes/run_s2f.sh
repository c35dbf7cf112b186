timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --config rpn 2>gpurun_out/s2f_rpn.err | tail -1 | cut -c1-600
