timeout 900 python -m pytest tests -x -q -m gpu -k "detector or rotate_nms or decode or voxelize or iou_3d" 2>&1 | tail -15
python bench.py --config rpn 2>gpurun_out/s2e_rpn.err | tail -1 | cut -c1-1500
