VERBOSE=1 python tools/batch_inference.py 2>&1 | tail -70
