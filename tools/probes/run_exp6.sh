for d in 0 2 1 3; do SCN_TC_DBG=$d python tools/dom_kernel.py --math bf16 2>&1 | tail -1 | sed "s/^/DBG=$d /"; done
for d in 0 2 1; do SCN_TC_DBG=$d python tools/layer_kernel.py --cin 32 --cout 32 2>&1 | tail -1 | sed "s/^/DBG=$d /"; done
