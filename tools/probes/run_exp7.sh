for t in 0 1; do
  SCN_TC_T=$t python tools/dom_kernel.py --math bf16 2>&1 | tail -1 | sed "s/^/T=$t /"
  SCN_TC_T=$t SCN_TC_PROF=1 python tools/dom_kernel.py --math bf16 --reps 2 2>&1 | grep tcprof | tail -1
  for c in "32 32" "64 64" "9 32"; do set -- $c; SCN_TC_T=$t python tools/layer_kernel.py --cin $1 --cout $2 2>&1 | tail -1 | sed "s/^/T=$t /"; done
  SCN_TC_T=$t python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('T=$t bench', round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['roofline']['ms_per_launch'],3))"
done
