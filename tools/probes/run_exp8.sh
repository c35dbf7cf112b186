for cfg in "2 0" "1 2" "1 3" "1 4"; do set -- $cfg
  SCN_TC_CTAS=$1 SCN_TC_T=$2 python tools/dom_kernel.py --math bf16 2>&1 | tail -1 | sed "s/^/CTAS=$1 T=$2 /"
  SCN_TC_CTAS=$1 SCN_TC_T=$2 SCN_TC_PROF=1 python tools/dom_kernel.py --math bf16 --reps 2 2>&1 | grep tcprof | tail -1
done
