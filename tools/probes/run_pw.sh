for pw in 0 3; do
  for c in "32 32" "9 32" "64 64" "128 128"; do set -- $c
    SCN_TC_PW8=$pw python tools/layer_kernel.py --cin $1 --cout $2 2>&1 | tail -1 | sed "s/^/PW8=$pw /"
  done
done
