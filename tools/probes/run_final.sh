set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final_tests.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || exit 1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err
python tools/dom_kernel.py --math bf16 --internal > gpurun_out/final_dom.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv_plan_tc -s 2 -c 1 -f -o gpurun_out/r2_dom_v2 python tools/dom_kernel.py --math bf16 --internal > gpurun_out/ncu_dom2.log 2>&1
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/final_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
