set -x
for w in conv deconv dw; do python tools/strided_kernels.py $w > gpurun_out/r2_plain_$w.log 2>&1 || exit 1; done
ncu --set full --clock-control none -k regex:conv_plan_tc -s 2 -c 1 -o gpurun_out/r2_strided_conv python tools/strided_kernels.py conv > gpurun_out/ncu_conv.log 2>&1
ncu --set full --clock-control none -k regex:conv_plan_tc -s 2 -c 1 -o gpurun_out/r2_deconv_lat python tools/strided_kernels.py deconv > gpurun_out/ncu_deconv.log 2>&1
ncu --set full --clock-control none -k regex:conv_dw_tc -s 2 -c 1 -o gpurun_out/r2_dw python tools/strided_kernels.py dw > gpurun_out/ncu_dw.log 2>&1
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
