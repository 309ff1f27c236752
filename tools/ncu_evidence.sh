# ncu evidence of the final kernels (run on the GPU box from the repository root; outputs under gpurun_out/)
set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:espcn_fused -s 2 -c 1 -o gpurun_out/r2_ef_ss python tools/run_espcn_fused.py 1 > gpurun_out/ev1.log 2>&1
$NCU -k regex:conv_strip -s 2 -c 1 -o gpurun_out/r2_conv_strip_4k python tools/profile_conv.py fwd 16x2160x242 > gpurun_out/ev2.log 2>&1
$NCU -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/r2_conv_tc_4k python tools/profile_conv.py fwd 16x2160x242 flat > gpurun_out/ev2b.log 2>&1
$NCU -k regex:wgrad_tc_batched -s 1 -c 1 -o gpurun_out/r2_wgrad_batched python tools/profile_step.py train > gpurun_out/ev3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_vdsr_train.csv python tools/profile_step.py train > gpurun_out/ev4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_espcn.csv python bench.py --steps 3 --warmup 3 --no-also > gpurun_out/ev5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_vdsr_infer.csv python tools/profile_step.py infer > gpurun_out/ev6.log 2>&1
tail -qn 2 gpurun_out/ev*.log
