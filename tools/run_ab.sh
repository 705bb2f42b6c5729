# scratch driver for one gpurun call (A/B timings + ncu captures); edited per experiment
set -x
JQC_LIB_PATH=$PWD/joltqc_b200/libjqc_regs.so python tools/class_profile.py gpurun_out/class_times_r2_24_regs.csv 2>&1 | tail -1
M=gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 240 ncu --clock-control none --metrics $M -k regex:jk_ -c 700 --csv --log-file gpurun_out/ncu_fp64_per_launch_r2_24.csv python tools/one_build.py valinomycin-tzvp 1 2>&1 | tail -2
python tools/ncu_fp64_classes.py gpurun_out/ncu_fp64_per_launch_r2_24.csv 34.2 > gpurun_out/ncu_fp64_per_kernel_r2_24.csv; wc -l gpurun_out/ncu_fp64_per_kernel_r2_24.csv; head -5 gpurun_out/ncu_fp64_per_kernel_r2_24.csv
