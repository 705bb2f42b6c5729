# scratch driver for one gpurun call (tests + bench + ncu launch lists); edited per experiment
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/tests_r2_30.txt 2>&1; tail -3 gpurun_out/tests_r2_30.txt
python bench.py --steps 1 --warmup 1 --class-profile gpurun_out/class_times_r2_30.csv > gpurun_out/bench_r2_30.json 2> gpurun_out/bench_r2_30.err; tail -c 600 gpurun_out/bench_r2_30.json
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 120 ncu --clock-control none --metrics $M -k regex:jk_ --csv --log-file gpurun_out/ncu_pipe_launches_taxol_r2_30.csv python tools/one_build.py taxol-svp 1 2>&1 | tail -2
timeout 220 ncu --clock-control none --metrics $M -k regex:jk_ --csv --log-file gpurun_out/ncu_pipe_launches_r2_30.csv python tools/one_build.py valinomycin-tzvp 1 2>&1 | tail -2
python tools/ncu_pipe_classes.py gpurun_out/ncu_pipe_launches_taxol_r2_30.csv > gpurun_out/ncu_pipe_per_kernel_taxol_r2_30.csv; tail -1 gpurun_out/ncu_pipe_per_kernel_taxol_r2_30.csv
python tools/ncu_pipe_classes.py gpurun_out/ncu_pipe_launches_r2_30.csv > gpurun_out/ncu_pipe_per_kernel_r2_30.csv; wc -l gpurun_out/ncu_pipe_per_kernel_r2_30.csv; tail -1 gpurun_out/ncu_pipe_per_kernel_r2_30.csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
