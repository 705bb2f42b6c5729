# scratch driver for one gpurun call (A/B timings + ncu captures); edited per experiment
set -x
python -m pytest tests/test_jk_gpu.py -x -q > gpurun_out/tests_r2_22.txt 2>&1; tail -3 gpurun_out/tests_r2_22.txt
python tools/class_profile.py gpurun_out/class_times_r2_22.csv 2>&1 | tail -1
JQC_BWARP=2 python tools/class_profile.py gpurun_out/class_times_r2_22_bw2.csv 2>&1 | tail -1
JQC_BWARP=0 python tools/class_profile.py gpurun_out/class_times_r2_22_bw0.csv 2>&1 | tail -1
