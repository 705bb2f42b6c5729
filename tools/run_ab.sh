# scratch driver for one gpurun call (A/B timings + ncu captures); edited per experiment
set -x
python -m pytest tests -m gpu -x -q -s > gpurun_out/tests_r2_19.txt 2>&1; tail -4 gpurun_out/tests_r2_19.txt
python tools/class_profile.py gpurun_out/class_times_r2_19_mixed.csv valinomycin-tzvp ones 1e-7 2>&1 | tail -1
NCU="ncu --set full --import-source on --clock-control none --kernel-name-base mangled"
$NCU -k regex:jk_brick_kernelIdLi1ELi0ELi1ELi0ELb1ELb1 -s 3 -c 1 -o gpurun_out/ncu_r2_19_brick_1010 -f python tools/one_build.py valinomycin-tzvp 1 2>&1 | tail -2
$NCU -k regex:jk_warp_kernelILi2ELi1ELi1ELi1ELb1ELb1 -s 2 -c 1 -o gpurun_out/ncu_r2_19_warp_2111 -f python tools/one_build.py valinomycin-tzvp 1 2>&1 | tail -2
$NCU -k regex:jk_bwarp_kernelIdLi3ELi1ELi2ELi1ELb1ELb1 -s 2 -c 1 -o gpurun_out/ncu_r2_19_bwarp_3121 -f python tools/one_build.py valinomycin-tzvp 1 2>&1 | tail -2
$NCU -k regex:jk_bwarp_kernelIfLi3ELi1ELi2ELi1ELb1ELb1 -s 2 -c 1 -o gpurun_out/ncu_r2_19_bwarp32_3121 -f python tools/one_build.py valinomycin-tzvp 1 ones 1e-7 2>&1 | tail -2
ls -la gpurun_out/*r2_19*
