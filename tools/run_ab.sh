# scratch driver for one gpurun call; edited per experiment.  This one: final state (two-pass brick classes on by default)
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/tests_r2_34.txt 2>&1; tail -n 2 gpurun_out/tests_r2_34.txt
python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --class-profile gpurun_out/class_times_r2_34.csv > gpurun_out/bench_r2_34.json 2> gpurun_out/bench_r2_34.err; head -c 300 gpurun_out/bench_r2_34.json
