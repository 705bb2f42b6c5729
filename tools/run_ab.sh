# scratch driver for one gpurun call; edited per experiment.  This one: 2 GPUs, real NCCL path at the final state
set -x
python -m pytest tests/test_nccl_gpu.py -x -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r2_31_2gpu.json 2> gpurun_out/bench_r2_31_2gpu.err; tail -c 900 gpurun_out/bench_r2_31_2gpu.json
