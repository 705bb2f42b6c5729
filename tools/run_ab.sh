# scratch driver for one gpurun call; edited per experiment.  This one: 2 GPUs, shard interleave in alternating direction
set -x
python -m pytest tests/test_jk_gpu.py tests/test_nccl_gpu.py -x -q -m gpu -k "shards or nccl or multi_chunk or quartet_list or benzene or loose_cutoff" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r2_32_2gpu.json 2> gpurun_out/bench_r2_32_2gpu.err; tail -c 400 gpurun_out/bench_r2_32_2gpu.json
