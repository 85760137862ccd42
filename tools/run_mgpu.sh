#!/bin/bash
# bash tools/run_mgpu.sh N  -- bench.py, config 5 (video) and config 4 (batch) on N GPUs; outputs under gpurun_out/
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err
tail -c 600 gpurun_out/r2b_bench_n$N.json
timeout 300 $TR tools/bench_sharded.py video --reps 3 > gpurun_out/r2b_video_n$N.json 2> gpurun_out/r2b_video_n$N.err
tail -c 700 gpurun_out/r2b_video_n$N.json
timeout 400 $TR tools/bench_sharded.py batch --images ${2:-64} --reps 2 > gpurun_out/r2b_batch_n$N.json 2> gpurun_out/r2b_batch_n$N.err
tail -c 700 gpurun_out/r2b_batch_n$N.json
