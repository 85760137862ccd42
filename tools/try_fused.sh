#!/bin/bash
# first runs of the fused stream kernel (encode_variant 8): parity worker, then A/B timing against the default kernels
mkdir -p gpurun_out
timeout 300 python tests/_variant_worker.py 8 > gpurun_out/fused_worker.log 2>&1; echo "worker rc=$?"
tail -15 gpurun_out/fused_worker.log
timeout 300 python tools/ab_quick.py --combos 2:2,8:2 > gpurun_out/fused_ab.log 2>&1; echo "ab rc=$?"
tail -8 gpurun_out/fused_ab.log
