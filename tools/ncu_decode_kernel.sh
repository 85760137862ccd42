#!/bin/bash
# ncu --set full of the block-decode kernel on the config-2 stream (run under gpurun)
python tools/prof_decode3.py 8192 matrix8_1.txt > gpurun_out/pd.log 2>&1 || exit 1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:decode_blocks_lean_kernel --launch-skip 1 -c 1 \
  -o gpurun_out/r2_decode_lean -f python tools/prof_decode3.py 8192 matrix8_1.txt > gpurun_out/ncu_decode_lean.log 2>&1
tail -2 gpurun_out/ncu_decode_lean.log
