#!/bin/bash
python tools/prof_video.py 240 > gpurun_out/pv.log 2>&1 || exit 1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:vparse_chain_kernel -c 1 \
  -o gpurun_out/r2_vchain -f python tools/prof_video.py 240 > gpurun_out/ncu_vchain.log 2>&1
tail -2 gpurun_out/ncu_vchain.log
