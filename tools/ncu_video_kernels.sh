#!/bin/bash
# ncu --set full of the P-frame tile kernel and the ADD-mode block decoder on a 48-frame clip (run under gpurun)
python tools/prof_video.py 48 > gpurun_out/pv.log 2>&1 || exit 1
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:encode_tiles_kernel<\(int\)4, \(int\)4, \(bool\)1|decode_blocks_fast_kernel<\(int\)4, \(bool\)1' \
  --launch-skip 10 -c 3 -o gpurun_out/r2_video_full2 -f python tools/prof_video.py 48 > gpurun_out/ncu_video_full.log 2>&1
tail -3 gpurun_out/ncu_video_full.log
