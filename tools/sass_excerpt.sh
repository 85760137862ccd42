#!/bin/bash
# SASS evidence of the hot kernels (no GPU needed): per-kernel resource usage, instruction histogram and the first 120
# instructions of each, from the built library.  bash tools/sass_excerpt.sh > profiles/r2_sass_excerpt.txt
LIB=imageencoder_b200/libimageencoder_b200.so
echo "# cuobjdump -sass of $LIB ($(date -u +%F)), sm_100a only:"
cuobjdump -lelf $LIB | head -5
for k in _ZN2ie19encode_tiles_kernelILi8ELi1ELb0ELb1ELi2EEEvNS_12EncodeParamsE _ZN2ie24tile_copyout_fast_kernelILi4ELj8EEEvNS_12EncodeParamsE \
         _ZN2ie25decode_blocks_lean_kernelILi8EEEvNS_12DecodeParamsE _ZN2ie17me_search8_kernelENS_8MEParamsE \
         _ZN2ie19encode_tiles_kernelILi4ELi4ELb1ELb1ELi2EEEvNS_12EncodeParamsE _ZN2ie15parse_spec_walkILi8192ELi64EEEvNS_11ParseParamsE; do
  echo; echo "## $(echo $k | c++filt)"
  cuobjdump --dump-resource-usage $LIB 2>/dev/null | grep -A1 "Function $k:" | tail -1
  cuobjdump -sass -fun $k $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" > /tmp/sass_$$.txt
  echo "instructions: $(wc -l < /tmp/sass_$$.txt)"
  echo "histogram (top 24):"
  awk '{print $2}' /tmp/sass_$$.txt | sed 's/;$//' | sed 's/^@!\?U\?P[0-9T]*$/(predicated)/' | sort | uniq -c | sort -rn | head -24 | awk '{printf "  %6d %s\n", $1, $2}'
  echo "first 120 instructions:"
  head -120 /tmp/sass_$$.txt | cut -c1-120
done
rm -f /tmp/sass_$$.txt
