#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference (ThenTech/ImageEncoder, /root/reference) into oracle/_ref/ so that the
# C restatement in oracle/ can be pinned against it and so that bench.py has a "reference" CPU arm.
#
#  * sources are compiled from a scratch copy under $TMPDIR (the reference tree is read-only and the
#    8x8 variant needs `BlockSize = 8u` in Block.hpp:13, a compile-time constant); nothing but the
#    binaries is written into the repo, and oracle/_ref/ is git-ignored.
#  * flags: the reference's own (makefile:28: -std=c++17 -O3 -fopenmp, ECFLAGS makefile:13) plus
#      -mlzcnt : makes util::ffs(0) (utils.hpp:210-216, __builtin_clz(0) is UB) deterministic = 0
#    and NEVER -march=native / -mfma (FMA contraction changes the bitstream).
#  * main.cpp is replaced by oracle/ref_harness.cpp (timing + _exit before the crashing dtor).
#
# Produces: oracle/_ref/ref_n{4,8}_{plain,huff}   (all OpenMP builds; OMP_NUM_THREADS=1 gives the serial order,
#           which is byte-identical), and oracle/_ref/ref_cli_{encoder,decoder}: the reference's own
#           CLIs (its main.cpp with -DENCODER / -DDECODER as its makefile builds them; 4x4, Huffman off), used only to pin the
#           drop-in CLI contract: .conf / matrix parsing and exit codes (tests/test_cli_contract_cpu.py).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
    echo "[build_ref] $REF not present (GPU box?) - keeping prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
CXX="${ORACLE_CXX:-/usr/bin/g++}"   # NOT $CXX: this image exports CXX=/opt/gcc/bin/g++, a wrapper without libgomp.spec
FLAGS="-std=c++17 -O3 -mlzcnt -fopenmp -DENABLE_OPENMP -DENCODER -DDECODER -w"

stamp="$OUT/.stamp"
newest=$(find "$REF" -maxdepth 1 \( -name '*.cpp' -o -name '*.hpp' \) -newer "$stamp" 2>/dev/null | head -1 || true)
if [ -f "$stamp" ] && [ -z "$newest" ] && [ "$HERE/ref_harness.cpp" -ot "$stamp" ] && [ "$HERE/build_ref.sh" -ot "$stamp" ] \
   && [ -x "$OUT/ref_n4_plain" ] && [ -x "$OUT/ref_n4_huff" ] && [ -x "$OUT/ref_n8_plain" ] && [ -x "$OUT/ref_n8_huff" ] \
   && [ -x "$OUT/ref_cli_encoder" ] && [ -x "$OUT/ref_cli_decoder" ]; then
    echo "[build_ref] up to date"
    exit 0
fi

SCR="$(mktemp -d)"
trap 'rm -rf "$SCR"' EXIT
pids=()
for N in 4 8; do
    d="$SCR/n$N"
    mkdir -p "$d"
    cp "$REF"/*.cpp "$REF"/*.hpp "$d"/
    mkdir -p "$d/cli" && mv "$d/main.cpp" "$d/cli/main.cpp"
    cp "$HERE/ref_harness.cpp" "$d/"
    if [ "$N" = 8 ]; then
        sed -i 's/BlockSize      =  4u/BlockSize      =  8u/' "$d/Block.hpp"
        grep -q 'BlockSize      =  8u' "$d/Block.hpp"
    fi
    # common objects (everything that does not look at ENABLE_HUFFMAN) are compiled once per N
    for f in "$d"/*.cpp; do
        b=$(basename "$f" .cpp)
        case "$b" in
            ImageEncoder|VideoEncoder|ref_harness) ;;
            *) ( cd "$d" && $CXX $FLAGS -c "$b.cpp" -o "$b.o" ) & pids+=($!) ;;
        esac
    done
    for v in plain huff; do
        def=""; [ "$v" = huff ] && def="-DENABLE_HUFFMAN"
        for b in ImageEncoder VideoEncoder ref_harness; do
            ( cd "$d" && $CXX $FLAGS $def -c "$b.cpp" -o "${b}_$v.o" ) & pids+=($!)
        done
    done
done
# the reference's CLIs (4x4, Huffman off): main.cpp once per target macro, as makefile:5-8 does
CLIFLAGS="-std=c++17 -O3 -mlzcnt -fopenmp -DENABLE_OPENMP -w"
( cd "$SCR/n4" && $CXX $CLIFLAGS -DENCODER -I. -c cli/main.cpp -o cli/main_encoder.o ) & pids+=($!)
( cd "$SCR/n4" && $CXX $CLIFLAGS -DDECODER -I. -c cli/main.cpp -o cli/main_decoder.o ) & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
for N in 4 8; do
    d="$SCR/n$N"
    common=$(ls "$d"/*.o | grep -v -e '_plain.o' -e '_huff.o')
    for v in plain huff; do
        $CXX -fopenmp $common "$d"/ImageEncoder_$v.o "$d"/VideoEncoder_$v.o "$d"/ref_harness_$v.o -o "$OUT/ref_n${N}_$v"
    done
done
common4=$(ls "$SCR/n4"/*.o | grep -v -e '_plain.o' -e '_huff.o')
for t in encoder decoder; do
    $CXX -fopenmp $common4 "$SCR/n4"/ImageEncoder_plain.o "$SCR/n4"/VideoEncoder_plain.o "$SCR/n4/cli/main_$t.o" -o "$OUT/ref_cli_$t"
done
touch "$stamp"
echo "[build_ref] built: $(ls "$OUT")"
