#!/bin/sh
# TEST INFRASTRUCTURE ONLY: compile the libpvk .cu sources with g++ against the SIMT
# emulator (tests/emu/cuda_emu.h) -> tests/emu/libpvk_emu.so.  Never loaded by the product.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
SRC="$ROOT/pypevoc_b200/csrc"
OUT="$HERE/libpvk_emu.so"
g++ -std=c++17 -O2 -g -fPIC -shared -ffp-contract=off -DPVK_EMU -Wall -Wno-unknown-pragmas -Wno-unused-function \
    -I"$HERE" -I"$ROOT/include" -I"$SRC" \
    -x c++ $(ls "$SRC"/*.cu) "$HERE/cuda_emu.cc" -o "$OUT" -lpthread
echo "$OUT"
