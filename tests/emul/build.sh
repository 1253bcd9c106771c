#!/bin/sh
# Builds the CPU-interpreted copy of the library used by the `not gpu` logic tests.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
g++ -std=c++17 -O2 -g -fPIC -shared -U_FORTIFY_SOURCE -D_FORTIFY_SOURCE=0 -ffp-contract=off -fno-strict-aliasing -Wno-unused-function \
    -Wno-attributes -o "$HERE/libb200lap_emul.so" "$HERE/emul_lib.cpp"
