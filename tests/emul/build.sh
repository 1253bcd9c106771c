#!/bin/sh
# Builds the CPU-interpreted copy of the library used by the `not gpu` logic tests.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
g++ -std=c++17 -O2 -g -fPIC -shared -ffp-contract=off -fno-strict-aliasing -Wno-unused-function \
    -Wno-attributes -o "$HERE/libb200lap_emul.so" "$HERE/emul_lib.cpp"
