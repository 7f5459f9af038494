#!/bin/sh
# builds the test-only emulation library (see grimb_emu.cpp)
set -e
cd "$(dirname "$0")"
g++ -O2 -g -std=c++17 -DGRIMB_EMU -ffp-contract=off -fPIC -shared -Wall -Wno-unused-variable \
    -o libgrimb_emu.so grimb_emu.cpp
