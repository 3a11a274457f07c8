#!/bin/bash
# Experiment build of the library with per-CTA phase timestamps in the tiled index kernel (tools/tiled_timeline.py).
set -e
cd "$(dirname "$0")/.."
C=efficientlo-net_b200/csrc
python -c "
import importlib, sys
sys.path.insert(0, '.')
importlib.import_module('efficientlo-net_b200.build').build()"
env -u CC -u CXX nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
    -DELO_TILED_TS -c -o /tmp/fused_conv_tiled_ts.o $C/fused_conv_tiled.cu
OBJS=$(ls $C/build/*.o | grep -v fused_conv_tiled.o)
env -u CC -u CXX nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/micro/libelo_b200_ts.so $OBJS /tmp/fused_conv_tiled_ts.o
ls -la tools/micro/libelo_b200_ts.so
