#!/bin/bash
# compute-sanitizer over the index ops (both work decompositions) and one forward; run under gpurun.
mkdir -p gpurun_out
echo "==== MEMCHECK index"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_index_gpu.py -k "raster_queries_vs_oracle or near_ties or demo_known or tie_order or golden" -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
echo "==== RACECHECK index"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_index_gpu.py -k "raster_queries_vs_oracle and tiled" -x -q 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head -8
echo "==== MEMCHECK forward"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python tools/one_forward.py 1 1 2>&1 | grep -E "ERROR SUMMARY|Invalid|^q " | head -5
echo "==== SYNCCHECK forward"; timeout 900 compute-sanitizer --tool synccheck --error-exitcode 0 python tools/one_forward.py 1 1 2>&1 | grep -E "ERROR SUMMARY|Barrier|^q " | head -5
