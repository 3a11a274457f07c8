#!/bin/bash
# Round-2 evidence, one B200 (run under gpurun from the repo root).  Everything lands in gpurun_out/ev2/;
# tools/summarise_profiles_r2.py turns it into the text / json summaries under profiles/r2/.
set -u
O=gpurun_out/ev2
mkdir -p $O
T="timeout 300"
$T python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_k20.json 2> $O/bench_ref.err
$T python bench.py --steps 20 --warmup 3 > $O/bench_k20.json 2> $O/bench_k20.err
$T python bench.py > $O/bench_default.json 2> $O/bench_default.err
$T python bench.py --streams 1 --steps 300 --no-cpu --kernel-times > $O/bench_serial.json 2> $O/bench_serial.err
$T python bench.py --batch 8 --streams 4 --steps 100 --warmup 5 --no-cpu > $O/bench_b8.json 2> $O/bench_b8.err
$T python bench.py --hw 128x2048 --steps 200 --warmup 5 --no-cpu > $O/bench_128x2048.json 2> $O/bench_128.err
$T python tools/index_bench.py 30 > $O/index_bench.jsonl 2> $O/index_bench.err
if [ -z "${LINES_ONLY:-}" ]; then       # LINES_ONLY=1: the bench lines and the index-op captures only (after an index-kernel change)
$T python tools/train_bench.py --batch 8 --steps 10 > $O/train_b8.json 2> $O/train.err
$T python tools/train_bench.py --batch 32 --steps 10 > $O/train_b32.json 2>> $O/train.err
# every launch of ~4 un-graphed forwards with its device time (cold, serialised: shares, not absolutes)
$T ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $O/launches.csv \
    python bench.py --steps 4 --warmup 3 --pool 2 --no-cpu --no-graph --streams 1 > $O/bench_under_ncu.log 2>&1
# ncu --set full of EVERY kernel of one forward at B = 1 (second forward of the process) ...
timeout 600 ncu --set full --clock-control none --import-source on -s 42 -c 42 -o $O/forward_b1 \
    python tools/one_forward.py 1 2 > $O/ncu_forward_b1.log 2>&1
# ... and config 3: B = 8, cost volume + set-conv / set-upconv kernels only
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cost_volume|group_mlp|set_conv_small|row_mlp" -s 24 -c 24 \
    -o $O/config3_b8 python tools/one_forward.py 8 2 > $O/ncu_config3.log 2>&1
fi
# the index op, configs[0]: without the store warp, and the default form (store warp from 128 cells, bulk-copy staging)
ELO_STORE_WARP_KT=1000000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_conv_tiled -s 3 -c 1 -o $O/index_7x25 \
    python tools/index_one.py 7 25 5 > $O/ncu_index.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_conv_tiled -s 3 -c 1 \
    -o $O/index_7x25_storewarp python tools/index_one.py 7 25 5 >> $O/ncu_index.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_conv_tiled -s 3 -c 1 -o $O/index_11x41 \
    python tools/index_one.py 11 41 5 >> $O/ncu_index.log 2>&1
if [ -z "${LINES_ONLY:-}" ]; then
# memory checker over the index op (both kernels, incl. the store warp) and one forward
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_index_gpu.py -m gpu -q -x \
    -k "golden or model_sites or unaligned or near_tie or full_frame" > $O/sanitize_index.log 2>&1; echo "memcheck index rc=$?" >> $O/sanitize_index.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/index_one.py 7 25 1 > $O/sanitize_storewarp.log 2>&1; echo "memcheck storewarp rc=$?" >> $O/sanitize_storewarp.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/one_forward.py 1 1 > $O/sanitize_forward.log 2>&1; echo "memcheck forward rc=$?" >> $O/sanitize_forward.log
fi
# gpurun brings back at most 64 MiB: turn the reports into text here and drop them
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,launch__occupancy_limit_shared_mem,launch__shared_mem_per_block_dynamic
for r in forward_b1 config3_b8 index_7x25 index_7x25_storewarp index_11x41; do
  [ -f $O/$r.ncu-rep ] || continue
  ncu -i $O/$r.ncu-rep --page raw --csv --metrics $METRICS > $O/${r}_metrics.csv 2>/dev/null
  ncu -i $O/$r.ncu-rep --page details > $O/${r}_details.txt 2>/dev/null
  case $r in index_*) ncu -i $O/$r.ncu-rep --page source --csv > $O/${r}_source.csv 2>/dev/null;; esac
  rm -f $O/$r.ncu-rep
done
gzip -f $O/*_details.txt $O/*_source.csv 2>/dev/null
cat $O/sanitize_*.log | tail -30
du -sh $O; ls -la $O
