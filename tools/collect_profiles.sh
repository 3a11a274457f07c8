#!/bin/bash
# Round evidence, one B200 (run under gpurun from the repo root): bench lines, launch list, ncu summaries.
# Everything lands in gpurun_out/evidence/; copy what should be judged into profiles/.
set -u
O=gpurun_out/evidence
mkdir -p $O
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --streams 1 --steps 300 --no-cpu > $O/bench_serial.json 2>> $O/bench.err
python bench.py --streams 12 --tile-policy 0 --steps 500 --no-cpu > $O/bench_spread.json 2>> $O/bench.err
python bench.py --batch 8 --streams 4 --steps 100 --warmup 5 --no-cpu > $O/bench_b8.json 2>> $O/bench.err
python bench.py --impl reference --steps 8 --warmup 1 > $O/bench_ref.json 2>> $O/bench.err
python tools/index_bench.py 20 > $O/index_bench.jsonl 2>> $O/bench.err
python tools/train_bench.py --batch 8 --steps 5 > $O/train_b8.json 2> $O/train.err
# launch list of ~4 un-graphed forwards (graph replays cannot be profiled launch by launch)
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $O/launches.csv \
    python bench.py --steps 4 --warmup 3 --pool 2 --no-cpu --no-graph --streams 1 > $O/bench_under_ncu.log 2>&1
# full captures: the tiled index kernel (config 1) and the two dominant fused kernels of the forward
ncu --set full --clock-control none --import-source on -k regex:fused_conv_tiled -s 3 -c 1 -o $O/index_tiled \
    python tools/index_op.py > $O/ncu_index.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:group_mlp_max_tc -s 4 -c 1 -o $O/group_tc_l0 \
    python tools/one_forward.py 1 1 > $O/ncu_group.log 2>&1
for r in index_tiled group_tc_l0; do
  ncu -i $O/$r.ncu-rep --page details > $O/ncu_$r.txt 2>/dev/null
  ncu -i $O/$r.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > $O/ncu_${r}_dram.csv 2>/dev/null
done
ls -la $O
