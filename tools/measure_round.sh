#!/bin/bash
# The measurement set behind profiles/<TAG>_*: run on ONE B200 (gpurun -- 'bash tools/measure_round.sh r02').
# Bench lines are un-profiled runs; the ncu passes run afterwards on the same build.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err
for w in c1 c2 c4; do python bench.py --workload $w --no-cpu > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; done
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_reference_arm.err
# launch list of a C3 bench run (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_launches_c3.log 2>&1
# full capture of the frame's kernels (steady-state C3 frames of tools/replay.py)
ncu --set full --clock-control none --import-source on -k regex:'assoc_tc|lap_stream|frame_|ctrl_upload' -c 24 \
    -f -o $O/${TAG}_frame_full python tools/replay.py 0 1 > $O/${TAG}_frame_full.log 2>&1
ls -la $O | tail -12
