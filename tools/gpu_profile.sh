#!/bin/bash
# Profiling pass of one build on the GPU box (run through gpurun; everything lands in gpurun_out/<tag>_*).
#   tools/gpu_profile.sh <tag> [traversal=2] [spp=16]
# 1. plain timing (no profiler)  2. per-kernel DRAM bytes + durations of every wavefront kernel of one render
# 3. ncu --set full of the first wavefront kernels (raygen, extend/shade of bounces 0..4) and of wf_tail.
set -u
TAG=${1:-r2x}; TRAV=${2:-2}; SPP=${3:-16}
OUT=gpurun_out; mkdir -p $OUT
python tools/quick_bench.py --spp 64 --integrators 1 --traversal $TRAV --reps 3 --count > $OUT/${TAG}_plain.log 2>&1 || exit 1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wf_ --csv \
    --log-file $OUT/${TAG}_dram.csv python tools/quick_bench.py --spp $SPP --integrators 1 --traversal $TRAV --reps 1 > $OUT/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_ -c 11 -f -o $OUT/${TAG}_full \
    python tools/quick_bench.py --spp $SPP --integrators 1 --traversal $TRAV --reps 1 > $OUT/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_tail -c 1 -f -o $OUT/${TAG}_tail \
    python tools/quick_bench.py --spp 64 --integrators 1 --traversal $TRAV --reps 2 > $OUT/${TAG}_ncu3.log 2>&1
ls -la $OUT | grep $TAG
