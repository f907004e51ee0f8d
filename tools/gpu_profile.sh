#!/bin/bash
# Profiling pass of one build on the GPU box (run through gpurun; everything lands in gpurun_out/<tag>_*).
#   tools/gpu_profile.sh <tag> [traversal=3] [spp=16] [scene=book1] [what=all|dram]
# 1. plain timing (no profiler)  2. per-kernel DRAM bytes + durations of every wavefront kernel of one render
# 3. ncu --set full of the first wavefront kernels (raygen, extend/shade of bounces 0..4) and of wf_tail.
# Numbers printed under ncu are never bench values; the plain log is.
set -u
TAG=${1:-r2x}; TRAV=${2:-3}; SPP=${3:-16}; SCENE=${4:-book1}; WHAT=${5:-all}
OUT=gpurun_out; mkdir -p $OUT
QB="python tools/quick_bench.py --scene $SCENE --integrators 1 --traversal $TRAV"
$QB --spp $((SPP*4)) --reps 3 --count > $OUT/${TAG}_plain.log 2>&1 || exit 1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:wf_ --csv \
    --log-file $OUT/${TAG}_dram.csv $QB --spp $SPP --reps 1 > $OUT/${TAG}_ncu1.log 2>&1
if [ "$WHAT" = "all" ]; then
ncu --set full --clock-control none --import-source on -k regex:wf_ -c 11 -f -o $OUT/${TAG}_full \
    $QB --spp $SPP --reps 1 > $OUT/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_tail -c 1 -f -o $OUT/${TAG}_tail \
    $QB --spp $((SPP*4)) --reps 2 > $OUT/${TAG}_ncu3.log 2>&1
fi
ls -la $OUT | grep $TAG
