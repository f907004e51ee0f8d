#!/bin/bash
# A/B timing of library variants on one GPU box (development aid):
#   tools/ab.sh "<variant suffixes, '' = product _lib>" [rounds=2] [quick_bench args...]
# Variants are built with `make OUT=_lib_<x> BUILD=build_<x> EXTRA=-D...` in zig-raytracing-weekend_b200/.
V="$1"; R=${2:-2}; shift 2
ARGS=${@:---spp 256 --integrators 1 --traversal 3 --reps 3}
for r in $(seq $R); do for v in $V; do
  [ "$v" = "-" ] && d=_lib || d=_lib_$v
  echo "variant [$v]"
  RTB_LIB_DIR=zig-raytracing-weekend_b200/$d python tools/quick_bench.py $ARGS 2>&1 | tail -1
done; done
