#!/bin/bash
# development aid: time alternative builds of the library (scratch_libs/*.so) under a few tile shapes
cp autotune-v2_b200/lib/libat2_b200.so /tmp/libat2_orig.so
for lib in /tmp/libat2_orig.so scratch_libs/*.so; do
  cp $lib autotune-v2_b200/lib/libat2_b200.so
  for cfg in "1 1 8" "1 3 8"; do
    set -- $cfg
    echo "== $lib team_warps=$1 reps_per_tile=$2 warps_per_cta=$3"
    AT2_HB_TEAM_WARPS=$1 AT2_HB_REPS_PER_TILE=$2 AT2_HB_WARPS_PER_CTA=$3 python scripts/quick_timing.py 2>&1 | grep -E "runs/s|Error|error" | head -3
  done
done
cp /tmp/libat2_orig.so autotune-v2_b200/lib/libat2_b200.so
