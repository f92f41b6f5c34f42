#!/bin/bash
# development aid: time alternative builds of the library (scratch_libs/*.so)
cp autotune-v2_b200/lib/libat2_b200.so /tmp/libat2_orig.so
for lib in /tmp/libat2_orig.so scratch_libs/*.so; do
  cp $lib autotune-v2_b200/lib/libat2_b200.so
  echo "== $lib"
  python scripts/quick_timing.py 2>&1 | grep -E "runs/s|Error|error" | head -3
done
cp /tmp/libat2_orig.so autotune-v2_b200/lib/libat2_b200.so
