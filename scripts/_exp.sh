for cfg in "1 3 8" "1 2 6" "1 2 8" "1 1 8" "1 2 7" "1 1 6" "1 2 4"; do
  set -- $cfg
  echo "== team_warps=$1 reps_per_tile=$2 warps_per_cta=$3"
  AT2_HB_TEAM_WARPS=$1 AT2_HB_REPS_PER_TILE=$2 AT2_HB_WARPS_PER_CTA=$3 python scripts/quick_timing.py 2>&1 | grep -E "runs/s|Error|error" | head -2
done
