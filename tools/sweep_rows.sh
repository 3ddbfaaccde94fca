#!/bin/bash
# sweep of the row_strips launch knobs on double input / lacunar input
# usage: tools/sweep_rows.sh COLS
COLS=${1:-200000}
echo "== defaults"
python tools/profile_ops.py --cols $COLS --ops colVarsD,rowSumsD,rowVarsD 2>&1 | tail -3
for W in 8 12 16; do for NT in 2 3 4; do for U in 2 3 4; do
  echo "== W=$W NT=$NT U=$U"
  SVTGPU_ROW_WARPS=$W SVTGPU_ROW_NTILES=$NT SVTGPU_ROW_SLOTS=$U timeout 120 python tools/profile_ops.py --cols $COLS --ops rowSumsD,rowVarsD 2>&1 | tail -2
done; done; done
