#!/bin/bash
# warps per env of the FP64 re-solve kernel (chain warp + W-1 workers) against the strict FP32 step time
python -m wfcrl_b200.build > /dev/null 2>&1
for cfg in "HornsRev1_ 8192 8 4" "Turb32_Row5_ 8192 8 4 2" "Turb_TCRWP_ 16384 8 4 2 1" "Ablaincourt_ 4096 4 2 1" "HornsRev1_ 1024 8 4" "Turb32_Row5_ 1024 8 4 2"; do
  set -- $cfg; name=$1; B=$2; shift 2
  for w in "$@"; do
    WFCRL_B200_FIX_WARPS=$w TAG="fixwarps=$w" python tools/quick_bench.py $name $B f32 20 | sed 's/regs=.*tag=/tag=/; s/| back-to-back.*//'
  done
done
