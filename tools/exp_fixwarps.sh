#!/bin/bash
for w in 2 4 8; do
  WFCRL_NVCC_EXTRA="-DWF_FIX_WARPS=$w" python -m wfcrl_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
  TAG="fixwarps=$w" python tools/quick_bench.py HornsRev1_ 8192 f32 10
  TAG="fixwarps=$w" python tools/quick_bench.py Turb_TCRWP_ 16384 f32 10
done
python -m wfcrl_b200.build --force > /dev/null 2>&1
