#!/bin/bash
for pf in 1 2 4; do
  WFCRL_NVCC_EXTRA="-DWF_VTAB64_PF_DIST=$pf" python -m wfcrl_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
  TAG="pf64=$pf" python tools/quick_bench.py HornsRev1_ 8192 f64 10
  TAG="pf64=$pf" python tools/quick_bench.py Turb32_Row5_ 8192 f64 10
done
python -m wfcrl_b200.build --force > /dev/null 2>&1
