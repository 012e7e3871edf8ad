#!/bin/bash
# table-path tuning experiments (run on the GPU box): prefetch distance x eviction hint
for cfg in "1 -DWF_VTAB_NO_EVICT_HINT" "1 -DWF_DUMMY" "2 -DWF_DUMMY" "3 -DWF_DUMMY" "0 -DWF_DUMMY"; do
  set -- $cfg
  WFCRL_NVCC_EXTRA="-DWF_VTAB_PF_DIST=$1 $2" python -m wfcrl_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
  grep -A3 "wf_step_fast_kernelILb1ELb1" wfcrl_b200/build.log | grep -o "Used [0-9]* registers" | head -1
  TAG="pf=$1,$2" python tools/quick_bench.py HornsRev1_ 8192 f32 10
done
python -m wfcrl_b200.build --force > /dev/null 2>&1
