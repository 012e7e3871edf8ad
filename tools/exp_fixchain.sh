#!/bin/bash
# timing-only ablations of the FP64 re-solve kernel (results are WRONG with any of these defined): where does its latency go?
for def in "" "-DWF_DBG_SKIP_TRIG" "-DWF_DBG_SKIP_POW" "-DWF_DBG_SKIP_D" "-DWF_DBG_SKIP_V" "-DWF_DBG_SKIP_CBRT" "-DWF_DBG_SKIP_TRIG -DWF_DBG_SKIP_POW -DWF_DBG_SKIP_D -DWF_DBG_SKIP_V -DWF_DBG_SKIP_CBRT"; do
  WFCRL_NVCC_EXTRA="$def" python -m wfcrl_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
  TAG="$def" python tools/quick_bench.py HornsRev1_ 8192 f32 10
done
python -m wfcrl_b200.build --force > /dev/null 2>&1
