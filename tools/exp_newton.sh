#!/bin/bash
# one Newton step instead of two in rcp64 / sqrt64 / cbrt64: time and parity
for n in 1 2; do
  WFCRL_NVCC_EXTRA="-DWF_LEAN_NEWTON=$n" python -m wfcrl_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
  TAG="newton=$n" python tools/quick_bench.py HornsRev1_ 8192 f32 20
  TAG="newton=$n" python tools/quick_bench.py HornsRev1_ 8192 f64 10
  TAG="newton=$n" python tools/quick_bench.py Turb32_Row5_ 8192 f64 10
  python tools/parity_sweep.py 2>&1 | grep -E "f64|f32 \{" | sed "s/^/newton=$n /" | cut -c1-420
done
python -m wfcrl_b200.build --force > /dev/null 2>&1
