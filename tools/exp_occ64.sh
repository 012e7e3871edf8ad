#!/bin/bash
# FP64 kernel throughput against resident CTAs per SM (dynamic shared memory padded to limit the occupancy).
WFCRL_NVCC_EXTRA="-DWF_EXP_SMEM_PAD" python -m wfcrl_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
for pad in 0 11560 14060 16560 24560; do
  WFCRL_B200_F64_SMEM_PAD=$pad TAG="pad=$pad" python tools/quick_bench.py Turb32_Row5_ 8192 f64 10
done
for pad in 0 2500 6500 10500; do
  WFCRL_B200_F64_SMEM_PAD=$pad TAG="pad=$pad" python tools/quick_bench.py HornsRev1_ 8192 f64 10
done
python -m wfcrl_b200.build --force > /dev/null 2>&1
