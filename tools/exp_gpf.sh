#!/bin/bash
# gather kernel: L2 prefetch distance of the target rows (0 = none) against time and DRAM bytes
for pf in 0 1 2 3; do
  WFCRL_NVCC_EXTRA="-DWF_FAST64_GATHER_MINB=16 -DWF_GATHER_PF_DIST=$pf" python -m wfcrl_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
  TAG="gpf=$pf" python tools/quick_bench.py HornsRev1_ 8192 f64 10
  TAG="gpf=$pf" python tools/quick_bench.py Turb32_Row5_ 8192 f64 10
  ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:wf_step_fast64 -s 3 -c 1 python tools/quick_bench.py HornsRev1_ 8192 f64 2 2>&1 | grep -E "dram__bytes|gpu__time|hit_rate"
done
python -m wfcrl_b200.build --force > /dev/null 2>&1
