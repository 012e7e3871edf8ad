#!/bin/bash
# Round-end evidence of the committed binary: ncu --set full of the three hot kernels, the launch list of a bench run, the
# bench lines themselves.  Outputs under gpurun_out/ (copied to profiles/ by hand after reading them).
set -x
python -m wfcrl_b200.build > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'wf_step_fast_kernel|wf_fixup64' -s 8 -c 2 -f -o gpurun_out/r2_fast python tools/quick_bench.py HornsRev1_ 8192 f32 2 > gpurun_out/ncu_fast.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_step_fast64 -s 4 -c 1 -f -o gpurun_out/r2_fast64 python tools/quick_bench.py HornsRev1_ 8192 f64 2 > gpurun_out/ncu_fast64.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_fast.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/launches_bench.log 2>&1
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/bench.err
python bench.py --precision f64 --quick > gpurun_out/r2_bench_1gpu_f64.json 2> gpurun_out/bench_f64.err
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/bench_ref.err
tail -c 300 gpurun_out/r2_bench_1gpu.json
