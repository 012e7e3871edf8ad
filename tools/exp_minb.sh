for mb in 12 14 16; do
  WFCRL_NVCC_EXTRA="-DWF_FAST_MINB=$mb" python -m wfcrl_b200.build --force > /dev/null 2>&1
  grep -A2 "wf_step_fast_kernelILb1" wfcrl_b200/build.log | grep -o "Used [0-9]* registers" | head -1
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('MINB', $mb, d['value'], d['ms_per_step'], d['roofline']['occupancy']['ctas_per_sm'], d['roofline']['occupancy']['regs_per_thread'])"
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --envs-per-gpu 9472 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  B=9472', d['value'], d['ms_per_step'])"
done
