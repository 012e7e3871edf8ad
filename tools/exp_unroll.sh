for cfg in "-DWF_FAST_UNROLL_D=1" "-DWF_FAST_UNROLL_D=2" "-DWF_FAST_UNROLL_D=2 -DWF_FAST_UNROLL_V=3" "-DWF_FAST_UNROLL_V=1 -DWF_FAST_UNROLL_D=2"; do
  WFCRL_NVCC_EXTRA="$cfg" python -m wfcrl_b200.build --force 2>&1 | grep -E "error" | head -3
  echo "== $cfg : $(grep -A1 'wf_step_fast_kernelILb1' wfcrl_b200/build.log | grep -o 'Used [0-9]* registers' | head -1)"
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['occupancy']['ctas_per_sm'], d['roofline']['occupancy']['regs_per_thread'])"
done
