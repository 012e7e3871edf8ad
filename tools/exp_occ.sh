# occupancy experiment: pad dynamic smem to force fewer resident env-CTAs per SM (227 KB / (smem + pad + 1 KB))
for pad in 0 1024 2048 3072 4608; do
  for B in 8192 65536; do
    WFCRL_SMEM_PAD=$pad python bench.py --steps 30 --warmup 5 --no-cpu-baseline --envs-per-gpu $B 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pad', $pad, 'B', $B, round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],4))"
  done
done
