#!/bin/bash
# copy what tools/final_profiles.sh brought back in gpurun_out/ into profiles/ (tracked) and refresh profiles/ncu_traffic.json
set -e
ncu -i gpurun_out/r2_fast.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > /tmp/fast_all.txt
n=$(grep -n "^## " /tmp/fast_all.txt | sed -n 2p | cut -d: -f1)
head -n $((n-1)) /tmp/fast_all.txt > profiles/r2_ncu_fast.txt
tail -n +$n /tmp/fast_all.txt > profiles/r2_ncu_fixup.txt
ncu -i gpurun_out/r2_fast64.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > profiles/r2_ncu_fast64.txt
cp gpurun_out/r2_launches_fast.csv profiles/
for f in r2_bench_1gpu r2_bench_1gpu_f64 r2_bench_reference_arm; do cp gpurun_out/$f.json profiles/$f.json; done
python - <<'PY'
import json, re
def grab(path, key):
    for line in open(path):
        if line.startswith(key + " "):
            parts = line.split()
            unit, val = parts[-2], float(parts[-1])
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "inst": 1.0}.get(unit, 1.0)
            return val * mult
    raise KeyError(key)
p = "profiles/ncu_traffic.json"
d = json.load(open(p))
f, x, g = "profiles/r2_ncu_fast.txt", "profiles/r2_ncu_fixup.txt", "profiles/r2_ncu_fast64.txt"
d["fast"].update(dram_bytes_read=int(grab(f, "dram__bytes_read.sum")), dram_bytes_write=int(grab(f, "dram__bytes_write.sum")),
                 warp_instructions_per_env_step=round(grab(f, "smsp__inst_executed.sum") / 8192))
d["fast"]["fixup_kernel"].update(dram_bytes_read=int(grab(x, "dram__bytes_read.sum")), dram_bytes_write=int(grab(x, "dram__bytes_write.sum")),
                                 warp_instructions_per_launch=int(grab(x, "smsp__inst_executed.sum")),
                                 duration_us=round(grab(x, "gpu__time_duration.sum") * 1e6, 1))
d["fast64"].update(dram_bytes_read=int(grab(g, "dram__bytes_read.sum")), dram_bytes_write=int(grab(g, "dram__bytes_write.sum")),
                   warp_instructions_per_env_step=round(grab(g, "smsp__inst_executed.sum") / 8192))
json.dump(d, open(p, "w"), indent=1)
for name in ("r2_bench_1gpu", "r2_bench_1gpu_f64"):
    b = json.loads(open(f"profiles/{name}.json").read().strip().splitlines()[-1])
    print(name, round(b["value"]), b["ms_per_step"], round(b["e2e"]["value"]), round(b["roofline"]["frac"], 4))
print(json.dumps(d["fast"]["fixup_kernel"])[:300])
PY
