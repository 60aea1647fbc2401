#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fwd_last_fast|k_inv_last_fast" -s 6 -c 2 -f -o /tmp/r2s_prof python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2s_ncu_run.log 2>&1
ncu -i /tmp/r2s_prof.ncu-rep --page source --csv > gpurun_out/r2s_source.csv 2>/dev/null
ncu -i /tmp/r2s_prof.ncu-rep --page raw --csv > gpurun_out/r2s_raw.csv 2>/dev/null
ls -la gpurun_out/r2s*
timeout 600 python bench.py --steps 20 --no-extras > gpurun_out/r2s_bench_e2e.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_e2e.json')); print(d['e2e'])"
