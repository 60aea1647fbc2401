#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2w_stage255.log
: > $L
for envs in "FH_X=1" "FH_ODD_T=8" "FH_ODD_TRW=4"; do
  env $envs BN=255 BD=3 BA=sym timeout 300 python tests/stage_time.py >> $L 2>&1
done
env BN=255 BD=6 BA=phase timeout 300 python tests/stage_time.py >> $L 2>&1
cut -c1-400 $L
for k in k_inv_last_odd k_fwd_last_odd k_mid_green_odd; do
  BN=255 BD=3 BA=sym timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r2w_$k python tests/stage_time.py > gpurun_out/r2w_ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests/test_gpu_odd.py -q --timeout 900 -x > gpurun_out/r2w_pytest_odd.log 2>&1; tail -n 5 gpurun_out/r2w_pytest_odd.log
timeout 1500 python -m pytest tests/test_gpu_bench_sizes.py -q --timeout 900 -x -k "255" > gpurun_out/r2w_pytest_255.log 2>&1; tail -n 3 gpurun_out/r2w_pytest_255.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['seconds'], d['e2e']['breakdown'])
print('config2', d.get('config2_255'))
PY
