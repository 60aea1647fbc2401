#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2v_stage255.log
: > $L
for envs in "" "FH_ODD=0" "FH_ODD_T=8" "FH_ODD_TRW=4"; do
  env $envs BN=255 BD=3 BA=sym timeout 300 python tests/stage_time.py >> $L 2>&1
done
env BN=255 BD=3 BA=phase timeout 300 python tests/stage_time.py >> $L 2>&1
env BN=255 BD=6 BA=phase timeout 300 python tests/stage_time.py >> $L 2>&1
env FH_ODD=0 BN=255 BD=6 BA=phase timeout 300 python tests/stage_time.py >> $L 2>&1
cut -c1-400 $L
timeout 1500 python -m pytest tests/test_gpu_odd.py -q --timeout 900 -x > gpurun_out/r2v_pytest_odd.log 2>&1; tail -n 5 gpurun_out/r2v_pytest_odd.log
timeout 1500 python -m pytest tests/test_gpu_bench_sizes.py -q --timeout 900 -x -k "255" > gpurun_out/r2v_pytest_255.log 2>&1; tail -n 3 gpurun_out/r2v_pytest_255.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['seconds'], d['e2e']['breakdown'])
print('config2', d.get('config2_255'))
PY
