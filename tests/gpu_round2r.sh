#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; cut -c1-200 gpurun_out/r2q_bench.json
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/r2q_bench_reference.json 2> gpurun_out/r2q_bench_reference.err
for nt in 768 512 384; do echo "== FH_REG3_NT=$nt" >> gpurun_out/r2q_stage512.log; BN=512 FH_REG3_NT=$nt timeout 300 python tests/stage_time.py >> gpurun_out/r2q_stage512.log 2>&1; done
BN=255 BD=3 BA=sym timeout 300 python tests/stage_time.py > gpurun_out/r2q_stage255.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2q_ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_mid_green_pipe|k_inv_last_fast|k_cg_update|k_fwd_last_fast|k_c2c_fast" -s 18 -c 6 -f -o /tmp/r2q_prof python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2q_ncu_full_run.log 2>&1
ncu -i /tmp/r2q_prof.ncu-rep --page raw --csv > gpurun_out/r2q_prof_raw.csv 2>/dev/null
ls -la gpurun_out
