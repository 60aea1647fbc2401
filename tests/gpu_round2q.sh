#!/bin/bash
# final 1-GPU validation of round 2: full GPU suite, bench lines, ncu launch list + full-set capture, sanitizer
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r2q_pytest_gpu.log 2>&1; tail -n 6 gpurun_out/r2q_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; cut -c1-200 gpurun_out/r2q_bench.json
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/r2q_bench_reference.json 2> gpurun_out/r2q_bench_reference.err
for nt in 768 512 384; do echo "== FH_REG3_NT=$nt" >> gpurun_out/r2q_stage512.log; BN=512 FH_REG3_NT=$nt timeout 300 python tests/stage_time.py >> gpurun_out/r2q_stage512.log 2>&1; done
BN=255 BD=3 BA=sym timeout 300 python tests/stage_time.py > gpurun_out/r2q_stage255.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2q_ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_mid_green_pipe|k_inv_last_fast|k_cg_update|k_fwd_last_fast|k_c2c_fast" -s 18 -c 6 -f -o gpurun_out/r2q_prof python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2q_ncu_full_run.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_sanitizer_memcheck.log 2>&1; tail -n 3 gpurun_out/r2q_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_sanitizer_racecheck.log 2>&1; tail -n 3 gpurun_out/r2q_sanitizer_racecheck.log
