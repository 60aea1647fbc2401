#!/bin/bash
# final 1-GPU evidence of round 2 (second session): full GPU suite, bench lines, ncu launch list, full-set captures of the
# kernels that changed (odd-length family at 255^3, 512-length last-axis kernels, two-stage S3 at 512)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r2af_pytest_gpu.log 2>&1; tail -n 4 gpurun_out/r2af_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; cut -c1-160 gpurun_out/r2af_bench.json; tail -2 gpurun_out/r2af_bench.err
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/r2af_bench_reference.json 2> gpurun_out/r2af_bench_reference.err
BN=512 BD=3 timeout 300 python tests/stage_time.py > gpurun_out/r2af_stage512_scalar.log 2>&1; cut -c1-330 gpurun_out/r2af_stage512_scalar.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2af_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2af_ncu_launch_run.log 2>&1
for k in k_fwd_last_odd k_inv_last_odd k_mid_green_odd k_c2c_fast; do
  BN=255 BD=3 BA=sym timeout 600 ncu --set full --clock-control none -k regex:$k -s 3 -c 1 -f -o /tmp/r2af_255_$k python tests/stage_time.py > gpurun_out/r2af_ncu_255_$k.log 2>&1
  ncu -i /tmp/r2af_255_$k.ncu-rep --page raw --csv > gpurun_out/r2af_raw_255_$k.csv 2>/dev/null
done
for k in k_fwd_last_reg3 k_inv_last_reg3; do
  BN=512 timeout 600 ncu --set full --clock-control none -k regex:$k -s 3 -c 1 -f -o /tmp/r2af_512_$k python tests/stage_time.py > gpurun_out/r2af_ncu_512_$k.log 2>&1
  ncu -i /tmp/r2af_512_$k.ncu-rep --page raw --csv > gpurun_out/r2af_raw_512_$k.csv 2>/dev/null
done
ls -la gpurun_out | grep r2af
