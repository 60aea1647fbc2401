#!/bin/bash
# first GPU session of round 2: S3 kernel variants, new parity tests, T6 drop-in, bench line, full GPU suite
mkdir -p gpurun_out
L=gpurun_out/r2a_stage_times.log
: > $L
nvidia-smi --query-gpu=name,memory.total --format=csv >> $L 2>&1
for p in 3 2 1 0; do echo "== FH_MID2_PREF=$p" >> $L; FH_MID2_PREF=$p timeout 300 python tests/stage_time.py >> $L 2>&1; done
echo "== FH_MID2=0 (round-1 kernel)" >> $L; FH_MID2=0 timeout 300 python tests/stage_time.py >> $L 2>&1
echo "== scalar D=3 new / old" >> $L; BD=3 timeout 300 python tests/stage_time.py >> $L 2>&1; BD=3 FH_MID2=0 timeout 300 python tests/stage_time.py >> $L 2>&1
echo "== 128^3 new / old" >> $L; BN=128 timeout 300 python tests/stage_time.py >> $L 2>&1; BN=128 FH_MID2=0 timeout 300 python tests/stage_time.py >> $L 2>&1
echo "== sym / rand coefficient modes" >> $L; BA=sym timeout 300 python tests/stage_time.py >> $L 2>&1; BA=rand timeout 300 python tests/stage_time.py >> $L 2>&1
timeout 1500 python -m pytest tests/test_gpu_bench_sizes.py -q --timeout 900 > gpurun_out/r2a_pytest_new.log 2>&1
timeout 900 python tests/t6_dropin.py > gpurun_out/r2a_t6.log 2>&1
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 900 python bench.py --impl reference --steps 3 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_bench_sizes.py > gpurun_out/r2a_pytest_all.log 2>&1
tail -3 gpurun_out/r2a_pytest_new.log gpurun_out/r2a_pytest_all.log gpurun_out/r2a_t6.log
