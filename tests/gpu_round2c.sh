#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2c_l2fetch.log
: > $L
for g in 0 128 32; do
echo "== FH_L2FETCH=$g round-1 S3 kernel" >> $L; FH_L2FETCH=$g FH_MID2=0 timeout 300 python tests/stage_time.py >> $L 2>&1
echo "== FH_L2FETCH=$g copy-only S3 (FH_MID_PIPE=9)" >> $L; FH_L2FETCH=$g FH_MID2=0 FH_MID_PIPE=9 timeout 300 python tests/stage_time.py >> $L 2>&1
done
echo "== 512 FH_L2FETCH=0 / 128" >> $L
BN=512 FH_L2FETCH=0 timeout 300 python tests/stage_time.py >> $L 2>&1
BN=512 FH_L2FETCH=128 timeout 300 python tests/stage_time.py >> $L 2>&1
timeout 900 python -m pytest tests/test_gpu_bench_sizes.py -q --timeout 900 -x -k "511 or c3 or random_spd" > gpurun_out/r2c_pytest.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "full_size or deferred or single_rank or golden_examples" > gpurun_out/r2c_pytest2.log 2>&1
tail -n 3 gpurun_out/r2c_pytest.log gpurun_out/r2c_pytest2.log
