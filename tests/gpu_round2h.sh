#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
for cfg in "4 1" "4 2" "8 1" "2 1"; do
set -- $cfg
FH_PUSH_CHUNKS=$1 FH_PUSH_S2_CTAS=$2 SLAB_X=push timeout 900 $TR tests/slab_check.py --notest --time 512 --profile > gpurun_out/r2h_J$1_c$2.log 2>&1
echo "J=$1 S2 ctas/SM=$2"; grep "^mode\|^profile" gpurun_out/r2h_J$1_c$2.log | cut -c1-260
done
