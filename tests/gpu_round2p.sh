#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $TR tests/slab_check.py --notest --time 512 --profile > gpurun_out/r2p_$name.log 2>&1
  echo "== $name"; grep "^mode\|^profile" gpurun_out/r2p_$name.log | cut -c1-260
}
run k2s7 FH_KBLOCKS=2 SLAB_X=kblock
run k3s7 FH_KBLOCKS=3 SLAB_X=kblock
run k4s7 FH_KBLOCKS=4 SLAB_X=kblock
run k3s7nt512 FH_KBLOCKS=3 FH_REG3_NT=512 SLAB_X=kblock
run p2ps7 SLAB_X=p2p
