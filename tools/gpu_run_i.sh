#!/bin/bash
# Round-2 GPU run I (N GPUs): bench at N ranks with the row push inside K2 (mode 3) and in the normalise kernel (mode 2).
mkdir -p gpurun_out
N=${1:-8}
T=I${N}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 \
      bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err
  echo "bench $name rc=$?"; tail -c 200 gpurun_out/${T}_bench_${name}.err
}
run push SM3_PEER_PUSH=1
run nopush SM3_PEER_PUSH=0
run push_b SM3_PEER_PUSH=1
ls -la gpurun_out | tail -5
