#!/bin/bash
# Round-2 GPU run F (2 GPUs): multi-rank parity (NCCL vs fused vs fused+push vs fused+symmetric, peer host pipeline,
# sharded k-means), then the bench at N=2 with the exchange modes side by side.
mkdir -p gpurun_out
T=${1:-F}
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_pytest_2gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest_2gpu.log
tail -30 gpurun_out/${T}_pytest_2gpu.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 \
      bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench2_${name}.json 2> gpurun_out/${T}_bench2_${name}.err
  echo "bench2 $name rc=$?"; tail -c 300 gpurun_out/${T}_bench2_${name}.err
}
run sym SM3_PEER_SYM=1
run nosym SM3_PEER_SYM=0
run nccl SM3_COMM=nccl
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/*_bench2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline',{})
        print(f, round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d.get('stages_ms'), d['parity'].get('ok'), d['parity'].get('loss_relerr'), d['parity'].get('fused_vs_nccl_grad_relerr'), d['exchange'][-60:])
    except Exception as e: print(f,'ERR',e)
PY
