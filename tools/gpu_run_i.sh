#!/bin/bash
# Round-2 GPU run I (N GPUs): bench at N ranks with the symmetric forward across ranks (mode 4, default) and without
# (mode 2), back to back on the same box.
mkdir -p gpurun_out
N=${1:-8}
T=${2:-I}${N}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 \
      bench.py --gpus $N --steps 40 --warmup 8 > gpurun_out/${T}_bench_${name}.json 2> gpurun_out/${T}_bench_${name}.err
  echo "bench $name rc=$?"; tail -c 200 gpurun_out/${T}_bench_${name}.err
}
run nosym SM3_PEER_SYM=0
run sym SM3_PEER_SYM=1
if [ "$3" == "more" ]; then
run nosym_b SM3_PEER_SYM=0
run sym_b SM3_PEER_SYM=1
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/*_bench_*sym*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']), d.get('stages_ms'), d['parity'].get('ok'), d['parity'].get('loss_relerr'), d['parity'].get('grad_relerr_rowblock'), d['parity'].get('fused_vs_nccl_grad_relerr'))
    except Exception as e: print(f,'ERR',e)
PY
