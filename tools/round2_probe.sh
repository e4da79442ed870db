#!/usr/bin/env bash
# First GPU call of the next round (one B200, ~2 GPU-minutes): the measurements round 1 ran out of budget for.
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'bash tools/round2_probe.sh'
# Everything lands in gpurun_out/ ; summarise with tools/ncu_summary.py and copy what matters to profiles/.
set -u
mkdir -p gpurun_out
# 1. regression: full parity suite + the default bench line
timeout 200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 150 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
# 2. D = 128 shapes: why are K2 / K3 at 1.7x / 2x their MMA / MUFU bounds at cfg2?
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"infonce_tc_(fwd2|bwd)_kernel" -s 8 -c 2 \
    -o gpurun_out/cfg2_tc python bench.py --workload cfg2 --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_cfg2_full.log 2>&1
# (round 1 only saw these knobs through the enqueue-bound eager step; read the CUDA-graph replay time instead.
#  SM3_TC_GROUPS=2 = two softmax warp groups alternating tiles: hides the tcgen05.ld / st latency of the per-warp chain,
#  which is the suspected limiter of K3 at D = 128 now that 4 S/H stages showed the MMA side is not.)
for poly in 0 2; do for ns in 4 2; do for grp in 1 2; do
  SM3_TC_POLY=$poly SM3_TC_BWD_NS=$ns SM3_TC_GROUPS=$grp timeout 40 python bench.py --workload cfg2 --steps 30 --no-extras 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('poly=$poly ns=$ns groups=$grp', d['ms_per_step'], d['cuda_graph'])"
done; done; done > gpurun_out/cfg2_knobs.txt 2>&1
# 2b. K2 / K3 timed on their own over every knob combination, one process (D = 128 shapes + cfg4)
timeout 150 python tools/tc_sweep.py --out gpurun_out/tc_sweep.jsonl > gpurun_out/tc_sweep.log 2>&1
# 3. N4 timing: cluster_memory at the Derm7pt bank size (413 train samples x 512, K = 5) and at 100k x 512
timeout 60 python - > gpurun_out/kmeans_timing.txt 2>&1 <<'PY'
import sys, time, types, torch
sys.path.insert(0, ".")
import skin_sm3_b200 as sm3
for n, d, k in ((413, 512, 5), (100000, 512, 5)):
    emb = torch.nn.functional.normalize(torch.randn(n, d, device="cuda"), dim=1)
    init = torch.randperm(n)[:k].cuda()
    for _ in range(3):
        sm3.spherical_kmeans(emb, init)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        sm3.spherical_kmeans(emb, init)
    torch.cuda.synchronize()
    print(n, d, k, "ms per clustering (10 iterations + final assignment):", (time.perf_counter() - t0) * 100)
PY
ls -la gpurun_out
