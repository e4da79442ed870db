#!/bin/bash
# Round-2 GPU run K (1 GPU): symmetric forward bring-up -- debug dump, its parity test, then the full suite and benches.
mkdir -p gpurun_out
T=${1:-K}
for nd in "384 128" "2304 128" "640 256"; do timeout 120 python tools/sym_debug.py $nd; done > gpurun_out/${T}_symdebug.log 2>&1
echo "symdebug rc=$?"; tail -45 gpurun_out/${T}_symdebug.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "symmetric_forward or cfg2_full or known_answers" -p no:cacheprovider > gpurun_out/${T}_pytest_sym.log 2>&1
echo "pytest sym rc=$?"; tail -30 gpurun_out/${T}_pytest_sym.log
if [ "$2" == "full" ]; then
timeout 2000 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -12 gpurun_out/${T}_pytest.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${T}_bench.err
SM3_TC_FWD_SYM=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_nosym.json 2> /dev/null
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_cfg2.json 2>/dev/null
SM3_TC_FWD_SYM=0 timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-extras > gpurun_out/${T}_bench_cfg2_nosym.json 2>/dev/null
ls -la gpurun_out | tail -8
