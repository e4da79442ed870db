#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/F2_pytest_2gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/F2_pytest_2gpu.log
grep -n "rank[01]\]:.*Error\|passed\|failed" gpurun_out/F2_pytest_2gpu.log | tail -8
