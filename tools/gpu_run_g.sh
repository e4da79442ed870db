#!/bin/bash
mkdir -p gpurun_out
T=${1:-G}
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "projector or tail_cal or owner_ordered" > gpurun_out/${T}_pytest_new.log 2>&1
echo "new tests rc=$?"; tail -25 gpurun_out/${T}_pytest_new.log
timeout 2000 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
