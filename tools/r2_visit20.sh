#!/bin/bash
# round 2, visit 20: preprocess / decompress drop-ins (splice3) end to end
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider -k "drop_in or end_to_end or long_contigs" > gpurun_out/pytest_v20.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -30 gpurun_out/pytest_v20.log
