#!/bin/bash
# round 2, visit 33 (2 GPUs): `Reads:` bytes of the spliced reference binary on 1 and 2 GPUs against the reference (ratio drift)
set +e
mkdir -p gpurun_out
timeout 500 python tests/tools/ratio_check.py 4000000 --gpus > gpurun_out/ratio_gpus_4M.txt 2>&1
echo "exit $?"; cat gpurun_out/ratio_gpus_4M.txt
