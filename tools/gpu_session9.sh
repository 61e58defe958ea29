#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_snowfall_gpu.py -m gpu -q -k "azimuth or golden" > gpurun_out/r2s9_gpu_tests.txt 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2s9_gpu_tests.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 4 -c 1 -f -o gpurun_out/r2s9_solve python tools/profile_step.py --steps 6 > gpurun_out/r2s9_ncu_solve.log 2>&1
echo "ncu solve rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 4 -c 1 -f -o gpurun_out/r2s9_scan python tools/profile_step.py --steps 6 > gpurun_out/r2s9_ncu_scan.log 2>&1
echo "ncu scan rc=$?"
