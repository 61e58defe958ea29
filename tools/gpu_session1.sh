#!/bin/bash
# round-2 GPU session 1: parity suite, bench A/B (new vs round-1 solve kernel), launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke.txt 2>&1
echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.txt 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2_gpu_tests.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_new.json 2> gpurun_out/r2_bench_new.err
echo "bench new rc=$?"; head -c 1500 gpurun_out/r2_bench_new.json; echo
LSS_OLD_SOLVE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_old.json 2> gpurun_out/r2_bench_old.err
echo "bench old rc=$?"; head -c 600 gpurun_out/r2_bench_old.json; echo
