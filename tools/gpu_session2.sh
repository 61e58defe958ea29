#!/bin/bash
# round-2 GPU session 2: full parity suite, ncu launch list + full captures of the solve and scan kernels, config 2, CPU arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.txt 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r2_gpu_tests.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 100 --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py --steps 10 > gpurun_out/r2_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 4 -c 1 -f -o gpurun_out/r2_solve python tools/profile_step.py --steps 6 > gpurun_out/r2_ncu_solve.log 2>&1
echo "ncu solve rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_snowfall -s 8 -c 1 -f -o gpurun_out/r2_scan python tools/profile_step.py --steps 6 > gpurun_out/r2_ncu_scan.log 2>&1
echo "ncu scan rc=$?"
timeout 600 python bench.py --config 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err
echo "bench cfg2 rc=$?"; head -c 400 gpurun_out/r2_bench_cfg2.json; echo
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "bench ref rc=$?"; cat gpurun_out/r2_bench_ref.json | head -c 1500; tail -5 gpurun_out/r2_bench_ref.err
