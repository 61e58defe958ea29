#!/bin/bash
# final single-GPU evidence run of round 2: parity suite, bench lines (config 1, 2, streams 2, CPU arm), ncu launch list + full captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2f_smoke.txt 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_gpu_tests.txt 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2f_gpu_tests.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --config 2 --no-cpu-baseline > gpurun_out/r2f_bench_cfg2.json 2> gpurun_out/r2f_bench_cfg2.err; echo "bench cfg2 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --streams 2 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_streams2.json 2> gpurun_out/r2f_bench_streams2.err; echo "bench s2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 100 --csv --log-file gpurun_out/r2f_launches.csv python tools/profile_step.py --steps 10 > gpurun_out/r2f_ncu_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_scan|k_list_sort|k_solve|k_snowfall" -s 16 -c 4 -f -o gpurun_out/r2f_beam python tools/profile_step.py --steps 6 > gpurun_out/r2f_ncu_beam.log 2>&1; echo "ncu beam rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"
