#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s5_gpu_tests.txt 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r2s5_gpu_tests.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-ms 400"
timeout 600 $B > gpurun_out/r2s5_bench.json 2> gpurun_out/r2s5_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 100 --csv --log-file gpurun_out/r2s5_launches.csv python tools/profile_step.py --steps 10 > gpurun_out/r2s5_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 4 -c 1 -f -o gpurun_out/r2s5_solve python tools/profile_step.py --steps 6 > gpurun_out/r2s5_ncu_solve.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan -s 4 -c 1 -f -o gpurun_out/r2s5_scan python tools/profile_step.py --steps 6 > gpurun_out/r2s5_ncu_scan.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s5_bench*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(b['ms_per_step'], 4), {k: round(v, 4) for k, v in b['roofline']['kernel_ms_all'].items()}, 'e2e', round(b['e2e']['ms_per_step'],3))
    except Exception as e:
        print(f, 'ERR', e)
PY
