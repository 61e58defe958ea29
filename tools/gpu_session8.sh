#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s8_gpu_tests.txt 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/r2s8_gpu_tests.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-ms 400"
timeout 600 $B > gpurun_out/r2s8_bench.json 2> gpurun_out/r2s8_bench.err; echo "bench rc=$?"
for V in "6 192" "8 160" "8 128"; do
  set -- $V
  LSS_NVCC_FLAGS="-DLSS_SOLVE_CTAS=$1 -DLSS_SOLVE_ARENA=$2" python -m lidar_snow_sim_b200.build > gpurun_out/r2s8_build_$1_$2.log 2>&1
  timeout 600 $B --no-e2e > gpurun_out/r2s8_bench_c$1_a$2.json 2> gpurun_out/r2s8_bench_c$1_a$2.err; echo "c$1 a$2 rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s8_bench*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(b['ms_per_step'], 4), {k: round(v, 4) for k, v in b['roofline']['kernel_ms_all'].items()}, b.get('theta_label_mismatch', {}).get('rate'))
    except Exception as e:
        print(f, 'ERR', e)
PY
