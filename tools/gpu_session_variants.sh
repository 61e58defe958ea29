#!/bin/bash
# 1 GPU: solve list sorted by occluder count (default build) vs by target range; snowfall parity tests on the default build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --min-timed-ms 400"
timeout 600 python -m pytest tests/test_snowfall_gpu.py tests/test_reference_replay_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 $B > gpurun_out/r2s17_bench_byL.json 2> gpurun_out/r2s17_bench_byL.err; echo "byL rc=$?"
LSS_NVCC_FLAGS="-DLSS_SOLVE_ARENA=192" python -m lidar_snow_sim_b200.build > gpurun_out/r2s17_build_a192.log 2>&1
timeout 600 $B > gpurun_out/r2s17_bench_byL_a192.json 2> gpurun_out/r2s17_bench_byL_a192.err; echo "byL a192 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s17_bench*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(b['ms_per_step'], 4), {k: round(v, 4) for k, v in b['roofline']['kernel_ms_all'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
