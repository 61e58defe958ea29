#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --min-timed-ms 400"
timeout 600 $B > gpurun_out/r2s12_bench_sc8.json 2> gpurun_out/r2s12_bench_sc8.err; echo "sc8 rc=$?"
for C in 10 12; do
  LSS_NVCC_FLAGS="-DLSS_SCAN_CTAS=$C" python -m lidar_snow_sim_b200.build > gpurun_out/r2s12_build_sc$C.log 2>&1
  timeout 600 $B > gpurun_out/r2s12_bench_sc$C.json 2> gpurun_out/r2s12_bench_sc$C.err; echo "sc$C rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s12_bench*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(b['ms_per_step'], 4), {k: round(v, 4) for k, v in b['roofline']['kernel_ms_all'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
