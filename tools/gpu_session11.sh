#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s11_gpu_tests.txt 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2s11_gpu_tests.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-ms 400"
timeout 600 $B > gpurun_out/r2s11_bench.json 2> gpurun_out/r2s11_bench.err; echo "bench rc=$?"
timeout 600 $B --config 2 > gpurun_out/r2s11_bench_cfg2.json 2> gpurun_out/r2s11_bench_cfg2.err; echo "bench cfg2 rc=$?"
timeout 600 $B --streams 2 --no-e2e > gpurun_out/r2s11_bench_s2.json 2> gpurun_out/r2s11_bench_s2.err; echo "bench s2 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s11_bench*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(b['ms_per_step'], 4), {k: round(v, 4) for k, v in b['roofline']['kernel_ms_all'].items()}, 'e2e', b['e2e'] and round(b['e2e']['ms_per_step'], 3))
    except Exception as e:
        print(f, 'ERR', e)
PY
