#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s6_gpu_tests.txt 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/r2s6_gpu_tests.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-ms 400"
timeout 600 $B > gpurun_out/r2s6_bench.json 2> gpurun_out/r2s6_bench.err; echo "bench rc=$?"
LSS_FUSE_WINDOW=1 timeout 600 $B --no-e2e > gpurun_out/r2s6_bench_fusewin.json 2> gpurun_out/r2s6_bench_fusewin.err; echo "bench fuse rc=$?"
timeout 600 python tools/sweep.py > gpurun_out/r2s6_sweep.json 2> gpurun_out/r2s6_sweep.err; echo "sweep rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s6_bench*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(b['ms_per_step'], 4), {k: round(v, 4) for k, v in b['roofline']['kernel_ms_all'].items()}, b['engine'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -c 1500 gpurun_out/r2s6_sweep.json
