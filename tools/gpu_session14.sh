#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "host or precompute or integration" > gpurun_out/r2s14_gpu_tests.txt 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2s14_gpu_tests.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-timed-ms 200"
timeout 600 $B > gpurun_out/r2s14_bench_kernelout.json 2> gpurun_out/r2s14_bench_kernelout.err; echo "kernel-out rc=$?"
LSS_PIPE_DMA_OUT=1 timeout 600 $B > gpurun_out/r2s14_bench_dmaout.json 2> gpurun_out/r2s14_bench_dmaout.err; echo "dma-out rc=$?"
timeout 600 $B --e2e-chunks 1 > gpurun_out/r2s14_bench_kernelout_c1.json 2> gpurun_out/r2s14_bench_kernelout_c1.err; echo "c1 rc=$?"
timeout 600 $B --e2e-chunks 4 > gpurun_out/r2s14_bench_kernelout_c4.json 2> gpurun_out/r2s14_bench_kernelout_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s14_bench*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(b['ms_per_step'], 4), 'e2e', round(b['e2e']['ms_per_step'], 3), 'sync', round(b['e2e']['sync_call']['ms_per_step'], 3))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-600:])
PY
