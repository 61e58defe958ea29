#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --min-timed-ms 400"
for S in 1 2; do
  timeout 600 $B --streams $S > gpurun_out/r2s7_bench_c1_s$S.json 2> gpurun_out/r2s7_bench_c1_s$S.err; echo "c1 s$S rc=$?"
  timeout 600 $B --streams $S --config 2 > gpurun_out/r2s7_bench_c2_s$S.json 2> gpurun_out/r2s7_bench_c2_s$S.err; echo "c2 s$S rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s7_bench*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(b['ms_per_step'], 4), round(b['ms_per_step_min'],4), round(b['ms_per_step_max'],4), {k: round(v, 4) for k, v in b['roofline']['kernel_ms_all'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
