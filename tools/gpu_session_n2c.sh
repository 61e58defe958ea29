#!/bin/bash
# 2 GPUs: push kernel with the chunk cursor: single-GPU test, every gather kind on real peers, one bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 300 python -m pytest tests/test_snowfall_gpu.py -q -m gpu -k "gather_push" 2>&1 | tail -2
timeout 300 $T tools/check_gather_ranks.py 2> gpurun_out/r2n${N}e_gather_check.err | grep '^{' > gpurun_out/r2n${N}e_gather_check.json; echo "gather check rc=${PIPESTATUS[0]}"; cat gpurun_out/r2n${N}e_gather_check.json
LSS_GATHER=push timeout 300 $T bench.py --gpus $N --steps 20 --warmup 5 --min-timed-ms 400 --no-e2e 2> gpurun_out/r2n${N}e_bench_push.err | grep '^{' > gpurun_out/r2n${N}e_bench_push.json; echo "bench rc=${PIPESTATUS[0]}"
python - <<PY
import json
b = json.loads(open('gpurun_out/r2n${N}e_bench_push.json').read().strip().splitlines()[-1])
print('ms', round(b['ms_per_step'], 4), 'value', '%.3e' % b['value'], {k: round(v, 3) for k, v in b['roofline']['kernel_ms_all'].items() if 'snow' in k}, b['engine'])
PY
