#!/bin/bash
# 8-GPU session, second pass: copy-engine gather on one stream per peer; e2e with DMA copy-out vs kept-rows copy-out kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
LSS_GATHER=ce timeout 400 $T bench.py --gpus $N --steps 20 --warmup 5 --min-timed-ms 400 > gpurun_out/r2n${N}b_bench_ce.json 2> gpurun_out/r2n${N}b_bench_ce.err; echo "bench ce rc=$?"
LSS_GATHER=ce LSS_PIPE_KERNEL_OUT=1 timeout 400 $T bench.py --gpus $N --steps 20 --warmup 5 --min-timed-ms 400 > gpurun_out/r2n${N}b_bench_ce_kernelout.json 2> gpurun_out/r2n${N}b_bench_ce_kernelout.err; echo "bench ce kernel-out rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/r2n${N}b_*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(b['ms_per_step'], 4), 'value', '%.3e' % b['value'], {k: round(v, 3) for k, v in b['roofline']['kernel_ms_all'].items() if 'snow' in k}, b['engine'].get('gather'), b['engine'].get('gather_fallback'), 'e2e', b['e2e'] and (round(b['e2e']['ms_per_step'], 3), '%.3e' % b['e2e']['value']))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-1500:])
PY
