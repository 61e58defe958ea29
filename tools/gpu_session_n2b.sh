#!/bin/bash
# N GPUs: gather kinds compared (push kernel with / without multicast, block counts), after the single-GPU push test
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 300 python -m pytest tests/test_snowfall_gpu.py -q -m gpu -k "gather_push or kat_channel" 2>&1 | tail -3
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 $T bench.py --gpus $N --steps 20 --warmup 5 --min-timed-ms 400 --no-e2e > gpurun_out/r2n${N}c_bench_$name.json 2> gpurun_out/r2n${N}c_bench_$name.err; echo "bench $name rc=$?"
}
run push LSS_GATHER=push
run push_nomc LSS_GATHER=push LSS_GATHER_MULTICAST=0
run push_nomc_b32 LSS_GATHER=push LSS_GATHER_MULTICAST=0 LSS_GATHER_BLOCKS=32
run push_nomc_b296 LSS_GATHER=push LSS_GATHER_MULTICAST=0 LSS_GATHER_BLOCKS=296
run push_b16 LSS_GATHER=push LSS_GATHER_BLOCKS=16
run ce LSS_GATHER=ce
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/r2n${N}c_*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(b['ms_per_step'], 4), 'value', '%.3e' % b['value'], {k: round(v, 3) for k, v in b['roofline']['kernel_ms_all'].items() if 'snow' in k}, b['engine'].get('gather'), 'mc', b['engine'].get('gather_multicast'), b['engine'].get('gather_fallback'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-1500:])
PY
