#!/bin/bash
# 8-GPU session: raw host-side probe + engine pipeline per rank, bench with the copy-engine gather (+ NCCL gather for comparison)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 300 $T tools/e2e_probe_ranks.py --bind 0 > gpurun_out/r2n${N}_probe_unbound.json 2> gpurun_out/r2n${N}_probe_unbound.err; echo "probe unbound rc=$?"
timeout 300 $T tools/e2e_probe_ranks.py --bind 1 --engine 0 > gpurun_out/r2n${N}_probe_bound.json 2> gpurun_out/r2n${N}_probe_bound.err; echo "probe bound rc=$?"
LSS_GATHER=ce timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --min-timed-ms 400 > gpurun_out/r2n${N}_bench_ce.json 2> gpurun_out/r2n${N}_bench_ce.err; echo "bench ce rc=$?"
LSS_GATHER=nccl timeout 600 $T bench.py --gpus $N --steps 20 --warmup 5 --min-timed-ms 400 --no-e2e > gpurun_out/r2n${N}_bench_nccl.json 2> gpurun_out/r2n${N}_bench_nccl.err; echo "bench nccl rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/r2n${N}_*.json')):
    try:
        b = json.loads(open(f).read().strip().splitlines()[-1])
        if 'ms_per_step' in b:
            print(f, 'ms', round(b['ms_per_step'], 4), 'value', '%.3e' % b['value'], {k: round(v, 3) for k, v in b['roofline']['kernel_ms_all'].items() if 'snow' in k}, b['engine'].get('gather'), b['engine'].get('gather_fallback'), 'e2e', b['e2e'] and (round(b['e2e']['ms_per_step'], 3), '%.3e' % b['e2e']['value']))
        else:
            print(f, {k: (v if not isinstance(v, list) else [round(x, 2) for x in v]) for k, v in b.items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-1500:])
PY
