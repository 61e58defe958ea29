#!/bin/bash
# 1 GPU: full GPU suite, compute-sanitizer over the small parity tests of the round-2 kernels, bench lines (both arms + config 2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s15_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s15_pytest.log
SEL='test_golden_kat_channel or test_golden_channel_cases or test_golden_augment_api or test_batch_ragged_and_empty or test_errors or test_degenerate_rows or test_beams_with_dozens or test_more_than_128 or test_device_fixed_seed or test_batch_matches_the_oracle or test_vs_oracle_given_plane or test_prepass_replays_reference or test_wet_ground_replays_reference'
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -k "$SEL and not augment_full and not augment_cfg1" > gpurun_out/r2s15_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|error" gpurun_out/r2s15_memcheck.log | head -20
RSEL='test_golden_kat_channel or test_golden_channel_cases or test_beams_with_dozens or test_device_fixed_seed or (test_batch_matches_the_oracle and 3000)'
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests -m gpu -q -k "$RSEL" > gpurun_out/r2s15_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r2s15_racecheck.log | head -20
timeout 600 python bench.py > gpurun_out/r2s15_bench.json 2> gpurun_out/r2s15_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
b = json.loads(open('gpurun_out/r2s15_bench.json').read().strip().splitlines()[-1])
print('ms', round(b['ms_per_step'], 4), 'frac', round(b['roofline']['frac'], 4), 'traffic', b['roofline']['traffic'], 'e2e', round(b['e2e']['ms_per_step'], 3), 'cpu', b['cpu_baseline']['value'])
PY
