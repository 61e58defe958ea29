"""Turn the raw ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.
    python tools/make_profiles.py <tag> <launches.csv> <beam.ncu-rep> [bench.json ...]

<beam.ncu-rep>: `ncu --set full -k "regex:k_scan|k_list_sort|k_solve|k_overflow" -c 4`: the four launches of the beam stage
of one step (scan, list sort, solve, overflow)."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
benches = sys.argv[4:]
out_dir = os.path.join(ROOT, 'profiles')
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel share of the step (cold-cache, serialised: compare SHARES) -------------------------------
rows = [r for r in csv.reader(l for l in open(launches) if not l.startswith('=='))]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
d = defaultdict(list)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 'nsecond': 1.0, 'usecond': 1e3, 'msecond': 1e6}.get(r[ui], 1.0)
        d[r[ki]].append(v)
    except Exception:
        pass
tot = sum(sum(v) for v in d.values())
with open(os.path.join(out_dir, f'{tag}_launch_list.txt'), 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 100  python tools/profile_step.py --steps 10\n')
    f.write('# per-launch device time, cold-cache and serialised under ncu: shares, not absolutes\n')
    f.write(f'{"kernel":72s} {"launches":>8s} {"avg_us":>10s} {"share":>7s}\n')
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        f.write(f'{k[:72]:72s} {len(v):8d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / tot:7.3f}\n')

# ---- full capture of the beam stage: headline metrics -----------------------------------------------------------------------
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u = rr[0], rr[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_static', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.per_cycle_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']


def to_bytes(unit, val):
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    return float(val.replace(',', '')) * mult


traffic = 0.0
per_launch = []
with open(os.path.join(out_dir, f'{tag}_beam_stage_ncu.txt'), 'w') as f:
    f.write('# ncu --set full --clock-control none --import-source on -k "regex:k_scan|k_list_sort|k_solve|k_overflow" -s 16 -c 4\n')
    f.write('#     python tools/profile_step.py --steps 6\n')
    f.write('# the four launches of the beam stage of one step: k_scan, k_list_sort, k_solve (dominant), k_overflow\n')
    for v in rr[2:]:
        if len(v) != len(h):
            continue
        f.write('\n')
        t = {}
        for i, col in enumerate(h):
            if col in want:
                f.write(f'{col:80s} {u[i]:16s} {v[i]}\n')
                t[col] = (u[i], v[i])
        b = to_bytes(*t['dram__bytes_read.sum']) + to_bytes(*t['dram__bytes_write.sum'])
        traffic += b
        per_launch.append({'kernel': t['Kernel Name'][1], 'dram_bytes': b,
                           'duration_under_ncu': f"{t['gpu__time_duration.sum'][1]} {t['gpu__time_duration.sum'][0]}"})
json.dump({'beam_stage_dram_bytes_per_launch': traffic, 'per_launch': per_launch,
           'source': f'profiles/{tag}_beam_stage_ncu.txt',
           'note': 'dram__bytes_read.sum + dram__bytes_write.sum summed over the four launches of the beam stage of one step '
                   '(32 clouds x 131072 points): scan + list sort + solve + overflow'},
          open(os.path.join(out_dir, 'traffic.json'), 'w'), indent=1)

# ---- per source line + per phase -----------------------------------------------------------------------------------------------
names = [p['kernel'] for p in per_launch]
for kname, fname in (('k_solve', 'solve'), ('k_scan', 'scan')):
    idx = next((i for i, n in enumerate(names) if kname in n), None)
    if idx is None:
        continue
    lines = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), rep, '30', f'--launch={idx}'],
                           capture_output=True, text=True).stdout
    open(os.path.join(out_dir, f'{tag}_k_{fname}_source_lines.txt'), 'w').write(
        f'# hottest source lines of {kname} (warp-stall samples, executed warp-instructions, average active threads)\n' + lines)

for b in benches:
    txt = open(b).read().strip().splitlines()[-1]
    json.loads(txt)
    name = os.path.basename(b)
    for pre in ('r2f_', 'r2n2_', 'r2n8_'):
        if name.startswith(pre):
            name = name[len(pre):] if pre == 'r2f_' else name[2:]
    open(os.path.join(out_dir, f'{tag}_' + name), 'w').write(txt + '\n')
print('profiles written:', sorted(os.listdir(out_dir)))
