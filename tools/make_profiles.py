"""Turn the raw ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.
    python tools/make_profiles.py <tag> <launches.csv> <full.ncu-rep> [bench.json ...]
"""
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
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
d = defaultdict(list)
for r in rows[1:]:
    try:
        d[r[ki]].append(float(r[vi].replace(',', '')))
    except Exception:
        pass
tot = sum(sum(v) for v in d.values())
with open(os.path.join(out_dir, f'{tag}_launch_list.txt'), 'w') as f:
    f.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3\n')
    f.write(f'# per-launch device time (ns), cold-cache and serialised under ncu: shares, not absolutes\n')
    f.write(f'{"kernel":72s} {"launches":>8s} {"avg_us":>10s} {"share":>7s}\n')
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        f.write(f'{k[:72]:72s} {len(v):8d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / tot:7.3f}\n')

# ---- full capture of the beam stage (scan, solve, overflow launches of one step): headline metrics ---------------------------
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u = rr[0], rr[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_static', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']


def to_bytes(unit, val):
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    return float(val.replace(',', '')) * mult


traffic = 0.0
per_launch = []
with open(os.path.join(out_dir, f'{tag}_k_snowfall_ncu.txt'), 'w') as f:
    f.write('# ncu --set full --clock-control none --import-source on -k regex:k_snowfall -s 9 -c 3  python bench.py --steps 1 --warmup 3\n')
    f.write('# the three launches of the beam stage of one step: k_snowfall<24,0> scan, <24,1> solve (dominant), <128,1> overflow\n')
    for v in rr[2:]:
        if len(v) != len(h):
            continue
        f.write('\n')
        name = ''
        t = {}
        for i, col in enumerate(h):
            if col in want:
                f.write(f'{col:80s} {u[i]:16s} {v[i]}\n')
                t[col] = (u[i], v[i])
        b = to_bytes(*t['dram__bytes_read.sum']) + to_bytes(*t['dram__bytes_write.sum'])
        traffic += b
        per_launch.append({'kernel': t['Kernel Name'][1], 'dram_bytes': b, 'duration_ms_under_ncu': t['gpu__time_duration.sum'][1]})
json.dump({'k_snowfall_dram_bytes_per_launch': traffic, 'per_launch': per_launch,
           'source': f'profiles/{tag}_k_snowfall_ncu.txt',
           'note': 'dram__bytes_read.sum + dram__bytes_write.sum summed over the three launches of the beam stage of one '
                   'step (32 clouds x 131072 points): scan + solve + overflow'},
          open(os.path.join(out_dir, 'traffic.json'), 'w'), indent=1)

# ---- per source line ---------------------------------------------------------------------------------------------------------
lines = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), rep, '40', '--launch=1'],
                       capture_output=True, text=True).stdout
open(os.path.join(out_dir, f'{tag}_k_snowfall_source_lines.txt'), 'w').write(
    '# hottest source lines of the solve kernel k_snowfall<24,1> (warp-stall samples, executed warp-instructions, average active threads)\n' + lines)

for b in benches:
    txt = open(b).read().strip().splitlines()[-1]
    json.loads(txt)
    open(os.path.join(out_dir, f'{tag}_' + os.path.basename(b)), 'w').write(txt + '\n')
print('profiles written:', sorted(os.listdir(out_dir)))
