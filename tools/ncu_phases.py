"""Phase summary of a kernel from an ncu report (needs -lineinfo and --import-source on): sums warp-stall samples,
executed warp-instructions and active lanes between marker comments of the source file.
    python tools/ncu_phases.py <report.ncu-rep> <source file name> <first line> "<name>=<marker text>" ..."""
import csv
import subprocess
import sys

rep, fname, first = sys.argv[1], sys.argv[2], int(sys.argv[3])
markers = [m.split('=', 1) for m in sys.argv[4:]]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur, lines = None, '', []
for r in rows:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and r and r[0].isdigit() and len(r) == len(hdr) and cur == fname:
        def f(name):
            try:
                return float(r[hdr.index(name)])
            except ValueError:
                return 0.0
        lines.append((int(r[0]), r[1], f('# Samples'), f('Instructions Executed'), f('Thread Instructions Executed')))
tot_s = sum(l[2] for l in lines) or 1
tot_i = sum(l[3] for l in lines) or 1
idx = []
for name, pat in markers:
    c = [l[0] for l in lines if pat in l[1] and l[0] >= first]
    if c:
        idx.append((name, c[0]))
idx.sort(key=lambda t: t[1])
bounds = [i for _, i in idx] + [10 ** 9]
print(f'{fname}: {tot_i:.3e} warp-instructions, {tot_s:.0f} samples in this file')
pre = [l for l in lines if l[0] < idx[0][1]]
print(f'{"(device helpers above the kernel)":40s} samples {sum(l[2] for l in pre) / tot_s * 100:5.1f}%  inst '
      f'{sum(l[3] for l in pre) / tot_i * 100:5.1f}%  lanes {sum(l[4] for l in pre) / max(sum(l[3] for l in pre), 1):5.1f}')
for k, (name, start) in enumerate(idx):
    seg = [l for l in lines if start <= l[0] < bounds[k + 1]]
    s, i, t = sum(l[2] for l in seg), sum(l[3] for l in seg), sum(l[4] for l in seg)
    print(f'{name:40s} samples {s / tot_s * 100:5.1f}%  inst {i / tot_i * 100:5.1f}%  lanes {t / max(i, 1):5.1f}')
